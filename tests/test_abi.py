"""The C-ABI library builds, loads and exports exactly what include/cpb200.h declares.
No compute is attempted here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from cosmopower_b200 import _engine

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "cpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cpb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_bound_and_exported(built_library):
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    bound = sorted(name for name, _, _ in _engine.SYMBOLS)
    assert declared == bound
    for name in declared:
        assert getattr(built_library, name) is not None


def test_abi_version(built_library):
    assert built_library.cpb200_abi_version() == 3


def test_no_cpu_fallback_without_device(built_library):
    if built_library.cpb200_device_count() > 0:
        pytest.skip("a B200 is visible")
    import cosmopower_b200 as cp
    from cosmopower_b200 import standins
    _, path = standins.model_path("PKLIN_NN")
    m = cp.cosmopower_NN(restore=True, restore_filename=path)
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        m.predictions_np({k: [0.1] for k in m.parameters})


def test_create_rejects_bad_description(built_library):
    h = ctypes.c_void_p()
    rc = built_library.cpb200_create(None, 0, 1, ctypes.byref(h))
    assert rc == -1 and built_library.cpb200_last_error()
    d = _engine.ModelDesc()
    rc = built_library.cpb200_create(ctypes.byref(d), 0, 7, ctypes.byref(h))
    assert rc == -1
    assert built_library.cpb200_destroy(None) == 0


def test_product_package_never_imports_oracle():
    """Only tests/ (incl. the manual checkers in tests/manual/), __graft_entry__.smoke() and bench.py's CPU arm may touch
    oracle/: neither the package nor the developer tools do."""
    for top in ("cosmopower_b200", "tools", "include"):
        for root, _, files in os.walk(os.path.join(REPO, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    src = open(os.path.join(root, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(root, f)
