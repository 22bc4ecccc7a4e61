"""The ctypes stub INTEGRATION.md shows a reference maintainer (section 2) is executed as written - only the library path is
pointed at the in-tree build - against a restored model: it must create an engine from the model's public attributes and
reproduce the drop-in class's own results through cpb200_forward_host / cpb200_forward_host_f64."""
import os
import re
import types

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = [b for b in blocks if "cosmopower/_b200.py" in b]
    assert len(src) == 1, "INTEGRATION.md no longer holds the _b200.py stub"
    return src[0]


def test_stub_is_valid_python_and_names_only_exported_symbols():
    src = _stub_source()
    compile(src, "INTEGRATION.md:_b200.py", "exec")
    header = open(os.path.join(REPO, "include", "cpb200.h")).read()
    for sym in set(re.findall(r"_lib\.(cpb200_[a-z0-9_]+)", src)):
        assert re.search(r"\b%s\s*\(" % sym, header), sym


@pytest.mark.gpu
def test_stub_runs_against_the_library():
    import cosmopower_b200 as cp
    from cosmopower_b200 import _engine, priors, standins
    src = _stub_source().replace('ctypes.CDLL("libcpb200.so")', "ctypes.CDLL(%r)" % _engine.LIB_PATH)
    mod = types.ModuleType("_b200")
    exec(compile(src, "INTEGRATION.md:_b200.py", "exec"), mod.__dict__)
    # the error path of the stub: a bad precision code must come back as the library's message, not as a crash
    bad = getattr(cp, standins.model_path("PKLIN_NN")[0])(restore=True, restore_filename=standins.model_path("PKLIN_NN")[1], device=0)
    with pytest.raises(RuntimeError) as err:
        mod.create(bad, precision=77)
    assert "precision" in str(err.value).lower()
    bad.close()
    for name in ("PKLIN_NN", "cmb_PP_PCAplusNN"):
        kind, path = standins.model_path(name)
        model = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        handle = mod.create(model)
        x64 = priors.sample_prior(model.parameters, 300, seed=8)
        y64 = mod.forward(handle, x64, model.n_modes)
        assert y64.dtype == np.float64 and np.array_equal(y64, model.forward_pass_np(x64))
        x32 = x64.astype(np.float32)
        y32 = mod.forward(handle, x32, model.n_modes, ten_to=True)
        d32 = {k: x32[:, i] for i, k in enumerate(model.parameters)}
        assert y32.dtype == np.float32 and np.array_equal(y32, model.ten_to_predictions_np(d32))
        assert mod._lib.cpb200_destroy(handle) == 0
        model.close()
