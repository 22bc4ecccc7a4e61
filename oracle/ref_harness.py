"""ORACLE — test infrastructure only.

Executes the *unmodified* reference source files from /root/reference in a process that has no
TensorFlow: a stub `tensorflow` module satisfies the module-level uses
(cosmopower/cosmopower_NN.py:4-8,14 and the @tf.function decorators), instances are made with
object.__new__ and filled by the reference's own `restore` (cosmopower_NN.py:310-327,
cosmopower_PCAplusNN.py:306-326), and then the reference's own `predictions_np` /
`ten_to_predictions_np` run.  Only usable where /root/reference exists (the build container);
it generates tests/golden/ and validates oracle_np.py.  Nothing here is copied from the reference.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("COSMOPOWER_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "cosmopower", "cosmopower_NN.py"))


def _install_stub_tf():
    if "tensorflow" in sys.modules and not getattr(sys.modules["tensorflow"], "_cpb200_stub", False):
        return
    tf = types.ModuleType("tensorflow")
    tf._cpb200_stub = True
    tf.float32 = "float32"

    def function(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f
    tf.function = function
    keras = types.ModuleType("tensorflow.keras")
    keras.Model = object
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    chain = ["tensorflow.python", "tensorflow.python.training",
             "tensorflow.python.training.tracking",
             "tensorflow.python.training.tracking.data_structures"]
    parent = tf
    for name in chain:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
        setattr(parent, name.rsplit(".", 1)[1], mod)
        parent = mod
    parent.ListWrapper = list
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except Exception:
            tq = types.ModuleType("tqdm")
            tq.trange = range
            sys.modules["tqdm"] = tq


_classes = {}


def reference_class(name):
    """name in {'cosmopower_NN', 'cosmopower_PCAplusNN'} -> the reference class object."""
    if name not in _classes:
        _install_stub_tf()
        path = os.path.join(REF_ROOT, "cosmopower", name + ".py")
        spec = importlib.util.spec_from_file_location("_ref_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _classes[name] = getattr(mod, name)
    return _classes[name]


def restore_reference(name, filename_without_suffix):
    cls = reference_class(name)
    obj = object.__new__(cls)
    cls.restore(obj, filename_without_suffix)
    return obj
