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

# The upstream checkout when it exists (the build container); else the byte copies of the two class files that
# __graft_entry__.build() stages in the git-ignored oracle/_ref/ so that the executed reference travels to the GPU box.
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _ref_root():
    env = os.environ.get("COSMOPOWER_REFERENCE")
    for root in ([env] if env else []) + ["/root/reference", _STAGED]:
        if os.path.isfile(os.path.join(root, "cosmopower", "cosmopower_NN.py")):
            return root
    return "/root/reference"


REF_ROOT = _ref_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "cosmopower", "cosmopower_NN.py"))


def stage(dst=_STAGED, src="/root/reference"):
    """Copy the two reference class files (unmodified) into oracle/_ref/ - git-ignored, shipped by gpurun - so that
    bench.py's CPU arm can execute the reference itself on the GPU box.  Called by __graft_entry__.build()."""
    import shutil
    os.makedirs(os.path.join(dst, "cosmopower"), exist_ok=True)
    for name in ("cosmopower_NN.py", "cosmopower_PCAplusNN.py"):
        shutil.copyfile(os.path.join(src, "cosmopower", name), os.path.join(dst, "cosmopower", name))
    for name in ("LICENSE",):
        if os.path.isfile(os.path.join(src, name)):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))


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


def reference_tf_twin(name, filename_without_suffix, method, x):
    """Execute one of the reference's TensorFlow-side methods (`forward_pass_tf`, `predictions_tf`, `ten_to_predictions_tf`;
    cosmopower_NN.py:160-211, cosmopower_PCAplusNN.py:163-239) on a restored reference object, with the handful of tf.*
    calls they make bound to NumPy and the TF-side attributes (W, b, ... - variables that only __init__ creates) set to
    the restored arrays.  The arithmetic is the reference's own source, evaluated in x's dtype."""
    import numpy as np
    obj = restore_reference(name, filename_without_suffix)
    tf = type(obj).restore.__globals__["tf"]          # the module object the reference source bound at import
    np_ops = {"divide": lambda a, b: a / b, "subtract": lambda a, b: a - b, "add": lambda a, b: a + b,
              "multiply": lambda a, b: a * b, "matmul": lambda a, b: np.matmul(a, b),
              "sigmoid": lambda v: 1. / (1. + np.exp(-v)), "pow": lambda a, b: np.power(a, b)}
    saved = {k: getattr(tf, k, None) for k in np_ops}
    for k, f in np_ops.items():
        setattr(tf, k, f)
    try:
        obj.W, obj.b, obj.alphas, obj.betas = obj.W_, obj.b_, obj.alphas_, obj.betas_
        obj.parameters_mean, obj.parameters_std = obj.parameters_mean_, obj.parameters_std_
        obj.features_mean, obj.features_std = obj.features_mean_, obj.features_std_
        if hasattr(obj, "pca_mean_"):
            obj.pca_mean, obj.pca_std, obj.pca_transform_matrix = obj.pca_mean_, obj.pca_std_, obj.pca_transform_matrix_
        with np.errstate(over="ignore"):
            return getattr(obj, method)(x)
    finally:
        for k, f in saved.items():
            if f is None:
                delattr(tf, k)
            else:
                setattr(tf, k, f)


# ---------------------------------------------------------------------------------------------------
# Planck-lite likelihood: execute the unmodified reference source with a NumPy-backed TensorFlow stub
# ---------------------------------------------------------------------------------------------------
def _numpy_tf():
    """A tensorflow / tensorflow_probability look-alike covering exactly the calls made by
    cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py, on float32 NumPy arrays."""
    import numpy as np

    tf = types.ModuleType("tensorflow")
    tf._cpb200_stub = True
    tf.float32 = np.float32
    tf.string = str

    def function(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f
    tf.function = function
    tf.convert_to_tensor = lambda x: np.asarray(x)
    tf.reshape = lambda x, shape: np.reshape(x, [int(s) for s in shape])
    tf.shape = lambda x: np.array(np.shape(x))
    tf.constant = lambda x, shape=None: np.asarray(x) if shape is None else np.asarray(x).reshape(shape)
    tf.concat = lambda xs, axis=0: np.concatenate(xs, axis=axis)
    tf.eye = lambda n: np.eye(int(n), dtype=np.float32)
    tf.transpose = lambda x: np.transpose(x)
    tf.zeros = lambda shape: np.zeros(shape, dtype=np.float32)
    tf.scalar_mul = lambda s, x: (np.float32(s) * x).astype(np.float32)
    tf.tile = lambda x, reps: np.tile(x, reps)
    tf.gather = lambda x, idx, axis=0: np.take(x, idx, axis=axis)
    tf.divide = lambda a, b: a / b
    tf.square = np.square
    tf.subtract = lambda a, b: a - b
    tf.add = lambda a, b: a + b
    tf.matmul = lambda a, b: np.matmul(a, b)
    tf.where = np.where
    tf.equal = lambda a, b: a == b

    linalg = types.ModuleType("tensorflow.linalg")
    linalg.cholesky = np.linalg.cholesky
    linalg.cholesky_solve = lambda chol, rhs: np.linalg.solve(chol.T, np.linalg.solve(chol, rhs)).astype(chol.dtype)
    linalg.diag_part = np.diagonal
    tf.linalg = linalg

    math = types.ModuleType("tensorflow.math")
    math.multiply = lambda a, b: a * b
    math.log = np.log
    math.is_inf = np.isinf

    def segment_sum(data, segment_ids):
        out = np.zeros((int(np.max(segment_ids)) + 1,) + data.shape[1:], dtype=data.dtype)
        np.add.at(out, np.asarray(segment_ids), data)
        return out
    math.segment_sum = segment_sum
    tf.math = math

    class DenseHashTable:
        def __init__(self, key_dtype=None, value_dtype=None, empty_key=None, deleted_key=None, default_value=None):
            self._d, self._default = {}, default_value

        def insert(self, keys, values):
            for k, v in zip(list(keys), values):
                self._d[str(k)] = v

        def lookup(self, keys):
            return np.stack([self._d.get(str(k), self._default) for k in np.asarray(keys).reshape(-1)])
    lookup = types.ModuleType("tensorflow.lookup")
    lookup.experimental = types.ModuleType("tensorflow.lookup.experimental")
    lookup.experimental.DenseHashTable = DenseHashTable
    tf.lookup = lookup

    tfp = types.ModuleType("tensorflow_probability")
    dist = types.ModuleType("tensorflow_probability.distributions")

    class Uniform:
        def __init__(self, low, high):
            self.low, self.high = low, high

        def prob(self, x):
            return np.where((x >= self.low) & (x <= self.high), 1.0 / (self.high - self.low), 0.0)

    class Normal:
        def __init__(self, loc, scale):
            self.loc, self.scale = loc, scale

        def prob(self, x):
            return np.exp(-0.5 * ((x - self.loc) / self.scale) ** 2) / (self.scale * np.sqrt(2 * np.pi))

    class Blockwise:
        def __init__(self, parts):
            self.parts = parts

        def prob(self, x):
            p = np.ones(x.shape[0])
            for i, d in enumerate(self.parts):
                p = p * d.prob(x[:, i])
            return p.astype(np.float32)
    dist.Uniform, dist.Normal, dist.Blockwise = Uniform, Normal, Blockwise
    tfp.distributions = dist
    return tf, tfp


def reference_planck_lite_class():
    """The reference class tf_planck2018_lite, imported from its own source file."""
    if "tf_planck2018_lite" not in _classes:
        tf, tfp = _numpy_tf()
        saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow_probability", "cosmopower")}
        sys.modules["tensorflow"] = tf
        sys.modules["tensorflow_probability"] = tfp
        sys.modules["cosmopower"] = types.ModuleType("cosmopower")       # imported by the file, not used by the class
        try:
            path = os.path.join(REF_ROOT, "cosmopower", "likelihoods", "tf_planck2018_lite", "tf_planck2018_lite.py")
            spec = importlib.util.spec_from_file_location("_ref_tf_planck2018_lite", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _classes["tf_planck2018_lite"] = mod.tf_planck2018_lite
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return _classes["tf_planck2018_lite"]


def reference_planck_lite_path():
    return os.path.join(REF_ROOT, "cosmopower", "likelihoods", "tf_planck2018_lite")
