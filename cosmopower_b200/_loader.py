"""TensorFlow-free reader for CosmoPower model files.

The reference stores a trained emulator as one pickled Python list
(`cosmopower_NN.save`, reference cosmopower/cosmopower_NN.py:277-306, order at :288-302;
`cosmopower_PCAplusNN.save`, cosmopower/cosmopower_PCAplusNN.py:267-302, order at :278-297).
Because the saving object was a `tf.keras.Model`, every list-typed attribute was wrapped and
the stream names the global
`tensorflow.python.training.tracking.data_structures.ListWrapper`.  This module reads those
bytes without TensorFlow by mapping that single global to `list`, and refuses every global
that is not needed to rebuild plain numpy arrays (a pickle is otherwise arbitrary code).

It also defines the flat `.npz` container that `restore()` accepts when no
`.pkl` exists: same attribute names, one array per attribute, no code execution on load.
"""
import io
import pickle

import numpy as np

NN_FIELDS = ("W_", "b_", "alphas_", "betas_",
             "parameters_mean_", "parameters_std_",
             "features_mean_", "features_std_",
             "n_parameters", "parameters",
             "n_modes", "modes",
             "n_hidden", "n_layers", "architecture")          # cosmopower_NN.py:322-327

PCANN_FIELDS = ("W_", "b_", "alphas_", "betas_",
                "parameters_mean_", "parameters_std_",
                "pca_mean_", "pca_std_",
                "features_mean_", "features_std_",
                "parameters", "n_parameters",
                "modes", "n_modes",
                "n_pcas", "pca_transform_matrix_",
                "n_hidden", "n_layers", "architecture")        # cosmopower_PCAplusNN.py:318-325

_LISTWRAPPER_MODULES = (
    "tensorflow.python.training.tracking.data_structures",
    "tensorflow.python.trackable.data_structures",
)

_ALLOWED = {
    ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"),
    ("numpy", "dtype"),
    ("builtins", "list"),
    ("builtins", "tuple"),
    ("builtins", "dict"),
    ("collections", "OrderedDict"),
}


class _SafeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == "ListWrapper" and module in _LISTWRAPPER_MODULES:
            return list
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(
            "refusing to load global %s.%s from a CosmoPower model file" % (module, name))


def load_pickle_list(path):
    """Return the top-level list stored in a reference `.pkl` model file."""
    with open(path, "rb") as f:
        data = f.read()
    obj = _SafeUnpickler(io.BytesIO(data)).load()
    if not isinstance(obj, (list, tuple)):
        raise ValueError("model file %s does not hold a list of attributes" % path)
    return list(obj)


def attributes_from_list(items, fields):
    """Same unpacking (and same ValueError on arity mismatch) as the reference `restore`."""
    if len(items) != len(fields):
        if len(items) > len(fields):
            raise ValueError("too many values to unpack (expected %d)" % len(fields))
        raise ValueError("not enough values to unpack (expected %d, got %d)"
                         % (len(fields), len(items)))
    return dict(zip(fields, items))


def save_npz(path, attrs, fields):
    """Write the flat container: per-layer arrays are stored as W_0, W_1, ..."""
    out = {"__fields__": np.array(list(fields))}
    for k in fields:
        v = attrs[k]
        if k in ("W_", "b_", "alphas_", "betas_"):
            out["len_" + k] = np.int64(len(v))
            for i, a in enumerate(v):
                out["%s%d" % (k, i)] = np.asarray(a)
        elif k == "parameters":
            out[k] = np.array([str(s) for s in v])
        elif k in ("n_hidden", "architecture"):
            out[k] = np.asarray(list(v), dtype=np.int64)
        else:
            out[k] = np.asarray(v)
    np.savez(path, **out)


def load_npz(path, fields):
    z = np.load(path, allow_pickle=False)
    stored = tuple(str(s) for s in z["__fields__"])
    if stored != tuple(fields):
        # mirror the reference's failure when the wrong class opens a file
        raise ValueError("model file %s holds %d attributes, expected %d"
                         % (path, len(stored), len(fields)))
    attrs = {}
    for k in fields:
        if k in ("W_", "b_", "alphas_", "betas_"):
            attrs[k] = [z["%s%d" % (k, i)] for i in range(int(z["len_" + k]))]
        elif k == "parameters":
            attrs[k] = [str(s) for s in z[k]]
        elif k in ("n_hidden", "architecture"):
            attrs[k] = [int(x) for x in z[k]]
        elif k in ("n_parameters", "n_modes", "n_layers", "n_pcas"):
            attrs[k] = int(z[k])
        else:
            attrs[k] = z[k]
    return attrs
