"""Architecture-identical SYNTHETIC STAND-INS for the three CP_paper emulators whose pickles are
stripped from the reference checkout (cmb_TT_NN, cmb_EE_NN, cmb_TE_PCAplusNN; reference
.MISSING_LARGE_BLOBS:2-4).  They reuse the real trained trunk of cmb_PP_PCAplusNN (a genuine
6 -> 4x512 network on the same parameters and prior) and add a seeded synthetic head, and are
written in the reference's own 15-/19-item pickle layout (including the TensorFlow ListWrapper
global) so the normal `restore` path reads them.  Every report must label them "synthetic
stand-in": they exercise shapes, throughput and numerics, not cosmology.

Architectures: TT 6->512x4->2507 (reference docs .../getting_started_with_cosmopower_NN.md:61,72-73,83);
TE 6->512x4->512 PCA->2507 (.../getting_started_with_cosmopower_PCAplusNN.md:68,79-80,90,102);
EE assumed equal to TT (hidden sizes unverified).
"""
import os
import pickle
import sys
import types

import numpy as np

from . import _loader

HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_DIR = os.path.join(HERE, "trained_models")
STANDIN_DIR = os.path.join(MODEL_DIR, "standins")

REAL_MODELS = {
    "cmb_PP_PCAplusNN": ("cosmopower_PCAplusNN", os.path.join(MODEL_DIR, "CP_paper", "CMB", "cmb_PP_PCAplusNN")),
    "PKLIN_NN": ("cosmopower_NN", os.path.join(MODEL_DIR, "CP_paper", "PK", "PKLIN_NN")),
    "PKNLBOOST_NN": ("cosmopower_NN", os.path.join(MODEL_DIR, "CP_paper", "PK", "PKNLBOOST_NN")),
}
STANDIN_MODELS = {
    "cmb_TT_NN": ("cosmopower_NN", os.path.join(STANDIN_DIR, "cmb_TT_NN")),
    "cmb_EE_NN": ("cosmopower_NN", os.path.join(STANDIN_DIR, "cmb_EE_NN")),
    "cmb_TE_PCAplusNN": ("cosmopower_PCAplusNN", os.path.join(STANDIN_DIR, "cmb_TE_PCAplusNN")),
}
ALL_MODELS = dict(REAL_MODELS, **STANDIN_MODELS)

_TF_LW_MODULE = "tensorflow.python.training.tracking.data_structures"


def _dump_reference_layout(path, items):
    """pickle.dump a list whose list-typed members are reduced through the TF ListWrapper global,
    byte-compatible with what a tf.keras.Model-based `save` emits."""
    created = []
    parts = _TF_LW_MODULE.split(".")
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            created.append(name)

    class ListWrapper(list):
        def __reduce__(self):
            return (ListWrapper, (list(self),))
    ListWrapper.__module__ = _TF_LW_MODULE
    ListWrapper.__qualname__ = "ListWrapper"
    mod = sys.modules[_TF_LW_MODULE]
    had = getattr(mod, "ListWrapper", None)
    mod.ListWrapper = ListWrapper
    try:
        wrapped = [ListWrapper(x) if isinstance(x, list) else x for x in items]
        with open(path, "wb") as f:
            pickle.dump(wrapped, f, protocol=3)
    finally:
        if had is None:
            del mod.ListWrapper
        else:
            mod.ListWrapper = had
        for name in created:
            sys.modules.pop(name, None)


def load_attrs(kind, path_without_suffix):
    fields = _loader.NN_FIELDS if kind == "cosmopower_NN" else _loader.PCANN_FIELDS
    if os.path.isfile(path_without_suffix + ".pkl"):
        return _loader.attributes_from_list(_loader.load_pickle_list(path_without_suffix + ".pkl"), fields)
    return _loader.load_npz(path_without_suffix + ".npz", fields)


def _trunk_from_pp():
    pp = load_attrs(*REAL_MODELS["cmb_PP_PCAplusNN"])
    return pp


def _last_hidden_rms(pp, seed=7):
    """RMS of the trunk's last hidden activation over in-prior inputs (to scale synthetic heads)."""
    from .priors import sample_prior
    x = sample_prior(pp["parameters"], 2048, seed)
    h = (x - pp["parameters_mean_"]) / pp["parameters_std_"]
    with np.errstate(over="ignore"):
        for i in range(4):
            a = h @ pp["W_"][i].astype(np.float64) + pp["b_"][i]
            h = (pp["betas_"][i] + (1. - pp["betas_"][i]) / (1. + np.exp(-pp["alphas_"][i] * a))) * a
    return float(np.sqrt(np.mean(h * h)))


def _tt_feature_stats():
    """Smooth log10 C_ell mean/std curves over ell in [2,2508] (shape of a CAMB TT spectrum:
    range about [-18.4,-8.8] like reference cosmopower/tests/tt_data, but synthetic)."""
    ell = np.arange(2, 2509, dtype=np.float64)
    mean = -9.6 - 2.0 * np.log10(ell / 2.0 + 1.0) - 1.8 * (ell / 2508.0) ** 1.5 \
        + 0.12 * np.sin(ell / 95.0) * np.exp(-ell / 1500.0)
    std = 0.35 + 0.25 * (ell / 2508.0) + 0.05 * np.cos(ell / 130.0)
    return ell.astype(np.int64), mean.astype(np.float32), std.astype(np.float32)


def build_standins(force=False):
    """Create the three stand-in model files (idempotent, deterministic)."""
    os.makedirs(STANDIN_DIR, exist_ok=True)
    done = all(os.path.isfile(p + ".pkl") for _, p in STANDIN_MODELS.values())
    if done and not force:
        return
    pp = _trunk_from_pp()
    rms = _last_hidden_rms(pp)
    modes, tt_mean, tt_std = _tt_feature_stats()
    trunk_W = [np.array(w, dtype=np.float32) for w in pp["W_"][:4]]
    trunk_b = [np.array(b, dtype=np.float32) for b in pp["b_"][:4]]
    alphas = [np.array(a, dtype=np.float32) for a in pp["alphas_"]]
    betas = [np.array(b, dtype=np.float32) for b in pp["betas_"]]
    params = list(pp["parameters"])

    for name, seed, shift in (("cmb_TT_NN", 20260101, 0.0), ("cmb_EE_NN", 20260102, -1.6)):
        rng = np.random.default_rng(seed)
        head = (rng.standard_normal((512, 2507)) / (8.0 * np.sqrt(512.0) * rms)).astype(np.float32)
        bias = (0.1 * rng.standard_normal(2507)).astype(np.float32)
        items = [trunk_W + [head], trunk_b + [bias], alphas, betas,
                 np.array(pp["parameters_mean_"], dtype=np.float32),
                 np.array(pp["parameters_std_"], dtype=np.float32),
                 (tt_mean + np.float32(shift)).astype(np.float32), tt_std,
                 6, params, 2507, modes, [512, 512, 512, 512], 5, [6, 512, 512, 512, 512, 2507]]
        _dump_reference_layout(STANDIN_MODELS[name][1] + ".pkl", items)

    rng = np.random.default_rng(20260103)
    n_pcas = 512
    head = (rng.standard_normal((512, n_pcas)) / (8.0 * np.sqrt(512.0) * rms)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(n_pcas)).astype(np.float32)
    q, _ = np.linalg.qr(rng.standard_normal((2507, n_pcas)))
    basis = np.ascontiguousarray(q.T).astype(np.float32)                 # (n_pcas, n_modes), orthonormal rows
    pca_std = (50.0 * 10.0 ** (-7.0 * np.arange(n_pcas) / (n_pcas - 1.0))).astype(np.float32)
    pca_mean = (1e-16 * rng.standard_normal(n_pcas)).astype(np.float32)
    ell = modes.astype(np.float64)
    te_std = (2e-11 / (1.0 + (ell / 600.0) ** 2) + 1e-14).astype(np.float32)
    te_mean = (5e-12 * np.sin(ell / 110.0) / (1.0 + (ell / 800.0) ** 2)).astype(np.float32)
    items = [trunk_W + [head], trunk_b + [bias], alphas, betas,
             np.array(pp["parameters_mean_"], dtype=np.float32),
             np.array(pp["parameters_std_"], dtype=np.float32),
             pca_mean, pca_std, te_mean, te_std,
             params, 6, modes, 2507, n_pcas, basis,
             [512, 512, 512, 512], 5, [6, 512, 512, 512, 512, n_pcas]]
    _dump_reference_layout(STANDIN_MODELS["cmb_TE_PCAplusNN"][1] + ".pkl", items)


def model_path(name):
    """(class name, restore_filename) for any of the six CP_paper emulators; builds stand-ins on demand."""
    if name in STANDIN_MODELS:
        build_standins()
    return ALL_MODELS[name]
