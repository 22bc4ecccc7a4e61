"""Fused multi-emulator evaluation: several emulators on the same rows in ONE launch.

Two reference call patterns are covered:

* emulators whose log-spectra are summed before 10**: the reference combines the linear matter power spectrum and the
  nonlinear boost on the host, `10.**(lin.predictions_np(p_lin) + boost.predictions_np(p_boost))`
  (cosmopower/trained_models/CP_paper/PK/README.md:62-65).  Here one launch evaluates both networks per row tile; the
  second emulator's output epilogue adds the first one's log-spectrum - written microseconds earlier by the same lane,
  still in L2 - and applies 10** in the same pass: no intermediate array in HBM, one device->host copy.
* emulators that share their parameter tensor: the Planck-lite likelihood calls TT, TE and EE on the same [N, 6] tensor
  (cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py:276-278); `predictions_multi_tf` runs them as one
  launch and returns the three spectra.
"""
import numpy as np

from . import _engine


def _multi_engine(models, device):
    """The cached multi-emulator engine of this tuple of models on `device` (kept on the first model)."""
    cache = models[0].__dict__.setdefault("_multi_engines", {})
    key = (tuple(id(m) for m in models), int(device))
    if key not in cache:
        cache[key] = (_engine.MultiEngine([m._attrs() for m in models], device=device), list(models))   # the models stay referenced
    return cache[key][0]


def _can_share_a_launch(models):
    return (1 < len(models) <= 3 and all(getattr(m, "_precision", "f16x3") == "f16x3" for m in models)
            and all(len(m.W_) >= 2 for m in models)
            and sum(len(m.W_) + (1 if hasattr(m, "pca_transform_matrix_") else 0) for m in models) <= 16)


def predictions_multi_tf(models, parameter_tensors, ten_to=False):
    """models: up to three emulators; parameter_tensors: one ordered CUDA float32 tensor per model with the same number of
    rows (pass the same tensor several times when the emulators share their parameters); ten_to: one bool per model or one
    for all.  Returns the list of CUDA output tensors - bit-identical to models[i].predictions_tf / ten_to_predictions_tf,
    from a single launch."""
    if len(models) != len(parameter_tensors):
        raise ValueError("need one parameter tensor per model")
    flags = [bool(ten_to)] * len(models) if isinstance(ten_to, (bool, int)) else [bool(v) for v in ten_to]
    dev = parameter_tensors[0].device
    if not _can_share_a_launch(models):
        return [m.engine(dev.index).forward_torch(x, ten_to=f) for m, x, f in zip(models, parameter_tensors, flags)]
    return _multi_engine(models, dev.index).forward_torch(list(parameter_tensors), ten_to=flags)


def ten_to_sum_predictions_tf(models, parameter_tensors):
    """models: emulators with identical `n_modes`; parameter_tensors: one ordered CUDA float32 tensor per model
    (same number of rows).  Returns 10**(sum of the models' predictions) as a CUDA tensor."""
    import torch
    if len(models) < 1 or len(models) != len(parameter_tensors):
        raise ValueError("need one parameter tensor per model")
    n_modes = int(models[0].n_modes)
    n = parameter_tensors[0].shape[0]
    for m, x in zip(models, parameter_tensors):
        if int(m.n_modes) != n_modes or x.shape[0] != n:
            raise ValueError("all emulators must share n_modes and the batch size")
    dev = parameter_tensors[0].device
    last = len(models) - 1
    if _can_share_a_launch(models):
        outs = _multi_engine(models, dev.index).forward_torch(
            list(parameter_tensors), ten_to=[i == last for i in range(len(models))], accumulate=[i > 0 for i in range(len(models))])
        return outs[-1]
    pitch = (n_modes + 7) // 8 * 8
    out = torch.empty((n, pitch), dtype=torch.float32, device=dev)[:, :n_modes]
    for i, (m, x) in enumerate(zip(models, parameter_tensors)):
        m.engine(dev.index).forward_torch(x, out=out, ten_to=(i == last), accumulate=(i > 0))
    return out


def ten_to_sum_predictions_np(models, parameter_dicts):
    """NumPy-in / NumPy-out twin: `10.**(sum_i models[i].predictions_np(parameter_dicts[i]))`."""
    import torch
    arrs = [m.dict_to_ordered_arr_np(d) for m, d in zip(models, parameter_dicts)]
    out_dtype = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
    dev = models[0].engine().device
    xs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda:%d" % dev) for a in arrs]
    y = ten_to_sum_predictions_tf(models, xs).cpu().numpy()
    return y if out_dtype == np.float32 else y.astype(np.float64)
