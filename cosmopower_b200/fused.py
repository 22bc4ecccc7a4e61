"""Fused multi-emulator evaluation: emulators whose log-spectra are summed before 10**.

The reference combines the linear matter power spectrum and the nonlinear boost on the host,
`10.**(lin.predictions_np(p_lin) + boost.predictions_np(p_boost))`
(cosmopower/trained_models/CP_paper/PK/README.md:62-65).  Here the second emulator's output epilogue adds
the first emulator's log-spectrum in place and applies 10** in the same pass: no intermediate array, one
device->host copy.
"""
import numpy as np


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
    pitch = (n_modes + 7) // 8 * 8
    out = torch.empty((n, pitch), dtype=torch.float32, device=dev)[:, :n_modes]
    last = len(models) - 1
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
