"""Streaming sweep over very large parameter sets (BASELINE.json configs[4]: all emulators, 100 M
parameter sets, sharded over the GPUs of a node).

100 M cmb_TT spectra are 1 TB of output: they are never materialised.  Each rank walks its shard in
chunks; the parameters of a chunk are drawn on the device (uniform over the model's training prior,
counter-based generator seeded per chunk, so any chunk can be regenerated independently), the fused
kernel writes the chunk's spectra into ONE reused device buffer, and only a float64 checksum of the
log-spectra leaves the GPU.  No collective is involved: ranks own disjoint chunk ranges.
"""
import numpy as np

from . import priors
from .sharding import shard_bounds


def chunk_parameters(parameters, chunk_index, rows, seed, device):
    """(rows, P) float32 CUDA tensor of in-prior parameters for one chunk (deterministic in (seed, chunk_index))."""
    import torch
    lo, hi = priors.prior_bounds(parameters)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1000003 + int(chunk_index))
    u = torch.rand((rows, len(parameters)), generator=g, device=device, dtype=torch.float32)
    lo_t = torch.tensor(lo, dtype=torch.float32, device=device)
    span = torch.tensor(hi - lo, dtype=torch.float32, device=device)
    return lo_t + u * span


def sweep(models, n_total, chunk_rows=1 << 20, seed=0, device=0, ten_to=False, checksum=True,
          world_size=1, rank=0):
    """Evaluate every emulator in `models` (dict name -> restored model) on `n_total` parameter sets.

    Returns {name: {"rows": rows done by this rank, "checksum": float64 sum of all outputs (or None),
                    "seconds": device time}}.  Chunks are dealt to ranks in contiguous ranges."""
    import torch
    dev = torch.device("cuda", device)
    n_chunks = (n_total + chunk_rows - 1) // chunk_rows
    c_lo, c_hi = shard_bounds(n_chunks, world_size, rank)
    out = {}
    for name, m in models.items():
        eng = m.engine(device)
        pitch = (int(m.n_modes) + 7) // 8 * 8
        buf = torch.empty((chunk_rows, pitch), dtype=torch.float32, device=dev)
        total = torch.zeros((), dtype=torch.float64, device=dev)
        rows_done = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for ci in range(c_lo, c_hi):
            rows = min(chunk_rows, n_total - ci * chunk_rows)
            x = chunk_parameters(m.parameters, ci, rows, seed, dev)
            view = buf[:rows, :int(m.n_modes)]
            eng.forward_torch(x, out=view, ten_to=ten_to)
            if checksum:
                total += view.sum(dtype=torch.float64)
            rows_done += rows
        e1.record()
        torch.cuda.synchronize(dev)
        out[name] = {"rows": rows_done, "checksum": float(total.item()) if checksum else None,
                     "seconds": e0.elapsed_time(e1) * 1e-3}
        del buf
    return out
