"""Row sharding of a parameter batch over the GPUs of one node.

The emulator forward pass is independent per row and the weights are a few MB, so each rank holds
a full engine and a contiguous slice of rows; there is NO collective on the hot path.  Only when
the caller wants the whole result on every rank is one all_gather issued (NCCL on GPUs, gloo in the
CPU tests)."""
import numpy as np


def shard_bounds(n_rows, world_size, rank):
    """Contiguous range [lo, hi) of rows owned by `rank`; sizes differ by at most one row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world size %r" % (rank, world_size))
    base, extra = divmod(int(n_rows), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def predict_sharded(forward, parameters_arr, n_modes, group=None, gather=False, forward_into=None):
    """Run `forward` (rows -> spectra; e.g. model.forward_pass_np or a device-tensor callable) on this
    rank's slice.  Returns (lo, hi, local_result) or, with gather=True, the full (N, n_modes) array on
    every rank: one all_gather, in place - every rank's shard sits at its final position of one buffer, and
    `forward_into(rows, out)` (e.g. lambda x, out: engine.forward_torch(x, out=out)), when given, lets the
    kernel write it there directly instead of through a copy."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = parameters_arr.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    if not gather or world == 1:
        local = forward(parameters_arr[lo:hi])
        return (lo, hi, local) if not gather else local
    rows_max = -(-n // world)
    is_np = isinstance(parameters_arr, np.ndarray)
    if is_np or forward_into is None:
        local = forward(parameters_arr[lo:hi])
        t = torch.from_numpy(np.ascontiguousarray(local)) if isinstance(local, np.ndarray) else local
        if not t.is_cuda and dist.get_backend(group) == "nccl":
            # NCCL moves device memory only: a host shard travels through this rank's GPU
            t = t.to(torch.device("cuda", torch.cuda.current_device()))
        full = torch.empty((world * rows_max, n_modes), dtype=t.dtype, device=t.device)
        mine = full[rank * rows_max:(rank + 1) * rows_max]
        mine[:hi - lo] = t
    else:
        # rows padded to a multiple of 8 floats, as the engine's own buffers: the kernel keeps its 32-byte stores, and
        # the padding (5 floats in 2512 for the CMB spectra) simply travels along; the caller gets the (N, n_modes) view
        pitch = (n_modes + 7) // 8 * 8
        full = torch.empty((world * rows_max, pitch), dtype=torch.float32, device=parameters_arr.device)
        mine = full[rank * rows_max:(rank + 1) * rows_max]
        forward_into(parameters_arr[lo:hi], mine[:hi - lo, :n_modes])
    if hi - lo < rows_max:
        mine[hi - lo:] = 0                       # the ragged tail of a short shard travels as zeros
    dist.all_gather_into_tensor(full, mine, group=group)
    if n == world * rows_max:
        out = full[:, :n_modes]                  # equal shards: already in place
    else:
        out = torch.cat([full[r * rows_max:r * rows_max + shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0], :n_modes]
                         for r in range(world)], dim=0)
    return out.cpu().numpy() if is_np else out
