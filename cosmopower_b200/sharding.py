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


def predict_sharded(forward, parameters_arr, n_modes, group=None, gather=False):
    """Run `forward` (rows -> spectra; e.g. model.forward_pass_np or a device-tensor callable) on this
    rank's slice.  Returns (lo, hi, local_result) or, with gather=True, the full (N, n_modes) array on
    every rank (one all_gather of the padded shards)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = parameters_arr.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = forward(parameters_arr[lo:hi])
    if not gather or world == 1:
        return (lo, hi, local) if not gather else local
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    rows_max = -(-n // world)
    pad = torch.zeros((rows_max, n_modes), dtype=t.dtype, device=t.device)
    pad[:hi - lo] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat([parts[r][:shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0]] for r in range(world)], dim=0)
    return out.numpy() if is_np else out
