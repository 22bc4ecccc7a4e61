"""SURVEY 8e's optional epilogue collective, timed: every rank evaluates its contiguous row shard of one batch on its
GPU and the shards are all-gathered over NVLink (NCCL), so that each rank ends up with the whole (N, n_modes) result.
Not part of the headline metric.   torchrun --nproc-per-node G tools/gather_bench.py [model] [rows]"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, sharding, standins  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=local)
    x = torch.from_numpy(priors.sample_prior(m.parameters, rows, seed=3, dtype=np.float32)).cuda()

    def timed(gather):
        for _ in range(2):
            sharding.predict_sharded(m.ten_to_predictions_tf, x, m.n_modes, gather=gather,
                                     forward_into=lambda rows, out: m.engine().forward_torch(rows, out=out, ten_to=True))
        dist.barrier(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = sharding.predict_sharded(m.ten_to_predictions_tf, x, m.n_modes, gather=gather,
                                     forward_into=lambda rows, out: m.engine().forward_torch(rows, out=out, ten_to=True))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    ms_local, _ = timed(False)
    ms_gather, full = timed(True)
    ok = tuple(full.shape) == (rows, m.n_modes)
    if rank == 0:
        shard_bytes = (rows // world) * m.n_modes * 4
        print(json.dumps({"model": name, "rows": rows, "world": world, "ms_shards_only": ms_local, "ms_with_all_gather": ms_gather,
                          "gathered_bytes_per_rank": shard_bytes * (world - 1), "gather_gbs_per_rank": shard_bytes * (world - 1) / max(ms_gather - ms_local, 1e-6) / 1e6,
                          "full_result_on_every_rank": ok}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
