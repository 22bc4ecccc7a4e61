"""N>1 host logic on CPU: world_size-2 gloo run of the row sharding + optional gather
(the GPU engine is replaced by the oracle as the per-rank forward; rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosmopower_b200 import priors, sharding, standins
from oracle import oracle_np


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 128, 1000003):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    attrs = standins.load_attrs(*standins.model_path("PKLIN_NN"))
    x = priors.sample_prior(attrs["parameters"], n, seed=99)
    fwd = lambda rows: oracle_np.forward_pass(attrs, rows)
    lo, hi, local = sharding.predict_sharded(fwd, x, attrs["n_modes"])
    assert (lo, hi) == sharding.shard_bounds(n, world, rank) and local.shape == (hi - lo, attrs["n_modes"])
    full = sharding.predict_sharded(fwd, x, attrs["n_modes"], gather=True)
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full)
    # tensor path with the shard written straight into its slot of the gathered buffer (forward_into), even and ragged
    import torch
    for m in (n - 1, n):
        xt = torch.from_numpy(x[:m].astype(np.float32))
        into = lambda rows, out: out.copy_(torch.from_numpy(oracle_np.forward_pass(attrs, rows.numpy().astype(np.float64)).astype(np.float32)))
        ft = sharding.predict_sharded(None, xt, attrs["n_modes"], gather=True, forward_into=into)
        ref = oracle_np.forward_pass(attrs, x[:m].astype(np.float32).astype(np.float64)).astype(np.float32)
        assert tuple(ft.shape) == ref.shape and np.array_equal(ft.numpy(), ref)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n = 301                                    # odd: ragged shards
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    attrs = standins.load_attrs(*standins.model_path("PKLIN_NN"))
    x = priors.sample_prior(attrs["parameters"], n, seed=99)
    ref = oracle_np.forward_pass(attrs, x)
    for r in range(2):
        full = np.load(tmp_path / ("full_%d.npy" % r))
        assert full.shape == ref.shape
        np.testing.assert_allclose(full, ref, rtol=1e-11, atol=1e-13)
