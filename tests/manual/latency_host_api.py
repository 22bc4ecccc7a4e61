"""Latency of the reference-facing NumPy API (dict in, ndarray out: ten_to_predictions_np) for small batches, next to the
NumPy/OpenBLAS oracle port on the host cores.  (B200 only)   Usage: python tests/manual/latency_host_api.py [model]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402
from oracle import oracle_np  # noqa: E402  (tools/ may time the oracle: it is the CPU baseline here, not the product)


def med(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    attrs = standins.load_attrs(kind, path)
    for n in (1, 16, 128, 1024, 8192):
        x = priors.sample_prior(m.parameters, n, seed=1)
        d = {k: x[:, i] for i, k in enumerate(m.parameters)}
        us_gpu = med(lambda: m.ten_to_predictions_np(d), 50)
        us_cpu = med(lambda: 10.0 ** oracle_np.forward_pass(attrs, np.stack([d[k] for k in m.parameters], axis=1)), 20 if n <= 1024 else 5)
        print("%s rows=%5d: B200 ten_to_predictions_np %8.1f us per call, NumPy port on %d host threads %9.1f us (x%.1f)"
              % (name, n, us_gpu, os.cpu_count(), us_cpu, us_cpu / us_gpu), flush=True)


if __name__ == "__main__":
    main()
