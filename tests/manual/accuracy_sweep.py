"""Max parity error of the tensor-core path against the float64 oracle over a large in-prior batch
(B200 + host CPU).  Usage: python tests/manual/accuracy_sweep.py [rows]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402
from oracle import oracle_np  # noqa: E402
from tests import parity  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    names = sys.argv[2].split(',') if len(sys.argv) > 2 else parity.MODELS
    for name in names:
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        attrs = standins.load_attrs(kind, path)
        n = rows if m.n_modes < 1000 else rows // 4
        x = priors.sample_prior(m.parameters, n, seed=4242)
        t0 = time.time()
        y = m.forward_pass_np(x)
        worst, p99 = 0.0, 0.0
        for s in range(0, n, 20000):
            ref = oracle_np.forward_pass(attrs, x[s:s + 20000])
            if name in parity.LINEAR_MODELS:
                e = np.abs(y[s:s + 20000] - ref) / np.max(np.abs(ref), axis=1, keepdims=True)
            else:
                e = np.abs(y[s:s + 20000] - ref) / np.maximum(np.abs(ref), 1.0)
            worst = max(worst, float(e.max()))
            p99 = max(p99, float(np.quantile(e.max(axis=1), 0.99)))
        print("%-18s rows=%d max err %.3g (tol %.0e)  99%%-ile of per-row max %.3g  [%.0fs]"
              % (name, n, worst, parity.model_tol(name), p99, time.time() - t0), flush=True)
        m.close()


if __name__ == "__main__":
    main()
