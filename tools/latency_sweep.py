"""Call latency of the device-tensor API against the batch size (B200 only): wall time per call of
ten_to_predictions_tf on a resident tensor, synchronised, median of 50.   Usage: python tools/latency_sweep.py [model]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    for n in (1, 128, 512, 2048, 8192, 37888, 65536, 262144, 1048576):
        x = torch.from_numpy(priors.sample_prior(m.parameters, n, seed=1, dtype=np.float32)).cuda()
        pitch = (m.n_modes + 7) // 8 * 8
        out = torch.empty((n, pitch), dtype=torch.float32, device="cuda")[:, :m.n_modes]
        eng = m.engine()
        eng.set_timing(True)
        for _ in range(5):
            eng.forward_torch(x, out=out, ten_to=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(50 if n <= 65536 else 10):
            t0 = time.perf_counter()
            eng.forward_torch(x, out=out, ten_to=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        us = 1e6 * float(np.median(ts))
        print("%s rows=%8d: %9.1f us per call, %10.3g spectra/s, %.3f M SM cycles" % (name, n, us, n / us * 1e6, eng.last_kernel_cycles() / 1e6), flush=True)


if __name__ == "__main__":
    main()
