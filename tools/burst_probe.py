"""Is the output phase limited by every cluster writing at the same time?  Runs the same number of
iterations per cluster with fewer clusters (CPB200_DEBUG_MAX_CLUSTERS) and prints CTA 0's mean
epilogue/MMA cycles per output job.  (B200 only)   Usage: python tools/burst_probe.py [model]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_PP_PCAplusNN"
    iters = 16
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    eng = m.engine()
    for cap in (74, 37, 18, 4):
        os.environ["CPB200_DEBUG_MAX_CLUSTERS"] = str(cap)
        rows = cap * 512 * iters
        x = torch.from_numpy(priors.sample_prior(m.parameters, rows, seed=1, dtype=np.float32)).cuda()
        pitch = (m.n_modes + 7) // 8 * 8
        out = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")[:, :m.n_modes]
        for _ in range(2):
            eng.forward_torch(x, out=out, ten_to=True)
        torch.cuda.synchronize()
        eng.debug_counters(enable=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward_torch(x, out=out, ten_to=True)
        e1.record()
        torch.cuda.synchronize()
        c = eng.debug_counters(enable=False, read=True).astype(np.float64)
        ms = e0.elapsed_time(e1)
        jobs = c.reshape(-1)[: c.size - 96 * 32].reshape(-1, 16)[2 * cap: 2 * cap + 96]
        tot = c[:2 * cap, 5].max()
        kinds = {}
        for r in jobs:
            if r[2] == 0:
                continue
            g = int(r[15]) & 255
            kinds.setdefault(g, []).append(r)
        print("%s clusters=%d rows=%d: %.3f ms, %.3g spectra/s per cluster-iteration %.0f cycles" % (
            name, cap, rows, ms, rows / ms * 1e3, tot / iters))
        for g in sorted(kinds):
            a = np.array(kinds[g]) / iters
            print("   contraction %d: %2d jobs  mma busy %7.0f (wait_full %6.0f wait_accempty %6.0f)  epi busy %7.0f wait %6.0f" % (
                g, len(a), a[:, 2].mean(), a[:, 0].mean(), a[:, 1].mean(), a[:, 4].mean(), a[:, 3].mean()))
        del x, out


if __name__ == "__main__":
    main()
