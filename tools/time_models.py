"""Single-launch device time of every CP_paper emulator at a given batch (production kernel, no counters)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(standins.ALL_MODELS)
    for name in names:
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        eng = m.engine()
        eng.set_timing(True)                 # the cycle figure is collected in timing mode only
        x = torch.from_numpy(priors.sample_prior(m.parameters, rows, seed=1, dtype=np.float32)).cuda()
        pitch = (m.n_modes + 7) // 8 * 8
        out = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")[:, :m.n_modes]
        for _ in range(3):
            eng.forward_torch(x, out=out, ten_to=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.forward_torch(x, out=out, ten_to=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print("%-18s rows=%d: %.3f ms (median of 5, min %.3f), %.3g spectra/s, %.3f M SM cycles (slowest CTA)" % (
            name, rows, ms, min(ts), rows / ms * 1e3, eng.last_kernel_cycles() / 1e6), flush=True)
        m.close()


if __name__ == "__main__":
    main()
