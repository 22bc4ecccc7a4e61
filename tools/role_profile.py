"""Where does each warp role of the fused kernel spend its cycles?  (B200 only)
Usage: python tools/role_profile.py [model] [rows]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402

NAMES = ["load_wait_empty", "load_wait_aready", "load_total", "mma_wait_full", "mma_wait_accempty", "mma_total",
         "epi_wait_accfull", "epi_act", "epi_out", "epi_total", "in_wait", "in_total"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    eng = m.engine()
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
    print("%s rows=%d: %.3f ms, %.3g spectra/s" % (name, rows, ms, rows / ms * 1e3))
    tot = c[:, 5].max()
    print("max over CTAs of the MMA role total = %.3g cycles -> %.0f MHz effective (leader CTAs only record MMA counters)" % (tot, tot / ms / 1e3))
    for i, n in enumerate(NAMES):
        print("  %-18s mean %12.0f  (%5.1f%% of mma_total)  min %12.0f max %12.0f" % (n, c[:, i].mean(), 100 * c[:, i].mean() / tot, c[:, i].min(), c[:, i].max()))


if __name__ == "__main__":
    main()
