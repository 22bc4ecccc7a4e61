"""How does the output phase depend on ten_to and on the row pitch?  (B200 only)
Usage: python tools/out_phase_probe.py [model] [rows]"""
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
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    eng = m.engine()
    x = torch.from_numpy(priors.sample_prior(m.parameters, rows, seed=1, dtype=np.float32)).cuda()
    for pitch in (2512, 2560, 2528, 2507):
        buf = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")
        out = buf[:, :m.n_modes]
        for ten_to in (True, False):
            for _ in range(3):
                eng.forward_torch(x, out=out, ten_to=ten_to)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.forward_torch(x, out=out, ten_to=ten_to)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("%s pitch=%d ten_to=%d: median %.3f ms min %.3f" % (name, pitch, ten_to, float(np.median(ts)), min(ts)), flush=True)
        del buf, out


if __name__ == "__main__":
    main()
