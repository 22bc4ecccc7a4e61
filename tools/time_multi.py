"""One launch for several emulators against one launch per emulator (B200 only), 1 048 576 rows by default.
Usage: python tools/time_multi.py [rows]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import fused, priors, standins  # noqa: E402


def load(name):
    kind, path = standins.model_path(name)
    return getattr(cp, kind)(restore=True, restore_filename=path, device=0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    lin, nl = load("PKLIN_NN"), load("PKNLBOOST_NN")
    x_nl = priors.sample_prior(nl.parameters, rows, seed=1, dtype=np.float32)
    cols = [nl.parameters.index(k) for k in lin.parameters]
    xb = torch.from_numpy(x_nl).cuda()
    xl = torch.from_numpy(np.ascontiguousarray(x_nl[:, cols])).cuda()
    out = torch.empty((rows, 424), dtype=torch.float32, device="cuda")[:, :420]

    def two_launches():
        lin.engine().forward_torch(xl, out=out)
        nl.engine().forward_torch(xb, out=out, ten_to=True, accumulate=True)
    me = fused._multi_engine([lin, nl], 0)
    me.set_timing(True)

    def one_launch():
        me.forward_torch([xl, xb], outs=[out, out], ten_to=[False, True], accumulate=[False, True])
    t2, t1 = timed(two_launches), timed(one_launch)
    print("P(k) = 10**(log lin + log boost), rows=%d: two launches %.3f ms, one launch %.3f ms (x%.3f), %.3f M SM cycles"
          % (rows, t2, t1, t2 / t1, me.last_kernel_cycles() / 1e6), flush=True)
    del out
    tt, te, ee = load("cmb_TT_NN"), load("cmb_TE_PCAplusNN"), load("cmb_EE_NN")
    n3 = min(rows, 1 << 19)
    x = torch.from_numpy(priors.sample_prior(tt.parameters, n3, seed=2, dtype=np.float32)).cuda()
    outs = [torch.empty((n3, 2512), dtype=torch.float32, device="cuda")[:, :2507] for _ in range(3)]

    def three_launches():
        tt.engine().forward_torch(x, out=outs[0], ten_to=True)
        te.engine().forward_torch(x, out=outs[1], ten_to=False)
        ee.engine().forward_torch(x, out=outs[2], ten_to=True)
    me3 = fused._multi_engine([tt, te, ee], 0)
    me3.set_timing(True)

    def one_launch3():
        me3.forward_torch([x, x, x], outs=outs, ten_to=[True, False, True])
    t3, t1 = timed(three_launches), timed(one_launch3)
    print("TT + TE + EE on one parameter tensor, rows=%d: three launches %.3f ms, one launch %.3f ms (x%.3f), %.3f M SM cycles"
          % (n3, t3, t1, t3 / t1, me3.last_kernel_cycles() / 1e6), flush=True)


if __name__ == "__main__":
    main()
