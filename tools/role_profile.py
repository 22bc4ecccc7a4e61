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

NJ = 320          # UM_MAX_JOBS

NAMES = ["load_wait_empty", "load_wait_aready", "load_total", "mma_wait_full", "mma_wait_accempty", "mma_total",
         "epi_wait_accfull", "epi_act", "epi_out", "epi_total", "in_wait", "in_total"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    if name == "PK_PAIR":            # PKLIN + PKNLBOOST in one launch, log-sum fused
        from cosmopower_b200 import fused
        ms_ = [getattr(cp, standins.model_path(n)[0])(restore=True, restore_filename=standins.model_path(n)[1], device=0) for n in ("PKLIN_NN", "PKNLBOOST_NN")]
        eng = fused._multi_engine(ms_, 0)
        xb = priors.sample_prior(ms_[1].parameters, rows, seed=1, dtype=np.float32)
        cols = [ms_[1].parameters.index(k) for k in ms_[0].parameters]
        xs = [torch.from_numpy(np.ascontiguousarray(xb[:, cols])).cuda(), torch.from_numpy(xb).cuda()]
        out = torch.empty((rows, 424), dtype=torch.float32, device="cuda")[:, :420]
        run = lambda: eng.forward_torch(xs, outs=[out, out], ten_to=[False, True], accumulate=[False, True])
    else:
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        eng = m.engine()
        x = torch.from_numpy(priors.sample_prior(m.parameters, rows, seed=1, dtype=np.float32)).cuda()
        pitch = (m.n_modes + 7) // 8 * 8
        out = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")[:, :m.n_modes]
        run = lambda: eng.forward_torch(x, out=out, ten_to=True)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    eng.debug_counters(enable=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    c = eng.debug_counters(enable=False, read=True).astype(np.float64)
    ms = e0.elapsed_time(e1)
    print("%s rows=%d: %.3f ms, %.3g spectra/s" % (name, rows, ms, rows / ms * 1e3))
    flat = c.reshape(-1)
    ntrace = NJ * 16 * 2
    trace = flat[flat.size - ntrace:].reshape(NJ, 16, 2)     # [job][stage][loader issue, MMA saw it landed]
    c = flat[:flat.size - ntrace].reshape(-1, 16)
    ncta = c.shape[0] - NJ                     # per-CTA rows, then NJ per-entry rows written by CTA 0
    jobs = c[ncta:]
    c = c[:ncta]
    tot = c[:, 5].max()
    lead = c[:, 5]
    lead = lead[lead > 0]
    print("median over leader CTAs of the MMA role total = %.0f cycles (clock-independent figure of merit)" % np.median(lead))
    print("max over CTAs of the MMA role total = %.3g cycles -> %.0f MHz effective (leader CTAs only record MMA counters)" % (tot, tot / ms / 1e3))
    for i, n in enumerate(NAMES):
        print("  %-18s mean %12.0f  (%5.1f%% of mma_total)  min %12.0f max %12.0f" % (n, c[:, i].mean(), 100 * c[:, i].mean() / tot, c[:, i].min(), c[:, i].max()))
    print("per job-table entry of CTA 0 (cycles summed over its iterations; time stamps of iteration 3 relative to the first):")
    print("  ji  g t  c nxt s0 ns FL slot | mma: wait_full wait_accempty     busy | epi:     wait     busy | load: wait_aready wait_empty | t_mma0  t_mma1  t_epi0  t_epi1 t_load0 t_load1")
    t0 = min(x for x in jobs[:, 8] if x > 0) if (jobs[:, 8] > 0).any() else 0
    for ji in range(NJ):
        r = jobs[ji]
        if r[2] == 0:
            continue
        jw = int(r[15])
        ts = ["%7d" % (x - t0) if x > 0 else "      -" for x in r[8:14]]
        print("  %3d %2d %d %2d  %d  %2d %2d %s%s %d%s | %14.0f %13.0f %8.0f | %13.0f %8.0f | %17.0f %10.0f | %s" % (
            ji, jw & 15, (jw >> 4) & 1, (jw >> 5) & 127, (jw >> 12) & 1, (jw >> 13) & 63, (jw >> 19) & 63,
            "F" if (jw >> 25) & 1 else "-", "L" if (jw >> 26) & 1 else "-", (jw >> 27) & 3, "+" if (jw >> 29) & 1 else " ",
            r[0], r[1], r[2], r[3], r[4], r[5], r[6], " ".join(ts)))
    print_trace(trace, jobs, t0)




def print_trace(trace, jobs, t0):
    print("stage trace of the traced iteration (cycles relative to its first MMA job): per job, loader issue times then MMA landed times")
    for ji in range(NJ):
        if jobs[ji][2] == 0 or trace[ji, 0, 1] == 0:
            continue
        ld = " ".join("%6d" % (x - t0) if x > 0 else "     -" for x in trace[ji, :, 0])
        mm = " ".join("%6d" % (x - t0) if x > 0 else "     -" for x in trace[ji, :, 1])
        print("  job %2d load %s" % (ji, ld))
        print("         mma  %s" % mm)


if __name__ == "__main__":
    main()
