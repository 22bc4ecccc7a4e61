"""Randomised soak of the device paths (B200 only): random batch sizes through every emulator - tensor path twice (run-to-run
identical) and against the CUDA-core fp32 path -, the multi-emulator launch against the single launches, and the likelihood
with fused binning against the unfused pipeline.  A stuck pipeline traps through the kernel's watchdog rather than hanging.
Usage: python tools/fuzz_sizes.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402
from cosmopower_b200.likelihoods import tf_planck2018_lite  # noqa: E402

NAMES = ["cmb_TT_NN", "cmb_EE_NN", "cmb_TE_PCAplusNN", "cmb_PP_PCAplusNN", "PKLIN_NN", "PKNLBOOST_NN"]


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    tens, simt = {}, {}
    for nm in NAMES:
        kind, path = standins.model_path(nm)
        tens[nm] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        simt[nm] = getattr(cp, kind)(restore=True, restore_filename=path, device=0, precision="fp32")
    like = {f: tf_planck2018_lite(parameters=None, tt_emu_model=tens["cmb_TT_NN"], te_emu_model=tens["cmb_TE_PCAplusNN"],
                                  ee_emu_model=tens["cmb_EE_NN"], fuse_binning=f) for f in (True, False)}
    t_end = time.time() + seconds
    it, worst = 0, {nm: 0.0 for nm in NAMES}
    worst_ll = 0.0
    while time.time() < t_end:
        it += 1
        n = int(np.exp(rng.uniform(0, np.log(70000.0))))
        nm = NAMES[rng.integers(len(NAMES))]
        m = tens[nm]
        x = torch.from_numpy(priors.sample_prior(m.parameters, n, seed=int(rng.integers(1 << 30)), dtype=np.float32)).cuda()
        y1 = m.predictions_tf(x).clone()
        y2 = m.predictions_tf(x)
        assert torch.equal(y1, y2), ("not reproducible", nm, n)
        yr = simt[nm].predictions_tf(x)
        if nm == "cmb_TE_PCAplusNN":
            err = float(((y1 - yr).abs() / yr.abs().amax(dim=1, keepdim=True)).max())
            tol = 1e-4
        else:
            err = float(((y1 - yr).abs() / yr.abs().clamp_min(1.0)).max())
            tol = 2e-5
        assert err <= tol and bool(torch.isfinite(y1).all()), (nm, n, err)
        worst[nm] = max(worst[nm], err)
        if it % 4 == 0:        # the three CMB emulators in one launch
            ms = [tens[k] for k in ("cmb_TT_NN", "cmb_TE_PCAplusNN", "cmb_EE_NN")]
            xs = torch.from_numpy(priors.sample_prior(ms[0].parameters, n, seed=it, dtype=np.float32)).cuda()
            outs = cp.predictions_multi_tf(ms, [xs, xs, xs], ten_to=[True, False, True])
            want = [ms[0].ten_to_predictions_tf(xs), ms[1].predictions_tf(xs), ms[2].ten_to_predictions_tf(xs)]
            for o, w in zip(outs, want):
                assert torch.equal(o, w), ("multi-emulator launch differs", n)
        if it % 5 == 0:        # likelihood: fused binning against the unfused pipeline
            nl = min(n, 20000)
            xs = priors.sample_prior(tens["cmb_TT_NN"].parameters, nl, seed=it + 7, dtype=np.float32)
            full = torch.from_numpy(np.concatenate([xs, np.ones((nl, 1), np.float32)], axis=1)).cuda()
            a, b = like[True].get_loglkl(full), like[False].get_loglkl(full)
            rel = float(((a - b).abs() / b.abs()).max())
            assert rel <= 1e-3 and bool(torch.isfinite(a).all()), ("likelihood", nl, rel)
            assert torch.equal(a, like[True].get_loglkl(full)), ("likelihood not reproducible", nl)
            worst_ll = max(worst_ll, rel)
    print("fuzz: %d iterations in %.0f s, worst tensor-vs-fp32 error per model: %s; likelihood fused-vs-unfused %.2e"
          % (it, seconds, ", ".join("%s %.2e" % kv for kv in worst.items()), worst_ll))


if __name__ == "__main__":
    main()
