"""Device time of the Planck-lite likelihood tail and of the full emulators + likelihood call.  (B200 only)
Usage: python tools/time_likelihood.py [rows]"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402
from cosmopower_b200.likelihoods import tf_planck2018_lite  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    models = {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    params = list(models["tt"].parameters)
    fold = os.environ.get("CPB200_FOLD_TE", "1") != "0"
    fuse = os.environ.get("CPB200_FUSE_BINNING", "1") != "0"
    like = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"], fold_te=fold,
                              fuse_binning=fuse)
    x = priors.sample_prior(params, rows, seed=5, dtype=np.float32)
    full = torch.from_numpy(np.concatenate([x, np.ones((rows, 1), np.float32)], axis=1)).cuda()
    cosmo = full[:, :len(params)].contiguous()
    if like._binned:      # TT / EE leave their emulators as per-4-multipole partial band powers
        tt = lambda: models["tt"].engine(0).forward_torch(cosmo, ten_to=True, binned=True)
        ee = lambda: models["ee"].engine(0).forward_torch(cosmo, ten_to=True, binned=True)
    else:
        tt = lambda: models["tt"].ten_to_predictions_tf(cosmo)
        ee = lambda: models["ee"].ten_to_predictions_tf(cosmo)
    cl_tt = tt()
    te = (lambda: like._te_binned.forward_torch(cosmo, ten_to=False)) if fold else (lambda: models["te"].predictions_tf(cosmo))
    cl_te = te()
    cl_ee = ee()
    cal = full[:, -1].contiguous()
    ms_tail = timed(lambda: like._engine.loglike_torch(cl_tt, cl_te, cl_ee, cal))
    ms_emu = timed(lambda: (tt(), te(), ee()))
    ms_all = timed(lambda: like.get_loglkl(full))
    n_ell, nb = like._engine.tt_ee_width, 613
    te_len = cl_te.shape[1]
    out = {"rows": rows, "fold_te": fold, "fuse_binning": bool(like._binned), "likelihood_tail_ms": ms_tail, "three_emulators_ms": ms_emu, "get_loglkl_ms": ms_all,
           "tail_hbm_gbs": rows * (2 * n_ell + te_len + 2 * 640) * 4 / ms_tail / 1e6,
           "tail_quadform_tflops_dense_equiv": rows * 2.0 * nb * nb / ms_tail / 1e9,
           "loglike_per_s": rows / ms_all * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
