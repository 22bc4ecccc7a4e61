"""Wall time of one get_loglkl call (three emulators + Planck-lite tail, device tensors in and out, synchronised) against
the batch size.  (B200 only)   Usage: python tools/latency_likelihood.py"""
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


def main():
    models = {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    params = list(models["tt"].parameters)
    like = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"])
    graphed = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"], cuda_graphs=True)
    for n in (1, 64, 1024, 8192, 65536):
        x = priors.sample_prior(params, n, seed=5, dtype=np.float32)
        full = torch.from_numpy(np.concatenate([x, np.ones((n, 1), np.float32)], axis=1)).cuda()
        res = {}
        for mode, rows in (("side by side", 8192), ("one after the other", 0)):
            like.concurrent_rows = rows
            for _ in range(5):
                ll = like.get_loglkl(full)
            torch.cuda.synchronize()
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                ll = like.get_loglkl(full)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            res[mode] = (1e6 * float(np.median(ts)), ll)
        for _ in range(5):
            ll = graphed.get_loglkl(full)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            ll = graphed.get_loglkl(full)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res["graph"] = (1e6 * float(np.median(ts)), ll)
        same = torch.equal(res["side by side"][1], res["one after the other"][1]) and torch.equal(res["graph"][1], res["side by side"][1])
        print("rows=%6d: get_loglkl %8.1f us as a CUDA-graph replay, %8.1f us with the three emulators side by side, %8.1f us one after the other (identical results: %s)"
              % (n, res["graph"][0], res["side by side"][0], res["one after the other"][0], same), flush=True)


if __name__ == "__main__":
    main()
