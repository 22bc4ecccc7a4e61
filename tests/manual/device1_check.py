"""Sanity check on a second GPU while device 0 stays the current device (needs 2 visible B200s):
an emulator and the likelihood tail created with device=1 must match the oracle."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import _engine, priors, standins  # noqa: E402
from cosmopower_b200.likelihoods import planck2018_lite as pl  # noqa: E402
from oracle import oracle_np, planck_lite_np  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2
    torch.cuda.set_device(0)
    kind, path = standins.model_path("PKLIN_NN")
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=1)
    x = priors.sample_prior(m.parameters, 5000, seed=3)
    y = m.predictions_np({k: x[:, i] for i, k in enumerate(m.parameters)})
    ref = oracle_np.forward_pass(standins.load_attrs(kind, path), x)
    err = float(np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1.0)))
    assert err <= 1e-5, err
    xt = torch.from_numpy(x.astype(np.float32)).to("cuda:1")
    yt = m.predictions_tf(xt)
    assert yt.device.index == 1 and float(np.max(np.abs(yt.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0))) <= 1e-5
    d = planck_lite_np.load_data()
    first, start, window, _, _ = pl.binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
    eng = _engine.PlikEngine(planck_lite_np.N_ELL, start, first, window, d["X_data"].astype(np.float32),
                             pl.fisher_from_covariance(d["cov"]), (2.7255e6) ** 2, device=1)
    with np.load(os.path.join(REPO, "tests", "golden", "planck_lite.npz")) as z:
        g = {k: z[k] for k in z.files}

    def dev(a):
        buf = torch.zeros((a.shape[0], 2512), dtype=torch.float32, device="cuda:1")
        buf[:, :a.shape[1]] = torch.from_numpy(a).to("cuda:1")
        return buf[:, :a.shape[1]]
    with torch.cuda.device(1):
        ll = eng.loglike_torch(dev(g["cl_tt"]), dev(g["cl_te"]), dev(g["cl_ee"]), torch.from_numpy(g["cal"]).to("cuda:1")).cpu().numpy()
    refl = g["loglkl_np64"].reshape(-1)
    assert (np.abs(ll - refl) <= 2e-4 * np.abs(refl) + 2e-3).all()
    assert torch.cuda.current_device() == 0
    print("device-1 check passed: emulator err %.3g, likelihood ok" % err)


if __name__ == "__main__":
    main()
