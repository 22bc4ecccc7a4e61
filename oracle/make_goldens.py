"""ORACLE — test infrastructure only.  Generates tests/golden/<model>.npz by EXECUTING the
reference's own source (oracle/ref_harness.py) on seeded in-prior inputs.  Run in the build
container (needs /root/reference):  python oracle/make_goldens.py

Each file holds: `params` float64 (n, P) in the model's column order, `parameters` (names),
`pred` float64 (n, n_modes) = reference predictions_np on the float64 inputs, and
`pred_f32in` float32 = reference predictions_np on the float32-cast inputs (the reference then
computes in float32).  Rows 0/1/2 are the prior's lower corner, upper corner and centre; row 3 is
the README example point (reference README.md:115-121) where the model has those parameters.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ref_harness  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402

ROWS = {"cmb_TT_NN": 24, "cmb_EE_NN": 24, "cmb_TE_PCAplusNN": 24, "cmb_PP_PCAplusNN": 24,
        "PKLIN_NN": 64, "PKNLBOOST_NN": 64}
README_POINT = {"omega_b": 0.0225, "omega_cdm": 0.113, "h": 0.7, "tau_reio": 0.055,
                "n_s": 0.96, "ln10^{10}A_s": 3.07, "z": 0.5, "c_min": 2.6, "eta_0": 0.7}


def golden_inputs(parameters, n, seed):
    lo, hi = priors.prior_bounds(parameters)
    x = priors.sample_prior(parameters, n, seed)
    x[0], x[1], x[2] = lo, hi, 0.5 * (lo + hi)
    x[3] = [README_POINT[k] for k in parameters]
    return x


def main():
    assert ref_harness.available(), "reference checkout not found"
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for idx, (name, n) in enumerate(ROWS.items()):
        kind, path = standins.model_path(name)
        ref = ref_harness.restore_reference(kind, path)
        x = golden_inputs(ref.parameters, n, 20260101 + idx)
        d64 = {k: x[:, i] for i, k in enumerate(ref.parameters)}
        d32 = {k: x[:, i].astype(np.float32) for i, k in enumerate(ref.parameters)}
        with np.errstate(over="ignore"):
            pred = ref.predictions_np(d64)
            pred32 = ref.predictions_np(d32)
            ten = ref.ten_to_predictions_np(d64)
        assert pred.dtype == np.float64 and pred32.dtype == np.float32
        assert np.array_equal(ten, 10. ** pred)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), params=x,
                            parameters=np.array(list(ref.parameters)), pred=pred, pred_f32in=pred32,
                            standin=np.bool_(name in standins.STANDIN_MODELS))
        print(name, x.shape, pred.shape, float(pred.min()), float(pred.max()))


if __name__ == "__main__":
    main()
