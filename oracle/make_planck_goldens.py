"""ORACLE — test infrastructure only.  Generates tests/golden/planck_lite.npz by EXECUTING the reference's own
tf_planck2018_lite source (oracle/ref_harness.py, NumPy-backed TensorFlow stub) and checks the NumPy
restatement (oracle/planck_lite_np.py) against it.  Run in the build container (needs /root/reference):
    python oracle/make_planck_goldens.py

Rows 0-15: spectra built so that the binned band powers sit within 0.1 % - 10 % of the Planck data vector
(chi^2 from ~1e2 to ~1e6); rows 16-31: the TT / TE / EE stand-in emulators on in-prior parameter draws (not
physical, chi^2 astronomically large: a range test).  `cal` = A_planck per row.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import planck_lite_np, ref_harness  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402

N_FIT, N_EMU = 16, 16


class TableEmulator:
    """Stands where the likelihood expects an emulator: returns prescribed rows (tf_planck2018_lite.py:276-278)."""
    def __init__(self, parameters, table, log):
        self.parameters, self._t, self._log = parameters, table, log

    def ten_to_predictions_tf(self, x):
        assert self._log
        return self._t

    def predictions_tf(self, x):
        assert not self._log
        return self._t


def near_fit_spectra(oracle, rng, n):
    nl = planck_lite_np.N_ELL
    cl = np.empty((n, 3 * nl))
    cl[:] = rng.uniform(1e-7, 1e-5, size=(n, 3 * nl)) / oracle.units_factor * 1e6     # multipoles outside every bin
    for r in range(n):
        eps = 10.0 ** rng.uniform(-3, -1)
        cl[r, oracle.indices] = oracle.X_data[oracle.indices_rep] * (1 + eps * rng.standard_normal(oracle.indices.size)) / oracle.units_factor
    return cl[:, :nl], cl[:, nl:2 * nl], cl[:, 2 * nl:]


def main():
    assert ref_harness.available(), "reference checkout not found"
    rng = np.random.default_rng(20260201)
    oracle64 = planck_lite_np.PlanckLiteOracle()
    oracle32 = planck_lite_np.PlanckLiteOracle(dtype=np.float32)
    tt, te, ee = near_fit_spectra(oracle64, rng, N_FIT)
    emu = {}
    names = {"tt": "cmb_TT_NN", "te": "cmb_TE_PCAplusNN", "ee": "cmb_EE_NN"}
    params = None
    for key, name in names.items():
        kind, path = standins.model_path(name)
        ref = ref_harness.restore_reference(kind, path)
        if params is None:
            parameters = list(ref.parameters)
            params = priors.sample_prior(parameters, N_EMU, 20260202).astype(np.float32)
        with np.errstate(over="ignore"):
            y = ref.forward_pass_np(params)
        emu[key] = (10.0 ** y if key != "te" else y).astype(np.float32)
    cl_tt = np.concatenate([tt.astype(np.float32), emu["tt"]])
    cl_te = np.concatenate([te.astype(np.float32), emu["te"]])
    cl_ee = np.concatenate([ee.astype(np.float32), emu["ee"]])
    n = cl_tt.shape[0]
    cal = (1.0 + 0.0025 * rng.standard_normal(n)).astype(np.float32)

    cls = ref_harness.reference_planck_lite_class()
    prior_spec = {k: (lo, hi, "uniform") for k, (lo, hi) in priors.prior_for(parameters).items()}
    prior_spec["A_planck"] = (1.0, 0.0025, "gaussian")
    like = cls(parameters=prior_spec, tf_planck2018_lite_path=ref_harness.reference_planck_lite_path(),
               tt_emu_model=TableEmulator(parameters, cl_tt, True), te_emu_model=TableEmulator(parameters, cl_te, False),
               ee_emu_model=TableEmulator(parameters, cl_ee, True))
    full = np.concatenate([np.concatenate([np.tile(params[:1], (N_FIT, 1)), params]), cal[:, None]], axis=1).astype(np.float32)
    ref_loglkl = np.asarray(like.get_loglkl(full), dtype=np.float32)

    # the restatement's tables must be the reference's own, element for element
    assert np.array_equal(np.asarray(like.indices), oracle64.indices)
    assert np.array_equal(np.asarray(like.indices_rep), oracle64.indices_rep)
    assert np.array_equal(np.asarray(like.window_ttteee).reshape(-1), oracle32.window)
    assert np.array_equal(np.asarray(like.X_data).reshape(-1), oracle32.X_data)
    np32 = oracle32.loglike_from_spectra(cl_tt, cl_te, cl_ee, cal)
    np64 = oracle64.loglike_from_spectra(cl_tt, cl_te, cl_ee, cal)
    rel32 = np.abs(np32 - ref_loglkl) / np.abs(ref_loglkl)
    rel64 = np.abs(np64 - ref_loglkl) / np.abs(ref_loglkl)
    print("float32 restatement vs executed reference: max rel %.3g" % rel32.max())
    print("float64 restatement vs executed reference (float32): max rel %.3g" % rel64.max())
    print("chi2 range: %.4g .. %.4g" % (float(-2 * np64.max()), float(-2 * np64.min())))
    # prior handling (:309-326, :349-365)
    pr = np.asarray(like.priors.prob(full)).reshape(-1, 1)
    ref_masked = np.asarray(like.loglkl(full, pr))
    out = os.path.join(REPO, "tests", "golden", "planck_lite.npz")
    np.savez_compressed(out, cl_tt=cl_tt, cl_te=cl_te, cl_ee=cl_ee, cal=cal, params=full,
                        parameters=np.array(parameters + ["A_planck"]),
                        loglkl_ref_f32=ref_loglkl, loglkl_np64=np64, prodpri=pr, loglkl_masked_ref=ref_masked)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
