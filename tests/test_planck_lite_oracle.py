"""CPU tests of the Planck-lite likelihood tail: the NumPy oracle against the executed reference's golden
vectors (tests/golden/planck_lite.npz, written by oracle/make_planck_goldens.py) and the host-side tables of
cosmopower_b200.likelihoods against the oracle's."""
import os

import numpy as np

from cosmopower_b200.likelihoods import planck2018_lite as pl
from oracle import planck_lite_np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden", "planck_lite.npz")


def _golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def test_float32_oracle_reproduces_the_executed_reference():
    g = _golden()
    o32 = planck_lite_np.PlanckLiteOracle(dtype=np.float32)
    ll = o32.loglike_from_spectra(g["cl_tt"], g["cl_te"], g["cl_ee"], g["cal"])
    # same algorithm in the same precision class; only the summation order of BLAS / segment sums differs
    np.testing.assert_allclose(ll, g["loglkl_ref_f32"], rtol=2e-5)


def test_float64_oracle_matches_its_golden_and_the_float32_reference():
    g = _golden()
    o64 = planck_lite_np.PlanckLiteOracle()
    ll = o64.loglike_from_spectra(g["cl_tt"], g["cl_te"], g["cl_ee"], g["cal"])
    np.testing.assert_allclose(ll, g["loglkl_np64"], rtol=1e-10)
    np.testing.assert_allclose(ll, g["loglkl_ref_f32"], rtol=1e-4)        # the reference's own float32 rounding
    masked = o64.loglkl(ll, g["prodpri"])
    np.testing.assert_allclose(masked, g["loglkl_masked_ref"], rtol=1e-4)
    assert (masked[g["prodpri"].reshape(-1) == 0] == -1e12).all()


def test_host_tables_equal_the_oracles():
    d = planck_lite_np.load_data()
    first, start, window, indices, indices_rep = pl.binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
    o = planck_lite_np.PlanckLiteOracle()
    assert np.array_equal(indices, o.indices) and np.array_equal(indices_rep, o.indices_rep)
    assert np.array_equal(window, o.window.astype(np.float32))
    assert start[0] == 0 and start[-1] == indices.size == 6413 and len(first) == 613
    # a bin is one contiguous run of multipoles
    for b in (0, 100, 214, 215, 413, 414, 612):
        assert np.array_equal(indices[start[b]:start[b + 1]], first[b] + np.arange(start[b + 1] - start[b]))
    f = pl.fisher_from_covariance(d["cov"])
    np.testing.assert_allclose(f, o.fisher, rtol=2e-6, atol=1e-30)
    assert f.dtype == np.float32 and np.abs(f - f.T).max() <= 1e-6 * np.abs(f).max()


def test_segment_sum_is_a_window_average():
    # every window sums to one (bweight.dat is normalised per bin): a flat spectrum bins to itself
    o = planck_lite_np.PlanckLiteOracle()
    flat = np.full((1, planck_lite_np.N_ELL), 3.0 / o.units_factor)
    binned = o.bin_spectra(flat, flat, flat)
    np.testing.assert_allclose(binned, 3.0, rtol=1e-6)


def test_folding_the_windows_into_the_te_basis_is_exact():
    """fold_binning_into_pca_emulator: the folded emulator's output (oracle forward pass on the folded attributes) equals
    the binned output of the original TE emulator - both steps are linear, so only float32 storage of the folded basis
    separates them; the folded basis is well conditioned for the engine's fp16 split (unit-peak columns)."""
    from cosmopower_b200 import priors, standins
    from oracle import oracle_np
    standins.build_standins()
    kind, path = standins.model_path("cmb_TE_PCAplusNN")
    attrs = standins.load_attrs(kind, path)
    d = planck_lite_np.load_data()
    first, start, window, _, _ = pl.binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
    first, start, window = np.asarray(first), np.asarray(start), np.asarray(window).reshape(-1)
    n_ell = planck_lite_np.N_ELL
    is_te = first // n_ell == 1
    bin0, n_te = int(np.argmax(is_te)), int(is_te.sum())
    folded = pl.fold_binning_into_pca_emulator(attrs, first, start, window, bin0, n_te, n_ell)
    assert folded["pca_transform_matrix_"].shape == (attrs["pca_transform_matrix_"].shape[0], n_te)
    assert np.allclose(np.abs(folded["pca_transform_matrix_"]).max(axis=0), 1.0)
    x = priors.sample_prior(attrs["parameters"], 64, seed=3)
    y = oracle_np.forward_pass(attrs, x)
    binned = np.stack([y[:, first[bin0 + b] - n_ell: first[bin0 + b] - n_ell + start[bin0 + b + 1] - start[bin0 + b]]
                       @ window[start[bin0 + b]:start[bin0 + b + 1]].astype(np.float64) for b in range(n_te)], axis=1)
    y_folded = oracle_np.forward_pass(folded, x)
    assert np.abs(y_folded - binned).max() <= 2e-7 * np.abs(binned).max()
