"""GPU parity of the Planck-lite likelihood tail (cpb200_plik_*) against the float64 NumPy oracle
(oracle/planck_lite_np.py, pinned to the executed reference by tests/golden/planck_lite.npz).

Tolerance: the reference computes this part in float32 (tf_planck2018_lite.py:108, :224, :281-304) and so does
the CUDA path; against the float64 oracle the log-likelihood must agree to
    |ll - ll_ref| <= 2e-4 * |ll_ref| + 2e-3
(the absolute term covers near-fit rows, where chi^2 ~ 1 is a small difference of large products).  The model
band powers must agree to 2e-6 of the row's largest |band power|.
"""
import os

import numpy as np
import pytest

from oracle import oracle_np, planck_lite_np

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_REL, TOL_ABS, TOL_BIN = 2e-4, 2e-3, 2e-6


def _close(ll, ref):
    ll = np.asarray(ll, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    err = np.abs(ll - ref)
    bound = TOL_REL * np.abs(ref) + TOL_ABS
    assert (err <= bound).all(), "worst %.3g of bound at row %d (ll %.6g ref %.6g)" % (
        float((err / bound).max()), int((err / bound).argmax()), ll[(err / bound).argmax()], ref[(err / bound).argmax()])
    return float((err / np.maximum(np.abs(ref), 1.0)).max())


@pytest.fixture(scope="module")
def plik():
    from cosmopower_b200 import _engine
    from cosmopower_b200.likelihoods import planck2018_lite as pl
    d = planck_lite_np.load_data()
    first, start, window, _, _ = pl.binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
    eng = _engine.PlikEngine(planck_lite_np.N_ELL, start, first, window, d["X_data"].astype(np.float32),
                             pl.fisher_from_covariance(d["cov"]), (2.7255e6) ** 2, device=0)
    yield eng
    eng.close()


def _padded(a, pitch=2512):
    import torch
    buf = torch.zeros((a.shape[0], pitch), dtype=torch.float32, device="cuda")
    buf[:, :a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    return buf[:, :a.shape[1]]


def test_golden_spectra(plik):
    import torch
    with np.load(os.path.join(REPO, "tests", "golden", "planck_lite.npz")) as z:
        g = {k: z[k] for k in z.files}
    ll, binned = plik.loglike_torch(_padded(g["cl_tt"]), _padded(g["cl_te"]), _padded(g["cl_ee"]),
                                    torch.from_numpy(g["cal"]).cuda(), return_binned=True)
    torch.cuda.synchronize()
    o = planck_lite_np.PlanckLiteOracle()
    ref_binned = o.bin_spectra(g["cl_tt"], g["cl_te"], g["cl_ee"]) / np.square(g["cal"].astype(np.float64))[:, None]
    berr = np.abs(binned.cpu().numpy() - ref_binned).max(axis=1) / np.abs(ref_binned).max(axis=1)
    assert berr.max() <= TOL_BIN, berr.max()
    worst = _close(ll.cpu().numpy(), g["loglkl_np64"])
    print("planck-lite golden: worst relative log-likelihood error %.3g, band powers %.3g" % (worst, berr.max()))
    assert plik.launch_count == 3


@pytest.mark.parametrize("n", [1, 127, 129, 1000])
def test_ragged_sizes_and_tight_pitch(plik, n):
    import torch
    rng = np.random.default_rng(n)
    o = planck_lite_np.PlanckLiteOracle()
    nl = planck_lite_np.N_ELL
    cl = np.abs(rng.standard_normal((n, 3 * nl))) * 1e-12
    cl[:, o.indices] = o.X_data[o.indices_rep] * (1 + 0.01 * rng.standard_normal((n, o.indices.size))) / o.units_factor
    cl = cl.astype(np.float32)
    cal = (1 + 0.0025 * rng.standard_normal(n)).astype(np.float32)
    ref = o.loglike_from_spectra(cl[:, :nl], cl[:, nl:2 * nl], cl[:, 2 * nl:], cal)
    parts = [torch.from_numpy(np.ascontiguousarray(cl[:, i * nl:(i + 1) * nl])).cuda() for i in range(3)]   # pitch 2507: scalar loads
    ll = plik.loglike_torch(*parts, torch.from_numpy(cal).cuda())
    _close(ll.cpu().numpy(), ref)
    ll2 = plik.loglike_torch(_padded(cl[:, :nl]), _padded(cl[:, nl:2 * nl]), _padded(cl[:, 2 * nl:]), torch.from_numpy(cal).cuda())
    assert torch.equal(ll, ll2)                       # the vector and scalar load paths agree bit for bit


def test_tensor_core_quadratic_form_against_the_fp32_cuda_core_one(plik):
    """Two independent implementations of chi^2 (split-fp16 tcgen05 contraction vs fp32 FMA kernel) on 20000 rows whose
    chi^2 spans twelve decades; both share the binning kernel."""
    import torch
    from cosmopower_b200 import _engine
    from cosmopower_b200.likelihoods import planck2018_lite as pl
    d = planck_lite_np.load_data()
    first, start, window, _, _ = pl.binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
    os.environ["CPB200_PLIK_FP32"] = "1"
    try:
        ref_eng = _engine.PlikEngine(planck_lite_np.N_ELL, start, first, window, d["X_data"].astype(np.float32),
                                     pl.fisher_from_covariance(d["cov"]), (2.7255e6) ** 2, device=0)
    finally:
        del os.environ["CPB200_PLIK_FP32"]
    n = 20000
    rng = np.random.default_rng(99)
    o = planck_lite_np.PlanckLiteOracle()
    nl = planck_lite_np.N_ELL
    cl = np.abs(rng.standard_normal((n, 3 * nl))) * 1e-12
    eps = 10.0 ** rng.uniform(-4, 2, size=(n, 1))
    cl[:, o.indices] = o.X_data[o.indices_rep] * (1 + eps * rng.standard_normal((n, o.indices.size))) / o.units_factor
    parts = [_padded(cl[:, i * nl:(i + 1) * nl]) for i in range(3)]
    cal = torch.ones(n, device="cuda")
    a = plik.loglike_torch(*parts, cal).cpu().numpy().astype(np.float64)
    b = ref_eng.loglike_torch(*parts, cal).cpu().numpy().astype(np.float64)
    ref_eng.close()
    assert np.isfinite(a).all() and np.isfinite(b).all()
    err = np.abs(a - b)
    assert (err <= TOL_REL * np.abs(b) + TOL_ABS).all(), float((err / (np.abs(b) + 1)).max())
    assert -2 * b.max() < 1e3 and -2 * b.min() > 1e9        # the range really was covered


def test_calibration_scaling_property_at_full_chunk_size(plik):
    """X_model = C_bin / A_planck^2 (tf_planck2018_lite.py:294): scaling every spectrum by a^2 and the calibration by a
    must not change the log-likelihood.  150 000 rows (more than one engine chunk), a = 2 is exact in binary."""
    import torch
    n = 150000
    g = torch.Generator(device="cuda").manual_seed(5)
    o = planck_lite_np.PlanckLiteOracle()
    nl = planck_lite_np.N_ELL
    base = torch.zeros((1, 3 * nl), dtype=torch.float32, device="cuda")
    base[0, torch.from_numpy(o.indices).cuda()] = torch.from_numpy((o.X_data[o.indices_rep] / o.units_factor).astype(np.float32)).cuda()
    noise = 1 + 0.05 * torch.randn((n, 3 * nl), generator=g, device="cuda")
    cl = base * noise
    parts = []
    for i in range(3):
        buf = torch.empty((n, 2512), dtype=torch.float32, device="cuda")[:, :nl]
        buf.copy_(cl[:, i * nl:(i + 1) * nl])
        parts.append(buf)
    cal = 1 + 0.0025 * torch.randn(n, generator=g, device="cuda")
    ll1 = plik.loglike_torch(*parts, cal)
    for p in parts:
        p.mul_(4.0)
    ll2 = plik.loglike_torch(*parts, cal * 2.0)
    assert torch.isfinite(ll1).all()
    assert torch.equal(ll1, ll2)          # powers of two commute with every rounding on the path


def test_empty_batch(plik):
    import torch
    e = torch.empty((0, planck_lite_np.N_ELL), dtype=torch.float32, device="cuda")
    assert plik.loglike_torch(e, e, e, torch.empty((0,), device="cuda")).shape == (0,)


def test_chunk_boundaries_and_determinism(plik):
    """More rows than one engine chunk (131072): every row must equal the same row evaluated in a small batch."""
    import torch
    n = 131072 + 300
    rng = np.random.default_rng(7)
    base = np.abs(rng.standard_normal((64, 3 * planck_lite_np.N_ELL))).astype(np.float32) * 1e-10
    rows = rng.integers(0, 64, size=n)
    nl = planck_lite_np.N_ELL
    small = [_padded(base[:, i * nl:(i + 1) * nl]) for i in range(3)]
    cal = torch.ones(64, device="cuda")
    ll_small = plik.loglike_torch(*small, cal)
    idx = torch.from_numpy(rows).cuda()
    big = [torch.empty((n, 2512), dtype=torch.float32, device="cuda")[:, :nl] for _ in range(3)]
    for b, s in zip(big, small):
        b.copy_(s[idx])
    ll_big = plik.loglike_torch(*big, torch.ones(n, device="cuda"))
    assert torch.equal(ll_big, ll_small[idx])
    assert torch.equal(ll_big, plik.loglike_torch(*big, torch.ones(n, device="cuda")))


@pytest.mark.parametrize("fold_te", [True, False])
def test_end_to_end_class_against_the_oracle_pipeline(fold_te):
    """tf_planck2018_lite mirror: emulators + likelihood on the device vs oracle emulators + oracle likelihood.
    fold_te=True (default): the TE emulator predicts its 199 band powers directly (window functions folded into its
    PCA basis); False: it writes multipoles like TT and EE."""
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import priors, standins
    from cosmopower_b200.likelihoods import tf_planck2018_lite
    models, attrs = {}, {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        attrs[key] = standins.load_attrs(kind, path)
    params = list(models["tt"].parameters)
    spec = {k: (lo, hi, "uniform") for k, (lo, hi) in priors.prior_for(params).items()}
    spec["A_planck"] = (1.0, 0.0025, "gaussian")
    like = tf_planck2018_lite(parameters=spec, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"],
                              fold_te=fold_te)
    assert (like._te_binned is not None) == fold_te
    n = 2048
    x = priors.sample_prior(params, n, seed=11)
    x[:8, 0] = 1.0                                     # omega_b outside its prior: zero prior probability
    cal = 1 + 0.0025 * np.random.default_rng(3).standard_normal(n)
    full = np.concatenate([x, cal[:, None]], axis=1).astype(np.float32)
    x32 = full[:, :len(params)].astype(np.float64)
    with np.errstate(over="ignore"):
        cl_tt = 10.0 ** oracle_np.forward_pass(attrs["tt"], x32)
        cl_te = oracle_np.forward_pass(attrs["te"], x32)
        cl_ee = 10.0 ** oracle_np.forward_pass(attrs["ee"], x32)
    o = planck_lite_np.PlanckLiteOracle()
    ref = o.loglike_from_spectra(cl_tt, cl_te, cl_ee, full[:, -1].astype(np.float64))
    ll = like.get_loglkl(torch.from_numpy(full).cuda()).cpu().numpy()
    assert ll.shape == (n, 1)
    # emulator tolerance (1e-5 in log10 C_ell) enters chi^2 twice: allow 1e-3 relative end to end
    rel = np.abs(ll[8:] - ref[8:]) / np.abs(ref[8:])          # the first 8 rows are far outside the prior (overflow)
    assert rel.max() <= 1e-3, rel.max()
    if fold_te:
        # the folded TE emulator against the oracle's TE spectrum binned in float64 (before units and calibration)
        te_b = like._te_binned.forward_torch(torch.from_numpy(full[:, :len(params)]).cuda(), ten_to=False).cpu().numpy()
        ref_b = o.bin_spectra(np.zeros_like(cl_te), cl_te, np.zeros_like(cl_te))[:, 215:414] / o.units_factor
        peak = np.abs(ref_b[8:]).max(axis=1, keepdims=True)
        assert te_b.shape == (n, 199) and (np.abs(te_b[8:] - ref_b[8:]) / peak).max() <= 1e-4
    post = like.posterior(torch.from_numpy(full).cuda()).cpu().numpy()
    assert (post[:8] <= -1e12).all() and np.isfinite(post[8:]).all()
    pr = like.priors.prob(torch.from_numpy(full).cuda()).cpu().numpy()
    assert (pr[:8] == 0).all() and (pr[8:] > 0).all()
    masked = like.loglkl(torch.from_numpy(full).cuda(), pr).cpu().numpy()
    assert (masked[:8] == -1e12).all() and np.array_equal(masked[8:], ll[8:])
    like.close()
    for m in models.values():
        m.close()


def test_cuda_graph_replay_equals_the_plain_call():
    """cuda_graphs=True: small batches replay a CUDA graph captured per batch size (emulators on three streams + tail);
    every replay must equal the plain call bit for bit, for fresh inputs and for several batch sizes."""
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import priors, standins
    from cosmopower_b200.likelihoods import tf_planck2018_lite
    models = {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    params = list(models["tt"].parameters)
    plain = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"])
    graphed = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"],
                                 cuda_graphs=True)
    for n in (1, 7, 1, 300, 7):
        for seed in (1, 2, 3):
            x = priors.sample_prior(params, n, seed=100 * n + seed, dtype=np.float32)
            full = torch.from_numpy(np.concatenate([x, np.full((n, 1), 1.001, np.float32)], axis=1)).cuda()
            a = graphed.get_loglkl(full)
            b = plain.get_loglkl(full)
            assert a.shape == (n, 1) and torch.equal(a, b), (n, seed)
    assert sorted(graphed._graphs) == [1, 7, 300]
    big = torch.from_numpy(np.concatenate([priors.sample_prior(params, 9000, seed=9, dtype=np.float32),
                                           np.ones((9000, 1), np.float32)], axis=1)).cuda()
    assert torch.equal(graphed.get_loglkl(big), plain.get_loglkl(big)) and 9000 not in graphed._graphs
    plain.close()
    graphed.close()


def test_binning_fused_into_the_emulators_matches_the_unfused_path_and_the_oracle():
    """VERDICT r1 item 5 / tf_planck2018_lite.py:281-291: with fuse_binning=True (default) the TT and EE emulators apply
    the window weights in their output epilogues and write per-4-multipole partial band powers (half the bytes of the
    spectra); the tail adds a bin's partials.  Against the unfused pipeline (spectra written, binned by the tail): band
    powers within 2e-6 of the row peak, log-likelihoods within the end-to-end tolerance of the oracle comparison; bit-stable
    from run to run; ragged batch sizes."""
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import priors, standins
    from cosmopower_b200.likelihoods import tf_planck2018_lite
    models, attrs = {}, {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
        attrs[key] = standins.load_attrs(kind, path)
    params = list(models["tt"].parameters)
    kw = dict(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"])
    fused = tf_planck2018_lite(fuse_binning=True, **kw)
    plain = tf_planck2018_lite(fuse_binning=False, **kw)
    assert fused._binned and not plain._binned
    o = planck_lite_np.PlanckLiteOracle()
    for n in (1, 300, 5001):
        x = priors.sample_prior(params, n, seed=40 + n)
        cal = 1 + 0.0025 * np.random.default_rng(n).standard_normal(n)
        full = torch.from_numpy(np.concatenate([x, cal[:, None]], axis=1).astype(np.float32)).cuda()
        a = fused.get_loglkl(full)
        assert torch.equal(a, fused.get_loglkl(full))                               # deterministic
        b = plain.get_loglkl(full)
        x32 = full[:, :len(params)].cpu().numpy().astype(np.float64)
        with np.errstate(over="ignore"):
            ref = o.loglike_from_spectra(10.0 ** oracle_np.forward_pass(attrs["tt"], x32), oracle_np.forward_pass(attrs["te"], x32),
                                         10.0 ** oracle_np.forward_pass(attrs["ee"], x32), full[:, -1].cpu().numpy().astype(np.float64))
        for ll in (a, b):
            rel = np.abs(ll.cpu().numpy() - ref) / np.abs(ref)
            assert rel.max() <= 1e-3, (n, rel.max())
        # the band powers themselves: fused partial sums against the tail's own binning of the written spectra
        cosmo = full[:, :len(params)].contiguous()
        ones = torch.ones(n, device="cuda")
        te = fused._te_binned.forward_torch(cosmo, ten_to=False)
        _, bp_f = fused._engine.loglike_torch(models["tt"].engine(0).forward_torch(cosmo, ten_to=True, binned=True), te,
                                              models["ee"].engine(0).forward_torch(cosmo, ten_to=True, binned=True), ones, return_binned=True)
        _, bp_p = plain._engine.loglike_torch(models["tt"].ten_to_predictions_tf(cosmo), te, models["ee"].ten_to_predictions_tf(cosmo), ones,
                                              return_binned=True)
        for lo, hi in ((0, 215), (414, 613)):
            peak = bp_p[:, lo:hi].abs().max(dim=1, keepdim=True).values
            assert float(((bp_f[:, lo:hi] - bp_p[:, lo:hi]).abs() / peak).max()) <= 2e-6, (n, lo)
        assert torch.equal(bp_f[:, 215:414], bp_p[:, 215:414])                       # TE goes the same way in both
    fused.close(); plain.close()
    for m in models.values():
        m.close()


def test_single_points_through_auto_precision_emulators():
    """Emulators built with precision="auto": batches of up to 8 rows go through the fp32 column-per-warp kernels and a tail of
    their own, larger ones through the tensor path with fused binning - both inside the end-to-end tolerance of the oracle
    pipeline, the small path also as a CUDA-graph replay."""
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import priors, standins
    from cosmopower_b200.likelihoods import tf_planck2018_lite
    models, attrs = {}, {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=0, precision="auto")
        attrs[key] = standins.load_attrs(kind, path)
    params = list(models["tt"].parameters)
    o = planck_lite_np.PlanckLiteOracle()
    for graphs in (False, True):
        like = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"],
                                  cuda_graphs=graphs)
        assert like._small is not None and like._binned
        for n in (1, 8, 9, 300):
            x = priors.sample_prior(params, n, seed=50 + n)
            cal = 1 + 0.0025 * np.random.default_rng(n).standard_normal(n)
            full = np.concatenate([x, cal[:, None]], axis=1).astype(np.float32)
            x64 = full[:, :len(params)].astype(np.float64)
            with np.errstate(over="ignore"):
                ref = o.loglike_from_spectra(10.0 ** oracle_np.forward_pass(attrs["tt"], x64), oracle_np.forward_pass(attrs["te"], x64),
                                             10.0 ** oracle_np.forward_pass(attrs["ee"], x64), full[:, -1].astype(np.float64))
            ll = like.get_loglkl(torch.from_numpy(full).cuda()).cpu().numpy()
            assert ll.shape == (n, 1)
            rel = np.abs(ll - ref) / np.abs(ref)
            assert rel.max() <= 1e-3, (graphs, n, rel.max())
        like.close()
    for m in models.values():
        m.close()
