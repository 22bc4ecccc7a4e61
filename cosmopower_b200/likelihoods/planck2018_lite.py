"""Planck 2018 plik-lite likelihood evaluated on the GPU behind the B200 emulators.

Host-side mirror of the reference class `tf_planck2018_lite`
(cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py): same constructor keywords (:38-47), same
public attributes (`parameters`, `parameter_keys`, `indices`, `indices_rep`, `window_ttteee`, `X_data`,
`fisher`, `nbin_tot`, ...), same methods `get_loglkl` (:257-304), `loglkl` (:309-326) and `posterior`
(:349-365), on torch CUDA tensors instead of TensorFlow ones.  The three emulators run through the fused
kernel (TT and EE with 10**, TE linear, :276-278) and the binning + Fisher quadratic form run in
cpb200_plik_loglike_device; there is no CPU path.

The Planck data vector, window functions and covariance come from the flat file written by
tools/convert_planck_lite.py (`tf_planck2018_lite_path` may instead point at a directory holding the
reference's own `data/planck2018_plik_lite/*.dat`, read as the reference reads them, :104-124, :206-216).
"""
import os

import numpy as np

from .. import _engine

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_NPZ = os.path.join(_HERE, "data", "planck2018_plik_lite.npz")


def _read_reference_dir(data_dir, nbin_hi):
    from scipy.io import FortranFile
    bval, x_data, x_sig = np.genfromtxt(os.path.join(data_dir, "cl_cmb_plik_v22.dat"), unpack=True)
    blmin = np.loadtxt(os.path.join(data_dir, "blmin.dat")).astype(int)
    blmax = np.loadtxt(os.path.join(data_dir, "blmax.dat")).astype(int)
    bin_w = np.loadtxt(os.path.join(data_dir, "bweight.dat"))
    f = FortranFile(os.path.join(data_dir, "c_matrix_plik_v22.dat"), "r")
    cov = f.read_reals(dtype=float).reshape((nbin_hi, nbin_hi))
    f.close()
    cov = np.tril(cov) + np.tril(cov, -1).T
    return dict(bval=bval, X_data=x_data, X_sig=x_sig, blmin=blmin, blmax=blmax, bin_w=bin_w, cov=cov)


def binning_tables(blmin, blmax, bin_w, n_ell=2507, plmin_tt=30, plmin=30, nbins=(215, 199, 199)):
    """Gather ranges of the three spectra in the concatenated [TT | TE | EE] vector, as the reference builds them
    (:126-202).  Returns (gather_first, bin_start, window, indices, indices_rep): a bin gathers the
    bin_start[b+1]-bin_start[b] consecutive multipoles that start at gather_first[b]; `indices`, `indices_rep`
    and `window` are the reference's flat tables (self.indices, self.indices_rep, self.window_ttteee)."""
    first, start, window, indices, indices_rep = [], [0], [], [], []
    b = 0
    for nb, pl, off in zip(nbins, (plmin_tt, plmin, plmin), (0, n_ell, 2 * n_ell)):
        for i in range(nb):
            lo, hi = int(blmin[i]) + pl - 2, int(blmax[i]) + pl + 1 - 2
            first.append(off + lo)
            start.append(start[-1] + hi - lo)
            window.append(bin_w[blmin[i]:blmax[i] + 1])
            indices.append(off + np.arange(lo, hi))
            indices_rep.append(np.full(hi - lo, b))
            b += 1
    return (np.array(first, dtype=np.int32), np.array(start, dtype=np.int32), np.concatenate(window).astype(np.float32),
            np.concatenate(indices), np.concatenate(indices_rep))


def fisher_from_covariance(cov):
    """Fisher matrix of the band powers (:206-230).  The reference inverts the float32 covariance by Cholesky in
    float32; here the float32-cast covariance is inverted in float64 and rounded once."""
    c = np.asarray(cov, dtype=np.float32).astype(np.float64)
    L = np.linalg.cholesky(c)
    y = np.linalg.solve(L, np.eye(c.shape[0]))
    return np.linalg.solve(L.T, y).T.astype(np.float32)


def fold_binning_into_pca_emulator(attrs, first, start, window, bin0, n_bins_te, n_ell):
    """Attributes of a cosmopower_PCAplusNN emulator that predicts the band powers of its (linear) spectrum directly.

    predictions = (c . pca_std + pca_mean) . P . features_std + features_mean   (cosmopower_PCAplusNN.py:381) is linear in
    the PCA coefficients c, and so is the binning sum_l w_l C_l (tf_planck2018_lite.py:284-291): folding the window
    functions into the basis, P'[p][b] = sum_{l in b} P[p][l] features_std[l] w_l and mean'[b] = sum_{l in b}
    features_mean[l] w_l, gives a 12x narrower output contraction and band powers that never exist as multipoles."""
    P = np.asarray(attrs["pca_transform_matrix_"], dtype=np.float64)
    fs = np.asarray(attrs["features_std_"], dtype=np.float64)
    fm = np.asarray(attrs["features_mean_"], dtype=np.float64)
    Pb = np.zeros((P.shape[0], n_bins_te))
    mb = np.zeros(n_bins_te)
    for b in range(n_bins_te):
        lo = int(first[bin0 + b]) - n_ell                    # TE is the second block of the concatenated row
        w = np.asarray(window[start[bin0 + b]:start[bin0 + b + 1]], dtype=np.float64)
        sl = slice(lo, lo + w.size)
        Pb[:, b] = (P[:, sl] * fs[sl]) @ w
        mb[b] = fm[sl] @ w
    col = np.abs(Pb).max(axis=0)
    col[col == 0] = 1.0
    out = dict(attrs)
    out["pca_transform_matrix_"] = (Pb / col).astype(np.float32)
    out["features_mean_"] = mb.astype(np.float32)
    out["features_std_"] = col.astype(np.float32)
    out["n_modes"] = n_bins_te
    out["modes"] = np.arange(n_bins_te)
    return out


class _Prior:
    """Product of independent uniform / gaussian densities (reference set_priors, :331-346, without tfp)."""

    def __init__(self, parameters):
        self.kinds, self.a, self.b = [], [], []
        for name in parameters:
            low, high, kind = parameters[name]
            if kind not in ("uniform", "gaussian"):
                raise ValueError("prior of %r must be 'uniform' or 'gaussian'" % name)
            self.kinds.append(kind); self.a.append(float(low)); self.b.append(float(high))

    def prob(self, x):
        import torch
        p = torch.ones(x.shape[0], dtype=torch.float32, device=x.device)
        for i, (kind, a, b) in enumerate(zip(self.kinds, self.a, self.b)):
            xi = x[:, i]
            if kind == "uniform":
                p = p * torch.where((xi >= a) & (xi <= b), torch.full_like(xi, 1.0 / (b - a)), torch.zeros_like(xi))
            else:
                p = p * torch.exp(-0.5 * ((xi - a) / b) ** 2) / (b * float(np.sqrt(2 * np.pi)))
        return p


class tf_planck2018_lite:
    def __init__(self, parameters=None, tf_planck2018_lite_path=None, spectra="TTTEEE", use_low_ell_bins=False,
                 units_factor=(2.7255e6) ** 2, tt_emu_model=None, te_emu_model=None, ee_emu_model=None, device=None,
                 chunk_rows=262144, fold_te=True, cuda_graphs=False, fuse_binning=True):
        self.parameters = parameters
        self.priors = _Prior(parameters) if parameters is not None else None
        self.spectra = spectra
        self.use_low_ell_bins = use_low_ell_bins
        self.units_factor = units_factor
        if tt_emu_model is None:
            raise ValueError("Unrecognised emulator for TT spectrum")
        if te_emu_model is None:
            raise ValueError("Unrecognised emulator for TE spectrum")
        if ee_emu_model is None:
            raise ValueError("Unrecognised emulator for EE spectrum")
        self.tt_emu_model, self.te_emu_model, self.ee_emu_model = tt_emu_model, te_emu_model, ee_emu_model
        if not (list(tt_emu_model.parameters) == list(te_emu_model.parameters) == list(ee_emu_model.parameters)):
            raise ValueError("different ordering in emulators")          # the reference prints and stops (:73-75)
        if spectra != "TTTEEE":
            raise ValueError("Spectra must be TTTEEE")                   # :93-95
        if use_low_ell_bins:
            # the reference sets nbintt_low_ell = 2 but never loads the low-ell data (:77-82): its Fisher matrix stays
            # 613 x 613 against 615 bins, so that branch cannot run there either
            raise NotImplementedError("use_low_ell_bins=True is not runnable in the reference; not provided")
        self.parameter_keys = list(parameters.keys()) if parameters is not None else list(tt_emu_model.parameters) + ["A_planck"]
        self.nbintt_low_ell, self.plmin_TT = 0, 30
        self.plmin, self.plmax = 30, 2508
        self.nbintt_hi, self.nbinte, self.nbinee = 215, 199, 199
        self.nbin_hi = self.nbintt_hi + self.nbinte + self.nbinee
        self.nbintt = self.nbintt_hi + self.nbintt_low_ell
        self.nbin_tot = self.nbintt + self.nbinte + self.nbinee

        if tf_planck2018_lite_path is not None and os.path.isdir(os.path.join(tf_planck2018_lite_path, "data", "planck2018_plik_lite")):
            self.data_directory = os.path.join(tf_planck2018_lite_path, "data")
            d = _read_reference_dir(os.path.join(self.data_directory, "planck2018_plik_lite"), self.nbin_hi)
        else:
            self.data_directory = os.path.dirname(DATA_NPZ)
            with np.load(DATA_NPZ) as z:
                d = {k: z[k] for k in z.files}
        self.bval, self.X_sig = d["bval"], d["X_sig"]
        self.blmin, self.blmax = d["blmin"].astype(int), d["blmax"].astype(int)
        self.bin_w = d["bin_w"].astype("float32")
        self.blmin_TT, self.blmax_TT, self.bin_w_TT = self.blmin, self.blmax, self.bin_w
        x_data = d["X_data"].astype("float32")

        n_ell = self.plmax - 1
        for m in (tt_emu_model, te_emu_model, ee_emu_model):
            if int(m.n_modes) != n_ell:
                raise ValueError("emulators must predict %d multipoles (ell = 2 .. %d)" % (n_ell, self.plmax))
        first, start, window, self.indices, self.indices_rep = binning_tables(
            self.blmin, self.blmax, self.bin_w, n_ell, self.plmin_TT, self.plmin, (self.nbintt, self.nbinte, self.nbinee))
        self.window_ttteee = window.reshape(1, -1)
        self.X_data = x_data.reshape(1, -1)
        self.fisher = self.get_inverse_covmat(d["cov"])
        self.device = device if device is not None else (getattr(tt_emu_model, "_device", None) or 0)
        self.chunk_rows = int(chunk_rows)
        # batches up to this size leave most SMs idle: the three emulators then run side by side on three streams
        self.concurrent_rows = 8192
        self._side_streams = None
        # cuda_graphs=True: get_loglkl for a small batch is captured once per batch size (side-by-side emulators, tail, all
        # of their launches) and replayed afterwards - a sampler's sequence of single points pays one graph launch per call
        self.cuda_graphs = bool(cuda_graphs)
        self._graphs = {}
        self._cols = None
        # A linear PCA emulator for TE gets the window functions folded into its basis (fold_te=True, the default):
        # it then predicts the 199 TE band powers directly and the likelihood kernel never sees TE multipoles.
        self._te_binned = None
        if fold_te and hasattr(te_emu_model, "_attrs") and "pca_transform_matrix_" in te_emu_model._attrs():
            folded = fold_binning_into_pca_emulator(te_emu_model._attrs(), first, start, self.window_ttteee.reshape(-1),
                                                    self.nbintt, self.nbinte, n_ell)
            te_prec = getattr(te_emu_model, "_precision", "f16x3")
            self._te_binned = _engine.Engine(folded, device=self.device, precision="f16x3" if te_prec == "auto" else te_prec)
        # fuse_binning=True (the default when TT and EE are tensor-core NN emulators): the window weights are applied and the
        # multipoles summed into per-4-multipole partial band powers inside the emulators' output epilogues
        # (cpb200_engine_set_binning): 5 KB per spectrum row leave the SM instead of 10 KB and the tail adds partials.
        self._binned = False
        if fuse_binning and all(hasattr(m, "engine") and getattr(m, "_precision", "") in ("f16x3", "auto") and not hasattr(m, "pca_transform_matrix_")
                                for m in (tt_emu_model, ee_emu_model)):
            w = self.window_ttteee.reshape(-1)
            for m, b0, nb, off in ((tt_emu_model, 0, self.nbintt, 0), (ee_emu_model, self.nbintt + self.nbinte, self.nbinee, 2 * n_ell)):
                weight, bin_of = np.zeros(n_ell, np.float32), np.full(n_ell, -1, np.int32)
                for b in range(b0, b0 + nb):
                    lo, ln = int(first[b]) - off, int(start[b + 1] - start[b])
                    weight[lo:lo + ln] = w[start[b]:start[b + 1]]
                    bin_of[lo:lo + ln] = b - b0
                m.engine(self.device).set_binning(weight, bin_of)
            self._binned = True
        self._engine = _engine.PlikEngine(n_ell, start, first, self.window_ttteee.reshape(-1), x_data, self.fisher,
                                          units_factor, device=self.device, te_binned=self._te_binned is not None, partials=self._binned)
        # Emulators built with precision="auto": batches of up to SMALL_ROWS rows (a sampler's single point) go through the
        # emulators' fp32 column-per-warp kernels (35-50 us per call instead of 100-120 us) and a tail of their own (spectra
        # in, no partials): with cuda_graphs=True one log-likelihood takes 70 us instead of 160 us.
        self._small = None
        if all(getattr(m, "_precision", "") == "auto" for m in (tt_emu_model, te_emu_model, ee_emu_model)):
            te_small = _engine.Engine(folded, device=self.device, precision="fp32") if self._te_binned is not None else None
            self._small = (te_small, _engine.PlikEngine(n_ell, start, first, self.window_ttteee.reshape(-1), x_data, self.fisher, units_factor,
                                                        device=self.device, te_binned=te_small is not None, partials=False))

    # ===== INVERSE COVARIANCE ======  (:206-230)
    def get_inverse_covmat(self, cov):
        return fisher_from_covariance(cov)

    def _split(self, parameters):
        import torch
        if not isinstance(parameters, torch.Tensor):
            parameters = torch.as_tensor(np.asarray(parameters, dtype=np.float32))
        parameters = parameters.to(device="cuda:%d" % self.device, dtype=torch.float32)
        keys = self.parameter_keys
        cal = parameters[:, keys.index("A_planck")]
        if self._cols is None or self._cols.device != parameters.device:     # device-side column index, built once
            self._cols = torch.tensor([keys.index(k) for k in self.tt_emu_model.parameters], dtype=torch.int64, device=parameters.device)
        return parameters, parameters.index_select(1, self._cols), cal

    def _emulate(self, x):
        """TT, TE and EE for the rows x (:276-278).  A small batch occupies a few clusters per emulator (the engine's
        small-batch mode), so the three launches go to three streams and share the device instead of queueing."""
        import torch
        te = (lambda: self._te_binned.forward_torch(x, ten_to=False)) if self._te_binned is not None else (lambda: self.te_emu_model.predictions_tf(x))
        if self._use_small(x.shape[0]):
            if self._small[0] is not None:
                te = lambda: self._small[0].forward_torch(x, ten_to=False)
            tt = lambda: self.tt_emu_model.ten_to_predictions_tf(x)          # precision="auto": the fp32 kernels at this size
            ee = lambda: self.ee_emu_model.ten_to_predictions_tf(x)
        elif self._binned:
            tt = lambda: self.tt_emu_model.engine(self.device).forward_torch(x, ten_to=True, binned=True)
            ee = lambda: self.ee_emu_model.engine(self.device).forward_torch(x, ten_to=True, binned=True)
        else:
            tt = lambda: self.tt_emu_model.ten_to_predictions_tf(x)
            ee = lambda: self.ee_emu_model.ten_to_predictions_tf(x)
        if x.shape[0] > self.concurrent_rows:
            return tt(), te(), ee()
        main = torch.cuda.current_stream(x.device)
        if self._side_streams is None:
            self._side_streams = [torch.cuda.Stream(device=x.device), torch.cuda.Stream(device=x.device)]
        s1, s2 = self._side_streams
        s1.wait_stream(main)                       # x is ready
        s2.wait_stream(main)
        cl_tt = tt()
        with torch.cuda.stream(s1):
            cl_te = te()
        with torch.cuda.stream(s2):
            cl_ee = ee()
        main.wait_stream(s1)
        main.wait_stream(s2)
        if not torch.cuda.is_current_stream_capturing():
            for t in (cl_te, cl_ee):
                t.record_stream(main)              # allocated on a side stream, consumed on the caller's
            x.record_stream(s1)
            x.record_stream(s2)
        return cl_tt, cl_te, cl_ee

    def _use_small(self, n_rows):
        return self._small is not None and n_rows <= type(self.tt_emu_model).SMALL_ROWS

    def get_loglkl(self, parameters):
        """(N, len(parameter_keys)) -> (N, 1) log-likelihood, as the reference's get_loglkl (:257-304)."""
        import torch
        if self.cuda_graphs and 0 < len(parameters) <= self.concurrent_rows:
            return self._get_loglkl_graphed(parameters)
        return self._get_loglkl(parameters)

    def _get_loglkl_graphed(self, parameters):
        """Replay of the CUDA graph captured for this batch size (captured on first use; at most 8 sizes are kept)."""
        import torch
        dev = "cuda:%d" % self.device
        if not isinstance(parameters, torch.Tensor):
            parameters = torch.as_tensor(np.asarray(parameters, dtype=np.float32))
        n = parameters.shape[0]
        entry = self._graphs.get(n)
        if entry is None:
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            static_in = torch.empty((n, len(self.parameter_keys)), dtype=torch.float32, device=dev)
            static_in.copy_(parameters)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(static_in.device))
            with torch.cuda.stream(side):            # warm-up outside the capture (lazy allocations, first launches)
                for _ in range(2):
                    self._get_loglkl(static_in)
            torch.cuda.current_stream(static_in.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._get_loglkl(static_in)
            entry = self._graphs[n] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(parameters)
        graph.replay()
        return static_out.clone()

    def _get_loglkl(self, parameters):
        import torch
        parameters, cosmo, cal = self._split(parameters)
        n = cosmo.shape[0]
        out = torch.empty((n, 1), dtype=torch.float32, device=cosmo.device)
        for r0 in range(0, n, self.chunk_rows):
            x = cosmo[r0:r0 + self.chunk_rows]
            cl_tt, cl_te, cl_ee = self._emulate(x)
            tail = self._small[1] if self._use_small(x.shape[0]) else self._engine
            out[r0:r0 + x.shape[0], 0] = tail.loglike_torch(cl_tt, cl_te, cl_ee, cal[r0:r0 + x.shape[0]])
        return out

    def loglkl(self, parameters, prodpri):
        """:309-326 - rows whose prior probability is zero get -1e12."""
        import torch
        loglike = self.get_loglkl(parameters)
        prodpri = torch.as_tensor(prodpri).to(loglike.device).reshape(-1, 1)
        return torch.where(prodpri == 0.0, torch.full_like(loglike, -0.5 * 2e12), loglike)

    def posterior(self, params):
        """:349-365 - log-likelihood + log-prior."""
        import torch
        parameters, _, _ = self._split(params)
        pr = self.priors.prob(parameters).reshape(-1, 1)
        logpr = torch.log(pr)
        loglike = self.loglkl(parameters, pr)
        logpr = torch.where(torch.isinf(logpr), torch.full_like(logpr, -1e32), logpr)
        return loglike + logpr

    def close(self):
        self._engine.close()
        if self._te_binned is not None:
            self._te_binned.close()
        if self._small is not None:
            if self._small[0] is not None:
                self._small[0].close()
            self._small[1].close()
            self._small = None



