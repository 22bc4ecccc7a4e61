"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py's CPU leg).

NumPy restatement of the reference's Planck 2018 plik-lite likelihood tail
(cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py): the consumer that sits directly behind the
TT / TE / EE emulators.  Follows, by reference line:

  :75-86    geometry (plmin 30, plmax 2508, 215 + 199 + 199 bins)
  :126-202  gather indices / window weights / bin ids of the three spectra in the concatenated C_ell vector
  :206-230  Fisher matrix = inverse of the (mirrored) covariance
  :257-304  get_loglkl: units factor, gather * window, segment sum, 1 / A_planck^2, chi^2, -chi^2 / 2
  :309-326  loglkl: rows with zero prior probability get -1e12

Pinned against the executed reference source (oracle/ref_harness.py: reference_planck_lite, NumPy-backed
TensorFlow stub) by oracle/make_goldens.py -> tests/golden/planck_lite.npz.  `dtype=np.float32` reproduces the
reference's float32 arithmetic class; the default float64 is the truth the CUDA path is compared with.
"""
import os

import numpy as np

DATA_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "cosmopower_b200", "likelihoods", "data", "planck2018_plik_lite.npz")

PLMIN, PLMAX = 30, 2508
NBIN_TT, NBIN_TE, NBIN_EE = 215, 199, 199
N_ELL = PLMAX - 1                  # 2507 multipoles per spectrum (ell = 2 .. 2508)


def load_data(path=DATA_NPZ):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def binning_tables(blmin, blmax, bin_w):
    """(indices, window, bin id) of every gathered element, in the reference's order (:131-202, 196-197)."""
    idx, win, rep = [], [], []
    offsets = (0, N_ELL, 2 * N_ELL)            # TT, TE, EE blocks of the concatenated vector (:157, :169)
    first_bin = (0, NBIN_TT, NBIN_TT + NBIN_TE)
    for nb, off, b0 in zip((NBIN_TT, NBIN_TE, NBIN_EE), offsets, first_bin):
        for i in range(nb):
            lo = blmin[i] + PLMIN - 2              # :126, :128, :130
            hi = blmax[i] + PLMIN + 1 - 2          # :127, :129, :131
            idx.append(off + np.arange(lo, hi))
            win.append(bin_w[blmin[i]:blmax[i] + 1])      # :133-138, :147, :159, :171
            rep.append(np.full(hi - lo, b0 + i))
    return np.concatenate(idx), np.concatenate(win), np.concatenate(rep)


def fisher_matrix(cov, dtype=np.float64):
    """:206-230 - cholesky_solve(cholesky(cov), I), transposed."""
    c = np.asarray(cov, dtype=dtype)
    L = np.linalg.cholesky(c)
    eye = np.eye(c.shape[0], dtype=dtype)
    y = np.linalg.solve(L, eye)
    return np.linalg.solve(L.T, y).T.astype(dtype)


class PlanckLiteOracle:
    def __init__(self, data=None, units_factor=(2.7255e6) ** 2, dtype=np.float64):
        d = data if data is not None else load_data()
        self.dtype = dtype
        self.units_factor = units_factor
        self.X_data = d["X_data"].astype(np.float32).astype(dtype)        # :108 casts the data vector to float32
        self.indices, window, self.indices_rep = binning_tables(d["blmin"], d["blmax"], d["bin_w"].astype(np.float32))
        self.window = window.astype(dtype)                                 # :123 reads the weights as float32
        self.fisher = fisher_matrix(d["cov"].astype(np.float32), dtype)    # :224 casts the covariance to float32
        self.nbin = NBIN_TT + NBIN_TE + NBIN_EE

    def bin_spectra(self, cl_tt, cl_te, cl_ee):
        """:281-291 - returns (N, 613) binned band powers (before the calibration)."""
        cl = self.units_factor * np.concatenate([cl_tt, cl_te, cl_ee], axis=1).astype(self.dtype)
        g = cl[:, self.indices] * self.window[None, :]
        out = np.zeros((cl.shape[0], self.nbin), dtype=self.dtype)
        np.add.at(out.T, self.indices_rep, g.T)            # segment_sum over the gathered axis
        return out

    def loglike_from_spectra(self, cl_tt, cl_te, cl_ee, cal):
        """:293-304 - cal is A_planck per row; returns (N, 1)."""
        cl_bin = self.bin_spectra(cl_tt, cl_te, cl_ee)
        x_model = cl_bin / np.square(np.asarray(cal, dtype=self.dtype)).reshape(-1, 1)
        diff = self.X_data[None, :] - x_model
        chi2 = np.einsum("ni,ij,nj->n", diff, self.fisher, diff)
        return (-0.5 * chi2).reshape(-1, 1)

    def loglkl(self, loglike, prodpri):
        """:309-326."""
        return np.where(np.asarray(prodpri).reshape(-1, 1) == 0.0, -0.5 * 2e12, loglike)
