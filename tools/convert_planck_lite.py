"""One-time conversion of the Planck 2018 plik-lite data files (public Planck PR3 data shipped beside the
reference's likelihood) into one flat .npz for cosmopower_b200.likelihoods.  Runs where the reference checkout
is mounted; the GPU box only sees the converted file.

Reads what tf_planck2018_lite.__init__ reads (tf_planck2018_lite.py:107-124, 206-216):
cl_cmb_plik_v22.dat (bval, X_data, X_sig), blmin.dat, blmax.dat, bweight.dat and the Fortran-unformatted
c_matrix_plik_v22.dat (613 x 613 float64, lower triangle valid, mirrored here as the reference does).

usage: python tools/convert_planck_lite.py [reference_likelihood_dir]
"""
import os
import sys

import numpy as np
from scipy.io import FortranFile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = "/root/reference/cosmopower/likelihoods/tf_planck2018_lite"


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else DEFAULT
    d = os.path.join(src, "data", "planck2018_plik_lite")
    bval, x_data, x_sig = np.genfromtxt(os.path.join(d, "cl_cmb_plik_v22.dat"), unpack=True)
    blmin = np.loadtxt(os.path.join(d, "blmin.dat")).astype(np.int32)
    blmax = np.loadtxt(os.path.join(d, "blmax.dat")).astype(np.int32)
    bin_w = np.loadtxt(os.path.join(d, "bweight.dat"))
    nbin = 215 + 199 + 199
    f = FortranFile(os.path.join(d, "c_matrix_plik_v22.dat"), "r")
    cov = f.read_reals(dtype=float).reshape((nbin, nbin))
    f.close()
    cov = np.tril(cov) + np.tril(cov, -1).T            # covmat[i, j] = covmat[j, i] for j >= i
    out = os.path.join(REPO, "cosmopower_b200", "likelihoods", "data", "planck2018_plik_lite.npz")
    np.savez_compressed(out, bval=bval, X_data=x_data, X_sig=x_sig, blmin=blmin, blmax=blmax, bin_w=bin_w, cov=cov)
    print("wrote", out, os.path.getsize(out), "bytes; bins", nbin, "cov cond", np.linalg.cond(cov))


if __name__ == "__main__":
    main()
