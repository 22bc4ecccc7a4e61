"""The converted Planck-lite data file equals what the reference's own loader reads (only where the reference
checkout is mounted: the build container)."""
import os

import numpy as np
import pytest

from oracle import planck_lite_np, ref_harness


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not mounted")
def test_converted_data_equals_the_reference_files():
    from scipy.io import FortranFile
    d = planck_lite_np.load_data()
    src = os.path.join(ref_harness.reference_planck_lite_path(), "data", "planck2018_plik_lite")
    bval, x_data, x_sig = np.genfromtxt(os.path.join(src, "cl_cmb_plik_v22.dat"), unpack=True)
    assert np.array_equal(d["bval"], bval) and np.array_equal(d["X_data"], x_data) and np.array_equal(d["X_sig"], x_sig)
    assert np.array_equal(d["blmin"], np.loadtxt(os.path.join(src, "blmin.dat")).astype(int))
    assert np.array_equal(d["blmax"], np.loadtxt(os.path.join(src, "blmax.dat")).astype(int))
    assert np.array_equal(d["bin_w"], np.loadtxt(os.path.join(src, "bweight.dat")))
    f = FortranFile(os.path.join(src, "c_matrix_plik_v22.dat"), "r")
    cov = f.read_reals(dtype=float).reshape((613, 613))
    f.close()
    for i in range(613):                      # the reference mirrors the lower triangle (tf_planck2018_lite.py:212-214)
        cov[i, i:] = cov[i:, i]
    assert np.array_equal(d["cov"], cov)


def test_data_file_shapes():
    d = planck_lite_np.load_data()
    assert d["X_data"].shape == (613,) and d["cov"].shape == (613, 613) and d["bin_w"].ndim == 1
    assert np.array_equal(d["cov"], d["cov"].T)
    assert (np.diff(d["blmin"][:215]) > 0).all() and (d["blmax"][:215] >= d["blmin"][:215]).all()
