"""Tolerance definitions shared by the parity tests (DESIGN.md "Tolerance").

Log-spectrum models (TT, EE, PP, PKLIN, PKNLBOOST): err = max |y - y_ref| / max(|y_ref|, 1) on the
log10 output of predictions_np against the float64 oracle; must be <= 1e-5 (north star).
Linear-spectrum model (TE): err = max |y - y_ref| / (row peak |y_ref|) <= 1e-4 (the reference's own
float32 path sits at 1.2e-5 by this metric on the stand-in).  SURVEY 7.2 proposed the stricter per-mode floor
`max |y - y_ref| / max(|y_ref|, features_std_) <= 1e-5` (`linear_error_strict`): the reference's OWN float32 pass
(what its TF twin computes) misses that at 7.7e-5 on the stand-in - the spectrum crosses zero and its
standard deviation is a tenth of the row peak - so no float32 implementation can be held to it.  The GPU tests
report it beside the row-peak metric and bound it by a small multiple of the reference's float32 figure
(tests/test_parity_gpu.py::test_te_tolerance_in_both_metrics).
ten_to_predictions_np: the same log-space metric applied to log10 of the returned spectrum (a
relative error of 10**y is ln10 * the log-space error, so it is the log error that is bounded).
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["cmb_TT_NN", "cmb_EE_NN", "cmb_TE_PCAplusNN", "cmb_PP_PCAplusNN", "PKLIN_NN", "PKNLBOOST_NN"]
LINEAR_MODELS = {"cmb_TE_PCAplusNN"}
TOL_LOG = 1e-5
TOL_LINEAR = 1e-4


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def log_error(y, y_ref):
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    return float(np.max(np.abs(y - y_ref) / np.maximum(np.abs(y_ref), 1.0)))


def linear_error(y, y_ref):
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    peak = np.max(np.abs(y_ref), axis=1, keepdims=True)
    return float(np.max(np.abs(y - y_ref) / peak))


def linear_error_strict(y, y_ref, features_std):
    """SURVEY 7.2's metric for the linear-spectrum model: |dy| / max(|y_ref|, features_std_) per mode."""
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    floor = np.maximum(np.abs(y_ref), np.asarray(features_std, np.float64)[None, :])
    return float(np.max(np.abs(y - y_ref) / floor))


def model_error(name, y, y_ref):
    return linear_error(y, y_ref) if name in LINEAR_MODELS else log_error(y, y_ref)


def model_tol(name):
    return TOL_LINEAR if name in LINEAR_MODELS else TOL_LOG


def ten_to_error(t, y_ref):
    """log-space error of a 10**y output (log models only)."""
    with np.errstate(divide="ignore"):
        return log_error(np.log10(np.asarray(t, np.float64)), y_ref)
