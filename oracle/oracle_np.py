"""ORACLE — test infrastructure only.  Never imported by the product package.

CPU restatement (NumPy) of the reference's emulator forward pass.  Each function names the
reference lines it follows (paths relative to the upstream checkout).  It is pinned by
`tests/golden/*.npz`, which were produced by *executing the reference's own source*
(`oracle/ref_harness.py` + `oracle/make_goldens.py`) in the build container; see
tests/test_oracle.py.  The reference's own test-suite holds no golden vectors for this path
(cosmopower/tests/test_NN.py:53 and test_PCAplusNN.py:64-65 are commented out), so parity is
pinned by those self-generated vectors, not by upstream known-answer tests.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import
this module.
"""
import numpy as np


def dict_to_ordered_arr_np(parameters, input_dict):
    """cosmopower/cosmopower_NN.py:333-350 and cosmopower/cosmopower_PCAplusNN.py:330-347."""
    if parameters is not None:
        return np.stack([input_dict[k] for k in parameters], axis=1)
    return np.stack([input_dict[k] for k in input_dict], axis=1)


def _trunk(m, parameters_arr):
    """Hidden stack shared by both classes: cosmopower_NN.py:370-378 / PCAplusNN.py:367-375."""
    x = (parameters_arr - m["parameters_mean_"]) / m["parameters_std_"]
    for i in range(m["n_layers"] - 1):
        a = np.dot(x, m["W_"][i]) + m["b_"][i]
        x = (m["betas_"][i] + (1. - m["betas_"][i]) * 1. / (1. + np.exp(-m["alphas_"][i] * a))) * a
    return x


def forward_pass_nn(m, parameters_arr):
    """cosmopower/cosmopower_NN.py:354-384."""
    x = _trunk(m, parameters_arr)
    y = np.dot(x, m["W_"][-1]) + m["b_"][-1]                       # :381
    return y * m["features_std_"] + m["features_mean_"]             # :384


def forward_pass_pcann(m, parameters_arr):
    """cosmopower/cosmopower_PCAplusNN.py:351-381."""
    x = _trunk(m, parameters_arr)
    c = np.dot(x, m["W_"][-1]) + m["b_"][-1]                       # :378
    return (np.dot(c * m["pca_std_"] + m["pca_mean_"], m["pca_transform_matrix_"])
            * m["features_std_"] + m["features_mean_"])            # :381


def pca_coefficients(m, parameters_arr):
    """cosmopower/cosmopower_PCAplusNN.py:163-193 (`forward_pass_tf`: the network up to the rescaled PCA coefficients),
    restated in NumPy with the arithmetic of the class's own NumPy pass (:367-378)."""
    with np.errstate(over="ignore"):
        x = _trunk(m, parameters_arr)
    c = np.dot(x, m["W_"][-1]) + m["b_"][-1]                       # :187
    return c * m["pca_std_"] + m["pca_mean_"]                      # :190-193


def forward_pass(m, parameters_arr):
    with np.errstate(over="ignore"):
        if "pca_transform_matrix_" in m:
            return forward_pass_pcann(m, parameters_arr)
        return forward_pass_nn(m, parameters_arr)


def predictions_np(m, parameters_dict):
    """cosmopower_NN.py:388-405 / cosmopower_PCAplusNN.py:384-401."""
    return forward_pass(m, dict_to_ordered_arr_np(m["parameters"], parameters_dict))


def ten_to_predictions_np(m, parameters_dict):
    """cosmopower_NN.py:409-425 / cosmopower_PCAplusNN.py:405-421."""
    return 10. ** predictions_np(m, parameters_dict)


def forward_chunked(m, parameters_arr, ten_to=False, chunk=16384):
    """Row-chunked driver: the reference keeps every intermediate alive
    (cosmopower_NN.py:370-378), so a 1M-row call does not fit in host memory."""
    outs = []
    for s in range(0, parameters_arr.shape[0], chunk):
        y = forward_pass(m, parameters_arr[s:s + chunk])
        outs.append(10. ** y if ten_to else y)
    return np.concatenate(outs, axis=0) if outs else np.zeros((0, m["n_modes"]))


def parity_error(y, y_ref, floor=1.0):
    """Error metric of DESIGN.md §Tolerance: max |y - y_ref| / max(|y_ref|, floor)."""
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    return float(np.max(np.abs(y - y_ref) / np.maximum(np.abs(y_ref), floor))) if y.size else 0.0
