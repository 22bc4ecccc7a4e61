"""CPU emulation of cheaper operand splits for the dense contractions: how far is each from the
1e-5 tolerance?  Test infrastructure (uses the float64 oracle); nothing here is on the product path.

Every contraction A.W is evaluated in float64 on operands rounded the way the tensor core would see them:
  f16x1        A_hi.W_hi                                    (1 pass)
  f16x2_a      A_hi.W_hi + A_lo.W_hi                        (2 passes, weights in plain fp16)
  f16x2_w      A_hi.W_hi + A_hi.W_lo                        (2 passes, activations in plain fp16)
  f16+f8x2     A_hi.W_hi + e4m3(A_lo).e4m3(W_hi) + e4m3(A_hi).e4m3(W_lo)   (1 fp16 pass + 2 fp8 passes = 2 pass-equivalents)
  f16x3        A_hi.W_hi + A_lo.W_hi + A_hi.W_lo            (3 passes, what the kernel runs)
with the same power-of-two range conditioning as the kernel (weights to [16384, 32768), activations x64, parameters
x1024); accumulation error is not modelled (fp64 sums), so these are lower bounds on the error of each format.

usage: python tests/manual/split_format_study.py [rows]
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from cosmopower_b200 import _loader, priors, standins  # noqa: E402
from oracle import oracle_np  # noqa: E402
from tests import parity  # noqa: E402


def f16(x):
    return x.astype(np.float16).astype(np.float64)


def f8(x):
    """e4m3 rounding; each operand is first scaled by a power of two per matrix so that its largest entry sits at the top
    of the e4m3 range (what a per-tensor scale would do), then scaled back."""
    m = np.max(np.abs(x))
    if m == 0:
        return x
    s = 2.0 ** np.floor(np.log2(256.0 / m))
    t = torch.from_numpy((x * s).astype(np.float32)).to(torch.float8_e4m3fn).to(torch.float32).numpy().astype(np.float64)
    return t / s


def contract(a, w, mode):
    wmax = np.max(np.abs(w))
    sw = 2.0 ** np.floor(np.log2(32768.0 / wmax)) if wmax > 0 else 1.0
    ws = w * sw
    a_hi, w_hi = f16(a), f16(ws)
    a_lo, w_lo = f16(a - a_hi), f16(ws - w_hi)
    y = a_hi @ w_hi
    if mode in ("f16x3", "f16x2_a"):
        y = y + a_lo @ w_hi
    if mode in ("f16x3", "f16x2_w"):
        y = y + a_hi @ w_lo
    if mode == "f16+f8x2":
        y = y + f8(a_lo) @ f8(w_hi) + f8(a_hi) @ f8(w_lo)
    return y / sw


def forward(m, x, mode):
    h = (x - m["parameters_mean_"].astype(np.float64)) / m["parameters_std_"].astype(np.float64) * 1024.0
    scale = 1024.0
    n = int(m["n_layers"])
    for i in range(n - 1):
        a = contract(h, m["W_"][i].astype(np.float64), mode) / scale + m["b_"][i]
        with np.errstate(over="ignore"):
            h = (m["betas_"][i] + (1. - m["betas_"][i]) / (1. + np.exp(-m["alphas_"][i].astype(np.float64) * a))) * a
        h = h * 64.0
        scale = 64.0
    y = contract(h, m["W_"][-1].astype(np.float64), mode) / scale + m["b_"][-1]
    if "pca_transform_matrix_" in m:
        c = (y * m["pca_std_"] + m["pca_mean_"]) * 64.0
        y = contract(c, m["pca_transform_matrix_"].astype(np.float64), mode) / 64.0
    return y * m["features_std_"] + m["features_mean_"]


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    print("worst parity error over %d in-prior rows (metric of tests/parity.py; tolerance 1e-5, TE 1e-4)" % rows)
    modes = ["f16x1", "f16x2_a", "f16x2_w", "f16+f8x2", "f16x3"]
    print("%-18s " % "model" + " ".join("%10s" % k for k in modes))
    for name in parity.MODELS:
        kind, path = standins.model_path(name)
        fields = _loader.PCANN_FIELDS if "PCA" in kind else _loader.NN_FIELDS
        m = _loader.load_npz(path + ".npz", fields) if not os.path.isfile(path + ".pkl") else \
            _loader.attributes_from_list(_loader.load_pickle_list(path + ".pkl"), fields)
        x = priors.sample_prior(m["parameters"], rows, seed=11, dtype=np.float64)
        ref = oracle_np.forward_pass(m, x)
        errs = [parity.model_error(name, forward(m, x, k), ref) for k in modes]
        print("%-18s " % name + " ".join("%10.2e" % e for e in errs), flush=True)


if __name__ == "__main__":
    main()
