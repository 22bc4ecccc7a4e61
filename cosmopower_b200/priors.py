"""Training priors of the shipped emulators and the synthetic parameter sampler used by the
parity tests and bench.py.

Ranges: reference cosmopower/trained_models/CP_paper/CMB/README.md:5-12 and
cosmopower/trained_models/CP_paper/PK/README.md:5-14.
"""
import numpy as np

CMB_PRIOR = {
    "omega_b": (0.005, 0.04),
    "omega_cdm": (0.001, 0.99),
    "h": (0.2, 1.0),
    "tau_reio": (0.01, 0.8),
    "n_s": (0.7, 1.3),
    "ln10^{10}A_s": (1.61, 5.0),
}

PK_PRIOR = {
    "omega_b": (0.01875, 0.02625),
    "omega_cdm": (0.05, 0.255),
    "h": (0.64, 0.82),
    "n_s": (0.84, 1.1),
    "ln10^{10}A_s": (1.61, 3.91),
    "c_min": (2.0, 4.0),
    "eta_0": (0.5, 1.0),
    "z": (0.0, 5.0),
}


def prior_for(parameters):
    """Pick the prior table that covers a model's `parameters` list."""
    table = PK_PRIOR if ("z" in parameters or "c_min" in parameters) else CMB_PRIOR
    return {k: table[k] for k in parameters}


def prior_bounds(parameters):
    pr = prior_for(parameters)
    lo = np.array([pr[k][0] for k in parameters], dtype=np.float64)
    hi = np.array([pr[k][1] for k in parameters], dtype=np.float64)
    return lo, hi


def sample_prior(parameters, n, seed, dtype=np.float64):
    """(n, P) array, each column i.i.d. uniform over the prior, column order = `parameters`."""
    lo, hi = prior_bounds(parameters)
    rng = np.random.default_rng(seed)
    u = rng.random((n, len(parameters)))
    return (lo + u * (hi - lo)).astype(dtype)


def sample_prior_dict(parameters, n, seed, dtype=np.float64):
    arr = sample_prior(parameters, n, seed, dtype)
    return {k: np.ascontiguousarray(arr[:, i]) for i, k in enumerate(parameters)}
