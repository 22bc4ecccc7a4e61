"""Smallest end-to-end case for compute-sanitizer: one emulator, a few hundred rows, tensor-core path, no torch."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "PKLIN_NN"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 700
kind, path = standins.model_path(name)
m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
x = priors.sample_prior(m.parameters, rows, seed=3, dtype=np.float32)
y = m.engine().forward_host(x, ten_to=True)
print(name, y.shape, float(np.abs(y).max()), bool(np.all(np.isfinite(y))))
