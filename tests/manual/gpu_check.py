"""Diagnostic run on a B200: parity of one precision mode against the oracle on all six models,
plus a quick timing.  Usage: python tests/manual/gpu_check.py <fp32|f16x3> [rows]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import priors, standins  # noqa: E402
from oracle import oracle_np  # noqa: E402
from tests import parity  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    import torch
    print("device:", torch.cuda.get_device_name(0), flush=True)
    for name in parity.MODELS:
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=0, precision=prec)
        attrs = standins.load_attrs(kind, path)
        x = priors.sample_prior(m.parameters, rows, seed=3)
        y_ref = oracle_np.forward_pass(attrs, x)
        t0 = time.time()
        y = m.forward_pass_np(x)
        dt = time.time() - t0
        err = parity.model_error(name, y, y_ref)
        t = m.engine().forward_host(x.astype(np.float32), ten_to=True)
        with np.errstate(over="ignore", invalid="ignore"):
            rel = parity.ten_to_error(t, y_ref) if name not in parity.LINEAR_MODELS else float("nan")
        xd = torch.from_numpy(x.astype(np.float32)).cuda()
        yd = m.predictions_tf(xd)
        torch.cuda.synchronize()
        same = bool(np.array_equal(yd.cpu().numpy(), y.astype(np.float32)))
        # timing on a bigger device-resident batch
        nb = 131072 if prec == "f16x3" else 16384
        xb = torch.from_numpy(priors.sample_prior(m.parameters, nb, seed=4, dtype=np.float32)).cuda()
        ob = torch.empty((nb, m.n_modes), dtype=torch.float32, device="cuda")
        eng = m.engine()
        eng.forward_torch(xb, out=ob)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.forward_torch(xb, out=ob)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("%-18s prec=%s rows=%d err=%.3g (tol %.0e) ten_to_logerr=%.3g dev==host:%s first_call=%.3fs | %d rows in %.3f ms = %.3g spectra/s"
              % (name, prec, rows, err, parity.model_tol(name), rel, same, dt, nb, ms, nb / ms * 1e3), flush=True)
        m.close()


if __name__ == "__main__":
    main()
