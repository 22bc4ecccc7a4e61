"""Where a mid-size host call spends its time (B200 only): the same batch through the dict API (float64 and float32), the
host-buffer C ABI call with pinned and pageable buffers, and the device-tensor call.   Usage: python tools/latency_breakdown.py [model]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import _engine, priors, standins  # noqa: E402


def med(fn, reps=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cmb_TT_NN"
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    eng = m.engine()
    for n in (1, 128, 1024, 8192):
        x64 = priors.sample_prior(m.parameters, n, seed=1)
        x32 = x64.astype(np.float32)
        d64 = {k: x64[:, i] for i, k in enumerate(m.parameters)}
        d32 = {k: x32[:, i] for i, k in enumerate(m.parameters)}
        xt = torch.from_numpy(x32).cuda()
        pin_in = _engine.PinnedArray((n, x32.shape[1])); pin_in.array[:] = x32
        pin_out = _engine.PinnedArray((n, m.n_modes))
        page_out = np.empty((n, m.n_modes), np.float32)
        def dev():
            m.ten_to_predictions_tf(xt); torch.cuda.synchronize()
        res = {
            "dict float64 -> new float64 array": med(lambda: m.ten_to_predictions_np(d64)),
            "dict float32 -> new float32 array": med(lambda: m.ten_to_predictions_np(d32)),
            "dict_to_ordered_arr_np alone": med(lambda: m.dict_to_ordered_arr_np(d64)),
            "forward_host pinned -> pinned": med(lambda: eng.forward_host(pin_in.array, out=pin_out.array, ten_to=True)),
            "forward_host pageable -> reused pageable float32": med(lambda: eng.forward_host(x32, out=page_out, ten_to=True)),
            "np.empty float64 + fill (first touch)": med(lambda: np.empty((n, m.n_modes), np.float64).fill(1.0)),
            "device tensor call, synchronised": med(dev),
        }
        print("%s rows=%d" % (name, n))
        for k, v in res.items():
            print("    %-52s %9.1f us" % (k, v))
        pin_in.free(); pin_out.free()


if __name__ == "__main__":
    main()
