import sys, numpy as np, torch
sys.path.insert(0, ".")
import cosmopower_b200 as cp
from cosmopower_b200 import priors, standins
for name in ("cmb_TT_NN", "PKLIN_NN"):
    kind, path = standins.model_path(name)
    m = getattr(cp, kind)(restore=True, restore_filename=path, device=0)
    n = 1 << 20
    x = torch.from_numpy(priors.sample_prior(m.parameters, n, seed=1, dtype=np.float32)).cuda()
    pitch = (m.n_modes + 7) // 8 * 8
    for label, out in (("padded pitch", torch.empty((n, pitch), dtype=torch.float32, device="cuda")[:, :m.n_modes]),
                       ("tight pitch", torch.empty((n, m.n_modes), dtype=torch.float32, device="cuda"))):
        eng = m.engine()
        eng.set_timing(True)
        for _ in range(2):
            eng.forward_torch(x, out=out, ten_to=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.forward_torch(x, out=out, ten_to=True); e1.record(); torch.cuda.synchronize()
        print("%s %s: %.3f ms, %.3f M cycles" % (name, label, e0.elapsed_time(e1), eng.last_kernel_cycles() / 1e6), flush=True)
        del out
    m.close()
