#!/usr/bin/env python
"""Benchmark of the CosmoPower batched emulator forward pass on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): cmb_TT_NN + cmb_EE_NN, ten_to_predictions on a 1,048,576-row
batch of in-prior synthetic LCDM parameters per model and per GPU.  One "step" = TT batch + EE
batch (2 kernel launches, 2,097,152 spectra per GPU).  TT/EE weights are SYNTHETIC STAND-INS with
the documented architecture (their pickles are stripped from the reference checkout).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = same metric through the host-buffer C-ABI call (pinned host in/out, H2D + D2H inside the
timed region) with the host's measured aggregate D2H ceiling beside it; `roofline` = tensor-pipe
roofline of the fused kernel; `per_model` = every BASELINE config's emulator at 1,048,576 rows (launch
time from CUDA events, tensor / HBM fractions, DRAM bytes of the committed ncu capture, parity of a
4096-row sample against the oracle computed in the same run, outside the timed region); `strong` = ONE
1,048,576-row batch per model split over the N ranks (the BASELINE metric read literally);
`likelihood` = the Planck-lite consumer of the CMB emulators (get_loglkl on 262,144 parameter sets, parity against the oracle
pipeline); `sweep` = BASELINE configs[4] at bench size (six emulators streamed over 8 M parameter sets per rank);
`cpu_baseline` = the reference's NumPy predictions path on the host cores, bounded sample.
`--impl reference` times that CPU path alone, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

BATCH = 1 << 20
MODELS = ["cmb_TT_NN", "cmb_EE_NN"]
ALL_MODELS = ["cmb_TT_NN", "cmb_EE_NN", "cmb_TE_PCAplusNN", "cmb_PP_PCAplusNN", "PKLIN_NN", "PKNLBOOST_NN"]
METRIC = "spectra/sec (batch 1M)"
# algorithmic work per spectrum, SURVEY 8(d): FLOP = 2*sum(in*out) over every matrix (1 pass), bytes = 4*(P + n_modes)
FLOP_PER_SPECTRUM = {"cmb_TT_NN": 4146176, "cmb_EE_NN": 4146176, "cmb_TE_PCAplusNN": 4670464,
                     "cmb_PP_PCAplusNN": 1965440, "PKLIN_NN": 1484800, "PKNLBOOST_NN": 1486848}
BYTES_PER_SPECTRUM = {"cmb_TT_NN": 10052, "cmb_EE_NN": 10052, "cmb_TE_PCAplusNN": 10052,
                      "cmb_PP_PCAplusNN": 10052, "PKLIN_NN": 1704, "PKNLBOOST_NN": 1712}
STANDINS = {"cmb_TT_NN", "cmb_EE_NN", "cmb_TE_PCAplusNN"}
PASSES = 3                                                            # fp16x3 split
CPU_SAMPLE_ROWS = 65536                                               # rows per model per CPU step


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def load_traffic():
    """DRAM bytes (read + write) of one 1,048,576-row launch per model, from the committed ncu --set full captures
    (profiles/traffic.json: {"models": {name: {"dram_bytes_read", "dram_bytes_write", "source"}}})."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return {k: (v["dram_bytes_read"] + v["dram_bytes_write"], v.get("source")) for k, v in d["models"].items()}
    except Exception:
        return {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                smax = float(f[1])
                if t0 <= t <= t1 + 0.1:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def emit(line):          # replaced in main(): writes to the real stdout
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# the reference's CPU implementation of the path
# --------------------------------------------------------------------------------------------------
def cpu_models(names, sample_rows, dtype):
    """[(callable doing one ten_to_predictions_np call, kind)] per model.  Prefers the executed, unmodified reference
    source staged by __graft_entry__.build() in oracle/_ref/ (kind "reference"); else the NumPy port (kind "port")."""
    from cosmopower_b200 import priors, standins
    from oracle import oracle_np, ref_harness
    use_ref = ref_harness.available()
    out = []
    for i, name in enumerate(names):
        kind, path = standins.model_path(name)
        attrs = standins.load_attrs(kind, path)
        x = priors.sample_prior(attrs["parameters"], sample_rows, seed=20260101 + i)
        d = {k: x[:, j].astype(dtype) for j, k in enumerate(attrs["parameters"])}
        if use_ref and os.path.isfile(path + ".pkl"):
            obj = ref_harness.restore_reference(kind, path)
            out.append((lambda obj=obj, d=d: obj.ten_to_predictions_np(d), "reference"))
        else:
            out.append((lambda attrs=attrs, d=d: oracle_np.ten_to_predictions_np(attrs, d), "port"))
    return out


def cpu_path(sample_rows, steps, warmup, dtype=np.float64):
    """Times the reference's CPU path (all BLAS threads).  float64 inputs are the reference's documented usage (dicts
    of lists); float32 arrays make its whole pass float32.  Returns (spectra/s, s/step, cores, kind)."""
    models = cpu_models(MODELS, sample_rows, dtype)
    kind = "reference" if all(k == "reference" for _, k in models) else "port"

    def step():
        for fn, _ in models:
            with np.errstate(over="ignore"):
                fn()

    def timed():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return time.perf_counter() - t0
    want = os.cpu_count() or 1
    try:   # torchrun exports OMP_NUM_THREADS=1: give BLAS every host core back for the CPU arm
        from threadpoolctl import threadpool_info, threadpool_limits
        with threadpool_limits(limits=want):
            dt = timed()
            cores = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        dt = timed()
        cores = int(os.environ.get("OMP_NUM_THREADS", want))
    return steps * len(MODELS) * sample_rows / dt, dt / steps, cores, kind


def cpu_kind_text(kind):
    if kind == "reference":
        return ("the unmodified reference source (cosmopower_NN.ten_to_predictions_np, staged in oracle/_ref, executed "
                "under a stub tensorflow module), NumPy/OpenBLAS")
    return "NumPy/OpenBLAS oracle port of cosmopower_NN.ten_to_predictions_np (oracle/_ref not staged on this box)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: 65536 rows per model per step at the driver's 20 + 5 steps is ~3 minutes of host time
    sample = CPU_SAMPLE_ROWS
    while sample > 8192 and (args.steps + args.warmup) * sample > 26 * CPU_SAMPLE_ROWS:
        sample //= 2
    v, s_per_step, cores, kind = cpu_path(sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cmb_TT_NN + cmb_EE_NN, ten_to_predictions, %d-row batch per model per GPU" % (1 << 20),
                   "models": "synthetic stand-ins (documented architecture 6-512x4-2507; pickles stripped upstream)",
                   "timed_on": "a bounded sample of that batch: %d rows per model per step" % sample, "sample_rows_per_model": sample},
        "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": cores, "kind": kind,
                         "sample": "%d rows per model per step, float64 inputs, %s" % (sample, cpu_kind_text(kind))},
        "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def d2h_ceiling(torch, dev, dist, world):
    """Aggregate pinned device->host copy rate of the host with all ranks copying at once (GB/s, whole job): the ceiling
    of the host-buffer `e2e` figure, which moves 10 KB per spectrum over PCIe."""
    n = 1 << 28                                                    # 1 GiB of float32 per copy
    src = torch.empty(n, dtype=torch.float32, device=dev)
    dst = torch.empty(n, dtype=torch.float32, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    reps = 4
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    del src, dst
    return world * reps * n * 4 / float(dt.item()) / 1e9


def per_model_section(torch, cp, dev, local, peaks, batch):
    """Every BASELINE config's emulator at `batch` rows on this GPU: launch time (CUDA events, 5 launches after 3
    warm-ups), roofline fractions, DRAM bytes of the committed ncu capture, parity of a 4096-row sample."""
    from cosmopower_b200 import priors, standins
    from oracle import oracle_np
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import parity
    traffic = load_traffic()
    peak = peaks["bf16_tflops_sustained"]
    out = {}
    buf = torch.empty((batch, 2512), dtype=torch.float32, device=dev)
    for i, name in enumerate(ALL_MODELS):
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=local, precision="f16x3")
        eng = m.engine(local)
        x64 = priors.sample_prior(m.parameters, batch, seed=20260401 + i)
        x = torch.from_numpy(x64.astype(np.float32)).to(dev)
        pitch = (m.n_modes + 7) // 8 * 8
        view = buf.view(-1)[:batch * pitch].view(batch, pitch)[:, :m.n_modes]
        ten_to = name != "cmb_TE_PCAplusNN"                         # TE is a linear-spectrum model (predictions_np)
        for _ in range(3):
            eng.forward_torch(x, out=view, ten_to=ten_to)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        evs[0].record()
        for j in range(5):
            eng.forward_torch(x, out=view, ten_to=ten_to)
            evs[j + 1].record()
        torch.cuda.synchronize()
        ms = [evs[j].elapsed_time(evs[j + 1]) for j in range(5)]
        avg = float(np.mean(ms)) * 1e-3
        # parity: the log10 / linear prediction of 4096 sampled rows against the float64 oracle (outside the timed region)
        eng.forward_torch(x, out=view, ten_to=False)
        rows = np.unique(np.concatenate([np.random.default_rng(7).choice(batch, 4096, replace=False), [0, batch - 1]]))
        got = view[torch.from_numpy(rows).to(dev)].cpu().numpy()
        attrs = standins.load_attrs(kind, path)
        ref = oracle_np.forward_chunked(attrs, x64[rows])
        err = parity.model_error(name, got, ref)
        flop = FLOP_PER_SPECTRUM[name] * batch
        tf = flop / avg / 1e12
        hbm = BYTES_PER_SPECTRUM[name] * batch / avg / 1e9
        tr = traffic.get(name)
        out[name] = {
            "rows": batch, "launch_ms_avg": avg * 1e3, "launch_ms_min": float(min(ms)), "spectra_per_s": batch / avg,
            "algorithmic_tflops": tf, "frac": tf / peak, "executed_frac": tf * PASSES / peak,
            "hbm_gbs": hbm, "hbm_frac": hbm / peaks["hbm_gbs"],
            "dram_bytes": tr[0] if tr else None, "dram_bytes_source": tr[1] if tr else None,
            "algorithmic_bytes": BYTES_PER_SPECTRUM[name] * batch,
            "parity_err": err, "parity_tol": parity.model_tol(name),
            "parity_metric": "max |y - y_ref| / row peak (linear spectrum)" if name in parity.LINEAR_MODELS else "max |dlog10| / max(|log10 ref|, 1)",
            "parity_rows": int(rows.size), "call": "ten_to_predictions" if ten_to else "predictions",
            "weights": "synthetic stand-in" if name in STANDINS else "reference pickle",
        }
        if name in parity.LINEAR_MODELS:
            out[name]["parity_err_strict"] = parity.linear_error_strict(got, ref, attrs["features_std_"])
        m.close()
        del x
    del buf
    return out


def likelihood_section(torch, cp, dev, local, rows=1 << 18):
    """SURVEY 8f.1 (the in-repo consumer of the CMB emulators): tf_planck2018_lite.get_loglkl on `rows` in-prior parameter
    sets - three emulators + Planck-lite tail on the device; parity of a 512-row sample against the oracle pipeline
    (oracle emulators + oracle likelihood, float64) computed outside the timed region."""
    from cosmopower_b200 import priors, standins
    from cosmopower_b200.likelihoods import tf_planck2018_lite
    from oracle import oracle_np, planck_lite_np
    models, attrs = {}, {}
    for key, name in (("tt", "cmb_TT_NN"), ("te", "cmb_TE_PCAplusNN"), ("ee", "cmb_EE_NN")):
        kind, path = standins.model_path(name)
        models[key] = getattr(cp, kind)(restore=True, restore_filename=path, device=local)
        attrs[key] = standins.load_attrs(kind, path)
    params = list(models["tt"].parameters)
    like = tf_planck2018_lite(parameters=None, tt_emu_model=models["tt"], te_emu_model=models["te"], ee_emu_model=models["ee"])
    x = priors.sample_prior(params, rows, seed=20260501, dtype=np.float32)
    cal = (1 + 0.0025 * np.random.default_rng(5).standard_normal(rows)).astype(np.float32)
    full_np = np.concatenate([x, cal[:, None]], axis=1)
    full = torch.from_numpy(full_np).to(dev)
    for _ in range(3):
        ll = like.get_loglkl(full)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ll = like.get_loglkl(full)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    n_par = 512
    x64 = full_np[:n_par, :len(params)].astype(np.float64)
    with np.errstate(over="ignore"):
        ref = planck_lite_np.PlanckLiteOracle().loglike_from_spectra(
            10.0 ** oracle_np.forward_pass(attrs["tt"], x64), oracle_np.forward_pass(attrs["te"], x64),
            10.0 ** oracle_np.forward_pass(attrs["ee"], x64), full_np[:n_par, -1].astype(np.float64))
    got = ll[:n_par].cpu().numpy().reshape(-1)
    rel = float(np.max(np.abs(got - ref.reshape(-1)) / np.abs(ref.reshape(-1))))
    med = float(np.median(ms))
    out = {"rows": rows, "get_loglkl_ms": med, "get_loglkl_ms_min": float(min(ms)), "loglkl_per_s": rows / med * 1e3,
           "call": "tf_planck2018_lite.get_loglkl (TT + TE + EE emulators and the plik-lite tail on the device)",
           "parity_rel_err": rel, "parity_tol": 1e-3, "parity_rows": n_par,
           "parity_metric": "max |ll - ll_ref| / |ll_ref| against oracle emulators + oracle likelihood (float64)",
           "weights": "synthetic stand-ins (TT / TE / EE pickles stripped upstream)"}
    like.close()
    for m in models.values():
        m.close()
    return out


def sweep_section(torch, cp, local, world, rank, rows_per_rank=8 << 20):
    """BASELINE configs[4] scaled to a bench-sized run: all six emulators streamed over `rows_per_rank` parameter sets
    per rank in 1,048,576-row chunks (parameters drawn on the device, one reused output buffer; the 100 M-set run is
    tools/run_sweep.py).  Returns this rank's device seconds per model."""
    from cosmopower_b200 import standins, sweep
    models = {}
    for name in ALL_MODELS:
        kind, path = standins.model_path(name)
        models[name] = getattr(cp, kind)(restore=True, restore_filename=path, device=local)
    res = sweep.sweep(models, rows_per_rank * world, chunk_rows=1 << 20, seed=3, device=local, ten_to=False, checksum=False,
                      world_size=world, rank=rank)
    for m in models.values():
        m.close()
    return {k: v["seconds"] for k, v in res.items()}, rows_per_rank


def run_ours(args):
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import _engine, affinity, priors, standins

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # host threads of this rank (and the pinned buffers they first touch) next to the GPU they feed
    pinned_cpus = affinity.pin_to_device(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    batch = args.batch
    models, engines, xs = [], [], []
    for i, name in enumerate(MODELS):
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=local, precision="f16x3")
        models.append(m)
        engines.append(m.engine(local))
        x = priors.sample_prior(m.parameters, batch, seed=20260101 + i + 1000 * rank, dtype=np.float32)
        xs.append(torch.from_numpy(x).to(dev))
    # 10.5 GB, reused by both models; row pitch 2512 floats (32-byte aligned rows), (batch, 2507) view
    out = torch.empty((batch, 2512), dtype=torch.float32, device=dev)[:, :2507]

    def step():
        for e, x in zip(engines, xs):
            e.forward_torch(x, out=out, ten_to=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = sum(e.launch_count() for e in engines)
    # per-launch events (dominant kernel = the fused MLP launch; both launches of a step are that kernel)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(MODELS) + 1)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    for s in range(args.steps):
        evs[s][0].record()
        for j, (e, x) in enumerate(zip(engines, xs)):
            e.forward_torch(x, out=out, ten_to=True)
            evs[s][j + 1].record()
    barrier()
    t_wall1 = time.time()
    launches = sum(e.launch_count() for e in engines) - launches0
    total_ms = evs[0][0].elapsed_time(evs[-1][-1])
    launch_ms = [evs[s][j].elapsed_time(evs[s][j + 1]) for s in range(args.steps) for j in range(len(MODELS))]
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_ms = max_over_ranks(total_ms)
    spectra_per_step = len(MODELS) * batch * world
    value = spectra_per_step * args.steps / (total_ms * 1e-3)

    # ---- strong scaling: ONE batch of `batch` rows per model split over the ranks (contiguous row shards)
    strong = None
    if world > 1:
        from cosmopower_b200.sharding import shard_bounds
        lo, hi = shard_bounds(batch, world, rank)
        k = max(3, min(args.steps, 10))

        def strong_steps(n_rows, reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for e, x in zip(engines, xs):
                    e.forward_torch(x[:n_rows], out=out[:n_rows], ten_to=True)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        strong_steps(hi - lo, 2)
        barrier()
        t_split = max_over_ranks(strong_steps(hi - lo, k))
        barrier()
        t_one = strong_steps(batch, k) if rank == 0 else 0.0      # the same batch on one GPU of this box, others idle
        barrier()
        if rank == 0:
            v_split = len(MODELS) * batch * k / (t_split * 1e-3)
            v_one = len(MODELS) * batch * k / (t_one * 1e-3)
            rows = hi - lo
            waves = rows / 512.0 / 74.0
            strong = {"value": v_split, "unit": "spectra/s", "batch_rows_total": batch, "rows_per_rank": rows, "steps": k,
                      "one_gpu_value": v_one, "efficiency": v_split / (world * v_one),
                      "note": "%d rows per rank = %.2f waves of 74 clusters x 512 rows: the last, partly filled wave "
                              "bounds the efficiency at %.3f" % (rows, waves, waves / np.ceil(waves))}

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    e2e_rows = args.e2e_batch
    try:   # pinned output = 10 KB per row: keep well inside the host's free memory
        avail = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0] * 1024
        while e2e_rows > 65536 and e2e_rows * 2507 * 4 * world * 3 > avail:
            e2e_rows //= 2
    except Exception:
        pass
    pin_in = [_engine.PinnedArray((e2e_rows, 6)) for _ in MODELS]
    pin_out = _engine.PinnedArray((e2e_rows, 2507))
    for p, x in zip(pin_in, xs):
        p.array[:] = x[:e2e_rows].cpu().numpy()
    del out
    torch.cuda.empty_cache()

    def e2e_step():
        for e, p in zip(engines, pin_in):
            e.forward_host(p.array, out=pin_out.array, ten_to=True)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = len(MODELS) * e2e_rows * world * e2e_steps / e2e_s
    for p in pin_in:
        p.free()
    pin_out.free()
    ceiling = None
    try:
        ceiling = d2h_ceiling(torch, dev, dist, world)
    except Exception as err:                                      # keep the JSON line whatever happens here
        sys.stderr.write("d2h ceiling probe failed: %r\n" % (err,))

    per_model = None
    if rank == 0 and not args.no_per_model:
        try:
            per_model = per_model_section(torch, cp, dev, local, load_peaks(), batch)
        except Exception as err:
            per_model = {"error": repr(err)}
    barrier()

    likelihood = None
    if rank == 0 and not args.no_per_model:
        try:
            likelihood = likelihood_section(torch, cp, dev, local)
        except Exception as err:
            likelihood = {"error": repr(err)}
    barrier()
    sweep_line = None
    if not args.no_per_model:
        # every rank streams its own shard of chunks (no collective on the path); the job's time is the slowest rank's.
        # A failure on one rank must not leave the others waiting in a reduction: the outcome is agreed on first.
        secs, rows_rank, sweep_err = None, 0, None
        try:
            secs, rows_rank = sweep_section(torch, cp, local, world, rank)
        except Exception as err:
            sweep_err = repr(err)
        all_ok = max_over_ranks(0.0 if sweep_err is None else 1.0) == 0.0
        if all_ok:
            tot = {k: max_over_ranks(secs[k]) for k in ALL_MODELS}
            total_s = max_over_ranks(sum(secs.values()))
            sweep_line = {"workload": "BASELINE configs[4] at bench size: six emulators x %d parameter sets per rank, 1,048,576-row chunks, "
                                      "parameters drawn on the device, predictions (log / linear spectra)" % rows_rank,
                          "rows_per_model": rows_rank * world, "seconds_per_model": tot, "seconds_total": total_s,
                          "spectra_per_s": len(ALL_MODELS) * rows_rank * world / total_s,
                          "extrapolated_100M_sets_s": total_s * 1e8 / (rows_rank * world)}
        else:
            sweep_line = {"error": sweep_err or "another rank failed"}
    barrier()

    if rank == 0:
        peaks = load_peaks()
        avg_launch_s = float(np.mean(launch_ms)) * 1e-3
        flop_launch = FLOP_PER_SPECTRUM["cmb_TT_NN"] * batch
        achieved = flop_launch / avg_launch_s / 1e12
        peak = peaks["bf16_tflops_sustained"]
        tr = load_traffic().get("cmb_TT_NN")
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": tr[0] if tr else None,
            "kernel": "cpb::fused_mlp_umma_kernel", "launch_ms_avg": avg_launch_s * 1e3,
            "passes": PASSES, "executed_tflops": achieved * PASSES, "executed_frac": achieved * PASSES / peak,
            "peak_kind": "cuBLAS bf16 sustained, %s (burst %.1f)" % (peaks["source"], peaks["bf16_tflops"]),
            "hbm_achieved_gbs": BYTES_PER_SPECTRUM["cmb_TT_NN"] * batch / avg_launch_s / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
            "algorithmic_bytes_per_launch": BYTES_PER_SPECTRUM["cmb_TT_NN"] * batch,
            "note": "algorithmic FLOP = 4,146,176 per spectrum (1 pass); the fp16x3 split executes 3 tensor passes, so frac <= 1/3 by construction and executed_frac is the tensor-pipe utilisation",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, _, cores, kind = cpu_path(CPU_SAMPLE_ROWS, 1, 1)
            v32, _, _, _ = cpu_path(8192, 2, 1, dtype=np.float32)
            cpu = {"value": v, "unit": "spectra/s", "cores": cores, "kind": kind,
                   "sample": "%d rows per model x 1 step after 1 warm-up, float64 inputs, %s" % (CPU_SAMPLE_ROWS, cpu_kind_text(kind)),
                   "value_float32_inputs": v32}
        d2h_step = len(MODELS) * e2e_rows * 2507 * 4
        e2e_gbs = e2e_value * 2507 * 4 / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16x3 split (fp16 hi+lo operands, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "cmb_TT_NN + cmb_EE_NN, ten_to_predictions, %d-row batch per model per GPU" % batch,
                       "models": "synthetic stand-ins (documented architecture 6-512x4-2507; pickles stripped upstream)",
                       "l2": "no explicit flush: each launch streams %.1f GB of output, far larger than the 126 MB L2" % (batch * 2507 * 4 / 1e9),
                       "parallelism": "rows sharded, %d rank(s), no collective" % world},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "spectra/s", "h2d_bytes_per_step": len(MODELS) * e2e_rows * 6 * 4,
                    "d2h_bytes_per_step": d2h_step, "rows_per_model": e2e_rows,
                    "api": "cpb200_forward_host via Engine.forward_host, pinned host buffers",
                    "d2h_gbs": e2e_gbs, "host_ceiling_gbs": ceiling,
                    "frac_of_host_ceiling": (e2e_gbs / ceiling) if ceiling else None,
                    "host_ceiling": "aggregate pinned D2H cudaMemcpyAsync rate, all %d rank(s) copying 1 GiB at once" % world,
                    "cpu_affinity": ("rank 0 pinned to the %d CPUs next to its GPU" % len(pinned_cpus)) if pinned_cpus else "not pinned"},
            "roofline": roofline,
        }
        if strong is not None:
            line["strong"] = strong
        if per_model is not None:
            line["per_model"] = per_model
        if likelihood is not None:
            line["likelihood"] = likelihood
        if sweep_line is not None:
            line["sweep"] = sweep_line
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--e2e-batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-model", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, for one) is sent to
    # stderr by pointing file descriptor 1 at stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit
    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
