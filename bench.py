#!/usr/bin/env python
"""Benchmark of the CosmoPower batched emulator forward pass on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): cmb_TT_NN + cmb_EE_NN, ten_to_predictions on a 1,048,576-row
batch of in-prior synthetic LCDM parameters per model and per GPU.  One "step" = TT batch + EE
batch (2 kernel launches, 2,097,152 spectra per GPU).  TT/EE weights are SYNTHETIC STAND-INS with
the documented architecture (their pickles are stripped from the reference checkout).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = same metric through the host-buffer C-ABI call (pinned host in/out, H2D + D2H inside the
timed region); `roofline` = tensor-pipe roofline of the fused kernel; `cpu_baseline` = the oracle
(NumPy port of the reference's predictions path) on the host cores, bounded sample.
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
METRIC = "spectra/sec (batch 1M)"
FLOP_PER_SPECTRUM = {"cmb_TT_NN": 4146176, "cmb_EE_NN": 4146176}       # 2*sum(in*out), SURVEY 8(d)
BYTES_PER_SPECTRUM = 4 * (6 + 2507)
PASSES = 3                                                            # fp16x3 split


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def load_traffic():
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed ncu --set full capture."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


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


def cpu_path(sample_rows, steps, warmup, dtype=np.float64):
    """The reference's CPU implementation of the path (oracle port: NumPy, all BLAS threads).  float64 inputs are the
    reference's documented usage (dicts of lists); float32 arrays make its whole pass float32."""
    from cosmopower_b200 import priors, standins
    from oracle import oracle_np
    models = []
    for i, name in enumerate(MODELS):
        kind, path = standins.model_path(name)
        attrs = standins.load_attrs(kind, path)
        x = priors.sample_prior(attrs["parameters"], sample_rows, seed=20260101 + i)
        d = {k: x[:, j].astype(dtype) for j, k in enumerate(attrs["parameters"])}
        models.append((attrs, d))

    def step():
        for attrs, d in models:
            with np.errstate(over="ignore"):
                oracle_np.ten_to_predictions_np(attrs, d)

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
    return steps * len(MODELS) * sample_rows / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8192
    v, s_per_step, cores = cpu_path(sample, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cmb_TT_NN + cmb_EE_NN, ten_to_predictions, %d-row batch per model per GPU" % (1 << 20),
                   "models": "synthetic stand-ins (documented architecture 6-512x4-2507; pickles stripped upstream)",
                   "timed_on": "a bounded sample of that batch: %d rows per model per step" % sample, "sample_rows_per_model": sample},
        "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port",
                         "sample": "%d rows per model per step, float64 inputs, NumPy/OpenBLAS oracle port of "
                                   "cosmopower_NN.ten_to_predictions_np (the reference checkout is absent on the GPU box)" % sample},
        "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_ours(args):
    import torch
    import cosmopower_b200 as cp
    from cosmopower_b200 import _engine, priors, standins

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    batch = args.batch
    engines, xs = [], []
    for i, name in enumerate(MODELS):
        kind, path = standins.model_path(name)
        m = getattr(cp, kind)(restore=True, restore_filename=path, device=local, precision="f16x3")
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
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    spectra_per_step = len(MODELS) * batch * world
    value = spectra_per_step * args.steps / (total_ms * 1e-3)

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
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = len(MODELS) * e2e_rows * world * e2e_steps / float(te.item())

    if rank == 0:
        peaks = load_peaks()
        avg_launch_s = float(np.mean(launch_ms)) * 1e-3
        flop_launch = FLOP_PER_SPECTRUM["cmb_TT_NN"] * batch
        achieved = flop_launch / avg_launch_s / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": load_traffic(),
            "kernel": "cpb::fused_mlp_umma_kernel", "launch_ms_avg": avg_launch_s * 1e3,
            "passes": PASSES, "executed_tflops": achieved * PASSES, "executed_frac": achieved * PASSES / peak,
            "peak_kind": "cuBLAS bf16 sustained, %s (burst %.1f)" % (peaks["source"], peaks["bf16_tflops"]),
            "hbm_achieved_gbs": BYTES_PER_SPECTRUM * batch / avg_launch_s / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
            "algorithmic_bytes_per_launch": BYTES_PER_SPECTRUM * batch,
            "note": "algorithmic FLOP = 4,146,176 per spectrum (1 pass); the fp16x3 split executes 3 tensor passes, so frac <= 1/3 by construction and executed_frac is the tensor-pipe utilisation",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, _, cores = cpu_path(8192, 2, 1)
            v32, _, _ = cpu_path(8192, 2, 1, dtype=np.float32)
            cpu = {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port",
                   "sample": "8192 rows per model x 2 steps, float64 inputs, NumPy/OpenBLAS oracle port of ten_to_predictions_np",
                   "value_float32_inputs": v32}
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
                    "d2h_bytes_per_step": len(MODELS) * e2e_rows * 2507 * 4, "rows_per_model": e2e_rows,
                    "api": "cpb200_forward_host via Engine.forward_host, pinned host buffers"},
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    for p in pin_in:
        p.free()
    pin_out.free()
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
