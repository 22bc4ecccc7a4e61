"""Extract the judged subset of an `ncu --set full` report into profiles/<tag>_ncu_full_summary.csv
and refresh profiles/traffic.json.

usage: python tools/ncu_summary.py --models r2k      (after tools/ncu_models.sh: gpurun_out/prof_<model>.ncu-rep for the six emulators
                                                      -> profiles/r2k_<model>_ncu_full_summary.csv + profiles/traffic.json, the per-model
                                                      DRAM bytes bench.py reports)
       python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep r1e
       python tools/ncu_summary.py gpurun_out/prof_full_pp.ncu-rep r1j_cmb_PP   (any tag with a model suffix: summary only,
                                                                                 profiles/traffic.json is left alone)
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r"pipe_tensor|dram__bytes_(read|write)\.sum$|dram__throughput|lts__t_(bytes|sectors)\.sum$|"
    r"lts__t_sector_hit_rate|lts__throughput|l1tex__throughput|gpu__time_duration\.sum|"
    r"sm__throughput|sm__cycles_elapsed\.(avg|max)$|sm__cycles_active\.avg$|sm__warps_active|"
    r"launch__(registers_per_thread|grid_size|block_size|cluster|shared_mem_per_block_dynamic|occupancy_limit)|"
    r"smsp__inst_executed\.sum$|sm__inst_executed_pipe_(xu|fma|alu|lsu|uniform)|smsp__issue_active|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
    r"smsp__average_warp.*issue_stalled.*_per_warp_active|sm__clock|gpc__cycles_elapsed\.max")


MODELS = ["cmb_TT_NN", "cmb_EE_NN", "cmb_TE_PCAplusNN", "cmb_PP_PCAplusNN", "PKLIN_NN", "PKNLBOOST_NN"]


def summarise(rep, path):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, launch = rows[0], rows[1], rows[2]
    out = [("metric", "unit", "value")]
    vals = {}
    for n, u, v in zip(names, units, launch):
        if KEEP.search(n):
            out.append((n, u, v))
        vals[n] = (u, v)
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows(out)

    def to_bytes(n):
        u, v = vals[n]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum"), vals


def all_models(tag, note):
    traffic = {"models": {}}
    for m in MODELS:
        path = f"profiles/{tag}_{m}_ncu_full_summary.csv"
        rd, wr, vals = summarise(f"gpurun_out/prof_{m}.ncu-rep", path)
        traffic["models"][m] = {"dram_bytes_read": rd, "dram_bytes_write": wr,
                                "launch": "1048576 rows, " + ("predictions" if m == "cmb_TE_PCAplusNN" else "ten_to"),
                                "source": f"{path} (ncu --set full --clock-control none, one launch, {note})"}
        print(m, "%.2f + %.2f GB" % (rd / 1e9, wr / 1e9), vals.get("gpu__time_duration.sum"),
              vals.get("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active") or
              [v for k, v in vals.items() if "pipe_tensor" in k and "pct_of_peak_sustained_active" in k][:1])
    with open("profiles/traffic.json", "w") as f:
        json.dump(traffic, f, indent=1)


def main():
    if sys.argv[1] == "--models":
        all_models(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "round-2 final kernel")
        return
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    launch = rows[2]
    kname = launch[names.index("Kernel Name")] if "Kernel Name" in names else "?"
    out = [("metric", "unit", "value")]
    vals = {}
    for n, u, v in zip(names, units, launch):
        if KEEP.search(n):
            out.append((n, u, v))
        vals[n] = (u, v)
    path = f"profiles/{tag}_ncu_full_summary.csv"
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows(out)

    if "_" in tag:          # a capture of another model: bench.py's roofline.traffic stays the cmb_TT figure
        print(path, len(out) - 1, "metrics")
        return

    def to_bytes(n):
        u, v = vals[n]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v.replace(",", "")) * scale

    traffic = {"kernel": kname, "launch": "cmb_TT_NN stand-in, 1048576 rows, ten_to",
               "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
               "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
               "source": f"{path} (ncu --set full, one launch)"}
    with open("profiles/traffic.json", "w") as f:
        json.dump(traffic, f)
    print(path, len(out) - 1, "metrics;", traffic)


if __name__ == "__main__":
    main()
