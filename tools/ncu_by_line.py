"""Aggregate ncu warp-stall samples of the fused kernel by CUDA source line.
Usage: python tools/ncu_by_line.py <report.ncu-rep> [min_pct]   (needs ncu, cuobjdump, nvdisasm; no GPU)"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = "fused_mlp_umma_kernel"


def line_table():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "cosmopower_b200/csrc/libcpb200.so")],
                   cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    table, cur, active = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = KERNEL in ln
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    base = min(int(r[idx["Address"]], 16) for r in data)
    table = line_table()
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    tot = 0.0
    for r in data:
        n = float(r[idx["# Samples"]] or 0)
        tot += n
        key = table.get(int(r[idx["Address"]], 16) - base, (None, ""))[0]
        a = agg.setdefault(key, {"n": 0.0, "inst": 0.0})
        a["n"] += n
        a["inst"] += float(r[idx["Instructions Executed"]] or 0)
        for c in stall:
            a[c] = a.get(c, 0.0) + float(r[idx[c]] or 0)
    src = open(os.path.join(REPO, "cosmopower_b200/csrc/umma_kernel.cuh")).read().splitlines()
    print("total samples %.0f" % tot)
    for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
        if a["n"] < tot * min_pct / 100:
            continue
        st = sorted(((a.get(c, 0), c[6:]) for c in stall), reverse=True)[:3]
        text = src[key[1] - 1].strip()[:80] if key and key[0] == "umma_kernel.cuh" and key[1] <= len(src) else str(key)
        print("%-5s %8.0f %5.1f%% inst=%-10.0f %-80s %s" % (key[1] if key else "?", a["n"], 100 * a["n"] / tot, a["inst"], text,
                                                  " ".join("%s=%.0f" % (c, v) for v, c in st)))


if __name__ == "__main__":
    main()
