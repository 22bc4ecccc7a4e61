"""Discrete-event model of the fused kernel's pipeline (loader / MMA issuer / epilogue warps / TMEM slots / stage
ring) for a given job table, with the per-job costs measured on B200.  Used to compare job
tables before spending GPU time; the numbers it prints are estimates, the ncu captures are the evidence.

    python tools/schedule_sim.py            # prints cycles per tile pair for every model and schedule variant
"""
import sys
from collections import namedtuple

STAGES = 6
L2_LAT = 1400           # cycles from a bulk copy's issue to its landing
SM_FILL = 64.0          # B/clk a single SM's loader sustains (L2 -> shared memory)

# entry of a job table: one PART of a job (g, t, c) = stages [s0, s0 + ns) of its K loop
Entry = namedtuple("Entry", "g t c nxt s0 ns first last slot nslots nc kind k_stages n_chunks")


def mma_stage_cycles(nc):
    # 6 MMAs (2 k-steps x 3 split products) of M = 256 (pair), N = nc, K = 16: nc / 2 cycles each at full rate;
    # below N = 64 the A-operand read (4 KB per MMA at 128 B/clk) is the floor
    return 6 * max(nc / 2.0, 32.0)


def epi_cycles(kind, nc, tt_like_out=False):
    if kind == "act":
        return 8500.0 * nc / 256.0
    if kind == "affine":
        return 1500.0 + 10.0 * nc
    if kind == "out":
        return 5300.0 * nc / 256.0
    if kind == "quad":
        return 3000.0 * nc / 256.0
    raise ValueError(kind)


def simulate(table, n_iter=8, verbose=False):
    """table: list of Entry for one iteration (pair `it` for nxt entries, pair `it - 1` for the others).
    Returns steady-state cycles per iteration (= per tile pair)."""
    load_free = 0.0
    mma_free = 0.0
    epi_free = 0.0
    ring_free = [0.0] * STAGES          # time the stage slot was released by the MMA issuer
    ring_i = 0
    prev_land = 0.0
    slot_free = [0.0] * 4               # TMEM 128-column slots
    aready = {}                         # (pair, t, g) -> time the operand written by contraction g of that tile is complete
    acc_done = {}                       # job key -> time its MMAs completed
    iter_end = []
    mma_busy = 0.0
    for it in range(n_iter + 1):
        for e in table:
            if e.nxt and it >= n_iter:
                continue
            if not e.nxt and it < 1:
                continue
            pair = it if e.nxt else it - 1
            key = (pair, e.g, e.t, e.c)
            # ---- loader
            dep = 0.0
            if e.first and e.g > 0 and e.c == 0:
                dep = aready.get((pair, e.t, e.g - 1), 0.0)
            t = mma_free
            if e.first:
                for k in range(e.nslots):
                    if slot_free[e.slot + k] == float("inf"):
                        raise RuntimeError("deadlock: entry %r wants TMEM slot %d, which an unfinished job owns" % (e, e.slot + k))
                    t = max(t, slot_free[e.slot + k])
            c_stage = mma_stage_cycles(e.nc)
            for s in range(e.ns):
                # loader: stage slot free (its MMAs of STAGES stages ago are complete), operand ready
                t_issue = max(load_free, ring_free[ring_i], dep)
                bytes_ = 16384 + e.nc * 64
                land = max(t_issue + L2_LAT, prev_land + bytes_ / SM_FILL)
                prev_land = land
                load_free = t_issue + 20
                # MMA issuer
                t = max(t, land) + c_stage
                mma_busy += c_stage
                ring_free[ring_i] = t
                ring_i = (ring_i + 1) % STAGES
            mma_free = t
            if e.last:
                acc_done[key] = t
                # ---- epilogue (in table order of last parts)
                t0 = max(epi_free, t)
                cost = epi_cycles(e.kind, e.nc)
                t1 = t0 + cost
                epi_free = t1
                if e.nslots == 2:
                    slot_free[e.slot] = t0 + 0.5 * cost + 100
                    slot_free[e.slot + 1] = t1
                else:
                    slot_free[e.slot] = t1
                for k in range(e.nslots):
                    pass
                if e.kind in ("act", "affine") and e.c == e.n_chunks - 1:
                    aready[(pair, e.t, e.g)] = t1 + 200
            else:
                # slots stay owned by the job
                for k in range(e.nslots):
                    slot_free[e.slot + k] = float("inf")
            # a lazily released ring: entries longer than the ring need their own earlier stages back
            # (handled because ring_free is updated stage by stage above; the inf placeholder only matters inside one entry)
        iter_end.append(max(mma_free, epi_free))
    # steady state: average over the middle iterations
    per = (iter_end[-2] - iter_end[2]) / (len(iter_end) - 4)
    return per, mma_busy / (n_iter)


# ---------------------------------------------------------------------------------------------------
# table builders (mirror build_jobs in cpb200_api.cu)
# ---------------------------------------------------------------------------------------------------
def model_gemms(name):
    """[(K, N, kind)] per contraction."""
    if name in ("TT", "EE"):
        return [(32, 512, "act")] + [(512, 512, "act")] * 3 + [(512, 2507, "out")]
    if name == "TE":
        return [(32, 512, "act")] + [(512, 512, "act")] * 3 + [(512, 512, "affine"), (512, 2507, "out")]
    if name == "PP":
        return [(32, 512, "act")] + [(512, 512, "act")] * 3 + [(512, 64, "affine"), (64, 2507, "out")]
    if name in ("PKLIN", "PKBOOST"):
        return [(32, 512, "act")] + [(512, 512, "act")] * 2 + [(512, 420, "out")]
    raise ValueError(name)


def chunking(N, max_nc):
    n_chunks = -(-N // max_nc)
    nc = -(-(-(-N // n_chunks)) // 32) * 32
    return n_chunks, nc


def table_round1(name):
    """The round-1 order: every job one part, two 256-column accumulators in turn, layer 0 of the next pair merged into
    the output phase."""
    G = model_gemms(name)
    out = []
    head, tail, nxt0 = [], [], []
    for g, (K, N, kind) in enumerate(G):
        n_chunks, nc = chunking(N, 256)
        ks = -(-K // 32)
        for t in range(2):
            for c in range(n_chunks):
                e = dict(g=g, t=t, c=c, nxt=int(g == 0), s0=0, ns=ks, first=1, last=1, slot=0, nslots=2, nc=nc, kind=kind,
                         k_stages=ks, n_chunks=n_chunks)
                (nxt0 if g == 0 else tail if g == len(G) - 1 else head).append(e)
    jobs = list(head)
    k = 0
    for i, e in enumerate(tail):
        jobs.append(e)
        while k < len(nxt0) and (k + 1) * len(tail) <= (i + 1) * (len(nxt0) + 1):
            jobs.append(nxt0[k]); k += 1
    jobs += nxt0[k:]
    res = []
    for j, e in enumerate(jobs):
        e["slot"] = 2 * (j & 1)
        res.append(Entry(**e))
    return res


def split_job(e, cuts):
    """split a job dict into parts at the stage indices `cuts`"""
    parts, s = [], 0
    bounds = [c for c in cuts if 0 < c < e["ns"]] + [e["ns"]]
    for i, b in enumerate(bounds):
        p = dict(e)
        p.update(s0=s, ns=b - s, first=int(i == 0), last=int(i == len(bounds) - 1))
        parts.append(p)
        s = b
    return parts


def table_v2_merge_l0(name, l0_nc=128):
    """Big-K output phase (TT-like, P(k)): layer 0 of the next pair as 128-column jobs, each issued in the MIDDLE of an
    output job (between two of its K stages) into the TMEM slot the previous output job's epilogue has already drained."""
    G = model_gemms(name)
    head, tail, nxt0 = [], [], []
    for g, (K, N, kind) in enumerate(G):
        n_chunks, nc = chunking(N, l0_nc if g == 0 else 256)
        ks = -(-K // 32)
        for t in range(2):
            for c in range(n_chunks):
                e = dict(g=g, t=t, c=c, nxt=int(g == 0), s0=0, ns=ks, first=1, last=1, slot=0, nslots=2 if nc > 128 else 1, nc=nc,
                         kind=kind, k_stages=ks, n_chunks=n_chunks)
                (nxt0 if g == 0 else tail if g == len(G) - 1 else head).append(e)
    res = []
    pair_sel = 0            # 2-slot jobs alternate (0,1) / (2,3)
    for e in head:
        e["slot"] = 2 * pair_sel; pair_sel ^= 1
        res.append(e)
    # distribute the layer-0 jobs of tile t over the output jobs of tile t (any for t = 0 except the first job of the phase)
    per_t = {0: [e for e in nxt0 if e["t"] == 0], 1: [e for e in nxt0 if e["t"] == 1]}
    outs_t = {0: [e for e in tail if e["t"] == 0], 1: [e for e in tail if e["t"] == 1]}
    for t in range(2):
        outs, l0s = outs_t[t], per_t[t]
        n_o, n_l = len(outs), len(l0s)
        # l0 job k goes into output job floor((k + 0.5) * n_o / n_l) (several per job when n_l > n_o)
        where = {}
        for k in range(n_l):
            where.setdefault(min(n_o - 1, int((k + 0.5) * n_o / n_l)), []).append(l0s[k])
        for i, o in enumerate(outs):
            o["slot"] = 2 * pair_sel
            other = 2 * (pair_sel ^ 1)
            pair_sel ^= 1
            ins = where.get(i, [])
            if not ins:
                res.append(o)
                continue
            cuts = [int(round((j + 1) * o["ns"] / (len(ins) + 1))) for j in range(len(ins))]
            parts = split_job(o, cuts)
            for j, p in enumerate(parts):
                res.append(p)
                if j < len(ins):
                    l0 = ins[j]
                    l0["slot"] = other + (j & 1)
                    res.append(l0)
    return [Entry(**e) for e in res]


def table_v2_overlap_out(name, per_h=2):
    """Small-K output phase (cmb_PP): the output jobs of pair it-1 (128 columns, one TMEM slot each) are issued between
    the K stages of the hidden jobs of pair it; what does not fit there surrounds the layer-0 and coefficient jobs."""
    G = model_gemms(name)
    trunk, outs = [], []
    for g, (K, N, kind) in enumerate(G):
        last = g == len(G) - 1
        n_chunks, nc = chunking(N, 128 if (last or g == 0) else 256)
        ks = -(-K // 32)
        for t in range(2):
            for c in range(n_chunks):
                e = dict(g=g, t=t, c=c, nxt=int(not last), s0=0, ns=ks, first=1, last=1, slot=0, nslots=2 if nc > 128 else 1, nc=nc,
                         kind=kind, k_stages=ks, n_chunks=n_chunks)
                (outs if last else trunk).append(e)
    res = []
    oi = 0
    pair_sel = 0
    ring1 = 0               # single-slot jobs outside the hidden phase walk the four slots in turn
    n_h = sum(1 for e in trunk if e["nslots"] == 2)
    n_small = len(trunk) - n_h
    # outputs placed inside hidden jobs: per_h each; the rest is spread around the single-slot trunk jobs
    n_in_h = min(len(outs), per_h * n_h)
    n_rest = len(outs) - n_in_h
    rest_per_small = -(-n_rest // max(1, n_small))
    for e in trunk:
        if e["nslots"] == 1:
            e["slot"] = ring1 % 4; ring1 += 1
            res.append(e)
            for _ in range(rest_per_small):
                if oi < len(outs) and n_rest > 0:
                    o = outs[oi]; oi += 1; n_rest -= 1
                    o["slot"] = ring1 % 4; ring1 += 1
                    res.append(o)
            continue
        # hidden job in slots (0,1) or (2,3); the other pair's slots are drained by the previous hidden job's epilogue:
        # first slot after half of it, second at its end
        e["slot"] = 2 * pair_sel
        other = 2 * (pair_sel ^ 1)
        pair_sel ^= 1
        ring1 = e["slot"] + 2       # keep the single-slot ring consistent afterwards
        k = min(per_h, len(outs) - oi) if n_in_h > 0 else 0
        if k == 0:
            res.append(e)
            continue
        cuts = [int(round((j + 1) * e["ns"] / (k + 1))) for j in range(k)]
        if k == 2:
            cuts = [6, 12] if e["ns"] == 16 else cuts
        parts = split_job(e, cuts)
        for j, p in enumerate(parts):
            res.append(p)
            if j < k:
                o = outs[oi]; oi += 1; n_in_h -= 1
                o["slot"] = other + (j & 1)
                res.append(o)
    while oi < len(outs):
        o = outs[oi]; oi += 1
        o["slot"] = ring1 % 4; ring1 += 1
        res.append(o)
    return [Entry(**e) for e in res]


def table_from_library(name):
    """The table the C++ builder actually produces (cpb200_debug_job_table), as Entry tuples."""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cosmopower_b200 import _engine
    G = model_gemms(name)
    kinds = {"act": 0, "affine": 1, "out": 2}
    K = [6] + [k for k, _, _ in G[1:]]
    entries, info = _engine.job_table(K, [n for _, n, _ in G], [kinds[k] for _, _, k in G])
    out = []
    for e in entries:
        ks, nc, n_chunks = info["gemms"][e["g"]]
        out.append(Entry(g=e["g"], t=e["t"], c=e["c"], nxt=e["nxt"], s0=e["s0"], ns=e["ns"], first=e["first"], last=e["last"],
                         slot=e["slot"], nslots=e["two"] + 1, nc=nc, kind=G[e["g"]][2], k_stages=ks, n_chunks=n_chunks))
    return out, info


def main():
    for name in ("TT", "TE", "PP", "PKLIN"):
        p1, busy = simulate(table_round1(name))
        tab, info = table_from_library(name)
        p2, busy2 = simulate(tab)
        print("%-6s round-1 table: %7.0f cycles/pair (MMA %.0f = %.0f%%);  library table (schedule %d, %d entries): %7.0f (MMA %.0f%%)"
              % (name, p1, busy, 100 * busy / p1, info["schedule"], len(tab), p2, 100 * busy2 / p2))


if __name__ == "__main__":
    main()
