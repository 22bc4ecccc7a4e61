"""Host logic of the fused kernel's static schedule (csrc/cpb200_api.cu build_jobs / assign_slots), checked without a GPU
through cpb200_debug_job_table: for the shipped architectures and a set of odd ones, the job table must

  * hold every job (contraction g, tile t, chunk c) exactly once, its parts in stage order and contiguous,
  * never hand a TMEM slot to a job while its previous owner has parts outstanding (the MMA issuer would deadlock),
  * keep every producer's last part ahead of its consumer's first part, and the waits / arrivals on each
    "operand ready" barrier strictly alternating (mbarrier phases),
  * never let an epilogue overwrite an activation buffer that a later-issued part still reads.
"""
import pytest

from cosmopower_b200 import _engine

ACT, AFFINE, OUT, QUAD = 0, 1, 2, 3
BUF_A0, BUF_ROT, BUF_PAR = 0x10, 0x20, 0x40

NETWORKS = {
    "cmb_TT_NN": ([6, 512, 512, 512, 512], [512, 512, 512, 512, 2507], [ACT] * 4 + [OUT]),
    "cmb_TE_PCAplusNN": ([6, 512, 512, 512, 512, 512], [512, 512, 512, 512, 512, 2507], [ACT] * 4 + [AFFINE, OUT]),
    "cmb_TE folded": ([6, 512, 512, 512, 512, 512], [512, 512, 512, 512, 512, 199], [ACT] * 4 + [AFFINE, OUT]),
    "cmb_PP_PCAplusNN": ([6, 512, 512, 512, 512, 64], [512, 512, 512, 512, 64, 2507], [ACT] * 4 + [AFFINE, OUT]),
    "PKLIN_NN": ([6, 512, 512, 512], [512, 512, 512, 420], [ACT] * 3 + [OUT]),
    "PKNLBOOST_NN": ([8, 512, 512, 512], [512, 512, 512, 420], [ACT] * 3 + [OUT]),
    "100-37-333": ([7, 100, 37], [100, 37, 333], [ACT, ACT, OUT]),
    "64-40": ([3, 64], [64, 40], [ACT, OUT]),
    "single": ([5], [70], [OUT]),
    "six hidden": ([6] + [512] * 6, [512] * 6 + [2507], [ACT] * 6 + [OUT]),
    "256-256-1000": ([40, 256, 256], [256, 256, 1000], [ACT, ACT, OUT]),
    "32-8": ([1, 32], [32, 8], [ACT, OUT]),
    "300-4001": ([33, 300], [300, 4001], [ACT, OUT]),
    "narrow tail 512-48-900": ([6, 512, 512, 48], [512, 512, 48, 900], [ACT, ACT, ACT, OUT]),
    "pca 16": ([6, 256, 256, 16], [256, 256, 16, 700], [ACT, ACT, AFFINE, OUT]),
    # several emulators in one launch (cpb200_create_multi): 4th element = emulator index of every contraction
    "PKLIN + PKNLBOOST": ([6, 512, 512, 512, 8, 512, 512, 512], [512, 512, 512, 420, 512, 512, 512, 420],
                          [ACT, ACT, ACT, OUT] * 2, [0] * 4 + [1] * 4),
    "TT + TE + EE": ([6, 512, 512, 512, 512] + [6, 512, 512, 512, 512, 512] + [6, 512, 512, 512, 512],
                     [512, 512, 512, 512, 2507] + [512, 512, 512, 512, 512, 199] + [512, 512, 512, 512, 2507],
                     [ACT] * 4 + [OUT] + [ACT] * 4 + [AFFINE, OUT] + [ACT] * 4 + [OUT], [0] * 5 + [1] * 6 + [2] * 5),
    "small pair": ([3, 64, 5, 100, 37], [64, 40, 100, 37, 333], [ACT, OUT, ACT, ACT, OUT], [0, 0, 1, 1, 1]),
}


def _buffer(code, pair, info):
    if code & BUF_ROT:
        return ("rot", ((code & 0xf) + pair * info["rot_mul"]) % 3)
    if code & BUF_PAR:
        return ("par", 2 * (pair & 1) + (code & 1))
    if code & BUF_A0:
        return ("a0", code & 0xf)
    return None


@pytest.mark.parametrize("name", sorted(NETWORKS))
def test_job_table_invariants(built_library, name):
    K, N, kinds = NETWORKS[name][:3]
    models = NETWORKS[name][3] if len(NETWORKS[name]) > 3 else [0] * len(K)
    entries, info = _engine.job_table(K, N, kinds, models=models)
    G = len(K)
    gem = info["gemms"]
    layer0 = [g == 0 or models[g] != models[g - 1] for g in range(G)]
    lastg = [g == G - 1 or models[g + 1] != models[g] for g in range(G)]
    assert len(entries) <= 320
    # ---- every job once, parts contiguous and in order, one slot assignment per job
    jobs = {}
    for pos, e in enumerate(entries):
        key = (e["g"], e["t"], e["c"])
        j = jobs.setdefault(key, dict(parts=[], slot=e["slot"], two=e["two"], nxt=e["nxt"]))
        assert (j["slot"], j["two"], j["nxt"]) == (e["slot"], e["two"], e["nxt"]), (name, key)
        j["parts"].append((pos, e))
    want = {(g, t, c) for g in range(G) for t in range(2) for c in range(gem[g][2])}
    assert set(jobs) == want, (name, sorted(want - set(jobs)), sorted(set(jobs) - want))
    for key, j in jobs.items():
        k_stages, nc, _ = gem[key[0]]
        s = 0
        for i, (pos, e) in enumerate(j["parts"]):
            assert e["s0"] == s and e["ns"] >= 1, (name, key)
            assert e["first"] == (i == 0) and e["last"] == (i == len(j["parts"]) - 1), (name, key)
            s += e["ns"]
        assert s == k_stages, (name, key, s, k_stages)
        assert j["two"] == (nc > 128) and j["slot"] + j["two"] <= 3 and (not j["two"] or j["slot"] % 2 == 0), (name, key)
    # ---- TMEM slots: walk the table twice (it repeats every iteration)
    open_by = [None] * 4
    for rep in range(2):
        for pos, e in enumerate(entries):
            key = (e["g"], e["t"], e["c"])
            if e["first"]:
                for k in range(e["two"] + 1):
                    assert open_by[e["slot"] + k] is None, "%s: entry %d %r takes slot %d still owned by %r" % (name, pos, key, e["slot"] + k, open_by[e["slot"] + k])
                    open_by[e["slot"] + k] = key
            if e["last"]:
                for k in range(e["two"] + 1):
                    assert open_by[e["slot"] + k] == key
                    open_by[e["slot"] + k] = None
    # ---- global time line over several pairs: (iteration, position); a `nxt` entry of pair p runs in iteration p, else p + 1
    n_pairs = 6
    events = []                     # (time, kind, ...) with time = iteration * len(entries) + pos
    L = len(entries)
    for pair in range(n_pairs):
        for pos, e in enumerate(entries):
            it = pair if e["nxt"] else pair + 1
            events.append((it * L + pos, pair, pos, e))
    events.sort(key=lambda x: x[0])
    # producers before consumers; barrier alternation
    last_part_time = {}
    first_part_time = {}
    for tm, pair, pos, e in events:
        key = (pair, e["g"], e["t"], e["c"])
        if e["first"]:
            first_part_time[key] = tm
        if e["last"]:
            last_part_time[key] = tm
    for (pair, g, t, c), tm in first_part_time.items():
        if not layer0[g]:
            for pc in range(gem[g - 1][2]):
                assert last_part_time[(pair, g - 1, t, pc)] < tm, (name, "consumer before producer", pair, g, t, c)
    # barrier ts = 2 * (pair & 1) + t: arrivals by the last chunk's epilogue of a producing contraction, waits by the first
    # part of chunk 0 of the consuming one
    for ts in range(4):
        seq = []
        for tm, pair, pos, e in events:
            if 2 * (pair & 1) + e["t"] != ts:
                continue
            g = e["g"]
            if not layer0[g] and e["first"] and e["c"] == 0:
                seq.append((tm - 0.25, "wait", pair, g))             # the loader waits before it loads the part ...
            if not lastg[g] and e["last"] and e["c"] == gem[g][2] - 1:
                seq.append((tm + 0.25, "arrive", pair, g))           # ... the epilogue arrives after the job's last part
        seq.sort()
        kinds_seq = [k for _, k, _, _ in seq]
        for i, k in enumerate(kinds_seq):
            assert k == ("arrive" if i % 2 == 0 else "wait"), (name, ts, seq[max(0, i - 2):i + 2])
        for i in range(1, len(seq) - 1, 2):         # wait i must precede the next arrival's job entirely finishing
            assert seq[i][0] < seq[i + 1][0]
    # ---- activation buffers: a write (after the writer's last part) must follow every read of the old contents
    reads = {}      # buffer -> list of (time of the reading part, contents id)
    contents = {}   # buffer -> contents id currently held
    for tm, pair, pos, e in events:
        g = e["g"]
        rb = _buffer(e["in_code"], pair, info)
        if layer0[g]:
            assert rb == ("a0", 2 * models[g] + e["t"]), (name, rb)
        if not layer0[g] and rb is not None:
            assert contents.get(rb) == (pair, g - 1, e["t"]), "%s: pair %d g%d t%d c%d reads %r holding %r" % (
                name, pair, g, e["t"], e["c"], rb, contents.get(rb))
        if e["last"] and not lastg[g] and e["c"] == 0:
            # chunk 0 is the first chunk of (g, t) to complete: from its epilogue on, the output buffer is being overwritten
            wb = _buffer(e["out_code"], pair, info)
            assert wb is not None
            old = contents.get(wb)
            if old is not None:
                # every part reading the old contents must have been issued before this job's last part
                og = old[1] + 1
                for oc in range(gem[og][2]):
                    assert last_part_time[(old[0], og, old[2], oc)] < tm, "%s: %r overwritten by pair %d g%d t%d before its reader g%d c%d was issued" % (
                        name, wb, pair, g, e["t"], og, oc)
            contents[wb] = (pair, g, e["t"])


def test_schedules_chosen_for_the_shipped_models(built_library):
    sched = {name: _engine.job_table(*NETWORKS[name][:3])[1]["schedule"] for name in NETWORKS if len(NETWORKS[name]) == 3}
    assert sched["cmb_TT_NN"] == 1 and sched["PKLIN_NN"] == 1 and sched["cmb_TE_PCAplusNN"] == 1      # layer 0 merged into the output jobs
    assert sched["cmb_PP_PCAplusNN"] == 1                  # (the overlapped output phase measured no faster: profiling builds only)
    assert sched["single"] == 0
