"""Tell-tale opcode counts per kernel of the built library (cuobjdump -sass), for profiles/<tag>_sass_grep.txt.
usage: python tools/sass_grep.py [cosmopower_b200/csrc/libcpb200.so] > profiles/r2h_sass_grep.txt"""
import collections
import re
import subprocess
import sys

KEEP = re.compile(r"^(UTC|LDTM|STTM|UBLKCP|SYNCS|MUFU|FENCE|MEMBAR|CCTL|F2FP|FHADD|FFMA2|FMUL2|FADD2|STG|LDG\.E\.(128|256|ENL2)|HMMA|IMMA|UTMA)")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "cosmopower_b200/csrc/libcpb200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    print("cuobjdump -sass %s, tell-tale opcodes per kernel" % lib)
    print("(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU = ex2/rcp,")
    print(" FFMA2 / FMUL2 / FADD2 = packed fp32 pairs, FHADD = mixed fp16 + fp32 add of the operand split)\n")
    name, counts, total = None, collections.Counter(), 0

    def flush():
        if name is None:
            return
        print("%s  (%d instructions)" % (name, total))
        for op in sorted(counts):
            print("    %-44s %d" % (op, counts[op]))
        print()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            total += 1
            op = m.group(1)
            if KEEP.match(op):
                counts[op] += 1
    flush()


if __name__ == "__main__":
    main()
