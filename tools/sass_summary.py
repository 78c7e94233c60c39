"""SASS opcode summary per kernel of libop3_b200.so (cuobjdump -sass): the tensor-core / TMA / bulk-copy mnemonics that prove
which hardware path a kernel uses (B200_PROFILING.md).  usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "op3_b200", "libop3_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "MMA.",
        "SYNCS", "UCGABAR", "SHFL", "MUFU", "FFMA", "LDS", "STS", "LDG", "STG", "REDG", "BAR.SYNC", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    name, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("op3::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for k in KEYS:
                if op.startswith(k) or (k == "MMA." and re.match(r"^[A-Z]*MMA\.", op) and not op.startswith("UTC")):
                    counts[name][k] += 1
                    break
    print("# SASS opcode counts per kernel of libop3_b200.so (static instruction counts; `cuobjdump -sass`, sm_100a)")
    print("# UTCHMMA = tcgen05.mma kind::f16/tf32, UTCCP = tcgen05.cp, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D),")
    print("# UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA tensor copies), LDGSTS = cp.async, HMMA = mma.sync")
    for n in sorted(counts, key=lambda k: -total[k]):
        c = counts[n]
        if total[n] < 200 and not any(c[k] for k in KEYS[:14]):
            continue
        print("%-58s %6d instr  %s" % (n[:58], total[n], "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))


if __name__ == "__main__":
    main()
