#!/usr/bin/env python
"""profiles/rNN_sass_summary.txt: which Blackwell-only instructions the shipped librqk_sm100a.so holds, per kernel
(`cuobjdump -sass`, no GPU needed).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "generative_ranking_recommender_b200", "librqk_sm100a.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "ATOMS", "ATOMG",
             "RED", "LDG.E.128", "STG.E.128", "HSET2", "HADD2", "FFMA", "REDUX", "MATCH"]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (("." in mn) and op.startswith(mn)):
                    per[cur][mn] += 1
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; counts of SASS instructions per kernel)\n")
        tot = collections.Counter()
        for k, c in per.items():
            tot.update(c)
        f.write("# whole library: " + ", ".join(f"{m} {tot[m]}" for m in MNEMONICS if tot[m]) + f", instructions {tot['_total']}\n")
        for k, c in per.items():
            f.write(f"{k}: instructions {c['_total']}; " + ", ".join(f"{m} {c[m]}" for m in MNEMONICS if c[m]) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt"))
