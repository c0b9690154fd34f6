#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / bulk-copy path (cuobjdump -sass of the shipped .so).
usage: python scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "articulated_pose_b200/libancsh_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "F2FP", "CREDUX", "REDUX", "VOTE", "DFMA", "MUFU", "FFMA", "LDG", "STS", "LDS"]
kernels, cur = [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = {"name": m.group(1), "total": 0, **{c: 0 for c in cols}}
        kernels.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                cur[c] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    except Exception:
        return n
print("# SASS summary of %s (cuobjdump -sass, sm_100a)" % lib)
print("# tcgen05 / TMEM / bulk-copy mnemonics per kernel: UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
print("# UBLKCP = cp.async.bulk (global -> shared, mbarrier complete_tx), SYNCS = mbarrier ops, F2FP = packed f32 -> f16x2 conversions,")
print("# CREDUX / REDUX = redux.sync, VOTE = ballot")
print("%-92s" % "kernel" + "".join("%8s" % c for c in cols) + "%8s" % "total")
for k in kernels:
    print("%-92s" % demangle(k["name"])[:91] + "".join("%8d" % k[c] for c in cols) + "%8d" % k["total"])
print("# totals: " + ", ".join("%s %d" % (c, sum(k[c] for k in kernels)) for c in cols[:5]))
