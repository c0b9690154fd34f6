#!/usr/bin/env python
"""Top stall sites of a kernel from `ncu -i rep --page source --csv` (SASS view).
usage: python scripts/ncu_stalls.py src.csv [kernel-substring] [topN]"""
import csv, sys
path = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# the file holds one table per kernel: a "Kernel Name" row, a header row, then instruction rows
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        i = j
        if sub not in name: continue
        H = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[H["# Samples"]] or 0) for r in body)
        inst = sum(int(r[H["Instructions Executed"]] or 0) for r in body)
        print("== %s\n   samples %d, warp instructions executed %d" % (name[:110], tot, inst))
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[H[h]] or 0) for r in body) for h in reasons}
        print("   " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
        order = sorted(range(len(body)), key=lambda k: -int(body[k][H["# Samples"]] or 0))[:top]
        for k in sorted(order):
            r = body[k]
            s = int(r[H["# Samples"]] or 0)
            why = sorted(((int(r[H[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
            print("   %5d %5.1f%%  x%-8s %-58s %s" % (k, 100.0 * s / max(tot, 1), r[H["Instructions Executed"]], r[H["Source"]].strip()[:58],
                                                  " ".join("%s:%d" % (n, v) for v, n in why if v)))
    else:
        i += 1
