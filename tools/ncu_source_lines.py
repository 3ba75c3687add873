#!/usr/bin/env python
"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by source line: stall samples, instructions, lanes."""
import collections
import csv
import sys


def main(path, top=36):
    rows = list(csv.reader(open(path)))
    cur, hdr = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0, "", 0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            iS, iI, iT = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            iG = hdr.index("L2 Theoretical Sectors Global") if "L2 Theoretical Sectors Global" in hdr else None
            iL = hdr.index("L2 Theoretical Sectors Local") if "L2 Theoretical Sectors Local" in hdr else None
            continue
        if hdr is None or not r[0].isdigit() or r[2] != "-":
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += int(r[iS] or 0); a[1] += int(r[iI] or 0); a[2] += int(r[iT] or 0); a[3] = r[1].strip()[:100]; a[4] += int(r[iG] or 0) if iG is not None else 0; a[5] += int(r[iL] or 0) if iL is not None else 0
    tot = [sum(v[i] for v in agg.values()) for i in (0, 1, 2, 4, 5)]
    print("totals: samples %d, warp instructions %.2f G, lanes %.1f, L2 sectors global %.1f M, local %.1f M" % (tot[0], tot[1] / 1e9, tot[2] / max(1, tot[1]), tot[3] / 1e6, tot[4] / 1e6))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp %5.1f%% inst lanes=%4.1f glob=%6.0fM loc=%5.0fM  %s:%d  %s" % (100 * v[0] / max(1, tot[0]), 100 * v[1] / max(1, tot[1]), v[2] / max(1, v[1]), v[4] / 1e6, v[5] / 1e6, k[0], k[1], v[3]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 36)
