#!/usr/bin/env python
"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...) into per-kernel totals of the LAST
device-resident step in it.  usage: launch_summary.py X.csv [first_kernel_of_a_step=k_good_len] > txt"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else "k_good_len"
    hdr, out = None, []
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            out.append((d["Kernel Name"].split("(")[0], float(d["Metric Value"].replace(",", "")) / 1e6))
    starts = [i for i, (n, _) in enumerate(out) if n.startswith(first)]
    # a device-resident step has ONE quality-floor launch; an end-to-end step has one per uploaded batch: take the last isolated one
    iso = [s for j, s in enumerate(starts) if (j + 1 == len(starts) or starts[j + 1] - s > 20)]
    lo = iso[-2] if len(iso) >= 2 else (iso[-1] if iso else 0)
    hi = starts[starts.index(lo) + 1] if lo in starts and starts.index(lo) + 1 < len(starts) else len(out)
    seg = out[lo:hi]
    agg, cnt = collections.OrderedDict(), collections.Counter()
    for n, m in seg:
        agg[n] = agg.get(n, 0.0) + m
        cnt[n] += 1
    tot = sum(agg.values())
    print("launches %d, sum %.2f ms (cold-cache serialised launch times: compare SHARES, not absolutes)" % (len(seg), tot))
    for n, m in sorted(agg.items(), key=lambda x: -x[1]):
        print("%-56s n=%4d %9.3f ms %5.1f%%" % (n[:56], cnt[n], m, 100 * m / tot))


if __name__ == "__main__":
    main()
