#!/usr/bin/env python3
"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_launch_shares.py launches.csv > profiles/xyz_shares.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    t = tot.setdefault(r[ki], [0.0, 0])
    t[0] += v; t[1] += 1
s = sum(t[0] for t in tot.values())
print("# per-kernel share of the launches in %s (ncu: cold-cache, serialised — compare shares, not absolutes)" % sys.argv[1])
for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%10.3f ms %5d launches %5.1f%%  %s" % (ms, n, 100 * ms / s, k[:110]))
