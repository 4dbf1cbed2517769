#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): python tools/launch_summary.py list.csv "title" > out.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    k = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
    tot[k] += ms
    cnt[k] += 1
allms = sum(tot.values())
print("# %s, ncu --metrics gpu__time_duration.sum --clock-control none" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per kernel: launches, summed gpu__time_duration [ms], share (ncu serialises launches and flushes caches: shares, not absolutes)")
for k, v in tot.most_common():
    print("%-44s %5d %12.3f %6.1f%%" % (k[:44], cnt[k], v, 100 * v / allms))
print("# total %.3f ms in %d launches" % (allms, sum(cnt.values())))
