"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share of the total.
   python scripts/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    ns = float(r[14])
    a = agg[name]
    a[0] += 1
    a[1] += ns
    a[2] = max(a[2], ns)
tot = sum(a[1] for a in agg.values())
print("# %s: %d launches, %.3f ms summed kernel time (cold-cache, serialised: compare shares)" % (sys.argv[1], len(rows), tot / 1e6))
print("%-46s %8s %12s %8s %12s" % ("kernel", "launches", "total ms", "share", "max ms"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s %8d %12.3f %7.1f%% %12.3f" % (k[:46], a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / 1e6))
