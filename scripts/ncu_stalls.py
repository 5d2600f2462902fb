"""Summarise an ncu source-page CSV: stall-reason totals and the most-sampled SASS instructions.
   ncu -i rep.ncu-rep --page source --csv [--kernel-name regex:..] > src.csv ; python scripts/ncu_stalls.py src.csv [ntop]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
data = data[:len(data) // 2] if len(data) > 1 and data[0][1] == data[len(data) // 2][1] else data
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = {s: 0 for s in stalls}
alls = 0
for r in data:
    alls += I(r[ix['# Samples']])
    for s in stalls:
        tot[s] += I(r[ix[s]])
print("kernel:", rows[0][1][:100])
print("samples", alls, "instructions", len(data))
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:9]:
    print("  %-24s %8d %5.1f%%" % (s, v, 100.0 * v / max(alls, 1)))
top = sorted(range(len(data)), key=lambda q: -I(data[q][ix['# Samples']]))[:ntop]
for q in top:
    r = data[q]
    st = {s[6:]: I(r[ix[s]]) for s in stalls}
    best = sorted(st.items(), key=lambda x: -x[1])[:2]
    print("%5d %7s  %-72s %s exec %s" % (q, r[ix['# Samples']], r[ix['Source']].strip()[:72], best, r[ix['Instructions Executed']]))
