"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of device time)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[start:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
    t = float(r[mv].replace(",", ""))
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
print("total device time %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%5d total_us=%10.1f share=%5.1f%% avg_us=%9.2f" % (k[:62], n, t, 100 * t / tot, t / n))
