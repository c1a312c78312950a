"""Warp-stall samples of one kernel of an .ncu-rep (captured with --import-source on) aggregated by SASS opcode, by stall reason and
by code region (between barriers): ncu -i rep --page source --csv --kernel-id ::regex:NAME:1 > src.csv; python tools/ncu_source_hot.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], rows[h + 1:]
ix = {c: i for i, c in enumerate(hdr)}
stallcols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot, agg, stot, out = 0, {}, {c: 0 for c in stallcols}, []
num = lambda v: int(v) if v.strip().isdigit() else 0
for k, r in enumerate(data):
    if len(r) < len(hdr): continue
    s = num(r[ix["# Samples"]]); tot += s
    toks = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    agg[op] = agg.get(op, 0) + s
    for c in stallcols: stot[c] += num(r[ix[c]])
    out.append((k, s, r[ix["Source"]].strip(), {c[6:]: num(r[ix[c]]) for c in stallcols if num(r[ix[c]]) > 0}))
print("total samples", tot)
print("by opcode:", sorted(agg.items(), key=lambda kv: -kv[1])[:14])
print("by reason:", sorted(stot.items(), key=lambda kv: -kv[1])[:10])
cum = 0
for k, s, src, st in out:
    cum += s
    if any(t in src for t in ("BAR.", "DEPBAR", "EXIT", "BRA ")) or s > 0.01 * tot:
        print("%5d cum %6d  %5d  %-60s %s" % (k, cum, s, src[:60], st))
