#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall reasons, instruction mix, hottest SASS lines."""
import csv, gzip, sys, collections, re
fn = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)
rows = list(csv.reader(f))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
mix = collections.Counter()
nsamp = 0
ninst = 0
recs = []
for r in data:
    if len(r) < len(hdr):
        continue
    s = int(r[ix["# Samples"]] or 0)
    e = int(r[ix["Instructions Executed"]] or 0)
    nsamp += s
    ninst += e
    for c in stall_cols:
        tot[c] += int(r[ix[c]] or 0)
    op = r[ix["Source"]].split()
    opn = op[0] if not op[0].startswith("@") else op[1]
    mix[opn.split(".")[0]] += e
    recs.append((s, e, r[ix["Address"]], r[ix["Source"]].strip(), r))
print("samples", nsamp, "warp-instructions", ninst)
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, sum(tot.values()))) for k, v in tot.most_common(10)))
print("instruction mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ninst) for k, v in mix.most_common(24)))
print("hottest lines:")
for i, (s, e, a, src, r) in enumerate(recs):
    pass
order = sorted(range(len(recs)), key=lambda i: -recs[i][0])[:top]
for i in order:
    s, e, a, src, r = recs[i]
    why = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%6d %5.2f%% exec=%9d  #%5d %-70s %s" % (s, 100.0 * s / nsamp, e, i, src[:70], why))
