#!/usr/bin/env python
"""List SASS lines of an ncu source-page CSV in program order with executed counts, between indices."""
import csv, gzip, sys, collections
fn = sys.argv[1]; lo = int(sys.argv[2]); hi = int(sys.argv[3])
f = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)
rows = list(csv.reader(f)); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
if lo < 0:
    # histogram of executed counts by bucket and opcode
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0e6
    mix = collections.Counter(); n = 0
    for r in data:
        e = int(r[ix["Instructions Executed"]] or 0)
        if e >= thr:
            op = r[ix["Source"]].split(); opn = op[0] if not op[0].startswith("@") else op[1]
            mix[opn.split(".")[0]] += 1; n += 1
    print(n, "static instructions with exec >=", thr); print(mix.most_common(40))
else:
    for i in range(lo, min(hi, len(data))):
        r = data[i]
        print("%5d %9s %6s  %s" % (i, r[ix["Instructions Executed"]], r[ix["# Samples"]], r[ix["Source"]].strip()[:110]))
