#!/usr/bin/env python
import csv, gzip, sys, collections
fn = sys.argv[1]; bs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
f = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)
rows = list(csv.reader(f)); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
for b in range(0, len(data), bs):
    chunk = data[b:b + bs]
    s = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
    e = sum(int(r[ix["Instructions Executed"]] or 0) for r in chunk)
    ops = collections.Counter()
    for r in chunk:
        op = r[ix["Source"]].split(); opn = op[0] if not op[0].startswith("@") else op[1]
        ops[opn.split(".")[0]] += 1
    st = collections.Counter()
    for r in chunk:
        for c in ("stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_branch_resolving", "stall_selected"):
            st[c[6:]] += int(r[ix[c]] or 0)
    print("%5d-%5d samples %5.1f%% exec %6.1fM  %s | %s" % (b, b + bs, 100.0 * s / tot, e / 1e6,
          " ".join("%s:%d" % kv for kv in st.most_common(3)), " ".join("%s%d" % kv for kv in ops.most_common(6))))
