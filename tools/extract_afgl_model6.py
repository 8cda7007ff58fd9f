#!/usr/bin/env python3
"""Extracts the AFGL U.S. Standard atmosphere (model 6) tables that LBLATM falls back to for unit flag '6' -- altitude,
pressure, temperature and the seven molecular mixing-ratio profiles (ppmv) on 50 levels -- from the DATA statements of the
reference's src/rrtatm.f (BLOCK DATA MLATMB: ALT, P6, T6, AMOL61..AMOL67) into rrtmg_lw_b200/data/afgl_us_standard.json.
Numeric constants of a published atmosphere (Anderson et al. 1986, AFGL-TR-86-0110), not code; run once, fixture committed.

    python tools/extract_afgl_model6.py /root/reference/src/rrtatm.f
"""
import json
import os
import re
import sys


def data_block(text, name, n=50):
    m = re.search(r"^\s+DATA %s\s*/" % name, text, flags=re.M)
    assert m, name
    vals = []
    start = text.rfind("\n", 0, m.end()) + 1
    first = True
    for line in text[start:].split("\n"):
        body = line[:72]                      # fixed form: columns 73-80 are sequence labels
        if body[:1] in "Cc*":
            continue
        body = body[m.end() - start:] if first else body[6:]   # after "DATA X /" / after the continuation column
        first = False
        for tok in re.findall(r"[-+]?\d*\.\d+(?:[Ee][-+]?\d+)?|[-+]?\d+\.(?:[Ee][-+]?\d+)?", body):
            vals.append(float(tok))
            if len(vals) == n:
                return vals
    raise ValueError(name)


def main(path):
    text = open(path).read()
    out = {"source": "src/rrtatm.f BLOCK DATA MLATMB (AFGL U.S. Standard 1976, model 6)",
           "alt_km": data_block(text, "ALT"), "p_mb": data_block(text, "P6"), "t_k": data_block(text, "T6"),
           "amol_ppmv": [data_block(text, "AMOL6%d" % k) for k in range(1, 8)]}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "rrtmg_lw_b200", "data", "afgl_us_standard.json")
    json.dump(out, open(dst, "w"), indent=0)
    print(dst, len(out["alt_km"]), out["alt_km"][:3], out["p_mb"][:3], out["t_k"][:3], [a[0] for a in out["amol_ppmv"]])


if __name__ == "__main__":
    main(sys.argv[1])
