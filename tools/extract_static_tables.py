#!/usr/bin/env python3
"""Extract the init-time constant tables that ARE present in the reference mount.

Reads the Fortran array-constructor assignments of
  src/rrtmg_lw_setcoef.f90  lwatmref        (:437-597)   pref, preflog, tref, chi_mls
  src/rrtmg_lw_setcoef.f90  lwavplank       (:600-1303)  totplnk, totplk16
  src/rrtmg_lw_setcoef.f90  lwavplankderiv  (:1306-2009) totplnkderiv, totplk16deriv
  src/rrtmg_lw_init.f90     lwcldpr         (:2037-2675) abscld1, absice0..3, absliq0, absliq1
and writes one little-endian float64 blob in the fixed order of STATIC_LAYOUT
(Fortran element order inside each array).  Only numbers are carried over; no
reference source text enters this repository.  Run in the build container
(needs /root/reference); the blob is committed because the GPU box has no
reference mount.

    python tools/extract_static_tables.py [/root/reference] [out.bin]
"""
import re
import sys
import numpy as np

# (name, Fortran shape) -- order == blob order; mirrored in csrc/static_tables.h
STATIC_LAYOUT = [
    ("pref", (59,)), ("preflog", (59,)), ("tref", (59,)), ("chi_mls", (7, 59)),
    ("totplnk", (181, 16)), ("totplk16", (181,)),
    ("totplnkderiv", (181, 16)), ("totplk16deriv", (181,)),
    ("abscld1", (1,)), ("absice0", (2,)), ("absice1", (2, 5)),
    ("absice2", (43, 16)), ("absice3", (46, 16)),
    ("absliq0", (1,)), ("absliq1", (58, 16)),
]

_num = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][-+]?\d+)?")


def _strip(src):
    out = []
    for line in src.splitlines():
        if line.lstrip().startswith("!"):
            continue
        out.append(line.split("!")[0].rstrip())
    text = "\n".join(out)
    return re.sub(r"&\s*\n", " ", text)


def _parse_sub(spec, extent):
    spec = spec.strip()
    if spec == ":":
        return list(range(extent))
    if ":" in spec:
        a, b = spec.split(":")
        return list(range(int(a) - 1, int(b)))
    return [int(spec) - 1]


def extract(text, name, shape):
    arr = np.full(shape, np.nan, dtype=np.float64)
    # scalar assignment
    if shape == (1,):
        m = re.search(r"^\s*%s\s*=\s*(%s)_rb" % (name, _num.pattern), text, re.M)
        arr[0] = float(m.group(1).replace("d", "e").replace("D", "e"))
        return arr
    pat = re.compile(r"^\s*%s\s*\(([^)]*)\)\s*=\s*\(/(.*?)/\)" % name, re.M | re.S)
    for m in pat.finditer(text):
        subs = [s for s in m.group(1).split(",")]
        vals = [float(v.replace("d", "e").replace("D", "e"))
                for v in _num.findall(m.group(2).replace("_rb", " "))]
        idx = [_parse_sub(s, e) for s, e in zip(subs, shape)]
        if len(shape) == 1:
            assert len(idx[0]) == len(vals), (name, m.group(1))
            arr[idx[0]] = vals
        else:
            # Fortran order: first subscript fastest
            n = len(idx[0]) * len(idx[1])
            assert n == len(vals), (name, m.group(1), n, len(vals))
            k = 0
            for j in idx[1]:
                for i in idx[0]:
                    arr[i, j] = vals[k]
                    k += 1
    assert not np.isnan(arr).any(), "unfilled elements in " + name
    return arr


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else "rrtmg_lw_b200/data/lw_static_tables.bin"
    text = _strip(open(root + "/src/rrtmg_lw_setcoef.f90").read()) + "\n" + \
        _strip(open(root + "/src/rrtmg_lw_init.f90").read())
    chunks = []
    for name, shape in STATIC_LAYOUT:
        a = extract(text, name, shape)
        chunks.append(np.asfortranarray(a).ravel(order="F"))
        print("%-14s %-10s min %.6e max %.6e" % (name, shape, a.min(), a.max()))
    blob = np.concatenate(chunks).astype("<f8")
    blob.tofile(out)
    print("wrote", out, blob.size, "doubles")


if __name__ == "__main__":
    main()
