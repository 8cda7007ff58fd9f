#!/usr/bin/env python3
"""Build tests/golden/std_atm.json from the reference's example suite (run in the build
container, needs /root/reference; the GPU box has no reference mount so the result is committed).

For every clear-sky iatm=0 case of run_examples_std_atm/ it stores
  * the input profile parsed with the record formats of src/rrtmg_lw.1col.f90:889-894, 1139-1142
    (layer pressure/temperature, level pressure/temperature, 7 molecular mixing ratios,
    broadening-gas column, surface temperature), and
  * the first (10-3250 cm-1) block of the shipped output file: level, pressure, upward flux,
    downward flux, net flux, heating rate (formats :737-743).
The k-dependent numbers of the outputs cannot be reproduced until the reference's k-distribution
data are available; the surface upward flux (Planck tables only) can, and is what the tests pin.
"""
import json
import os
import sys

ROOT = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EX = os.path.join(ROOT, "run_examples_std_atm")
CASES = ["MLS-clr", "MLW-clr", "SAW-clr", "TROP-clr"]


def parse_input(path):
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.startswith("$"))
    flags = lines[i + 1]
    # 9011 format (18x,i2,29x,i1,19x,i1,13x,i2,2x,i3,1x,i1,1x,i1,i1)
    I = lambda a, b: int(flags[a:b].strip() or 0)  # noqa: E731  (blank field reads as 0)
    iaer = I(18, 20); iatm = I(49, 50); ixsect = I(69, 70)
    iout = I(87, 90); idrv = I(91, 92); imca = I(93, 94); icld = I(94, 95)
    assert iatm == 0 and icld == 0, (iatm, icld)
    rec = lines[i + 2]
    tbound = float(rec[0:10])
    k = i + 3 + (1 if idrv == 1 else 0)
    hdr = lines[k]
    iform = int(hdr[1:2]); nlayers = int(hdr[2:5]); nmol = int(hdr[5:10] or 7)
    assert iform == 1 and nmol == 7
    k += 1
    pavel, tavel, pz, tz, wkl, wbrodl = [], [], [], [], [], []
    for l in range(nlayers):
        a = lines[k]; b = lines[k + 1]; k += 2
        pavel.append(float(a[0:15])); tavel.append(float(a[15:25]))
        if l == 0:
            pz.append(float(a[48:56])); tz.append(float(a[56:63]))
        pz.append(float(a[70:78])); tz.append(float(a[78:85]))
        vals = [float(b[15 * m:15 * (m + 1)]) for m in range(8)]
        wkl.append(vals[:7]); wbrodl.append(vals[7])
    return dict(iaer=iaer, ixsect=ixsect, iout=iout, idrv=idrv, imca=imca, icld=icld, tbound=tbound, nlayers=nlayers,
                pavel=pavel, tavel=tavel, pz=pz, tz=tz, wkl=wkl, wbrodl=wbrodl)


def parse_output(path):
    rows = []
    started = False
    for l in open(path):
        p = l.split()
        if len(p) == 6 and p[0].isdigit():
            started = True
            rows.append([int(p[0])] + [float(x) for x in p[1:]])
        elif started:
            break
    return rows


def main():
    out = {}
    for c in CASES:
        out[c] = dict(input=parse_input(os.path.join(EX, "input_rrtm_" + c)),
                      output=parse_output(os.path.join(EX, "output_rrtm_" + c)))
        o = out[c]["output"]
        print(c, "layers", out[c]["input"]["nlayers"], "levels", len(o), "sfc up", o[-1][2], "toa up", o[0][2])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "std_atm.json")
    json.dump(out, open(dst, "w"), indent=0)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
