#!/usr/bin/env python3
"""Standalone form of the host-side half of rrtmg_lw_ini (packaging item): reads the k-distribution data,
runs the library's restatement of the reference's init (src/rrtmg_lw_init.f90: constants, lookup tables, 16-g -> 140-g
reduction of every k-table) on the HOST and writes the versioned table image that ``rrtmg_lw_b200_ini_tables`` uploads.
Needs the built library but no GPU.

    python tools/make_table_blob.py --kdata rrtmg_lw.nc -o rrtmg_lw_tables.bin [--cpdair 1004.64] [--real-kind 4|8]
    python tools/make_table_blob.py --kdata synthetic -o tables.bin          # the deterministic test tables
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--kdata", required=True, help="rrtmg_lw.nc (src/rrtmg_lw_read_nc.f90 layout), a raw float64 blob "
                    "in the order of include/rrtmg_lw_kdata_layout.h, or 'synthetic'")
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--cpdair", type=float, default=1004.64, help="specific heat of dry air handed to rrtmg_lw_ini")
    ap.add_argument("--real-kind", type=int, default=4, choices=[4, 8],
                    help="default real kind of the Fortran build being replaced (8 = the shipped -r8 makefiles)")
    a = ap.parse_args()
    import numpy as np
    from rrtmg_lw_b200 import api, kdata
    if a.kdata == "synthetic":
        blob = kdata.synthetic_blob()
    elif a.kdata.endswith(".nc"):
        blob = kdata.load_netcdf(a.kdata)
    else:
        blob = np.fromfile(a.kdata, dtype="<f8")
    img = api.tables_build(a.cpdair, blob, a.real_kind)
    with open(a.output, "wb") as f:
        f.write(img)
    print("%s: %d bytes (format version %d)" % (a.output, len(img), int.from_bytes(img[8:16], "little")))


if __name__ == "__main__":
    main()
