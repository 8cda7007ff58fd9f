#!/usr/bin/env python3
"""Regenerates profiles/traffic.json (read by bench.py for roofline.traffic) from an ncu launch-list capture
(--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of one step of the default bench
workload: DRAM bytes per column of the dominant kernel group (taumol + sweep kernels on the two-phase path, the 16
band kernels on the fused path) and of the whole step.

    python tools/traffic_json.py profiles/r02_v1_traffic.csv 1048576 4
"""
import collections
import csv
import json
import os
import re
import sys


def main(path, ncol, band_kernel):
    rows = list(csv.reader(open(path)))
    per = collections.OrderedDict()
    for r in rows:
        if len(r) < 15 or r[0] == "ID":
            continue
        per.setdefault((r[0], r[4]), {})[r[12]] = float(r[14].replace(",", ""))
    group = re.compile(r"taumol_kernel|rtrnmc_kernel" if band_kernel == 4 else r"band_kernel<")
    g_bytes = g_ns = all_bytes = all_ns = 0.0
    kernels = collections.OrderedDict()
    for (_, name), v in per.items():
        b = v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
        t = v["gpu__time_duration.sum"]
        all_bytes += b
        all_ns += t
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("lw::", "")
        k = kernels.setdefault(short, [0.0, 0.0])
        k[0] += b
        k[1] += t
        if group.search(name):
            g_bytes += b
            g_ns += t
    out = {"band_kernel": band_kernel, "source": os.path.basename(path), "columns": ncol,
           "dram_bytes_per_column": g_bytes / ncol, "whole_step_dram_bytes_per_column": all_bytes / ncol,
           "group_ms": g_ns / 1e6, "whole_step_ms": all_ns / 1e6,
           "per_kernel": {k: {"dram_bytes_per_column": v[0] / ncol, "ms": v[1] / 1e6, "tb_per_s": v[0] / v[1] / 1e3}
                          for k, v in kernels.items()},
           "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one %d-column step (%s): %s = %.1f GB in %.1f ms = "
                   "%.2f TB/s; whole step %.1f GB" % (ncol, os.path.basename(path),
                                                     "taumol + sweep kernels" if band_kernel == 4 else "band kernels 1..16",
                                                     g_bytes / 1e9, g_ns / 1e6, g_bytes / g_ns / 1e3, all_bytes / 1e9)}
    dst = os.path.join(os.path.dirname(os.path.abspath(path)), "traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(out["note"])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 4)
