#!/usr/bin/env python3
"""Algorithmic FP64 operation counts per column of the BASELINE.json configurations, from the op-counting build of
the oracle (oracle/liborc_count.so: every double is a counting wrapper; +, -, *, / count 1 each, exp / log / pow are
listed separately).  Writes profiles/flops_per_column.json, which bench.py reads for its roofline numerators.

    python tools/count_flops.py [ncol_sample]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import orc  # noqa: E402
from rrtmg_lw_b200 import kdata, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    blob = kdata.synthetic_blob()
    out = {"sample_columns": n, "how": "oracle/liborc_count.so (orc.count_ops) on synth.make_batch samples, seed 20261017"}
    b = synth.make_batch(n, 72)
    out["C2"] = orc.count_ops(blob, b, icld=0, idrv=0)
    b = synth.make_batch(n, 72, cloudy=True)
    out["C3"] = orc.count_ops(blob, b, icld=2, idrv=0)
    b = synth.make_batch(n, 72, cloudy=True, aerosol=True)
    orc.ini(blob, 1004.0, 4)
    alpha = orc.get_alpha(b, 5, idcor=1, juldat=200)
    out["C4"] = orc.count_ops(blob, b, icld=5, idrv=1, alpha=alpha)
    b = synth.make_batch(n, 137, cloudy=True)
    out["C5"] = orc.count_ops(blob, b, icld=2, idrv=0)
    path = os.path.join(ROOT, "profiles", "flops_per_column.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: (v["flops"] if isinstance(v, dict) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
