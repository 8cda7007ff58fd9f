#!/usr/bin/env python3
"""Build tests/golden/std_atm_cases.json: every case of the reference's run_examples_std_atm suite
(script.run_std_atm) parsed with rrtmg_lw_b200/colmode.py -- the INPUT_RRTM record set, the
IN_CLD_RRTM / IN_AER_RRTM file the run script pairs it with -- plus the broadband block
(10-3250 cm-1) of the shipped OUTPUT_RRTM file.  Run in the build container (needs /root/reference;
the GPU box has no reference mount, so the result is committed).  iatm = 1 (ICRCCM_sonde) goes through
rrtmg_lw_b200/rrtatm.py (the layering of src/rrtatm.f)."""
import dataclasses
import json
import os
import sys

ROOT = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EX = os.path.join(ROOT, "run_examples_std_atm")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rrtmg_lw_b200 import colmode  # noqa: E402

# (output name, input file, cloud file, aerosol file) as script.run_std_atm pairs them
SUITE = [
    ("ICRCCM_sonde", "input_rrtm_ICRCCM_sonde", None, None),
    ("MLS-cld5-imca0-icld2", "input_rrtm_MLS-cld-imca0-icld2", "in_cld_rrtm-cld5", None),
    ("MLS-cld5-imca1-icld2", "input_rrtm_MLS-cld-imca1-icld2", "in_cld_rrtm-cld5", None),
    ("MLS-cld7-imca1-icld2", "input_rrtm_MLS-cld-imca1-icld2", "in_cld_rrtm-cld7", None),
    ("MLS-cld7-imca1-icld4-idcor0", "input_rrtm_MLS-cld-imca1-icld4-idcor0", "in_cld_rrtm-cld7", None),
    ("MLS-cld7-imca1-icld5-idcor0", "input_rrtm_MLS-cld-imca1-icld5-idcor0", "in_cld_rrtm-cld7", None),
    ("MLS-cld7-imca1-icld5-idcor1", "input_rrtm_MLS-cld-imca1-icld5-idcor1", "in_cld_rrtm-cld7", None),
    ("MLS-clr", "input_rrtm_MLS-clr", None, None),
    ("MLS-clr-aer12", "input_rrtm_MLS-clr-aer12", None, "in_aer_rrtm-aer12"),
    ("MLS-clr-xsec", "input_rrtm_MLS-clr-xsec", None, None),
    ("MLS-clr-idrv1", "input_rrtm_MLS-clr-idrv1", None, None),
    ("MLW-clr", "input_rrtm_MLW-clr", None, None),
    ("SAW-clr", "input_rrtm_SAW-clr", None, None),
    ("TROP-clr", "input_rrtm_TROP-clr", None, None),
]


def rd(name):
    return open(os.path.join(EX, name)).read() if name else None


def main():
    out = {}
    for name, inp, cld, aer in SUITE:
        blocks = colmode.parse_output_rrtm(rd("output_rrtm_" + name))
        b0 = blocks[0]
        gold = dict(nblocks=len(blocks), wn=[b0["wn1"], b0["wn2"]],
                    **{k: [float(x) for x in b0[k]] for k in ("pz", "uflx", "dflx", "fnet", "htr")})
        try:
            case = dataclasses.asdict(colmode.parse_input_rrtm(rd(inp), rd(cld), rd(aer)))
            print(name, "layers", case["nlayers"], "icld", case["icld"], "imca", case["imca"], "idrv", case["idrv"],
                  "iaer", case["iaer"], "xs", case["xsname"], "blocks", len(blocks), "sfc up", gold["uflx"][0])
        except NotImplementedError as e:
            case = None
            print(name, "unsupported:", e)
        out[name] = dict(files=[inp, cld, aer], case=case, golden=gold)
    dst = os.path.join(REPO, "tests", "golden", "std_atm_cases.json")
    json.dump(out, open(dst, "w"), indent=0)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
