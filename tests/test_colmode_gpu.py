"""The reference's example suite (run_examples_std_atm, fixture tests/golden/std_atm_cases.json) run by the
column-mode harness through the C ABI of the CUDA library, against the same harness over the CPU oracle:
FP64 fluxes within 1e-6 W m-2, heating rates within 1e-6 K day-1 (BASELINE.json north_star), for the
clear / aerosol / cross-section / idrv cases and the McICA cases (MT19937 stream, permuteseed = sample
number, icld 2 / 4 / 5, idcor 0 / 1), on host pointers and on device pointers."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from colmode_checker import OracleBackend  # noqa: E402
from rrtmg_lw_b200 import colmode  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = json.load(open(os.path.join(ROOT, "tests", "golden", "std_atm_cases.json")))
RUNNABLE = [n for n, d in SUITE.items() if d["case"] is not None]


@pytest.mark.parametrize("name", RUNNABLE)
def test_example_case_matches_oracle(gpu, oracle, name):
    d = SUITE[name]
    c = colmode.Case(**d["case"])
    nmca = 12 if c.imca == 1 else None
    ref = colmode.run_case(c, OracleBackend(oracle), nmca=nmca)
    for device in (None, "cuda:0"):
        got = colmode.run_case(c, colmode.GpuBackend(device), nmca=nmca)
        gpu.sync()
        for k in ("uflx", "dflx", "fnet", "htr"):
            err = float(np.abs(got[k] - ref[k]).max())
            assert err <= 1e-6, (name, device, k, err)
        # data-independent golden number: surface upward flux (Planck tables; idrv1: dtbound-adjusted)
        assert abs(got["uflx"][0] - d["golden"]["uflx"][0]) <= 1.5e-3
    text = colmode.format_output(got)
    assert len(text.split("\n")) == c.nlayers + 1 + 3 + 2


def test_full_200_sample_mcica_average(gpu, oracle):
    """The reference's nmca = 200 loop (1col:459-460, 640-702) on the GPU equals the oracle's."""
    c = colmode.Case(**SUITE["MLS-cld7-imca1-icld5-idcor1"]["case"])
    ref = colmode.run_case(c, OracleBackend(oracle))
    got = colmode.run_case(c, colmode.GpuBackend("cuda:0"))
    assert got["nmca"] == 200
    for k in ("uflx", "dflx", "fnet", "htr"):
        assert float(np.abs(got[k] - ref[k]).max()) <= 1e-6, k


def test_non_mcica_cloudy_case_all_17_blocks(gpu, oracle):
    """MLS-cld5-imca0-icld2 (iout = 99): cldprop + rtrnmr through the library's rrtmg_lw_nomcica entry, broadband block
    and the 16 per-band runs, against the oracle's restatement; and with icld switched to 1 (rtrn)."""
    import dataclasses
    c = colmode.Case(**SUITE["MLS-cld5-imca0-icld2"]["case"])
    assert (c.imca, c.icld, c.iout) == (0, 2, 99)
    for case in (c, dataclasses.replace(c, icld=1)):
        ref = colmode.run_blocks(case, OracleBackend(oracle))
        got = colmode.run_blocks(case, colmode.GpuBackend("cuda:0"))
        gpu.sync()
        assert len(ref) == len(got) == 17
        for k, (r, g) in enumerate(zip(ref, got)):
            for key in ("uflx", "dflx", "fnet", "htr"):
                err = float(np.abs(g[key] - r[key]).max())
                assert err <= 1e-6, (case.icld, "block", k, key, err)
        assert float(np.abs(got[0]["dflx"] - colmode.run_case(dataclasses.replace(case, icld=0), colmode.GpuBackend("cuda:0"))["dflx"]).max()) > 1.0


@pytest.mark.parametrize("name", ["MLS-clr", "MLS-clr-aer12", "MLS-clr-idrv1"])
def test_per_band_blocks_match_oracle(gpu, oracle, name):
    """iout = 99: all 17 output blocks (broadband + one run per band, band 16 with its own Planck table) through the
    library's band-range option equal the oracle's per-band runs to 1e-6."""
    c = colmode.Case(**SUITE[name]["case"])
    assert c.iout == 99
    ref = colmode.run_blocks(c, OracleBackend(oracle))
    got = colmode.run_blocks(c, colmode.GpuBackend("cuda:0"))
    gpu.sync()
    assert len(ref) == len(got) == 17
    for k, (r, g) in enumerate(zip(ref, got)):
        assert (r["istart"], r["iend"]) == (g["istart"], g["iend"])
        for key in ("uflx", "dflx", "fnet", "htr"):
            err = float(np.abs(g[key] - r[key]).max())
            assert err <= 1e-6, (name, "block", k, key, err)
    text = colmode.format_blocks(got)
    assert text.count("Wavenumbers:") == 17
