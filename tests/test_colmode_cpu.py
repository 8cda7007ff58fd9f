"""Example-suite harness (rrtmg_lw_b200/colmode.py) on the CPU: the parsers against the
reference's own example files (fixture tests/golden/std_atm_cases.json made by tools/make_golden_cases.py;
re-parsed from /root/reference when that exists), the writer against the shipped OUTPUT_RRTM text, and
the driver loop (McICA averaging, idrv dtbound adjustment) run over the oracle and pinned on the
data-independent numbers of the goldens (surface upward flux, incl. the idrv = 1 adjusted value)."""
import dataclasses
import json
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from colmode_checker import OracleBackend  # noqa: E402
from rrtmg_lw_b200 import colmode  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EX = "/root/reference/run_examples_std_atm"
have_ref = pytest.mark.skipif(not os.path.isdir(REF_EX), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def suite():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "std_atm_cases.json")))


def _supported(suite):
    return {n: d for n, d in suite.items() if d["case"] is not None}


def test_suite_inventory(suite):
    """14 runs in script.run_std_atm; all parse: 13 with layer records (iatm = 0), ICRCCM_sonde through the rrtatm
    layering (iatm = 1, N4)."""
    assert len(suite) == 14
    assert [n for n, d in suite.items() if d["case"] is None] == []
    assert [n for n, d in suite.items() if d["case"]["iatm"] == 1] == ["ICRCCM_sonde"]
    flags = {n: (d["case"]["icld"], d["case"]["imca"], d["case"]["idrv"], d["case"]["iaer"], d["case"]["ixsect"])
             for n, d in _supported(suite).items()}
    assert flags["MLS-clr"] == (0, 0, 0, 0, 0)
    assert flags["MLS-clr-idrv1"] == (0, 0, 1, 0, 0) and suite["MLS-clr-idrv1"]["case"]["dtbound"] == 5.0
    assert flags["MLS-clr-aer12"] == (0, 0, 0, 10, 0)
    assert flags["MLS-clr-xsec"] == (0, 0, 0, 0, 1)
    assert flags["MLS-cld5-imca0-icld2"] == (2, 0, 0, 0, 0)
    assert flags["MLS-cld7-imca1-icld5-idcor1"] == (5, 1, 0, 0, 0)
    c = suite["MLS-cld7-imca1-icld5-idcor1"]["case"]
    assert (c["idcor"], c["juldat"], c["lat"]) == (1, 1, 45.0)
    c = suite["MLS-cld7-imca1-icld4-idcor0"]["case"]
    assert c["idcor"] == 0 and c["decorr_con"] > 0
    assert (c["inflag"], c["iceflag"], c["liqflag"]) == (2, 3, 1)
    assert sum(1 for f in c["cldfrac"] if f > 0) == 13 and c["cldfrac"][3] == 0.5 and c["clddat"][11] == [2.0, 1.0, 25.0, 10.0]
    a = suite["MLS-clr-aer12"]["case"]["tauaer"]
    assert a[0][0] == 0.0005 and a[0][3] == 0.0015 and sum(1 for r in a if max(r) > 0) == 12
    x = suite["MLS-clr-xsec"]["case"]
    assert [colmode.xs_slot(n) for n in x["xsname"]] == [2, 4, 3, 1]   # F11 F22 F12 CCL4 -> cfc11 cfc22 cfc12 ccl4
    assert x["wx0"][0] == [3.5199330e-10, 1.5e-10, 6.3581760e-10, 1.3975950e-10]


def test_matches_first_golden_fixture(suite, std_atm):
    """Same profiles as the round's first fixture (tools/make_golden.py, an independent parser)."""
    for n in ("MLS-clr", "MLW-clr", "SAW-clr", "TROP-clr"):
        a, b = suite[n]["case"], std_atm[n]["input"]
        for k in ("pavel", "tavel", "pz", "tz", "wbrodl", "tbound", "nlayers"):
            assert a[k] == b[k], (n, k)
        assert [w[:7] for w in a["wkl"]] == b["wkl"]


@have_ref
def test_fixture_is_what_the_reference_files_parse_to(suite):
    rd = lambda f: open(os.path.join(REF_EX, f)).read() if f else None  # noqa: E731
    for n, d in _supported(suite).items():
        c = colmode.parse_input_rrtm(*[rd(f) for f in d["files"]])
        assert dataclasses.asdict(c) == d["case"], n


@have_ref
def test_writer_reproduces_the_shipped_output_text(suite):
    """format_output fed with the golden numbers gives the reference's block character for character
    (formats 9899-9958, src/rrtmg_lw.1col.f90:737-751), for every case of the suite."""
    for n, d in suite.items():
        text = open(os.path.join(REF_EX, "output_rrtm_" + n)).read()
        g = {k: np.array(v) for k, v in d["golden"].items() if k in ("pz", "uflx", "dflx", "fnet", "htr")}
        mine = colmode.format_output(g).split("\n")
        theirs = text.split("\n")[:len(mine) - 2]
        assert [ln.rstrip() for ln in mine[:len(theirs)]] == [ln.rstrip() for ln in theirs], n


def test_writer_parser_round_trip():
    rng = np.random.default_rng(7)
    pz = np.array([1013.0, 500.0, 99.5, 9.95, 0.995, 0.0995, 0.00995, 0.000995])
    res = dict(pz=pz, uflx=rng.uniform(0, 500, 8), dflx=rng.uniform(0, 400, 8), htr=rng.uniform(-30, 30, 8))
    res["fnet"] = res["uflx"] - res["dflx"]
    res["htr"][3] = -0.25
    b = colmode.parse_output_rrtm(colmode.format_output(res, iplon=3) + colmode.format_output(res, iplon=4))
    assert len(b) == 2 and b[0]["atm"] == 3 and (b[0]["wn1"], b[0]["wn2"]) == (10.0, 3250.0)
    assert list(b[0]["level"]) == list(range(8))
    np.testing.assert_allclose(b[0]["uflx"], res["uflx"], atol=5.1e-5)
    np.testing.assert_allclose(b[0]["fnet"], res["fnet"], atol=5.1e-8)
    np.testing.assert_allclose(b[0]["htr"], res["htr"], atol=5.1e-6)
    np.testing.assert_allclose(b[0]["pz"], pz, rtol=6e-4)


def test_cloud_and_aerosol_readers_on_synthetic_text():
    cld = "    0    0    0\n    2       1.0       3.5\n   10      0.25       0.5\n%\nignored\n"
    inflag, iceflag, liqflag, frac, dat = colmode.parse_in_cld(cld, 12)
    assert (inflag, iceflag, liqflag) == (0, 0, 0) and frac[1] == 1.0 and dat[1][0] == 3.5 and frac[9] == 0.25
    with pytest.raises(ValueError):
        colmode.parse_in_cld("    2    3    1\n   13       1.0       3.5\n%\n", 12)
    row = lambda lay, v: "  %3d" % lay + "".join("%7.4f" % v for _ in range(16))  # noqa: E731
    aer = "    2\n    2    1\n" + row(1, 0.01) + "\n" + row(2, 0.02) + "\n    1    1\n" + row(5, 0.05) + "\n"
    tau = colmode.parse_in_aer(aer, 6)
    assert tau[0][0] == 0.01 and tau[1][15] == 0.02 and tau[4][7] == 0.05 and max(tau[2]) == 0.0
    with pytest.raises(ValueError, match="MORE THAN ONE AEROSOL"):
        colmode.parse_in_aer("    2\n    1    1\n" + row(1, 0.01) + "\n    1    1\n" + row(1, 0.02) + "\n", 6)
    # implied decimals of a field without a decimal point (Fortran F/E edit descriptors)
    assert colmode._real("  950", 3) == pytest.approx(0.95) and colmode._real("1.5D2") == 150.0 and colmode._real("   ") == 0.0


def test_mixing_ratio_and_column_amount_inputs_agree(suite):
    """readprof's imix switch (1col:1006-1036): amounts > 1 are molecules/cm2 and divide by the dry column."""
    c = colmode.Case(**suite["MLS-clr"]["case"])
    b1 = colmode.case_to_batch(c)
    assert b1.meta["imix"] == 1
    summol = np.array([sum(w[1:7]) for w in c.wkl])
    coldry = np.array(c.wbrodl) / (1.0 - summol)
    c2 = dataclasses.replace(c, wkl=[[x * cd for x in w] for w, cd in zip(c.wkl, coldry)],
                             wbrodl=list(coldry * (1.0 - summol)))
    b2 = colmode.case_to_batch(c2)
    assert b2.meta["imix"] == 0
    # column amounts carry the water column too: coldry(imix=0) = wbrodl + sum(non-water) by the reference's rule
    np.testing.assert_allclose(b2.co2vmr.numpy(), b1.co2vmr.numpy(), rtol=1e-12)
    np.testing.assert_allclose(b2.h2ovmr.numpy(), b1.h2ovmr.numpy(), rtol=1e-12)


def test_driver_loop_over_the_oracle_pins_the_goldens_surface_flux(suite, oracle):
    """Every supported case through run_case with the CPU checker: level pressures equal the golden's, the
    surface upward flux (Planck tables only; emissivity 1) equals the golden's to 1.5e-3 W m-2 -- for
    MLS-clr-idrv1 that is the dtbound-adjusted value 453.6771, which pins dplankbnd_dt -- and the
    downward flux at the top is zero.  k-dependent numbers are not comparable (synthetic k-data)."""
    be = OracleBackend(oracle)
    for n, d in _supported(suite).items():
        c = colmode.Case(**d["case"])
        r = colmode.run_case(c, be, nmca=None if n == "MLS-cld7-imca1-icld5-idcor1" else 6)
        g = d["golden"]
        assert r["nmca"] == (1 if c.imca == 0 else (200 if n == "MLS-cld7-imca1-icld5-idcor1" else 6))
        np.testing.assert_allclose(r["pz"], g["pz"], rtol=6e-4)
        assert abs(r["uflx"][0] - g["uflx"][0]) <= 1.5e-3, (n, r["uflx"][0], g["uflx"][0])
        assert r["dflx"][-1] == 0.0 and r["htr"][-1] == 0.0
        np.testing.assert_allclose(r["fnet"], r["uflx"] - r["dflx"], atol=1e-12)
        blk = colmode.parse_output_rrtm(colmode.format_output(r))[0]
        np.testing.assert_allclose(blk["uflx"], r["uflx"], atol=5.1e-5)


def test_idrv_adjustment_is_the_reference_formula(suite, oracle):
    """1col:587-610: uflx += duflx_dt * dtbound, fnet and htr recomputed; dflx unchanged; dtbound = 0 is a no-op."""
    be = OracleBackend(oracle)
    c = colmode.Case(**suite["MLS-clr-idrv1"]["case"])
    r5 = colmode.run_case(c, be)
    r0 = colmode.run_case(dataclasses.replace(c, dtbound=0.0), be)
    base = colmode.run_case(dataclasses.replace(c, idrv=0), be)
    np.testing.assert_allclose(r0["uflx"], base["uflx"], atol=1e-9)
    np.testing.assert_allclose(r0["htr"], base["htr"], atol=1e-9)
    np.testing.assert_array_equal(r5["dflx"], base["dflx"])
    o = be.clear(colmode.case_to_batch(c), 1)
    np.testing.assert_allclose(r5["uflx"], base["uflx"] + 5.0 * o["duflx_dt"].reshape(-1), atol=1e-9)
    # the finite difference of two runs 5 K apart agrees with the derivative to second order
    hot = colmode.run_case(dataclasses.replace(c, idrv=0, tbound=c.tbound + 5.0), be)
    # (the second-order term 6 sigma T^2 dT^2 is 0.74 W m-2 at the surface)
    assert abs(hot["uflx"][0] - r5["uflx"][0]) < 1.0
    assert abs(hot["uflx"][-1] - r5["uflx"][-1]) < 1.0


def test_mcica_samples_are_seeded_like_the_column_mode_generator(suite):
    """The column-mode generator seeds sample ims with permuteseed = ims * nsubclw = 140 ims
    (src/mcica_subcol_gen_lw.1col.f90:251), not with ims; the harness must hand exactly those seeds to
    mcica_subcol_lw, with irng = 1 (src/rrtmg_lw.1col.f90:417), and use rrlw_con's heatfac for the idrv adjustment."""
    seen = []

    class Recorder:
        def heatfac(self):
            return 8.4391

        def get_alpha(self, *a):
            raise AssertionError("icld 2 needs no alpha")

        def clear(self, b, idrv):
            raise AssertionError("a cloudy McICA case must not take the clear path")

        def mcica(self, b, icld, idrv, permuteseed, irng, flags, alpha):
            seen.append((permuteseed, irng))
            L = b.nlay
            return dict(uflx=np.zeros(L + 1), dflx=np.zeros(L + 1), hr=np.zeros(L), duflx_dt=np.zeros(L + 1))

    c = colmode.Case(**suite["MLS-cld7-imca1-icld2"]["case"])
    r = colmode.run_case(c, Recorder(), nmca=5)
    assert seen == [(140 * ims, 1) for ims in range(1, 6)]
    assert r["nmca"] == 5
    ref_src = "/root/reference/src/mcica_subcol_gen_lw.1col.f90"
    if os.path.exists(ref_src):
        src = open(ref_src).read()
        assert "permuteseed = ims*nsubclw" in src


def test_iout_99_yields_the_17_blocks_of_the_goldens(suite, oracle):
    """iout = 99: broadband block + one block per band with the band limits of rrlw_wvn (init.f90:215-220), in the
    order and with the 'Wavenumbers' headers of the shipped 17-block outputs; the band blocks add up to the broadband
    block (fluxes are sums over bands) except for band 16, which alone takes the totplk16 Planck table
    (setcoef.f90:233-251) -- its block differs from its share of the broadband sum."""
    assert [b[:2] for b in colmode.output_blocks(99)] == [(1, 16)] + [(k, k) for k in range(1, 17)]
    assert colmode.output_blocks(0) == [(1, 16, 10.0, 3250.0)] and colmode.output_blocks(-1) == []
    assert colmode.output_blocks(7) == [(7, 7, 980.0, 1080.0)]
    be = OracleBackend(oracle)
    name = "MLS-clr-idrv1"
    c = colmode.Case(**suite[name]["case"])
    assert c.iout == 99
    blocks = colmode.run_blocks(c, be)
    assert len(blocks) == 17
    if os.path.isdir(REF_EX):
        gold = colmode.parse_output_rrtm(open(os.path.join(REF_EX, "output_rrtm_" + name)).read())
        assert len(gold) == 17
        assert [(g["wn1"], g["wn2"]) for g in gold] == [(b["wn1"], b["wn2"]) for b in blocks]
        mine = colmode.parse_output_rrtm(colmode.format_blocks(blocks))
        assert [(g["wn1"], g["wn2"], len(g["level"])) for g in gold] == [(g["wn1"], g["wn2"], len(g["level"])) for g in mine]
        # data-independent golden numbers, per band: surface upward flux (emissivity 1, Planck tables, dtbound-adjusted);
        # block 17 pins the totplk16 variant (its golden differs from the 2600 cm-1 -> infinity integral)
        for k, (g, b) in enumerate(zip(gold, blocks)):
            assert abs(b["uflx"][0] - g["uflx"][0]) <= 1.5e-3, (k, b["uflx"][0], g["uflx"][0])
    total = sum(b["uflx"] for b in blocks[1:16])
    be.band_range(16, 16)
    own16 = colmode.run_case(c, be)
    be.band_range(1, 16)
    np.testing.assert_array_equal(own16["uflx"], blocks[16]["uflx"])
    # bands 1..15 are additive; band 16's share of the broadband block is the 2600 -> infinity integral, larger than block 17
    share16 = blocks[0]["uflx"] - total
    assert (share16 > blocks[16]["uflx"] - 1e-12).all() and share16[0] > blocks[16]["uflx"][0]


def test_rrtatm_layering_of_the_sonde_case(suite, oracle):
    """IATM = 1: the sounding of input_rrtm_ICRCCM_sonde layered by rrtmg_lw_b200/rrtatm.py, the
    restatement of the part of src/rrtatm.f that case uses.  Pinned on the reference's own output file: all 32 level
    pressures of output_rrtm_ICRCCM_sonde to its print precision (the top four come from the U.S. Standard defaults of
    DEFALT, the rest from AMERGE's exponential interpolation), the surface upward flux, and -- trap T18 -- the fact that
    the golden is a DRY result: the program converts the sonde's g/kg water vapour with AIRMWT = 0.  Further checks are
    physical: every layer's air column equals the hydrostatic one, PBAR / TBAR lie between the boundary values."""
    d = suite["ICRCCM_sonde"]
    c = colmode.Case(**d["case"])
    g = d["golden"]
    assert c.nlayers == 31 and len(c.pz) == 32 and len(g["pz"]) == 32
    for mine, gold in zip(c.pz, g["pz"]):
        assert abs(mine - gold) <= 0.6 * 10.0 ** (math.floor(math.log10(gold)) - 3), (mine, gold)   # 4 significant digits
    wkl = np.array(c.wkl)
    altz = np.array(c.altz)
    sonde = altz[1:] <= 30.0                              # layers inside the sounding (g/kg water vapour)
    assert sonde.sum() == 27 and (wkl[sonde, 0] == 0.0).all()   # T18: the sonde's water vapour never reaches the radiation code
    assert (wkl[~sonde, 0] > 0.0).all() and wkl[~sonde, 0].sum() * 18.015 / 6.02214199e23 < 1e-4   # only the model default above it
    air = np.array(c.wbrodl) + wkl.sum(axis=1)            # molecules cm-2
    pz = np.array(c.pz)
    hydro = (pz[:-1] - pz[1:]) * 100.0 / 9.80665 / (28.964e-3 / 6.02214199e23) * 1.0e-4
    # LBLATM integrates rho dz on the sounding's OWN altitudes: where the sonde's z and p are not hydrostatic (its
    # 0.06 -> 0.187 km step drops 17 hPa over 127 m) or where the profile switches to the model default (30 km) the
    # layer differs from dp / g; the body of the troposphere agrees to 0.5 %, the whole column to 1 %.
    ratio = air / hydro
    assert np.abs(ratio[3:20] - 1.0).max() < 6e-3
    assert abs(ratio[0] - 1.0) < 1e-3 and abs(air.sum() / hydro.sum() - 1.0) < 1e-2
    assert np.abs(ratio[[1, 2]].mean() - 1.0) < 0.03      # the two non-hydrostatic sonde steps compensate
    assert ((np.array(c.pavel) < pz[:-1]) & (np.array(c.pavel) > pz[1:])).all()
    tz = np.array(c.tz)
    lo, hi = np.minimum(tz[:-1], tz[1:]) - 3.0, np.maximum(tz[:-1], tz[1:]) + 3.0
    assert ((np.array(c.tavel) > lo) & (np.array(c.tavel) < hi)).all()
    assert abs(wkl[:, 6].sum() / (air.sum() - wkl[:, 6].sum() * 0) - 0.209) < 2e-3      # O2 from the model default
    r = colmode.run_case(c, OracleBackend(oracle))
    assert abs(r["uflx"][0] - g["uflx"][0]) <= 1.5e-3     # 406.2213 W m-2: Planck tables at tbound = 290.93 K
    assert r["dflx"][-1] == 0.0
    if os.path.isdir(REF_EX):
        rd = lambda f: open(os.path.join(REF_EX, f)).read()  # noqa: E731
        wet = colmode.parse_input_rrtm(rd("input_rrtm_ICRCCM_sonde"), airmwt=28.964)
        h2o = np.array(wet.wkl)[:, 0]
        pwv_cm = h2o.sum() * 18.015 / 6.02214199e23      # g cm-2 = cm of precipitable water
        assert 1.0 < pwv_cm < 6.0                         # a humid mid-latitude summer sounding once AIRMWT is set
        assert np.array_equal(np.array(wet.pz), pz)


def test_rrtatm_pieces():
    """Pieces of the rrtatm restatement on their own: DEFALT's four-point Lagrange interpolation returns a model level
    exactly at that level and stays monotone in pressure between levels; WATVAP / CONVRT give the same number density for
    the same humidity written in different units (ppmv, g/kg with AIRMWT set, mb partial pressure, dew point); modes of
    LBLATM outside the restated subset are refused by name."""
    from rrtmg_lw_b200 import rrtatm
    A = rrtatm._us_standard()
    for k in (0, 7, 25, 49):
        w = [0.0] * 7
        p, t = rrtatm._defalt(A["alt_km"][k], 6, 6, [6] * 7, 0.0, 0.0, w)
        assert abs(p / A["p_mb"][k] - 1.0) < 1e-12 and abs(t - A["t_k"][k]) < 1e-9
        assert all(abs(w[m] - A["amol_ppmv"][m][k]) <= 1e-9 * max(1.0, abs(w[m])) for m in range(7))
    ps = [rrtatm._defalt(z, 6, 6, [10] * 7, 0.0, 0.0, [0.0] * 7)[0] for z in np.linspace(0.0, 100.0, 401)]
    assert all(b < a for a, b in zip(ps, ps[1:]))
    with pytest.raises(ValueError):
        rrtatm._defalt(121.0, 6, 6, [6] * 7, 0.0, 0.0, [0.0] * 7)
    # one humidity, four spellings: 10 000 ppmv at 900 mb, 285 K
    p, t = 900.0, 285.0
    rho = rrtatm.ALOSMT * (p / rrtatm.PZERO) * (rrtatm.TZERO / t)
    vmr = 1.0e-2
    d_ppmv = rrtatm._watvap(p, t, 10, vmr * 1e6, rho, 28.964)
    d_gkg = rrtatm._watvap(p, t, 12, vmr * 18.015 / 28.964 * 1e3, rho, 28.964)
    e_mb = d_ppmv / rho * p                                    # partial pressure of that density
    d_mb = rrtatm._watvap(p, t, 14, e_mb, rho, 28.964)
    assert abs(d_gkg / d_ppmv - 1.0) < 1e-12 and abs(d_mb / d_ppmv - 1.0) < 1e-12
    assert rrtatm._watvap(p, t, 12, 5.0, rho, 0.0) == 0.0     # trap T18: AIRMWT = 0 zeroes a g/kg profile
    sat = rrtatm._watvap(p, t, 17, 100.0, rho, 28.964)         # 100 % relative humidity
    assert abs(rrtatm._watvap(p, t, 15, t, rho, 28.964) / sat - 1.0) < 1e-12       # dew point = temperature
    assert 0.5 * sat < rrtatm._watvap(p, t, 17, 60.0, rho, 28.964) / 0.6 < 1.5 * sat
    assert 3.0e17 < sat < 4.5e17                               # e_s(285 K) ~ 14 hPa -> 3.6e17 cm-3
    rec31 = "    1    2   32    1    1    7    0"
    with pytest.raises(NotImplementedError):
        rrtatm.layer_atmosphere([rec31], 0)                    # MODEL = 1: model atmospheres are not restated
    with pytest.raises(NotImplementedError):
        rrtatm.layer_atmosphere(["    0    2    0    1    1    7    0"], 0)   # IBMAX = 0: automatic layering
