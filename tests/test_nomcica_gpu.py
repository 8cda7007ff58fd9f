"""Parity of the non-McICA entry (rrtmg_lw of src/rrtmg_lw_rad.nomcica.f90: cldprop + rtrn / rtrnmr) against the
CPU oracle's restatement (oracle/orc_nomcica.cpp) on identical seeded inputs, through the C ABI.

Bars: FP64 fluxes within 1e-6 W m-2, heating rates within 1e-6 K day-1 (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

from rrtmg_lw_b200 import driver, synth

pytestmark = pytest.mark.gpu

FLUX_TOL = 1e-6
HR_TOL = 1e-6
KEYS = ("uflx", "dflx", "uflxc", "dflxc", "hr", "hrc")


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else x


def _check(got, ref, idrv=0, what=""):
    assert ref["rc"] == 0
    for k in KEYS + (("duflx_dt", "duflxc_dt") if idrv else ()):
        tol = HR_TOL if k.startswith("hr") else FLUX_TOL
        err = float(np.abs(_np(got[k]) - ref[k]).max())
        assert err <= tol, (what, k, err)


def overlap_batch(ncol=192, nlay=40, seed=5):
    """Cloud-fraction profiles built to walk every branch of the overlap-factor passes of rtrnmr (:345-505): blocks of
    adjacent cloudy layers with rising, falling and EQUAL fractions, overcast layers (cldfrac = 1), single-layer
    clouds, clouds in the top and the bottom layer, fractions below the 1e-6 cloud threshold."""
    b = synth.make_batch(ncol, nlay, cloudy=True, aerosol=True, seed=seed)
    g = torch.Generator().manual_seed(seed)
    cf = torch.zeros(nlay, ncol, dtype=torch.float64)
    for i in range(ncol):
        k = 0
        while k < nlay:
            if torch.rand(1, generator=g).item() < 0.45:
                n = int(torch.randint(1, 7, (1,), generator=g).item())
                vals = torch.rand(n, generator=g, dtype=torch.float64)
                kind = i % 6
                if kind == 0:
                    vals, _ = torch.sort(vals)
                elif kind == 1:
                    vals, _ = torch.sort(vals, descending=True)
                elif kind == 2:
                    vals = torch.full((n,), float(vals[0]), dtype=torch.float64)  # equal fractions
                elif kind == 3:
                    vals[n // 2] = 1.0  # an overcast layer inside the block
                elif kind == 4:
                    vals = torch.round(vals * 4) / 4 + 0.125  # repeated values, some > 1 clipped below
                    vals = vals.clamp(0.125, 1.0)
                cf[k:k + n, i] = vals[: max(0, min(n, nlay - k))]
                k += n
            else:
                k += int(torch.randint(1, 4, (1,), generator=g).item())
    cf[0, ::7] = 0.6   # cloud in the surface layer
    cf[-1, ::5] = 0.3  # cloud in the top layer
    cf[3, ::11] = 5e-7  # below the cloudy-layer threshold of rtrn / rtrnmr, above cldprop's
    b.cldfrac[:] = cf
    on = cf > 0
    b.ciwp[:] = torch.where(on, 2.0 + 40.0 * torch.rand(nlay, ncol, generator=g, dtype=torch.float64), torch.zeros(()).double())
    b.clwp[:] = torch.where(on, 1.0 + 60.0 * torch.rand(nlay, ncol, generator=g, dtype=torch.float64), torch.zeros(()).double())
    b.ciwp[:, ::9] = 0.0   # liquid-only columns
    b.clwp[:, ::13] = 0.0  # ice-only columns
    b.rei[:] = 15.0 + 100.0 * torch.rand(nlay, ncol, generator=g, dtype=torch.float64)
    b.rel[:] = 3.0 + 50.0 * torch.rand(nlay, ncol, generator=g, dtype=torch.float64)
    b.tauc[:] = torch.where(on[..., None], 3.0 * torch.rand(nlay, ncol, 16, generator=g, dtype=torch.float64),
                            torch.zeros(()).double())
    return b


@pytest.mark.parametrize("icld,idrv", [(1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (0, 1)])
def test_nomcica_overlap_branches_match_oracle(gpu, oracle, icld, idrv):
    b = overlap_batch()
    ref = oracle.rrtmg_lw_nomcica(b, icld=icld, idrv=idrv, nthreads=8)
    got = driver.run_nomcica(b, icld=icld, idrv=idrv, device="cuda:0")
    gpu.sync()
    _check(got, ref, idrv, what=(icld, idrv))
    if icld == 0:
        assert np.array_equal(_np(got["uflx"]), _np(got["uflxc"]))
    else:
        assert float(np.abs(_np(got["dflx"]) - _np(got["dflxc"])).max()) > 1.0  # the clouds are seen


@pytest.mark.parametrize("inflag,iceflag,liqflag", [(0, 0, 0), (2, 0, 0), (2, 1, 1), (2, 2, 1), (2, 3, 1), (2, 1, 0), (1, 0, 0)])
@pytest.mark.parametrize("icld", [1, 2])
def test_nomcica_cldprop_options(gpu, oracle, icld, inflag, iceflag, liqflag):
    """Every cloud-optics option of cldprop (:196-290), including the 5-band ice case that changes ncbands."""
    b = overlap_batch(ncol=96, nlay=33, seed=11)
    if iceflag == 1:
        b.rei[:] = b.rei.clamp(13.0, 130.0)
    ref = oracle.rrtmg_lw_nomcica(b, icld=icld, idrv=0, inflglw=inflag, iceflglw=iceflag, liqflglw=liqflag, nthreads=8)
    got = driver.run_nomcica(b, icld=icld, inflglw=inflag, iceflglw=iceflag, liqflglw=liqflag, device="cuda:0")
    gpu.sync()
    _check(got, ref, what=(icld, inflag, iceflag, liqflag))


def test_nomcica_host_call_chunks_and_clear_limit(gpu, oracle):
    """Host pointers (staged passes), several device chunks, and: with no cloud the entry equals the McICA-free clear run."""
    b = overlap_batch(ncol=300, nlay=25, seed=3)
    ref = oracle.rrtmg_lw_nomcica(b, icld=2, idrv=1, nthreads=8)
    gpu.set_chunk(128)
    try:
        got = driver.run_nomcica(b, icld=2, idrv=1)
    finally:
        gpu.set_chunk(262144)
    _check(got, ref, 1, what="host")
    dev = driver.run_nomcica(b, icld=2, idrv=1, device="cuda:0")
    gpu.sync()
    for k in KEYS:
        assert np.array_equal(_np(dev[k]), got[k]), k
    b.cldfrac[:] = 0.0
    gpu.set_option("tail_group", 0)  # one sweep kernel per band, like the non-McICA path: same summation order
    try:
        clear = driver.run_clear_or_strict(b, idrv=1, device="cuda:0")
    finally:
        gpu.set_option("tail_group", 1)
    none = driver.run_nomcica(b, icld=2, idrv=1, device="cuda:0")
    gpu.sync()
    for k in KEYS + ("duflx_dt", "duflxc_dt"):
        assert torch.equal(clear[k], none[k]), k


def test_nomcica_error_codes_and_icld_clamp(gpu):
    from rrtmg_lw_b200 import api
    b = synth.make_batch(64, 30, cloudy=True)
    b.cldfrac[:] = 0.5; b.ciwp[:] = 5.0; b.clwp[:] = 5.0
    b.rei[:] = 150.0
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_nomcica(b, icld=2)
    assert e.value.code == 2
    b.rei[:] = 50.0; b.rel[:] = 70.0
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_nomcica(b, icld=1)
    assert e.value.code == 3
    b.rel[:] = 10.0
    out = driver.run_nomcica(b, icld=7)
    assert out["icld"] == 2  # rad.nomcica.f90:458
