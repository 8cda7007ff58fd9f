"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer work bit-exact (interpolation indices, table indices,
McICA masks); FP64 fluxes within 1e-6 W m-2 and heating rates within 1e-6 K day-1.
"""
import numpy as np
import pytest
import torch

from rrtmg_lw_b200 import driver, synth

pytestmark = pytest.mark.gpu

FLUX_TOL = 1e-6   # W m-2
HR_TOL = 1e-6     # K day-1
KEYS = ("uflx", "dflx", "uflxc", "dflxc", "hr", "hrc")


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else x


def _check(got, ref, idrv=0, what=""):
    keys = KEYS + (("duflx_dt", "duflxc_dt") if idrv else ())
    for k in keys:
        tol = HR_TOL if k.startswith("hr") else FLUX_TOL
        err = float(np.abs(_np(got[k]) - ref[k]).max())
        assert err <= tol, (what, k, err)


def test_clear_sky_batch_and_stage_ints(gpu, oracle):
    """Config 2 shape (clear, idrv=0, 72 layers) at a size the oracle does in seconds."""
    b = synth.make_batch(2048, 72)
    ref = oracle.rrtmg_lw(b, nthreads=8)
    gpu.debug_capture(1234)
    got = driver.run_clear_or_strict(b, device="cuda:0")
    gpu.sync()
    _check(got, ref, what="clear")
    assert np.array_equal(_np(got["uflx"]), _np(got["uflxc"]))
    s = oracle.column_stages(b, 1234)
    d = gpu.debug_stages(1234, 72)
    gpu.debug_capture(None)
    assert np.array_equal(s["ints6"], d["ints6"]), "jp/jt/jt1/indself/indfor/indminor must be bit-exact"
    assert s["laytrop"] == d["laytrop"]
    assert np.array_equal(s["itr"], d["itr"]), "rtrnmc table indices must be bit-exact"
    assert np.array_equal(s["fracs"], d["fracs"])
    rel = np.abs(s["taug"] - d["taug"]) / np.maximum(np.abs(s["taug"]), 1e-300)
    assert rel.max() < 1e-13   # only libm log/exp/pow rounding may differ
    assert (s["taug"] == d["taug"]).mean() > 0.9


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_mcica_mask_bit_exact(gpu, oracle, icld):
    b = synth.make_batch(300, 72, cloudy=True)
    ref = oracle.mcica_subcol_lw(b, icld, permuteseed=7, irng=0)
    got = driver.run_mcica_subcol(b, icld, permuteseed=7, device="cuda:0")
    for k in ("cldfmcl", "ciwpmcl", "clwpmcl", "taucmcl", "reicmcl", "relqmcl"):
        assert np.array_equal(_np(got[k]), ref[k]), (icld, k)
    assert ref["cldfmcl"].sum() > 0


@pytest.mark.parametrize("icld", [4, 5])
def test_exponential_overlap_masks_and_alpha(gpu, oracle, icld):
    b = synth.make_batch(200, 72, cloudy=True)
    al_ref = oracle.get_alpha(b, icld, idcor=1, juldat=200)
    al = np.zeros_like(al_ref)
    gpu.get_alpha(1, b.ncol, b.nlay, icld, 1, 2500.0, b.dz.numpy().copy(), b.lat.numpy().copy(), 200,
                  b.cldfrac.numpy().copy(), al)
    assert np.abs(al - al_ref).max() < 1e-15   # exp() rounding only
    ref = oracle.mcica_subcol_lw(b, icld, permuteseed=2, irng=0, alpha=al)
    got = driver.run_mcica_subcol(b, icld, permuteseed=2, alpha=al, device="cuda:0")
    assert np.array_equal(_np(got["cldfmcl"]), ref["cldfmcl"])


@pytest.mark.parametrize("icld,idrv,aer", [(2, 0, False), (1, 1, True), (3, 0, True), (5, 1, True)])
def test_fused_cloudy_matches_oracle(gpu, oracle, icld, idrv, aer):
    """Configs 3/4 shapes: generator + radiation fused on the device vs mcica_subcol_lw ->
    rrtmg_lw in the oracle (icld 5 with idcor = 1 alpha, aerosol tau, idrv = 1)."""
    b = synth.make_batch(512, 72, cloudy=True, aerosol=aer)
    alpha = oracle.get_alpha(b, icld, idcor=1, juldat=200) if icld >= 4 else None
    mc = oracle.mcica_subcol_lw(b, icld, permuteseed=1, irng=0, alpha=alpha)
    ref = oracle.rrtmg_lw(b, icld=icld, idrv=idrv, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    assert ref["rc"] == 0
    got = driver.run_fused(b, icld=icld, idrv=idrv, alpha=alpha, device="cuda:0")
    gpu.sync()
    _check(got, ref, idrv, what="fused icld=%d" % icld)
    # clouds really matter in this batch
    assert np.abs(ref["dflx"][0] - ref["dflxc"][0]).max() > 1.0


@pytest.mark.parametrize("iceflag,liqflag", [(0, 0), (1, 1), (2, 1), (3, 1)])
def test_strict_rrtmg_lw_with_mcica_arrays(gpu, oracle, iceflag, liqflag):
    """The literal rrtmg_lw() boundary: (140,ncol,nlay) sub-column arrays in, all ice/liquid options."""
    b = synth.make_batch(192, 72, cloudy=True, aerosol=True)
    b.rei = b.rei.clamp(15.0, 125.0)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=3, irng=0)
    ref = oracle.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=iceflag, liqflglw=liqflag, mc=mc, nthreads=8)
    assert ref["rc"] == 0
    got = driver.run_clear_or_strict(b, icld=2, idrv=1, inflglw=2, iceflglw=iceflag, liqflglw=liqflag, mc=mc,
                                     device="cuda:0")
    gpu.sync()
    _check(got, ref, 1, what="strict ice=%d" % iceflag)
    # host-pointer call gives the same numbers as the device-pointer call
    goth = driver.run_clear_or_strict(b, icld=2, idrv=1, inflglw=2, iceflglw=iceflag, liqflglw=liqflag, mc=mc)
    for k in KEYS:
        assert np.array_equal(goth[k], _np(got[k])), k


def test_strict_general_cloud_arrays(gpu, oracle):
    """rrtmg_lw() accepts arbitrary per-g cloud arrays: fractional cldfmcl (only == 1 counts as
    cloudy in rtrnmc, T9), per-g water paths, inflag = 0 with prescribed tau."""
    b = synth.make_batch(128, 40, cloudy=True)
    rng = np.random.default_rng(5)
    shp = (b.nlay, b.ncol, 140)
    cf = np.where(rng.random(shp) < 0.15, 1.0, np.where(rng.random(shp) < 0.1, rng.random(shp), 0.0))
    mc = dict(cldfmcl=cf, taucmcl=rng.random(shp) * 3 * (cf > 0), ciwpmcl=rng.random(shp) * 20 * (cf > 0),
              clwpmcl=rng.random(shp) * 40 * (cf > 0), reicmcl=30 + 50 * rng.random(shp[:2]),
              relqmcl=5 + 10 * rng.random(shp[:2]))
    for inflag in (0, 2):
        ref = oracle.rrtmg_lw(b, icld=1, inflglw=inflag, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
        assert ref["rc"] == 0
        got = driver.run_clear_or_strict(b, icld=1, inflglw=inflag, iceflglw=3, liqflglw=1, mc=mc, device="cuda:0")
        gpu.sync()
        _check(got, ref, what="general inflag=%d" % inflag)


def test_inflag0_fused_prescribed_tau(gpu, oracle):
    b = synth.make_batch(128, 72, cloudy=True)
    b.tauc = (2.0 * torch.rand(b.nlay, b.ncol, 16, dtype=torch.float64)) * (b.cldfrac > 0)[..., None]
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    ref = oracle.rrtmg_lw(b, icld=2, inflglw=0, mc=mc, nthreads=8)
    got = driver.run_fused(b, icld=2, inflglw=0, iceflglw=0, liqflglw=0, device="cuda:0")
    gpu.sync()
    _check(got, ref, what="inflag0")


def test_std_atm_profiles_and_planck_known_answer(gpu, oracle, std_atm):
    """Config 1: the reference's own example profiles (51-60 layers) through the same ABI; the
    surface upward flux reproduces the shipped output files (data-independent known answer)."""
    for case, d in std_atm.items():
        b = synth.batch_from_profile(d["input"], ncol=3)
        ref = oracle.rrtmg_lw(b)
        got = driver.run_clear_or_strict(b)   # host-pointer call
        _check(got, ref, what=case)
        assert abs(got["uflx"][0, 0] - d["output"][-1][2]) < 1.5e-3


def test_137_layers_and_stress_gases(gpu, oracle):
    """Config 5 layer count; CO2 3x / N2O 1.9x the reference mixing ratios exercise the pow() paths
    (taumol.f90:550,736,1355,...)."""
    b = synth.make_batch(256, 137, cloudy=True, stress_co2=True)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    ref = oracle.rrtmg_lw(b, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    got = driver.run_fused(b, icld=2, device="cuda:0")
    gpu.sync()
    _check(got, ref, what="137 layers")


def test_chunking_and_leading_dimension_invariance(gpu, oracle):
    """Results do not depend on the internal chunk size nor on ld > ncol (GCM pcols chunks)."""
    b = synth.make_batch(500, 72, cloudy=True)
    a = driver.run_fused(b, icld=2, device="cuda:0")
    gpu.set_chunk(96)
    try:
        c = driver.run_fused(b, icld=2, device="cuda:0")
        h = driver.run_fused(b, icld=2)
    finally:
        gpu.set_chunk(262144)
    for k in KEYS:
        assert torch.equal(a[k], c[k]), k
        assert np.array_equal(_np(a[k]), h[k]), k
    # ld > ncol: run on the first 321 columns of the 500-column arrays
    from rrtmg_lw_b200 import api
    conv = driver._conv(b, "cuda:0")
    at = driver.atm_args(b, conv)
    cl = [conv(x) for x in (b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc)]
    out = driver._outputs(b.nlay, b.ncol, like=at[0])
    api.rrtmg_lw_mcica(321, b.nlay, 2, 0, 1, 0, *at, 2, 3, 1, *cl, None, conv(b.tauaer), out["uflx"], out["dflx"],
                       out["hr"], out["uflxc"], out["dflxc"], out["hrc"], None, None, ld=500)
    gpu.sync()
    for k in KEYS:
        assert torch.equal(out[k][:, :321], a[k][:, :321]), k
        assert float(out[k][:, 321:].abs().max()) == 0.0


def test_error_codes_match_reference_stops(gpu):
    from rrtmg_lw_b200 import api
    b = synth.make_batch(64, 30, cloudy=True)
    b.cldfrac[:] = 0.5; b.ciwp[:] = 5.0; b.clwp[:] = 5.0
    b.rei[:] = 150.0
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_fused(b, icld=1)
    assert e.value.code == 2
    b.rei[:] = 50.0; b.rel[:] = 70.0
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_fused(b, icld=1)
    assert e.value.code == 3
    b.rel[:] = 10.0
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_fused(b, icld=1, inflglw=1)
    assert e.value.code == 1
    b.play[0], b.play[1] = b.play[1].clone(), b.play[0].clone()
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_fused(b, icld=1)
    assert e.value.code == 4
    with pytest.raises(api.RrtmgLwError) as e:
        driver.run_fused(synth.make_batch(8, 30, cloudy=True), icld=9)
    assert e.value.code == 1


def test_default_real_kind_8(gpu, oracle, kblob):
    """Semantics of the shipped -r8 makefiles: both sides switch their default-kind literals."""
    from oracle import orc
    from rrtmg_lw_b200 import api
    b = synth.make_batch(256, 72, cloudy=True)
    try:
        orc.ini(kblob, 1004.0, 8)
        api.rrtmg_lw_ini(1004.0, kblob, 8)
        assert np.array_equal(api.get_table("tau_tbl"), orc.get_table("tau_tbl"))
        mc = orc.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
        ref8 = orc.rrtmg_lw(b, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
        got8 = driver.run_fused(b, icld=2, device="cuda:0")
        api.sync()
        _check(got8, ref8, what="drk8")
    finally:
        orc.ini(kblob, 1004.0, 4)
        api.rrtmg_lw_ini(1004.0, kblob, 4)
    ref4 = orc.rrtmg_lw(b, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    # the two Fortran builds differ by more than the parity tolerance
    assert np.abs(ref4["uflx"] - ref8["uflx"]).max() > 1e-6


def test_init_tables_bit_identical(gpu, oracle):
    ngc = [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2]
    for n in ("tau_tbl", "exp_tbl", "tfn_tbl", "rwgt"):
        assert np.array_equal(gpu.get_table(n), oracle.get_table(n)), n
    from rrtmg_lw_b200 import kdata
    for band, name, lead, gfirst, _ in kdata.karrays():
        red = ("absa" if name == "kao" else "absb" if name == "kbo" else
               "ka" + name[3:] if name.startswith("kao") else "kb" + name[3:] if name.startswith("kbo") else name[:-1])
        a, o = gpu.get_table("%d:%s" % (band, red)), oracle.get_table("%d:%s" % (band, red))
        assert a.size == lead * ngc[band - 1] and np.array_equal(a, o), (band, red)


def test_full_size_column_independence(gpu, oracle):
    """At the bench's full size (BASELINE.json C3: 1 048 576 columns x 72 layers, two 524 288-column passes) every
    column's result equals the result of the same column run through the oracle alone (columns are independent): a
    random sample of 2048 columns, from both passes, is checked to the FP64 bars."""
    n = 1048576
    ns = 2048
    b = synth.make_batch(n, 72, cloudy=True, device="cuda:0")
    gpu.set_chunk(524288)
    try:
        got = driver.run_fused(b, icld=2, device="cuda:0")
        gpu.sync()
    finally:
        gpu.set_chunk(262144)
        gpu.set_option("release_scratch", 1)
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:ns]
    assert int((idx < 524288).sum()) > 500 and int((idx >= 524288).sum()) > 500
    sub = synth.Batch(ncol=ns, nlay=72, **{k: None for k in ()},
                      play=b.play[:, idx].cpu(), plev=b.plev[:, idx].cpu(), tlay=b.tlay[:, idx].cpu(),
                      tlev=b.tlev[:, idx].cpu(), tsfc=b.tsfc[idx].cpu(), h2ovmr=b.h2ovmr[:, idx].cpu(),
                      o3vmr=b.o3vmr[:, idx].cpu(), co2vmr=b.co2vmr[:, idx].cpu(), ch4vmr=b.ch4vmr[:, idx].cpu(),
                      n2ovmr=b.n2ovmr[:, idx].cpu(), o2vmr=b.o2vmr[:, idx].cpu(), cfc11vmr=b.cfc11vmr[:, idx].cpu(),
                      cfc12vmr=b.cfc12vmr[:, idx].cpu(), cfc22vmr=b.cfc22vmr[:, idx].cpu(),
                      ccl4vmr=b.ccl4vmr[:, idx].cpu(), emis=b.emis[:, idx].cpu(), tauaer=b.tauaer[:, :, idx].cpu(),
                      cldfrac=b.cldfrac[:, idx].cpu(), ciwp=b.ciwp[:, idx].cpu(), clwp=b.clwp[:, idx].cpu(),
                      rei=b.rei[:, idx].cpu(), rel=b.rel[:, idx].cpu(), tauc=b.tauc[:, idx].cpu())
    sub = synth.Batch(**{k: (v.contiguous() if hasattr(v, "contiguous") else v) for k, v in sub.__dict__.items()})
    mc = oracle.mcica_subcol_lw(sub, 2, permuteseed=1, irng=0)
    ref = oracle.rrtmg_lw(sub, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    for k in KEYS:
        tol = HR_TOL if k.startswith("hr") else FLUX_TOL
        err = float(np.abs(got[k][:, idx.to("cuda:0")].cpu().numpy() - ref[k]).max())
        assert err <= tol, (k, err)
    del b, got
    torch.cuda.empty_cache()


@pytest.mark.parametrize("kernel", [2, 4])
def test_both_band_kernel_mappings_match_oracle(gpu, oracle, kernel):
    """Both kernel structures (set_option("band_kernel", 4 = two-phase taumol + sweep kernels, 2 = fused per-band
    kernels) reproduce the oracle: clear sky
    with idrv = 1 and a McICA cloudy batch with aerosol."""
    try:
        gpu.set_option("band_kernel", kernel)
        b = synth.make_batch(384, 72, aerosol=True)
        ref = oracle.rrtmg_lw(b, idrv=1, nthreads=8)
        got = driver.run_clear_or_strict(b, idrv=1, device="cuda:0")
        gpu.sync()
        _check(got, ref, 1, what="kernel %d clear" % kernel)
        b = synth.make_batch(384, 72, cloudy=True, aerosol=True)
        mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
        ref = oracle.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
        got = driver.run_fused(b, icld=2, idrv=1, device="cuda:0")
        gpu.sync()
        _check(got, ref, 1, what="kernel %d cloudy" % kernel)
    finally:
        import os
        gpu.set_option("band_kernel", int(os.environ.get("RRTMG_LW_BAND_KERNEL", "4")))


@pytest.mark.parametrize("icld", [1, 2, 3, 4, 5])
def test_mersenne_twister_stream_masks_bit_exact(gpu, oracle, icld):
    """irng = 1: ONE global MT19937 stream over (sub-column, column, layer)
    (src/mcica_subcol_gen_lw.f90:523-530, 586-595; getRandomReal semantics of
    src/mcica_random_numbers.f90:276-295).  Device pointers, host pointers, and a column chunk
    smaller than the call (the words between a chunk's runs are stepped over) all reproduce the oracle."""
    b = synth.make_batch(300, 72, cloudy=True)
    al = oracle.get_alpha(b, icld, idcor=1, juldat=200) if icld >= 4 else None
    ref = oracle.mcica_subcol_lw(b, icld, permuteseed=11, irng=1, alpha=al)
    assert ref["cldfmcl"].sum() > 0
    got = driver.run_mcica_subcol(b, icld, permuteseed=11, irng=1, alpha=al, device="cuda:0")
    assert got["irng"] == 1
    for k in ("cldfmcl", "ciwpmcl", "clwpmcl", "taucmcl", "reicmcl", "relqmcl"):
        assert np.array_equal(_np(got[k]), ref[k]), (icld, k)
    try:
        gpu.set_chunk(96)   # 300 columns = 4 passes; every pass regenerates the stream up to its runs
        got = driver.run_mcica_subcol(b, icld, permuteseed=11, irng=1, alpha=al, device="cuda:0")
        assert np.array_equal(_np(got["cldfmcl"]), ref["cldfmcl"]), (icld, "chunked device")
        got = driver.run_mcica_subcol(b, icld, permuteseed=11, irng=1, alpha=al)
        assert np.array_equal(got["cldfmcl"], ref["cldfmcl"]), (icld, "chunked host")
    finally:
        gpu.set_chunk(262144)


def test_mersenne_twister_fused_matches_oracle(gpu, oracle):
    """mcica_subcol_lw(irng=1) -> rrtmg_lw in one pass, device and staged host calls (host passes of 128
    columns keep their place in the global stream)."""
    b = synth.make_batch(320, 72, cloudy=True, aerosol=True)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=3, irng=1)
    ref = oracle.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    assert ref["rc"] == 0
    got = driver.run_fused(b, icld=2, idrv=1, permuteseed=3, irng=1, device="cuda:0")
    gpu.sync()
    _check(got, ref, 1, what="MT fused device")
    try:
        gpu.set_option("host_chunk", 128)
        got = driver.run_fused(b, icld=2, idrv=1, permuteseed=3, irng=1)
        _check(got, ref, 1, what="MT fused host, 3 staged passes")
    finally:
        gpu.set_option("host_chunk", 32768)


@pytest.mark.parametrize("icld", [2, 5])
def test_mersenne_twister_blocks_of_a_sharded_batch(gpu, oracle, icld):
    """with set_mt_block(global_ncol, offset) a call on a column block draws exactly the words the
    unsharded call would draw for those columns (the other blocks' words are stepped over) -- masks bit-equal to the
    oracle's one-stream run over the whole batch, fused fluxes within tolerance, for an uneven 3-way split, on device
    and host pointers, also when the library cuts the block into passes."""
    from rrtmg_lw_b200 import sharding
    b = synth.make_batch(301, 72, cloudy=True)
    al = oracle.get_alpha(b, icld, idcor=1, juldat=200) if icld >= 4 else None
    mc = oracle.mcica_subcol_lw(b, icld, permuteseed=9, irng=1, alpha=al)
    ref = oracle.rrtmg_lw(b, icld=icld, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    assert ref["rc"] == 0
    try:
        for world, chunk, device in ((3, 262144, "cuda:0"), (2, 64, "cuda:0"), (3, 262144, None)):
            gpu.set_chunk(chunk)
            masks, flux = [], {k: [] for k in KEYS}
            for rank in range(world):
                lo, hi = sharding.block_range(b.ncol, world, rank)
                sh = sharding.shard_batch(b, world, rank)
                a = None if al is None else np.ascontiguousarray(al[:, lo:hi])
                gpu.set_mt_block(b.ncol, lo)
                m = driver.run_mcica_subcol(sh, icld, permuteseed=9, irng=1, alpha=a, device=device)
                masks.append(_np(m["cldfmcl"]))
                o = driver.run_fused(sh, icld=icld, permuteseed=9, irng=1, alpha=a, device=device)
                gpu.sync()
                for k in KEYS:
                    flux[k].append(_np(o[k]))
            assert np.array_equal(np.concatenate(masks, axis=1), mc["cldfmcl"]), (icld, world, chunk, device)
            _check({k: np.concatenate(v, axis=1) for k, v in flux.items()}, ref, what="MT blocks %s" % ((world, chunk, device),))
        # a block that does not fit the declared batch is refused
        gpu.set_mt_block(100, 50)
        with pytest.raises(Exception):
            driver.run_fused(sharding.shard_batch(b, 2, 0), icld=icld, permuteseed=9, irng=1,
                             alpha=None if al is None else np.ascontiguousarray(al[:, :151]), device="cuda:0")
    finally:
        gpu.set_mt_block(0)
        gpu.set_chunk(262144)
    # switched off: a call is its own stream again
    got = driver.run_mcica_subcol(b, icld, permuteseed=9, irng=1, alpha=al, device="cuda:0")
    assert np.array_equal(_np(got["cldfmcl"]), mc["cldfmcl"])


def test_fp32_mode_within_north_star_tolerance(gpu, oracle):
    """Optional FP32 mode (set_option("precision", 32)): the float band kernels (same source as the FP64 ones,
    tools/gen_fp32.py) must stay within 0.05 W m-2 / 0.01 K day-1 of the FP64 oracle (BASELINE.json north_star);
    setcoef indices and McICA masks are produced by the unchanged FP64 / integer kernels."""
    try:
        gpu.set_option("precision", 32)
        worst = {}
        for name, b, kw in (
            ("clear", synth.make_batch(1024, 72, aerosol=True), None),
            ("cloudy", synth.make_batch(1024, 72, cloudy=True, aerosol=True), dict(icld=2)),
            ("cloudy137", synth.make_batch(256, 137, cloudy=True), dict(icld=1)),
        ):
            if kw is None:
                ref = oracle.rrtmg_lw(b, idrv=1, nthreads=8)
                got = driver.run_clear_or_strict(b, idrv=1, device="cuda:0")
            else:
                mc = oracle.mcica_subcol_lw(b, kw["icld"], permuteseed=1, irng=0)
                ref = oracle.rrtmg_lw(b, icld=kw["icld"], idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
                got = driver.run_fused(b, icld=kw["icld"], idrv=1, device="cuda:0")
            gpu.sync()
            assert ref["rc"] == 0
            for k in KEYS + ("duflx_dt", "duflxc_dt"):
                err = float(np.abs(_np(got[k]) - ref[k]).max())
                worst[(name, k)] = err
                tol = 0.01 if k.startswith("hr") else 0.05
                assert err <= tol, (name, k, err)
        # host pointers (staged passes) and the Mersenne-Twister stream go through the same float kernels
        b = synth.make_batch(320, 72, cloudy=True, aerosol=True)
        mc = oracle.mcica_subcol_lw(b, 2, permuteseed=3, irng=1)
        ref = oracle.rrtmg_lw(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
        gpu.set_option("host_chunk", 128)
        got = driver.run_fused(b, icld=2, idrv=0, permuteseed=3, irng=1)
        for k in KEYS:
            err = float(np.abs(_np(got[k]) - ref[k]).max())
            assert err <= (0.01 if k.startswith("hr") else 0.05), ("host+MT", k, err)
        # column passes inside one device-pointer call (first column of a pass != 0) do not change the result
        gpu.set_option("host_chunk", 32768)
        one = driver.run_fused(b, icld=2, idrv=0, device="cuda:0")
        gpu.sync()
        gpu.set_chunk(96)
        many = driver.run_fused(b, icld=2, idrv=0, device="cuda:0")
        gpu.sync()
        for k in KEYS:
            assert np.array_equal(_np(one[k]), _np(many[k])), ("fp32 chunking", k)
        print("fp32 worst errors:", {k: "%.2e" % v for k, v in worst.items()})
    finally:
        gpu.set_option("precision", 64)
        gpu.set_option("host_chunk", 32768)
        gpu.set_chunk(262144)


@pytest.mark.parametrize("ncol,nlay", [(1, 72), (33, 51), (129, 4), (7, 200)])
def test_ragged_and_extreme_sizes(gpu, oracle, ncol, nlay):
    """One column, column counts that leave partial warps / CTAs, the 4-layer minimum of the KISS seeds and
    a deep 200-layer grid (parrrtm mxlay = 603): fused McICA entry with aerosol and dF/dTs, device and host."""
    b = synth.make_batch(ncol, nlay, cloudy=True, aerosol=True)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    ref = oracle.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=4)
    assert ref["rc"] == 0
    got = driver.run_fused(b, icld=2, idrv=1, device="cuda:0")
    gpu.sync()
    _check(got, ref, 1, what="device %dx%d" % (ncol, nlay))
    got = driver.run_fused(b, icld=2, idrv=1)
    _check(got, ref, 1, what="host %dx%d" % (ncol, nlay))
    gm = driver.run_mcica_subcol(b, 2, device="cuda:0")
    assert np.array_equal(_np(gm["cldfmcl"]), mc["cldfmcl"])


def test_empty_and_invalid_calls_are_refused(gpu):
    """ncol = 0, ld < ncol, nlay = 0 or above the compiled maximum: RRTMG_LW_ERR_ARG, nothing is launched."""
    from rrtmg_lw_b200 import api
    b = synth.make_batch(8, 72, cloudy=True)
    c = driver._conv(b, None)
    at = driver.atm_args(b, c)
    cl = [c(x) for x in (b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc)]
    o = driver._outputs(72, 8)
    outs = [o[k] for k in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")]
    for ncol, nlay in ((0, 72), (8, 0), (8, 100000)):
        with pytest.raises(api.RrtmgLwError) as e:
            api.rrtmg_lw_mcica(ncol, nlay, 2, 0, 1, 0, *at, 2, 3, 1, *cl, None, c(b.tauaer), *outs, None, None)
        assert e.value.code == 7


# ---------------------------------------------------------------------------------------------------------------
# integer bit-exactness on many columns and branch coverage of the batches the CUDA path is compared on
COVERAGE_BATCHES = {
    "cloudy72": dict(ncol=256, nlay=72, cloudy=True, aerosol=True),
    "cloudy137": dict(ncol=96, nlay=137, cloudy=True),
    "stress72": dict(ncol=256, nlay=72, cloudy=True, stress_co2=True),
    "spread72": dict(ncol=256, nlay=72, cloudy=True, gas_spread=1.2),
}


@pytest.mark.parametrize("name", sorted(COVERAGE_BATCHES))
def test_stage_dumps_bit_exact_on_32_columns(gpu, oracle, name):
    """jp / jt / jt1 / indself / indfor / indminor, laytrop, the McICA mask and the rtrnmc table indices itr / itgas /
    ittot (rtrnmc.f90:386, 404, 412, 435) of cloudy layers, bit for bit, on 32 columns spread over a cloudy 72-layer,
    a cloudy 137-layer, a stress-gas (pow paths) and a gas-spread (all eta branches) batch; Planck fractions equal,
    gas optical depths to libm rounding."""
    kw = dict(COVERAGE_BATCHES[name])
    b = synth.make_batch(**kw)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    cols = np.unique(np.linspace(0, b.ncol - 1, 32).astype(int))
    assert len(cols) == 32
    cloudy_layers_seen = 0
    try:
        for col in cols:
            col = int(col)
            gpu.debug_capture(col)
            driver.run_fused(b, icld=2, idrv=0, device="cuda:0")
            gpu.sync()
            d = gpu.debug_stages(col, b.nlay, want_mask=True)
            s = oracle.column_stages(b, col, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc)
            assert s["rc"] == 0
            assert np.array_equal(s["ints6"], d["ints6"]), (name, col, "setcoef integers")
            assert s["laytrop"] == d["laytrop"], (name, col)
            mask_ref = (mc["cldfmcl"][:, col, :] == 1.0).T.astype(np.uint8)       # (140, nlay)
            assert np.array_equal(mask_ref, d["mask"]), (name, col, "McICA mask")
            if not np.array_equal(s["itr"], d["itr"]):
                w = np.argwhere(s["itr"] != d["itr"])
                g_, l_, k_ = (int(x) for x in w[0])
                raise AssertionError((name, col, "itr / itgas / ittot", len(w), "first at (g, layer, which)", (g_, l_, k_),
                                      "oracle", s["itr"][g_, l_].tolist(), "gpu", d["itr"][g_, l_].tolist(),
                                      "cloud bit", int(mask_ref[g_, l_]), "layer cloudy", int(mask_ref[:, l_].max())))
            assert np.array_equal(s["fracs"], d["fracs"]), (name, col, "Planck fractions")
            rel = np.abs(s["taug"] - d["taug"]) / np.maximum(np.abs(s["taug"]), 1e-300)
            assert rel.max() < 1e-12, (name, col, rel.max())   # libm log / pow rounding only
            cloudy_layers_seen += int((mask_ref.sum(0) > 0).sum())
            assert (s["itr"][:, :, 1] >= 0).sum() > 0 or mask_ref.sum() == 0
    finally:
        gpu.debug_capture(None)
    assert cloudy_layers_seen >= 100   # ittot really was compared in cloudy layers


def test_parity_batches_cover_every_branch(gpu, oracle):
    """The batches of this suite, run through the oracle with its branch counters armed and through the CUDA path
    (fluxes within 1e-6 / heating rates within 1e-6), exercise every data-dependent branch of the reference >= 100
    times: the three eta branches of both (p, T) corners of every binary-species band (taumol.f90:569-598), both
    atmosphere halves of every band, all ten adjfac sites (:550, 736, 1355, 1483, 1639, 1721, 1768, 1879, 2036, 2497),
    the three cloudy-layer sub-branches (rtrnmc.f90:358-425) and both clear-layer branches (:429-440)."""
    oracle.branch_arm(True)
    oracle.branch_counts(reset=True)
    for name in sorted(COVERAGE_BATCHES):
        kw = dict(COVERAGE_BATCHES[name])
        b = synth.make_batch(**kw)
        mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
        ref = oracle.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
        assert ref["rc"] == 0
        got = driver.run_fused(b, icld=2, idrv=1, device="cuda:0")
        gpu.sync()
        _check(got, ref, 1, what="coverage batch " + name)
    c = oracle.branch_counts(reset=True)
    oracle.branch_arm(False)
    for band in oracle.BINARY_LOWER_BANDS:
        assert (c["eta"][band - 1] >= 100).all(), ("eta branches of band", band, c["eta"][band - 1].tolist())
    assert (c["half"] >= 100).all(), c["half"].tolist()
    assert (c["adjfac"] >= 100).all(), dict(zip(oracle.BRANCH_ADJFAC_SITES, c["adjfac"].tolist()))
    assert (c["cloudy"] >= 100).all() and (c["clear"] >= 100).all(), (c["cloudy"], c["clear"])


@pytest.mark.parametrize("icld", [1, 2])
def test_kiss_stepping_over_clear_layers_is_bit_exact(gpu, oracle, icld):
    """The KISS generator stops at the highest layer with a cloud fraction and jumps over runs of cloud-free layers
    (closed-form n-step maps of the four component generators); the masks must equal the oracle's, which draws every
    number, and the library's own exhaustive mode -- on columns with clear runs of every length from 1 to > 100."""
    b = synth.make_batch(700, 137, cloudy=True, seed=77)
    g = torch.Generator().manual_seed(9)
    cf = torch.zeros(137, 700, dtype=torch.float64)
    for i in range(700):
        l = int(torch.randint(0, 4, (1,), generator=g))
        while l < 137:
            n = int(torch.randint(1, 4, (1,), generator=g))
            cf[l:l + n, i] = torch.rand(min(n, 137 - l), generator=g, dtype=torch.float64)
            l += n + 1 + (i % 120)  # clear run of 1 + i % 120 layers
    b.cldfrac[:] = cf
    ref = oracle.mcica_subcol_lw(b, icld, permuteseed=140, irng=0)
    got = driver.run_mcica_subcol(b, icld, permuteseed=140, device="cuda:0")
    gpu.set_option("kiss_sparse", 0)
    try:
        full = driver.run_mcica_subcol(b, icld, permuteseed=140, device="cuda:0")
    finally:
        gpu.set_option("kiss_sparse", 1)
    assert ref["cldfmcl"].sum() > 1000
    assert np.array_equal(_np(got["cldfmcl"]), ref["cldfmcl"])
    assert torch.equal(got["cldfmcl"], full["cldfmcl"])


def test_init_from_the_table_image_is_bit_identical(gpu, oracle, kblob):
    """rrtmg_lw_b200_ini_tables(image built on the host) leaves the library exactly where rrtmg_lw_b200_ini(k-data)
    does: every reduced table readable through get_table and a cloudy + dF/dT run are bit-identical."""
    from rrtmg_lw_b200 import api
    b = synth.make_batch(512, 72, cloudy=True, aerosol=True)
    ref = driver.run_fused(b, icld=2, idrv=1, device="cuda:0")
    names = ["tau_tbl", "exp_tbl", "tfn_tbl", "rwgt", "heatfac", "fluxfac", "3:absa", "3:absb", "16:fracrefa", "9:kb_mn2o"]
    tabs = {n: api.get_table(n) for n in names}
    try:
        api.rrtmg_lw_ini_tables(api.tables_build(1004.0, kblob, 4))
        got = driver.run_fused(b, icld=2, idrv=1, device="cuda:0")
        gpu.sync()
        for k in KEYS + ("duflx_dt", "duflxc_dt"):
            assert torch.equal(got[k], ref[k]), k
        for n in names:
            assert tabs[n].size > 0 and np.array_equal(api.get_table(n), tabs[n]), n
    finally:
        api.rrtmg_lw_ini(1004.0, kblob, 4)


@pytest.mark.parametrize("idrv", [0, 1])
def test_schedule_variants_are_bit_identical(gpu, oracle, idrv):
    """The scheduling switches change WHERE kernels run, never what they compute: serial (concurrency 0), masks beside
    taumol only (2), sweeps over streams only (3) and the resident McICA work-queue kernel with the cloud optics fused
    (mcica_beside 1 / 2, a second launch of the same kernel taking the items left) give the default schedule's fluxes
    bit for bit, on a ragged batch cut into three passes; and the default's agree with the oracle to 1e-6."""
    b = synth.make_batch(1300, 72, cloudy=True, aerosol=True, seed=21)
    keys = KEYS + (("duflx_dt", "duflxc_dt") if idrv else ())
    gpu.set_chunk(512)
    try:
        ref = driver.run_fused(b, icld=2, idrv=idrv, device="cuda:0")
        gpu.sync()
        ref = {k: ref[k].clone() for k in keys}
        for name, value in (("concurrency", 0), ("concurrency", 2), ("concurrency", 3), ("mcica_beside", 1), ("mcica_beside", 2)):
            gpu.set_option(name, value)
            try:
                got = driver.run_fused(b, icld=2, idrv=idrv, device="cuda:0")
                gpu.sync()
                for k in keys:
                    assert torch.equal(got[k], ref[k]), (name, value, k)
            finally:
                gpu.set_option(name, 1 if name == "concurrency" else 0)
    finally:
        gpu.set_chunk(262144)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    want = oracle.rrtmg_lw(b, icld=2, idrv=idrv, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
    assert want["rc"] == 0
    _check(ref, want, idrv, what="default schedule, three passes")
