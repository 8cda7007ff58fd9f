"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- never imported by rrtmg_lw_b200).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Arrays are numpy float64 in the reference's Fortran layout seen from C
(``(nlay, ncol)`` row-major == Fortran ``(ncol, nlay)``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")
_STATIC = os.path.join(os.path.dirname(_HERE), "rrtmg_lw_b200", "data", "lw_static_tables.bin")

_lib = None
_plain = None      # liborc.so while the coverage build is switched in
_ini_args = None   # (blob, cpdair, default_real_kind) of the last ini(), replayed into the coverage build
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False):
    """(Re)build liborc.so / liborc_count.so when a source is newer than the library (make decides)."""
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(_LIB):   # no toolchain and no prebuilt library: nothing to check against
            raise


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_get_table.restype = C.c_long
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_dp)


def ini(kdata_blob: np.ndarray, cpdair: float = 1004.0, default_real_kind: int = 4):
    global _ini_args
    _ini_args = (kdata_blob, cpdair, default_real_kind)
    st = np.fromfile(_STATIC, dtype="<f8")
    rc = lib().orc_ini(C.c_double(cpdair), _p(st), _p(np.ascontiguousarray(kdata_blob)), C.c_int(default_real_kind))
    assert rc == 0, rc


def _np(t):
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float64)


def set_band_range(istart=1, iend=16):
    """Band range of the following runs (per-band output blocks of the column-mode program)."""
    rc = lib().orc_set_band_range(int(istart), int(iend))
    assert rc == 0, rc


def rrtmg_lw(b, icld=0, idrv=0, inflglw=0, iceflglw=0, liqflglw=0, mc=None, nthreads=1):
    """Run the oracle on a synth.Batch.  mc = dict(cldfmcl, taucmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl)
    with the (nlay, ncol, 140) / (nlay, ncol) layouts.  Returns dict of outputs + rc + icld."""
    ncol, nlay = b.ncol, b.nlay
    z3 = np.zeros((nlay, ncol, 140)) if (mc is None and icld > 0) else None
    z2 = np.zeros((nlay, ncol)) if (mc is None and icld > 0) else None
    m = mc or dict(cldfmcl=z3, taucmcl=z3, ciwpmcl=z3, clwpmcl=z3, reicmcl=z2, relqmcl=z2)
    if icld == 0 and mc is None:
        dummy3 = np.zeros((1,)); dummy2 = np.zeros((1,))
        m = dict(cldfmcl=dummy3, taucmcl=dummy3, ciwpmcl=dummy3, clwpmcl=dummy3, reicmcl=dummy2, relqmcl=dummy2)
    out = {k: np.zeros((nlay + 1, ncol)) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
    out["hr"] = np.zeros((nlay, ncol)); out["hrc"] = np.zeros((nlay, ncol))
    ic = C.c_int(icld)
    keep = [_np(x) for x in (b.play, b.plev, b.tlay, b.tlev, b.tsfc, b.h2ovmr, b.o3vmr, b.co2vmr, b.ch4vmr, b.n2ovmr,
                             b.o2vmr, b.cfc11vmr, b.cfc12vmr, b.cfc22vmr, b.ccl4vmr, b.emis)]
    mk = [_np(m[k]) for k in ("cldfmcl", "taucmcl", "ciwpmcl", "clwpmcl", "reicmcl", "relqmcl")]
    ta = _np(b.tauaer)
    rc = lib().orc_rrtmg_lw(ncol, ncol, nlay, C.byref(ic), idrv, *[_p(x) for x in keep], inflglw, iceflglw, liqflglw,
                            *[_p(x) for x in mk], _p(ta), _p(out["uflx"]), _p(out["dflx"]), _p(out["hr"]),
                            _p(out["uflxc"]), _p(out["dflxc"]), _p(out["hrc"]), _p(out["duflx_dt"]),
                            _p(out["duflxc_dt"]), nthreads)
    out["rc"] = rc
    out["icld"] = ic.value
    return out


def rrtmg_lw_nomcica(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, nthreads=1):
    """The non-McICA GCM entry (rrtmg_lw_rad.nomcica.f90: cldprop + rtrn / rtrnmr) on a cloudy synth.Batch."""
    ncol, nlay = b.ncol, b.nlay
    out = {k: np.zeros((nlay + 1, ncol)) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
    out["hr"] = np.zeros((nlay, ncol)); out["hrc"] = np.zeros((nlay, ncol))
    ic = C.c_int(icld)
    keep = [_np(x) for x in (b.play, b.plev, b.tlay, b.tlev, b.tsfc, b.h2ovmr, b.o3vmr, b.co2vmr, b.ch4vmr, b.n2ovmr,
                             b.o2vmr, b.cfc11vmr, b.cfc12vmr, b.cfc22vmr, b.ccl4vmr, b.emis)]
    cl = [_np(x) for x in (b.cldfrac, b.tauc, b.ciwp, b.clwp, b.rei, b.rel)]
    ta = _np(b.tauaer)
    rc = lib().orc_rrtmg_lw_nomcica(ncol, ncol, nlay, C.byref(ic), idrv, *[_p(x) for x in keep], inflglw, iceflglw,
                                    liqflglw, *[_p(x) for x in cl], _p(ta), _p(out["uflx"]), _p(out["dflx"]),
                                    _p(out["hr"]), _p(out["uflxc"]), _p(out["dflxc"]), _p(out["hrc"]),
                                    _p(out["duflx_dt"]), _p(out["duflxc_dt"]), nthreads)
    out["rc"] = rc
    out["icld"] = ic.value
    return out


def column_stages(b, icol, icld=0, idrv=0, inflglw=0, iceflglw=0, liqflglw=0, mc=None):
    ncol, nlay = b.ncol, b.nlay
    d = np.zeros((1,))
    m = mc or dict(cldfmcl=d, taucmcl=d, ciwpmcl=d, clwpmcl=d, reicmcl=d, relqmcl=d)
    keep = [_np(x) for x in (b.play, b.plev, b.tlay, b.tlev, b.tsfc, b.h2ovmr, b.o3vmr, b.co2vmr, b.ch4vmr, b.n2ovmr,
                             b.o2vmr, b.cfc11vmr, b.cfc12vmr, b.cfc22vmr, b.ccl4vmr, b.emis)]
    mk = [_np(m[k]) for k in ("cldfmcl", "taucmcl", "ciwpmcl", "clwpmcl", "reicmcl", "relqmcl")]
    ta = _np(b.tauaer)
    ints6 = np.zeros((nlay, 6), dtype=np.int32)
    facs = np.zeros((nlay, 4))
    taug = np.zeros((140, nlay)); fracs = np.zeros((140, nlay))
    itr = np.zeros((140, nlay, 2), dtype=np.int32)
    scal = np.zeros(2)
    rc = lib().orc_column_stages(icol, ncol, ncol, nlay, icld, idrv, *[_p(x) for x in keep], inflglw, iceflglw,
                                 liqflglw, *[_p(x) for x in mk], _p(ta), ints6.ctypes.data_as(_ip), _p(facs),
                                 _p(taug), _p(fracs), itr.ctypes.data_as(_ip), _p(scal))
    return dict(rc=rc, ints6=ints6, facs=facs, taug=taug, fracs=fracs, itr=itr, laytrop=int(scal[0]), pwvcm=scal[1])


def mcica_subcol_lw(b, icld, permuteseed=1, irng=0, alpha=None):
    """Oracle sub-column generator on a synth.Batch; returns the McICA dict for rrtmg_lw()."""
    ncol, nlay = b.ncol, b.nlay
    o3 = {k: np.zeros((nlay, ncol, 140)) for k in ("cldfmcl", "ciwpmcl", "clwpmcl", "taucmcl")}
    o2 = {k: np.zeros((nlay, ncol)) for k in ("reicmcl", "relqmcl")}
    ir = C.c_int(irng)
    al = _np(alpha) if alpha is not None else np.zeros((nlay, ncol))
    args = [_np(x) for x in (b.play, b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc)]
    rc = lib().orc_mcica_subcol_lw(ncol, ncol, nlay, icld, permuteseed, C.byref(ir), *[_p(x) for x in args], _p(al),
                                   _p(o3["cldfmcl"]), _p(o3["ciwpmcl"]), _p(o3["clwpmcl"]), _p(o2["reicmcl"]),
                                   _p(o2["relqmcl"]), _p(o3["taucmcl"]))
    d = dict(o3); d.update(o2); d["rc"] = rc; d["irng"] = ir.value
    return d


def get_alpha(b, icld, idcor=0, decorr_con=2500.0, juldat=1):
    alpha = np.zeros((b.nlay, b.ncol))
    rc = lib().orc_get_alpha(b.ncol, b.ncol, b.nlay, icld, idcor, C.c_double(decorr_con), _p(_np(b.dz)),
                             _p(_np(b.lat)), juldat, _p(_np(b.cldfrac)), _p(alpha))
    assert rc == 0
    return alpha


def count_ops(kdata_blob, b, icld=0, idrv=0, irng=0, alpha=None, default_real_kind=4):
    """Algorithmic FP64 operation count per column (op-counting build liborc_count.so):
    returns dict(add, mul, div, exp, log, pow, flops) averaged over the batch's columns for
    rrtmg_lw() itself (the McICA generator's 2 flops per draw are integer-dominated and excluded)."""
    global _lib
    path = os.path.join(_HERE, "liborc_count.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc_count.so"])
    saved = _lib
    try:
        _lib = C.CDLL(path)
        _lib.orc_get_table.restype = C.c_long
        ini(kdata_blob, 1004.0, default_real_kind)
        mc = mcica_subcol_lw(b, icld, 1, irng, alpha) if icld > 0 else None
        out = (C.c_longlong * 6)()
        _lib.orc_opcount(out, 1)
        r = rrtmg_lw(b, icld=icld, idrv=idrv, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=1)
        assert r["rc"] == 0
        _lib.orc_opcount(out, 1)
    finally:
        _lib = saved
    names = ("add", "mul", "div", "exp", "log", "pow")
    d = {n: out[i] / b.ncol for i, n in enumerate(names)}
    d["flops"] = d["add"] + d["mul"] + d["div"]
    return d


BRANCH_ADJFAC_SITES = ("taumol:550 b3 lower n2o", "taumol:736 b3 upper n2o", "taumol:1355 b6 co2", "taumol:1483 b7 lower co2",
                       "taumol:1639 b7 upper co2", "taumol:1721 b8 lower co2", "taumol:1768 b8 upper co2",
                       "taumol:1879 b9 lower n2o", "taumol:2036 b9 upper n2o", "taumol:2497 b13 co2")
BINARY_LOWER_BANDS = (3, 4, 5, 7, 9, 12, 13, 15, 16)


def branch_arm(on=True):
    """Switches every following oracle call to the coverage build liborc_cov.so (same arithmetic + the 143 branch
    counters) / back to the plain library, which carries no instrumentation because it is what gets timed."""
    global _lib, _plain
    if on and _plain is None:
        build()
        _plain = lib()
        cov = C.CDLL(os.path.join(_HERE, "liborc_cov.so"))
        cov.orc_get_table.restype = C.c_long
        _lib = cov
        if _ini_args is not None:
            ini(*_ini_args)
    elif not on and _plain is not None:
        _lib, _plain = _plain, None


def branch_counts(reset=False):
    """Branch-coverage counters of the oracle since the last reset (orc_common.h NBRANCH), decoded:
    eta[band][corner][branch] (branch 0: specparm < 0.125, 1: > 0.875, 2: linear), half[band][0 lower | 1 upper]
    (layers evaluated), adjfac[10] (sites in source order), cloudy[3], clear[2]."""
    n = lib().orc_branch_counts(None, 0, 0)
    raw = (C.c_longlong * n)()
    lib().orc_branch_counts(raw, n, 1 if reset else 0)
    v = np.array(list(raw), dtype=np.int64)
    return dict(eta=v[0:96].reshape(16, 2, 3), half=v[96:128].reshape(16, 2), adjfac=v[128:138], cloudy=v[138:141],
                clear=v[141:143])


def mt_stream(seed, n):
    ints = np.zeros(n, dtype=np.uint32); reals = np.zeros(n)
    lib().orc_mt_stream(seed, n, ints.ctypes.data_as(C.POINTER(C.c_uint)), _p(reals))
    return ints, reals


def kiss_stream(seeds, n):
    reals = np.zeros(n); st = np.zeros(4, dtype=np.int32)
    lib().orc_kiss_stream(*[int(s) for s in seeds], n, _p(reals), st.ctypes.data_as(_ip))
    return reals, st


def get_table(name):
    n = lib().orc_get_table(name.encode(), None, 0)
    out = np.zeros(n)
    lib().orc_get_table(name.encode(), _p(out), n)
    return out
