"""Batch-level convenience wrappers: run a ``synth.Batch`` through the C ABI, on host (numpy)
or device (CUDA torch) buffers.  Pure plumbing; all arithmetic happens in the CUDA library."""
from __future__ import annotations

import numpy as np

from . import api


def _outputs(nlay, ncol, like=None):
    if like is not None and hasattr(like, "new_zeros"):
        mk = lambda n: like.new_zeros((n, ncol))  # noqa: E731
    else:
        mk = lambda n: np.zeros((n, ncol))  # noqa: E731
    out = {k: mk(nlay + 1) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
    out["hr"] = mk(nlay)
    out["hrc"] = mk(nlay)
    return out


def _conv(b, device):
    """Batch fields as numpy (device None) or tensors on `device`."""
    import torch

    def c(t):
        if t is None:
            return None
        if device is None:
            return np.ascontiguousarray(t.detach().cpu().numpy()) if hasattr(t, "detach") else np.ascontiguousarray(t)
        t = t if hasattr(t, "detach") else torch.from_numpy(np.ascontiguousarray(t))
        return t.to(device).contiguous()
    return c


def atm_args(b, c):
    return [c(x) for x in (b.play, b.plev, b.tlay, b.tlev, b.tsfc, b.h2ovmr, b.o3vmr, b.co2vmr, b.ch4vmr, b.n2ovmr,
                           b.o2vmr, b.cfc11vmr, b.cfc12vmr, b.cfc22vmr, b.ccl4vmr, b.emis)]


def run_clear_or_strict(b, icld=0, idrv=0, inflglw=0, iceflglw=0, liqflglw=0, mc=None, device=None, out=None):
    """rrtmg_lw() on a Batch.  mc: dict of McICA arrays (nlay,ncol,140)/(nlay,ncol) when icld >= 1."""
    c = _conv(b, device)
    a = atm_args(b, c)
    ta = c(b.tauaer)
    if out is None:
        out = _outputs(b.nlay, b.ncol, like=a[0] if device is not None else None)
    m = [c(mc[k]) if mc is not None else None for k in ("cldfmcl", "taucmcl", "ciwpmcl", "clwpmcl", "reicmcl", "relqmcl")]
    icld_out = api.rrtmg_lw(b.ncol, b.nlay, icld, idrv, *a, inflglw, iceflglw, liqflglw, *m, ta, out["uflx"],
                            out["dflx"], out["hr"], out["uflxc"], out["dflxc"], out["hrc"], out["duflx_dt"],
                            out["duflxc_dt"])
    out["icld"] = icld_out
    return out


def run_fused(b, icld=2, idrv=0, permuteseed=1, irng=0, inflglw=2, iceflglw=3, liqflglw=1, alpha=None, device=None,
              out=None):
    """rrtmg_lw_mcica() on a Batch (generator + radiation in one device pass)."""
    c = _conv(b, device)
    a = atm_args(b, c)
    cl = [c(x) for x in (b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc)]
    al = c(alpha) if alpha is not None else None
    ta = c(b.tauaer)
    if out is None:
        out = _outputs(b.nlay, b.ncol, like=a[0] if device is not None else None)
    api.rrtmg_lw_mcica(b.ncol, b.nlay, icld, idrv, permuteseed, irng, *a, inflglw, iceflglw, liqflglw, *cl, al, ta,
                       out["uflx"], out["dflx"], out["hr"], out["uflxc"], out["dflxc"], out["hrc"], out["duflx_dt"],
                       out["duflxc_dt"])
    return out


def run_nomcica(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, device=None, out=None):
    """rrtmg_lw() of the non-McICA build (grid-mean clouds; rtrn for icld = 1, rtrnmr otherwise) on a Batch."""
    c = _conv(b, device)
    a = atm_args(b, c)
    cl = [c(x) for x in (b.cldfrac, b.tauc, b.ciwp, b.clwp, b.rei, b.rel)]
    ta = c(b.tauaer)
    if out is None:
        out = _outputs(b.nlay, b.ncol, like=a[0] if device is not None else None)
    out["icld"] = api.rrtmg_lw_nomcica(b.ncol, b.nlay, icld, idrv, *a, inflglw, iceflglw, liqflglw, *cl, ta,
                                       out["uflx"], out["dflx"], out["hr"], out["uflxc"], out["dflxc"], out["hrc"],
                                       out["duflx_dt"], out["duflxc_dt"])
    return out


def run_mcica_subcol(b, icld, permuteseed=1, irng=0, alpha=None, device=None):
    c = _conv(b, device)
    args = [c(x) for x in (b.play, b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc)]
    al = c(alpha) if alpha is not None else None
    if device is None:
        mk3 = lambda: np.zeros((b.nlay, b.ncol, 140))  # noqa: E731
        mk2 = lambda: np.zeros((b.nlay, b.ncol))  # noqa: E731
    else:
        mk3 = lambda: args[0].new_zeros((b.nlay, b.ncol, 140))  # noqa: E731
        mk2 = lambda: args[0].new_zeros((b.nlay, b.ncol))  # noqa: E731
    o = dict(cldfmcl=mk3(), ciwpmcl=mk3(), clwpmcl=mk3(), reicmcl=mk2(), relqmcl=mk2(), taucmcl=mk3())
    o["irng"] = api.mcica_subcol_lw(1, b.ncol, b.nlay, icld, permuteseed, irng, *args, al, o["cldfmcl"], o["ciwpmcl"],
                                    o["clwpmcl"], o["reicmcl"], o["relqmcl"], o["taucmcl"])
    return o
