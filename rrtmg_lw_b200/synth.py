"""Synthetic GCM column batches of the shapes named in BASELINE.json.

All arrays are float64 torch tensors in the reference's Fortran layout seen from C:
``(nlay, ncol)`` row-major == Fortran ``(ncol, nlay)`` (column index fastest), layer 1 = surface
(``src/rrtmg_lw_rad.f90:785-787``).  Works on CPU (tests, CPU baseline) and on CUDA (bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

NBND = 16
NGPT = 140


@dataclass
class Batch:
    ncol: int
    nlay: int
    play: torch.Tensor   # (nlay, ncol) hPa
    plev: torch.Tensor   # (nlay+1, ncol)
    tlay: torch.Tensor
    tlev: torch.Tensor
    tsfc: torch.Tensor   # (ncol,)
    h2ovmr: torch.Tensor
    o3vmr: torch.Tensor
    co2vmr: torch.Tensor
    ch4vmr: torch.Tensor
    n2ovmr: torch.Tensor
    o2vmr: torch.Tensor
    cfc11vmr: torch.Tensor
    cfc12vmr: torch.Tensor
    cfc22vmr: torch.Tensor
    ccl4vmr: torch.Tensor
    emis: torch.Tensor    # (16, ncol)
    tauaer: torch.Tensor  # (16, nlay, ncol)
    # grid-mean cloud inputs of mcica_subcol_lw
    cldfrac: torch.Tensor = None
    ciwp: torch.Tensor = None
    clwp: torch.Tensor = None
    rei: torch.Tensor = None
    rel: torch.Tensor = None
    tauc: torch.Tensor = None   # (nlay, ncol, 16) == Fortran (16, ncol, nlay)
    lat: torch.Tensor = None
    dz: torch.Tensor = None
    meta: dict = field(default_factory=dict)

    def gas_fields(self):
        return [self.h2ovmr, self.o3vmr, self.co2vmr, self.ch4vmr, self.n2ovmr, self.o2vmr,
                self.cfc11vmr, self.cfc12vmr, self.cfc22vmr, self.ccl4vmr]


def sigma_interfaces(nlay: int, dtype=torch.float64, device="cpu") -> torch.Tensor:
    """Fixed sigma-like interface grid from 1.0 down to 1e-5 (top ~0.01 hPa); about one third
    of the layers lie below ~100 hPa."""
    x = torch.arange(nlay + 1, dtype=dtype, device=device) / nlay
    return torch.exp(-math.log(1.0e5) * x ** 1.5)


def _t_profile(p: torch.Tensor) -> torch.Tensor:
    """MLS-like temperature (K) as a piecewise-linear function of ln p (hPa)."""
    lp = torch.log(p)
    knots_p = [1013.25, 150.0, 50.0, 1.0, 0.01]
    knots_t = [294.0, 216.0, 216.0, 270.0, 220.0]
    t = torch.full_like(p, knots_t[0])
    for i in range(4):
        a, b = math.log(knots_p[i]), math.log(knots_p[i + 1])
        w = ((lp - a) / (b - a)).clamp(0.0, 1.0)
        seg = knots_t[i] + w * (knots_t[i + 1] - knots_t[i])
        t = torch.where(lp <= a, seg, t)
    return t


def make_batch(ncol: int, nlay: int = 72, cloudy: bool = False, aerosol: bool = False, seed: int = 20261017,
               device="cpu", stress_co2: bool = False, col_offset: int = 0, gas_spread: float = 0.0) -> Batch:
    """Synthetic GCM batch of SURVEY 8(d).  ``stress_co2`` triples CO2 / doubles N2O (the ``adjfac`` pow paths);
    ``gas_spread`` > 0 multiplies every absorber by an independent 10**(gas_spread * N(0,1)) per (layer, column)
    (clipped to [1e-3, 50]), which drives the binary-species interpolation parameter through all three eta branches
    of every band (taumol.f90:569-598) -- a coverage batch for the parity tests, not a physical atmosphere."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7919 * col_offset)
    f64 = torch.float64

    def U(*shape):
        return torch.rand(*shape, generator=g, dtype=f64, device=dev)

    def N(*shape):
        return torch.randn(*shape, generator=g, dtype=f64, device=dev)

    sig = sigma_interfaces(nlay, f64, dev)                      # (nlay+1,)
    ps = 1013.25 * (0.85 + 0.2 * U(ncol))                       # (ncol,)
    plev = sig[:, None] * ps[None, :]
    play = 0.5 * (plev[:-1] + plev[1:])
    toff = (8.0 * N(ncol)).clamp(-25.0, 25.0)
    tlay = (_t_profile(play) + toff[None, :]).clamp(180.0, 320.0)
    tlev = (_t_profile(plev) + toff[None, :]).clamp(180.0, 320.0)
    tsfc = tlev[0] + 2.0 * N(ncol)

    prel = play / ps[None, :]
    h2o = (0.02 * prel ** 3 * torch.exp(0.5 * N(nlay, ncol))).clamp_min(3.0e-6)
    o3 = 3.0e-8 + 8.0e-6 * torch.exp(-0.5 * torch.log(play / 10.0) ** 2)
    ones = torch.ones(nlay, ncol, dtype=f64, device=dev)
    co2 = (1.2e-3 if stress_co2 else 4.0e-4) * ones
    n2o = (6.0e-7 if stress_co2 else 3.2e-7) * ones
    ch4 = 1.8e-6 * ones
    if gas_spread > 0.0:
        def spread(x, cap):
            return (x * torch.pow(torch.tensor(10.0, dtype=f64, device=dev), gas_spread * N(nlay, ncol)).clamp(1.0e-3, 50.0)).clamp_max(cap)
        h2o = spread(h2o, 0.04)
        o3 = spread(o3, 1.0e-3)
        co2 = spread(co2, 0.05)
        n2o = spread(n2o, 1.0e-3)
        ch4 = spread(ch4, 1.0e-3)
    emis = 0.96 + 0.04 * U(NBND, ncol)
    if aerosol:
        tauaer = 0.01 * U(NBND, nlay, ncol) * (play > 500.0).to(f64)[None]
    else:
        tauaer = torch.zeros(NBND, nlay, ncol, dtype=f64, device=dev)

    b = Batch(ncol=ncol, nlay=nlay, play=play.contiguous(), plev=plev.contiguous(), tlay=tlay, tlev=tlev, tsfc=tsfc,
              h2ovmr=h2o, o3vmr=o3, co2vmr=co2, ch4vmr=ch4, n2ovmr=n2o, o2vmr=0.209 * ones,
              cfc11vmr=2.3e-10 * ones, cfc12vmr=5.2e-10 * ones, cfc22vmr=2.3e-10 * ones, ccl4vmr=8.0e-11 * ones,
              emis=emis, tauaer=tauaer.contiguous())
    if cloudy:
        inband = (play < 900.0) & (play > 150.0)
        has = (U(nlay, ncol) < 0.3) & inband
        cf = torch.where(has, 0.05 + 0.95 * U(nlay, ncol), torch.zeros_like(play))
        clear_col = U(ncol) < 0.10
        cf = torch.where(clear_col[None, :], torch.zeros_like(cf), cf)
        # 5 % of the columns carry an overcast deck in one layer of the cloud band
        deck_col = (U(ncol) < 0.05) & ~clear_col
        nband = inband.to(torch.int64).sum(0).clamp_min(1)
        first = torch.argmax(inband.to(torch.int64), dim=0)
        deck_lay = first + (U(ncol) * nband.to(f64)).to(torch.int64).clamp_max(nlay - 1)
        lay_idx = torch.arange(nlay, device=dev)[:, None]
        deck = deck_col[None, :] & (lay_idx == deck_lay[None, :])
        cf = torch.where(deck, torch.ones_like(cf), cf)
        cloudy_pt = cf > 0
        clwp = torch.where(cloudy_pt & (tlay > 253.0), 30.0 * torch.exp(0.7 * N(nlay, ncol)), torch.zeros_like(cf))
        ciwp = torch.where(cloudy_pt & (tlay < 273.0), 10.0 * torch.exp(0.7 * N(nlay, ncol)), torch.zeros_like(cf))
        b.cldfrac = cf
        b.clwp = clwp
        b.ciwp = ciwp
        b.rel = 4.0 + 16.0 * U(nlay, ncol)
        b.rei = 20.0 + 100.0 * U(nlay, ncol)
        b.tauc = torch.zeros(nlay, ncol, NBND, dtype=f64, device=dev)
        b.lat = -90.0 + 180.0 * U(ncol)
        # hydrostatic thickness (m): dz = R_d T / g * ln(p_bot / p_top)
        b.dz = 287.05 * tlay / 9.8066 * torch.log(plev[:-1] / plev[1:])
    b.meta = dict(seed=seed, cloudy=cloudy, aerosol=aerosol, stress_co2=stress_co2, gas_spread=gas_spread)
    return b


def batch_from_profile(prof: dict, ncol: int = 1, device="cpu") -> Batch:
    """GCM-style inputs from one parsed column-mode profile (tests/golden/std_atm.json, made by
    tools/make_golden.py from run_examples_std_atm/input_rrtm_*): pavel/tavel -> play/tlay,
    pz/tz -> plev/tlev, molecular mixing ratios (h2o, co2, o3, n2o, co, ch4, o2) -> vmr arrays.
    The column is replicated ncol times.  (The column-mode program derives coldry from the file's
    broadening-gas column, the GCM entry from the level pressures, src/rrtmg_lw_rad.f90:807-809, so
    the k-dependent goldens are comparable only approximately through this path.)"""
    f64 = torch.float64
    nlay = prof["nlayers"]

    def col(v):
        return torch.tensor(v, dtype=f64, device=device)[:, None].repeat(1, ncol).contiguous()

    wkl = torch.tensor(prof["wkl"], dtype=f64, device=device)   # (nlay, 7)
    z = torch.zeros(nlay, ncol, dtype=f64, device=device)
    g = lambda m: wkl[:, m][:, None].repeat(1, ncol).contiguous()  # noqa: E731
    return Batch(ncol=ncol, nlay=nlay, play=col(prof["pavel"]), plev=col(prof["pz"]), tlay=col(prof["tavel"]),
                 tlev=col(prof["tz"]), tsfc=torch.full((ncol,), prof["tbound"], dtype=f64, device=device),
                 h2ovmr=g(0), co2vmr=g(1), o3vmr=g(2), n2ovmr=g(3), ch4vmr=g(5), o2vmr=g(6),
                 cfc11vmr=z.clone(), cfc12vmr=z.clone(), cfc22vmr=z.clone(), ccl4vmr=z.clone(),
                 emis=torch.ones(NBND, ncol, dtype=f64, device=device),
                 tauaer=torch.zeros(NBND, nlay, ncol, dtype=f64, device=device), meta=dict(profile=True))
