"""k-distribution data handling for rrtmg_lw_ini.

The reference fills the 16-g "original" absorption-coefficient arrays in ``lw_kgb01..16``
(``src/rrtmg_lw_k_g.f90`` or ``src/rrtmg_lw_read_nc.f90`` + ``data/rrtmg_lw.nc``).  Both
sources are missing from the reference mount (``.MISSING_LARGE_BLOBS``), so this module
provides

* the blob layout (parsed from ``include/rrtmg_lw_kdata_layout.h`` -- single source of truth),
* ``synthetic_kdata``: a deterministic generator of tables with the exact reference shapes
  (``modules/rrlw_kg01..16.f90``), positive, smooth in (eta, T, p), increasing in g, with a
  small element-wise jitter so that any index mix-up changes the result,
* ``load_netcdf``: reader for a real ``rrtmg_lw.nc`` following the variable names and
  hyperslabs of ``src/rrtmg_lw_read_nc.f90`` / ``modules/rrlw_ncpar.f90`` (NetCDF-3 classic
  through ``scipy.io.netcdf_file``); round-trip tested against ``save_netcdf`` (same names, dimensions and
  hyperslabs), not against real data because none is available.
"""
from __future__ import annotations

import os
import re

import numpy as np

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                    "rrtmg_lw_kdata_layout.h")


def karrays():
    """[(band, name, lead, gfirst, weighted)] in blob order."""
    txt = open(_HDR).read()
    out = []
    for m in re.finditer(r'\{(\d+), "([a-z0-9_]+)", (\d+), ([01]), ([01])\}', txt):
        out.append((int(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4)), int(m.group(5))))
    n = int(re.search(r"#define RRTMG_LW_NKARRAYS (\d+)", txt).group(1))
    assert len(out) == n, (len(out), n)
    return out


KDATA_DOUBLES = sum(16 * a[2] for a in karrays())

# nspa/nspb per band, src/rrtmg_lw_init.f90:227-228
NSPA = (1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9)
NSPB = (1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0)

# typical scaled column amount (1e-20 molecules cm-2) of the key absorber of each band, used
# only to give the synthetic optical depths a realistic dynamic range
_UKEY_LO = (1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 1400., 76., 0.07, 1400.)
_UKEY_UP = (0.3, 0.3, 0.3, 0.05, 0.05, 1.0, 0.05, 0.05, 0.02, 0.3, 0.3, 1.0, 1.0, 4.0, 1.0, 0.02)


def pack(arrays: dict) -> np.ndarray:
    """dict[(band, name)] -> flat float64 blob (Fortran element order per array)."""
    chunks = []
    for band, name, lead, gfirst, _ in karrays():
        a = np.asarray(arrays[(band, name)], dtype=np.float64)
        shape = (16, lead) if gfirst else (lead, 16)
        assert a.shape == shape, (band, name, a.shape, shape)
        chunks.append(a.ravel(order="F"))
    blob = np.concatenate(chunks)
    assert blob.size == KDATA_DOUBLES
    return np.ascontiguousarray(blob)


def unpack(blob: np.ndarray) -> dict:
    out, off = {}, 0
    for band, name, lead, gfirst, _ in karrays():
        n = 16 * lead
        shape = (16, lead) if gfirst else (lead, 16)
        out[(band, name)] = blob[off:off + n].reshape(shape, order="F").copy()
        off += n
    return out


def synthetic_kdata(seed: int = 20261017, jitter: float = 0.08) -> dict:
    """Deterministic synthetic k-tables with the reference's exact shapes."""
    rng = np.random.default_rng(seed)
    g = (np.arange(16) + 0.5) / 16.0
    out = {}

    def jit(shape):
        return np.exp(jitter * rng.standard_normal(shape))

    for band, name, lead, gfirst, _ in karrays():
        b = band - 1
        # log-spaced k(g): from weak continuum-like to line-centre values
        span = np.exp(np.log(1e-3) + (np.log(3e3) - np.log(1e-3)) * g ** 1.5)
        if name in ("kao", "kbo"):
            lower = name == "kao"
            nsp = NSPA[b] if lower else max(NSPB[b], 1)
            npr = 13 if lower else 47
            u = _UKEY_LO[b] if lower else _UKEY_UP[b]
            eta = np.arange(nsp) / max(nsp - 1, 1)
            jt = np.arange(5) - 2.0
            jp = np.arange(npr) / (npr - 1.0)
            k = ((40.0 if lower else 0.05) * span / u)[None, None, None, :] \
                * (1.0 + 0.6 * np.sin(2.2 * eta + 0.3 * b))[:, None, None, None] \
                * np.exp(0.08 * jt)[None, :, None, None] \
                * np.exp(-(1.5 - 2.5 * g)[None, None, None, :] * jp[None, None, :, None])
            k = k * jit(k.shape)
            a = k.reshape((nsp * 5 * npr, 16), order="F")
            # a band without upper-atmosphere key species still owns 235 rows (band 16: nspb = 0
            # but kbo(5,13:59,16) exists)
            assert a.shape[0] == lead, (band, name, a.shape, lead)
        elif name == "selfrefo":
            t = np.arange(10)
            a = (30.0 * span / _UKEY_LO[b])[None, :] * np.exp(-0.06 * t)[:, None] * jit((10, 16))
        elif name == "forrefo":
            t = np.arange(4)
            a = (1.0 * span / _UKEY_LO[b])[None, :] * np.exp(-0.03 * t)[:, None] * jit((4, 16))
        elif name.startswith("fracref"):
            # Planck fractions: positive, each eta column sums to exactly ~1 over the 16 g
            w = np.exp(-3.0 * g)[:, None] * (1.0 + 0.3 * np.cos(np.arange(lead) * 0.7 + b))[None, :]
            w = w * jit((16, lead))
            a = w / w.sum(axis=0, keepdims=True)
        elif name in ("ccl4o", "cfc11adjo", "cfc12o", "cfc22adjo"):
            a = (2.0e3 * (0.2 + g))[None, :] * jit((1, 16))
        else:
            # minor-gas tables (19 T points, optionally x eta)
            neta = lead // 19
            umin = {"mn2": 1.5e5, "mn2o": 0.07, "mo3": 0.09, "mco2": 76., "mco": 0.03, "mo2": 4.0e4}[
                name.split("_", 1)[1]]
            if name in ("kao_mn2", "kbo_mn2", "kao_mo2", "kbo_mo2"):
                umin *= 1.0  # amounts are further scaled by p/T-type factors in taumol
            t = np.arange(19)
            eta = np.arange(neta) / max(neta - 1, 1)
            k = (0.3 * span / umin)[None, None, :] * np.exp(0.02 * t)[None, :, None] \
                * (1.0 + 0.4 * np.cos(1.7 * eta))[:, None, None]
            k = k * jit(k.shape)
            a = k.reshape((lead, 16), order="F")
        out[(band, name)] = np.ascontiguousarray(a, dtype=np.float64)
    return out


def synthetic_blob(seed: int = 20261017) -> np.ndarray:
    return pack(synthetic_kdata(seed))


# ----------------------------------------------------------------------------------------------
# real data: rrtmg_lw.nc  (src/rrtmg_lw_read_nc.f90, modules/rrlw_ncpar.f90:15-44)
_ABSORBER = {"n2": 1, "ccl4": 2, "cfc11": 3, "cfc12": 4, "cfc22": 5, "h2o": 6, "co2": 7, "o3": 8, "n2o": 9,
             "co": 10, "ch4": 11, "o2": 12}
_MINOR_OF = {"mn2": "n2", "mn2o": "n2o", "mo3": "o3", "mco2": "co2", "mco": "co", "mo2": "o2",
             "ccl4o": "ccl4", "cfc11adjo": "cfc11", "cfc12o": "cfc12", "cfc22adjo": "cfc22"}


def load_netcdf(path: str) -> dict:
    """Read ``rrtmg_lw.nc`` into the blob dictionary.  LW uses gPointSetNumber = 1.

    Variable names / dimension order follow ``src/rrtmg_lw_read_nc.f90`` (e.g. band 1 ``:47-110``,
    band 3 ``:168-232``): in the file (C order) the Fortran dimensions appear reversed.
    """
    from scipy.io import netcdf_file

    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    # Fortran dims (keylower=9, Tdiff=5, plower=13, g=16, band=16, set=2) -> C order reversed
    kl = np.asarray(v["KeySpeciesAbsorptionCoefficientsLowerAtmos"][:])[0]
    ku = np.asarray(v["KeySpeciesAbsorptionCoefficientsUpperAtmos"][:])[0]
    sf = np.asarray(v["H20SelfAbsorptionCoefficients"][:])[0]
    fr = np.asarray(v["H20ForeignAbsorptionCoefficients"][:])[0]
    pl = np.asarray(v["PlanckFractionLowerAtmos"][:])[0]
    pu = np.asarray(v["PlanckFractionUpperAtmos"][:])[0]
    ml = np.asarray(v["AbsorptionCoefficientsLowerAtmos"][:])[0]
    mu = np.asarray(v["AbsorptionCoefficientsUpperAtmos"][:])[0]
    out = {}
    for band, name, lead, gfirst, _ in karrays():
        b = band - 1
        if name == "kao":
            a = kl[b].transpose(3, 2, 1, 0)[:NSPA[b]]            # (key, T, p, g)
            a = a.reshape((lead, 16), order="F")
        elif name == "kbo":
            nsp = max(NSPB[b], 1)
            a = ku[b].transpose(3, 2, 1, 0)[:nsp].reshape((lead, 16), order="F")
        elif name == "selfrefo":
            a = sf[b].T
        elif name == "forrefo":
            a = fr[b].T
        elif name == "fracrefao":
            a = pl[b].T[:, :lead]                                # (g, eta)
        elif name == "fracrefbo":
            a = pu[b].T[:, :lead]
        else:
            key = name.split("_", 1)[1] if "_" in name else name
            ab = _ABSORBER[_MINOR_OF[key]] - 1
            src = ml if (name.startswith("kao") or name in _MINOR_OF) else mu
            blk = src[b, ab].transpose(2, 1, 0)                  # (key, T=19, g)
            if lead == 1:
                a = blk[0, 0][None, :]
            elif lead == 19:
                a = blk[0]
            else:
                a = blk[:lead // 19].reshape((lead, 16), order="F")
        out[(band, name)] = np.ascontiguousarray(a, dtype=np.float64)
    f.close()
    return out


def save_netcdf(path: str, arrays: dict) -> None:
    """Write a blob dictionary as an ``rrtmg_lw.nc``-shaped NetCDF-3 classic file: the eight variables of
    ``src/rrtmg_lw_read_nc.f90`` with the dimensions of ``modules/rrlw_ncpar.f90:31-44`` (Fortran order
    ``(keylower, Tdiff, plower, GPoint, band, GPointSet)`` etc., i.e. reversed in the file), the slabs each
    ``lw_kgbNN`` reads filled from ``arrays`` and everything else zero.  Used to exercise ``load_netcdf`` without
    the reference's data file and to hand k-tables to a reference build that reads NetCDF."""
    from scipy.io import netcdf_file

    dims = dict(keylower=9, keyupper=5, Tdiff=5, plower=13, pupper=47, Tself=10, Tforeign=4, T=19, band=16, GPoint=16,
                GPointSet=2, Absorber=12)
    spec = {   # file (C) order = reversed Fortran order
        "KeySpeciesAbsorptionCoefficientsLowerAtmos": ("GPointSet", "band", "GPoint", "plower", "Tdiff", "keylower"),
        "KeySpeciesAbsorptionCoefficientsUpperAtmos": ("GPointSet", "band", "GPoint", "pupper", "Tdiff", "keyupper"),
        "H20SelfAbsorptionCoefficients": ("GPointSet", "band", "GPoint", "Tself"),
        "H20ForeignAbsorptionCoefficients": ("GPointSet", "band", "GPoint", "Tforeign"),
        "PlanckFractionLowerAtmos": ("GPointSet", "band", "keylower", "GPoint"),
        "PlanckFractionUpperAtmos": ("GPointSet", "band", "keyupper", "GPoint"),
        "AbsorptionCoefficientsLowerAtmos": ("GPointSet", "band", "Absorber", "GPoint", "T", "keylower"),
        "AbsorptionCoefficientsUpperAtmos": ("GPointSet", "band", "Absorber", "GPoint", "T", "keyupper"),
    }
    data = {k: np.zeros([dims[d] for d in v]) for k, v in spec.items()}
    kl, ku = data["KeySpeciesAbsorptionCoefficientsLowerAtmos"][0], data["KeySpeciesAbsorptionCoefficientsUpperAtmos"][0]
    sf, fr = data["H20SelfAbsorptionCoefficients"][0], data["H20ForeignAbsorptionCoefficients"][0]
    pl, pu = data["PlanckFractionLowerAtmos"][0], data["PlanckFractionUpperAtmos"][0]
    ml, mu = data["AbsorptionCoefficientsLowerAtmos"][0], data["AbsorptionCoefficientsUpperAtmos"][0]
    for band, name, lead, gfirst, _ in karrays():
        b = band - 1
        a = np.asarray(arrays[(band, name)], dtype=np.float64)
        if name == "kao":
            kl[b][:, :, :, :NSPA[b]] = a.reshape((NSPA[b], 5, 13, 16), order="F").transpose(3, 2, 1, 0)
        elif name == "kbo":
            nsp = max(NSPB[b], 1)
            ku[b][:, :, :, :nsp] = a.reshape((nsp, 5, 47, 16), order="F").transpose(3, 2, 1, 0)
        elif name == "selfrefo":
            sf[b] = a.T
        elif name == "forrefo":
            fr[b] = a.T
        elif name == "fracrefao":
            pl[b][:lead] = a.T
        elif name == "fracrefbo":
            pu[b][:lead] = a.T
        else:
            key = name.split("_", 1)[1] if "_" in name else name
            ab = _ABSORBER[_MINOR_OF[key]] - 1
            dst = ml if (name.startswith("kao") or name in _MINOR_OF) else mu
            if lead == 1:
                dst[b, ab][:, 0, 0] = a[0]
            elif lead == 19:
                dst[b, ab][:, :, 0] = a.T
            else:
                dst[b, ab][:, :, :lead // 19] = a.reshape((lead // 19, 19, 16), order="F").transpose(2, 1, 0)
    f = netcdf_file(path, "w")
    for d, n in dims.items():
        f.createDimension(d, n)
    for k, v in spec.items():
        var = f.createVariable(k, "d", v)
        var[:] = data[k]
    f.close()
