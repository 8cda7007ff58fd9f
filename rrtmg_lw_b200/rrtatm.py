"""Layering of a user-supplied sounding for ``IATM = 1`` inputs of the column-mode program: the part of
``src/rrtatm.f`` (LBLATM as modified for RRTM, 7 939 lines) that the reference's one ``IATM = 1`` example,
``run_examples_std_atm/input_rrtm_ICRCCM_sonde``, exercises -- a user profile (``MODEL = 0``) on altitude levels, user
layer boundaries in altitude (``IBMAX > 0``) and the vertical path RRTM forces (``ITYPE = 2``, ``ANGLE = 0``,
``rrtatm.f:585-590, 829-831``).  Host-side test/harness code like ``colmode.py``; everything else of LBLATM (model
atmospheres, pressure boundaries, automatic layering, slant / refracted paths, cross-sections) raises
``NotImplementedError``.

What it restates, statement for statement where it matters for the numbers:

* records 3.1-3.6 (``rrtatm.f:585, 829, 877, 3103, 3329-3347``) and the unit flags of ``JOU`` / ``CHECK`` / ``CONVRT`` /
  ``WATVAP`` (``:3393-3479, 3868-3976, 3977-4110``);
* ``AMERGE`` (``:5075-5252``): boundaries and profile levels merged into the path levels, pressure and densities
  interpolated exponentially (``EXPINT``, ``extra.f:223``), temperature linearly;
* ``ALAYER`` (``:5253-5495``) for a vertical ray: within a sub-layer pressure, air density and every absorber density
  are exponentials of height (linear when the two ends agree to 1e-5), integrated in steps of at most ``DELTAS`` = 5 km;
* ``FPACK`` (``:5805-5930``): sub-layers packed into the output layers; ``PBAR`` / ``TBAR`` are the air-density weighted
  means, the broadening-gas column is the air column minus the listed absorbers.

**Trap T18.**  ``CONVRT`` / ``WATVAP`` turn a MASS mixing ratio (unit flag ``C``, g/kg) into a volume mixing ratio with
``AIRMWT / AMWT(k)``.  ``AIRMWT`` lives in ``COMMON /CONSTS/``; the RRTMG driver fills that block only up to ``RADCN2``
(``src/rrtmg_lw.1col.f90:792-904``) and the ``DATA AIRMWT / 28.964 /`` of LBLATM is commented out (``rrtatm.f:1791``), so
in the reference program ``AIRMWT = 0`` and every absorber given in g/kg comes out as ZERO.  The sonde example gives its
water vapour that way: the shipped ``output_rrtm_ICRCCM_sonde`` (106.6 W m-2 of downward flux at a 290.9 K, 974 hPa
surface) is a dry-atmosphere result.  ``airmwt=0.0`` (the default here) reproduces the program; ``airmwt=28.964`` gives
what the input file means.
"""
from __future__ import annotations

import math

ALOSMT = 2.6867775e19      # rrlw_con, src/rrtmg_lw_init.f90:257 (passed to /CONSTS/, 1col:901-903)
AVOGAD = 6.02214199e23
GASCON = 8.31447200e07
PZERO, TZERO = 1013.25, 273.15   # rrtatm.f:1764
DELTAS = 5.0                     # :1763
EPSILN = 1.0e-5                  # ALAYER
TOL = 5.0e-4                     # AMERGE
RE_DEFAULT = 6371.23             # :676
GCAIR = 1.0e-3 * GASCON / AVOGAD  # :524
AMWT = (18.015, 44.010, 47.998, 44.01, 28.011, 16.043, 31.999)   # :1792-1793
_US_STD = None


def _us_standard():
    """Model 6 of BLOCK DATA MLATMB (AFGL U.S. Standard), the only model atmosphere a unit flag may default to here;
    fixture made by tools/extract_afgl_model6.py."""
    global _US_STD
    if _US_STD is None:
        import json
        import os
        _US_STD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "afgl_us_standard.json")))
    return _US_STD


def _defalt(z: float, junitp: int, junitt: int, junit, p: float, t: float, wmol):
    """DEFALT, rrtatm.f:3480-3672: values flagged 1-6 are taken from the model atmosphere by four-point Lagrange
    interpolation in altitude (pressure in ln p); only model 6 is available.  Returns (p, t)."""
    flags = [junitp, junitt] + list(junit)
    if all(f > 6 for f in flags):
        return p, t
    if any(f <= 6 and f != 6 for f in flags):
        raise NotImplementedError("rrtatm: defaults from model atmospheres 1-5")
    A = _us_standard()
    alt = A["alt_km"]
    if z > alt[49]:
        raise ValueError("rrtatm: Z IS GREATER THAN 120 KM")
    i2 = 50
    for im in range(2, 51):                     # 1-based as in the source
        if alt[im - 1] >= z:
            i2 = im
            break
    i1, i0, i3 = i2 - 1, i2 - 2, i2 + 1
    if i0 < 1:                                  # lower endpoint correction
        i0 = i1; i1 = i2; i2 = i3; i3 = i3 + 1
    elif i3 > 50:                               # upper endpoint correction
        i3 = i2; i2 = i1; i1 = i0; i0 = i1 - 1
    z0, z1, z2, z3 = alt[i0 - 1], alt[i1 - 1], alt[i2 - 1], alt[i3 - 1]
    den1 = (z0 - z1) * (z0 - z2) * (z0 - z3)
    den2 = (z1 - z2) * (z1 - z3) * (z1 - z0)
    den3 = (z2 - z3) * (z2 - z0) * (z2 - z1)
    den4 = (z3 - z0) * (z3 - z1) * (z3 - z2)
    a1 = ((z - z1) * (z - z2) * (z - z3)) / den1
    a2 = ((z - z2) * (z - z3) * (z - z0)) / den2
    a3 = ((z - z3) * (z - z0) * (z - z1)) / den3
    a4 = ((z - z0) * (z - z1) * (z - z2)) / den4

    def val(tab, f=lambda x: x):
        return a1 * f(tab[i0 - 1]) + a2 * f(tab[i1 - 1]) + a3 * f(tab[i2 - 1]) + a4 * f(tab[i3 - 1])
    if junitp <= 6:
        p = math.exp(val(A["p_mb"], math.log))
    if junitt <= 6:
        t = val(A["t_k"])
    for m in range(len(junit)):
        if junit[m] <= 6:
            wmol[m] = val(A["amol_ppmv"][m])
            junit[m] = 10
    return p, t


def _f(s: str, d: int = 3) -> float:
    s = s.strip()
    if not s:
        return 0.0
    if "." in s or "E" in s.upper():
        return float(s)
    return int(s) * 10.0 ** (-d)


def _i(s: str) -> int:
    s = s.strip()
    return int(s) if s else 0


def _jou(ch: str) -> int:
    """JOU, rrtatm.f:3393-3427."""
    if ch in "123456":
        return int(ch)
    if ch in " A":
        return 10
    k = "BCDEFGHIJK".find(ch)
    if k < 0:
        raise ValueError("rrtatm: invalid unit flag %r" % ch)
    return 11 + k


def _expint(x1: float, x2: float, a: float) -> float:
    """EXPINT, extra.f:223-247."""
    if x1 == 0.0 or x2 == 0.0:
        return x1 + (x2 - x1) * a
    return x1 * (x2 / x1) ** a


def layer_atmosphere(lines, k: int, airmwt: float = 0.0):
    """Reads records 3.1-3.6 starting at ``lines[k]`` and returns ``(atm, next_k)`` where ``atm`` holds what
    ``readprof`` takes over from ``rrtatm`` through ``COMMON /PROFILE/`` and ``/SPECIES/`` (1col:795-797, 1000-1003):
    ``nlayers, pavel, tavel, pz, tz, altz, wkl`` (molecules cm-2, 7 per layer), ``wbrodl``."""
    rec = lines[k].ljust(100); k += 1                                                    # 3.1, format 900 (:1569)
    model, itype, ibmax_b, nozero, noprnt, nmol = (_i(rec[0:5]), _i(rec[5:10]), _i(rec[10:15]), _i(rec[15:20]),
                                                   _i(rec[20:25]), _i(rec[25:30]))
    re_km = _f(rec[40:50])
    itype, nozero = 2, 1                                                                 # forced, :588-590
    if nmol == 0:
        nmol = 7
    if model != 0:
        raise NotImplementedError("rrtatm: only MODEL = 0 (user-supplied profile) is restated")
    if ibmax_b <= 0:
        raise NotImplementedError("rrtatm: only user boundaries in altitude (IBMAX > 0) are restated")
    if nmol > 7:
        raise NotImplementedError("rrtatm: NMOL > 7")
    ibmax = ibmax_b
    if re_km == 0.0:
        re_km = RE_DEFAULT
    rec = lines[k].ljust(70); k += 1                                                     # 3.2, format 932 (:1602)
    h1, h2, len_ = _f(rec[0:10], 4), _f(rec[10:20], 4), _i(rec[50:55])
    angle = 0.0                                                                          # ANGLEF = 0., :831
    if not h1 < h2 or len_ != 0:
        raise NotImplementedError("rrtatm: only the upward vertical path H1 < H2 is restated")
    zbnd = []
    while len(zbnd) < ibmax:                                                             # 3.3B, format 940 (:1617)
        rec = lines[k].ljust(80); k += 1
        zbnd += [_f(rec[10 * j:10 * j + 10]) for j in range(min(8, ibmax - len(zbnd)))]
    if any(b <= a for a, b in zip(zbnd, zbnd[1:])):
        raise ValueError("rrtatm: boundary altitudes not in ascending order")            # :885-887
    rec = lines[k].ljust(29); k += 1                                                     # 3.4, format 905 (:3150)
    immax_b = _i(rec[0:5])
    if immax_b <= 0:
        raise NotImplementedError("rrtatm: only profiles on altitude levels (IMMAX > 0) are restated")
    immax = immax_b
    zm, pm, tm, denm, dryair = [], [], [], [], []
    for _ in range(immax):
        rec = lines[k].ljust(80); k += 1                                                 # 3.5, format 900 (:3384)
        z, p, t = _f(rec[0:10]), _f(rec[10:20]), _f(rec[20:30])
        junitp, junitt, jlong = _jou(rec[35]), _jou(rec[36]), rec[38]
        junit = [_jou(rec[40 + m]) for m in range(nmol)]
        rec = lines[k].ljust(120); k += 1                                                # 3.6, formats 905 / 906
        w = 15 if jlong == "L" else 10
        wmol = [_f(rec[w * m:w * (m + 1)], 8 if w == 15 else 3) for m in range(nmol)]
        if junitp == 11:                                                                 # CHECK, :3429-3478
            p *= PZERO
        elif junitp == 12:
            p *= PZERO / 760.0
        elif junitp > 12:
            raise ValueError("rrtatm: CHECK(P)")
        if junitt == 11:
            t += 273.15
        elif junitt > 11:
            raise ValueError("rrtatm: CHECK(T)")
        p, t = _defalt(z, junitp, junitt, junit, p, t, wmol)
        rhoair = ALOSMT * (p / PZERO) * (TZERO / t)                                      # CONVRT, :3906
        d = [0.0] * nmol
        d[0] = _watvap(p, t, junit[0], wmol[0], rhoair, airmwt)
        dry = rhoair - d[0]
        for m in range(1, nmol):
            b, r = AVOGAD / AMWT[m], airmwt / AMWT[m]
            if junit[m] <= 10:
                d[m] = wmol[m] * dry * 1.0e-6
            elif junit[m] == 11:
                d[m] = wmol[m]
            elif junit[m] == 12:
                d[m] = r * wmol[m] * 1.0e-3 * dry
            elif junit[m] == 13:
                d[m] = b * wmol[m] * 1.0e-6
            elif junit[m] == 14:
                d[m] = ALOSMT * (wmol[m] / PZERO) * (TZERO / t)
            else:
                raise ValueError("rrtatm: CONVRT unit %d" % junit[m])
        zm.append(z); pm.append(p); tm.append(t); denm.append(d); dryair.append(dry)
    if any(b <= a for a, b in zip(zm, zm[1:])):
        raise ValueError("rrtatm: input altitudes not in ascending order")               # :3131-3133
    if h2 > zm[-1] or h1 < zm[0]:
        raise NotImplementedError("rrtatm: path ends outside the profile (REDUCE)")
    # ---- FSCGEO for the vertical path looking up: HMIN = H1, LEN = 0 (:4111-4310, case 2A)
    hmin = h1
    # ---- AMERGE, :5075-5252
    zh = [hmin, h2]
    for b in zbnd:                                                                       # loop 30
        if abs(b - zh[0]) < TOL:
            zh[0] = b
    zout = [zh[0]]
    ib = next((j for j, b in enumerate(zbnd) if b > zh[0]), ibmax - 1)
    ih = 1
    while True:                                                                          # labels 50-80
        if ib < ibmax:
            if abs(zbnd[ib] - zh[ih]) < TOL:
                zh[ih] = zbnd[ib]
            if zbnd[ib] < zh[ih]:
                zout.append(zbnd[ib]); ib += 1
                continue
            if zbnd[ib] == zh[ih]:
                ib += 1
        zout.append(zh[ih]); ih += 1
        if ih > 1:
            break
    ioutmx = len(zout)
    zm = list(zm)
    im = next(j for j, z in enumerate(zm) if z >= hmin)                                  # loop 90
    zpth, pp, tp, denp = [], [], [], []
    iout = 0
    while True:                                                                          # labels 110-150
        if im < immax and abs(zout[iout] - zm[im]) < TOL:
            zm[im] = zout[iout]
        if im < immax and not zout[iout] < zm[im]:
            if zout[iout] == zm[im]:
                iout += 1
            zpth.append(zm[im]); pp.append(pm[im]); tp.append(tm[im]); denp.append(list(denm[im]))
            im += 1
            if abs(zpth[-1] - zout[ioutmx - 1]) < TOL:
                zout[ioutmx - 1] = zpth[-1]
        else:
            jm = max(im, 1)
            a = (zout[iout] - zm[jm - 1]) / (zm[jm] - zm[jm - 1])
            zpth.append(zout[iout])
            pp.append(_expint(pm[jm - 1], pm[jm], a))
            tp.append(tm[jm - 1] + (tm[jm] - tm[jm - 1]) * a)
            denp.append([_expint(denm[jm - 1][m], denm[jm][m], a) for m in range(nmol)])
            iout += 1
        if abs(zpth[-1] - zout[ioutmx - 1]) < TOL:
            zpth[-1] = zout[ioutmx - 1]
        if zpth[-1] == zout[ioutmx - 1]:
            break
    ipmax = len(zpth)
    # ---- RFPATH / ALAYER for SINAI = 0, COSAI = -1, :4882-5074, 5253-5495
    ppsum, tpsum, rhopsm = [0.0] * ipmax, [0.0] * ipmax, [0.0] * ipmax
    amtp = [[0.0] * nmol for _ in range(ipmax)]
    for j in range(ipmax - 1):
        z1, z2 = zpth[j], zpth[j + 1]
        pa, pb_end = pp[j], pp[j + 1]
        if pb_end == pa:
            raise ValueError("LBLATM: PRESSURES IN ADJOINING LAYERS MUST DIFFER")
        ta, tb_end = tp[j], tp[j + 1]
        rhoa, rhob_end = pa / (GCAIR * ta), pb_end / (GCAIR * tb_end)
        dz = z2 - z1
        hp = -dz / math.log(pb_end / pa)
        hrho = -dz / math.log(rhob_end / rhoa) if abs(rhob_end / rhoa - 1.0) >= EPSILN else 1.0e30
        dena, hden = list(denp[j]), [0.0] * nmol
        for m in range(nmol):
            da, db = denp[j][m], denp[j + 1][m]
            if not (da == 0.0 or db == 0.0 or abs(1.0 - da / db) <= EPSILN):
                hden[m] = -dz / math.log(db / da)
        hh1 = z1
        r1 = re_km + hh1
        dhmin = DELTAS ** 2 / (2.0 * r1)
        while True:
            dh = max(DELTAS, dhmin)                # -DELTAS * COSAI1 with COSAI1 = -1
            h3 = min(hh1 + dh, z2)
            dh = h3 - hh1
            r3 = re_km + h3
            ds = r3 - r1                           # X = -R COSAI = R and DS/DX = 1 for a vertical ray: DS = X3 - X1
            dsdz = ds / dh
            pb = pa * math.exp(-dh / hp)
            rhob = rhoa * math.exp(-dh / hrho)
            if dh / hrho >= EPSILN:
                ppsum[j] += dsdz * (hp / (1.0 + hp / hrho)) * (pa * rhoa - pb * rhob)
                tpsum[j] += dsdz * hp * (pa - pb) / GCAIR
                rhopsm[j] += dsdz * hrho * (rhoa - rhob)
            else:
                ppsum[j] += 0.5 * ds * (pa * rhoa + pb * rhob)
                tpsum[j] += 0.5 * ds * (pa + pb) / GCAIR
                rhopsm[j] += 0.5 * ds * (rhoa + rhob)
            for m in range(nmol):
                if hden[m] == 0.0 or abs(dh / hden[m]) < EPSILN:
                    denb = denp[j][m] + (denp[j + 1][m] - denp[j][m]) * (h3 - z1) / dz
                    amtp[j][m] += 0.5 * (dena[m] + denb) * ds * 1.0e5
                else:
                    denb = denp[j][m] * math.exp(-(h3 - z1) / hden[m])
                    amtp[j][m] += dsdz * hden[m] * (dena[m] - denb) * 1.0e5
                dena[m] = denb
            pa, rhoa = pb, rhob
            if h3 < z2:
                hh1, r1 = h3, r3
            else:
                break
    # ---- FPACK, :5805-5930 (NOZERO = 1: no amounts are zeroed)
    nl = ioutmx - 1
    pbar, tbar, rhosum = [0.0] * nl, [0.0] * nl, [0.0] * nl
    amount = [[0.0] * nmol for _ in range(nl)]
    pz, tz = [pp[0]], [tp[0]]
    io = 0
    for ip in range(ipmax - 1):
        pbar[io] += ppsum[ip]; tbar[io] += tpsum[ip]; rhosum[io] += rhopsm[ip]
        for m in range(nmol):
            amount[io][m] += amtp[ip][m]
        if zpth[ip + 1] == zout[io + 1]:
            pz.append(pp[ip + 1]); tz.append(tp[ip + 1])
            io += 1
    if io != nl:
        raise ValueError("rrtatm: ERROR FPACK")
    wn2l = []
    for l in range(nl):
        pbar[l] /= rhosum[l]
        tbar[l] /= rhosum[l]
        rhosum[l] *= 1.0e5
        wn2l.append(rhosum[l] - sum(amount[l]))
    wkl = [a + [0.0] * (7 - nmol) for a in amount]
    return dict(nlayers=nl, nmol=nmol, pavel=pbar, tavel=tbar, pz=pz, tz=tz, altz=list(zout), wkl=wkl, wbrodl=wn2l,
                angle=angle), k


def _watvap(p: float, t: float, junit: int, wmol: float, rhoair: float, airmwt: float) -> float:
    """WATVAP, rrtatm.f:3977-4110: water-vapour number density (cm-3) from the profile's unit."""
    a = TZERO / t
    b = AVOGAD / AMWT[0]
    r = airmwt / AMWT[0]

    def densat(atemp):
        return atemp * b * math.exp(18.9766 - 14.9595 * atemp - 2.4388 * atemp ** 2) * 1.0e-6
    if junit == 10:
        w = wmol * 1.0e-6
        return (w / (1.0 + w)) * rhoair
    if junit == 11:
        return wmol
    if junit == 12:
        w = wmol * r * 1.0e-3
        return (w / (1.0 + w)) * rhoair
    if junit == 13:
        return b * wmol * 1.0e-6
    if junit == 14:
        return ALOSMT * (wmol / PZERO) * (TZERO / t)
    if junit == 15:
        atd = TZERO / wmol
        return densat(atd) * wmol / t
    if junit == 16:
        atd = TZERO / (TZERO + wmol)
        return densat(atd) * (TZERO + wmol) / t
    if junit == 17:
        return densat(a) * (wmol / 100.0)
    raise ValueError("rrtatm: WATVAP unit %d" % junit)
