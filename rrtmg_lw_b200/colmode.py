"""Example-suite harness: the column-mode program's file formats and its
driver loop, restated as host code over the GCM entry points of the CUDA library.

What the reference does (``src/rrtmg_lw.1col.f90``):
  * ``readprof`` (:759-1131) reads one ``$`` record set of ``INPUT_RRTM``: the control flags
    (format 9011), the surface record (9012), the optional ``dtbound`` (idrv = 1), the overlap
    records (icld = 4/5: idcor, decorr_con | juldat, lat), the layer records (form1/form2/form3,
    iform 0 or 1) and the optional cross-section block (``xsident`` :1296-1362);
  * ``readcld`` (:1139-1206) reads ``IN_CLD_RRTM``, ``readaer`` (:1209-1291) reads ``IN_AER_RRTM``;
  * the main program (:440-705) runs ``nmca`` = 200 McICA samples with ``permuteseed = ims * 140`` (``mcica_subcol_gen_lw.1col.f90:251``) and
    ``irng = 1`` (:417), averages them (:640-702), applies the idrv ``dtbound`` adjustment
    (:587-610) and writes ``OUTPUT_RRTM`` blocks (formats 9899-9958, :737-751).

Everything here is parsing, bookkeeping and formatting; the radiation itself is done by whatever
``backend`` the caller hands to :func:`run_case` -- :class:`GpuBackend` (the C ABI of the CUDA library)
in the product; the tests hand in a CPU checker with the same three methods.  Known, documented differences from the
column-mode program (they come from driving the GCM entry, not from this file):
  * the GCM ``inatm`` derives the dry-air column from the level pressures
    (``src/rrtmg_lw_rad.f90:807-809``) where column mode takes it from the file's broadening-gas
    column (:1013-1021), and sets the CO amount to zero (``rad:816``);
  * per-band output (``iout = 99`` / ``1..16``): every block is its own run of the library over bands
    ``istart..iend`` (``set_option("band_lo" / "band_hi")``; band 16 alone takes the ``totplk16`` Planck variant of
    ``src/rrtmg_lw_setcoef.f90:233-251``), :func:`run_blocks`;
  * ``imca = 0`` with clouds is the non-McICA path (``cldprop`` + ``rtrn`` for icld = 1 / ``rtrnmr`` otherwise,
    ``src/rrtmg_lw.1col.f90:498,557-572``): one run of the library's ``rrtmg_lw_nomcica`` entry;
  * ``iatm = 1``: the sounding is layered by ``rrtatm.py``, a restatement of the part of ``src/rrtatm.f`` the example
    suite uses (user profile on altitude levels, user boundaries in altitude, vertical path); anything else of LBLATM
    is refused with ``NotImplementedError``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

NBND = 16
NGPT = 140            # nsubclw = ngptlw (modules/parrrtm.f90)
NMCA = 200            # 1col:459-460

_XS_ALIASES = (       # xsident's alias table, 1col:1313-1320 (index = slot in wx: ccl4, cfc11, cfc12, cfc22)
    ("CCL4", "CCL3F", "CCL2F2", "CHCLF2"),
    (None, "CFCL3", "CF2CL2", "CHF2CL"),
    (None, "CFC11", "CFC12", "CFC22"),
    (None, "F11", "F12", "F22"),
)


# ----------------------------------------------------------------------------------------------
# Fortran formatted-read helpers
# ----------------------------------------------------------------------------------------------
def _int(s: str) -> int:
    s = s.strip()
    return int(s) if s else 0


def _real(s: str, d: int = 0) -> float:
    """One Fw.d / Ew.d / Gw.d input field: blank = 0, a field without a decimal point carries d
    implied decimals, 'D' exponents allowed."""
    s = s.strip().replace("D", "E").replace("d", "e")
    if not s:
        return 0.0
    if "." in s or "E" in s.upper():
        if "." not in s:                      # e.g. "123E2" -> implied decimals still apply to the mantissa
            m, e = s.upper().split("E")
            return float(m) * 10.0 ** (-d) * 10.0 ** int(e)
        return float(s)
    return int(s) * 10.0 ** (-d)


def _fields(line: str, start: int, width: int, n: int, d: int = 0):
    line = line.rstrip("\n")
    return [_real(line[start + k * width:start + (k + 1) * width], d) for k in range(n)]


@dataclass
class Case:
    """One ``$`` record set of INPUT_RRTM (+ IN_CLD_RRTM / IN_AER_RRTM), as ``readprof`` returns it."""
    iaer: int = 0
    iatm: int = 0
    ixsect: int = 0
    numangs: int = 0
    iout: int = 0
    idrv: int = 0
    imca: int = 0
    icld: int = 0
    tbound: float = 0.0
    iemiss: int = 0
    ireflect: int = 0
    semiss: list = field(default_factory=lambda: [1.0] * NBND)
    dtbound: float = 0.0
    idcor: int = 0
    decorr_con: float = 0.0
    juldat: int = 0
    lat: float = 0.0
    nlayers: int = 0
    nmol: int = 7
    pavel: list = field(default_factory=list)
    tavel: list = field(default_factory=list)
    altz: list = field(default_factory=list)    # km, nlayers+1
    pz: list = field(default_factory=list)
    tz: list = field(default_factory=list)
    wkl: list = field(default_factory=list)     # [layer][molecule] as read (mixing ratio or molecules/cm2)
    wbrodl: list = field(default_factory=list)
    xsname: list = field(default_factory=list)
    wx0: list = field(default_factory=list)     # [layer][species] as read
    # clouds (readcld) and aerosols (readaer)
    inflag: int = 0
    iceflag: int = 0
    liqflag: int = 0
    cldfrac: list = field(default_factory=list)
    clddat: list = field(default_factory=list)  # [layer][4]
    tauaer: list = field(default_factory=list)  # [layer][16]


def parse_in_cld(text: str, nlayers: int):
    """``readcld`` (1col:1139-1206): flags (3x,i2,4x,i1,4x,i1); records (a1,1x,i3,5e10.5) until '%'."""
    lines = text.split("\n")
    inflag, iceflag, liqflag = _int(lines[0][3:5]), _int(lines[0][9:10]), _int(lines[0][14:15])
    cldfrac = [0.0] * nlayers
    clddat = [[0.0] * 4 for _ in range(nlayers)]
    for ln in lines[1:]:
        if not ln.strip():
            continue
        if ln[0] == "%":
            break
        lay = _int(ln[2:5])
        v = _fields(ln, 5, 10, 5, d=5)
        if not 1 <= lay <= nlayers:
            raise ValueError("IN_CLD_RRTM: layer %d outside 1..%d" % (lay, nlayers))
        cldfrac[lay - 1] = v[0]
        clddat[lay - 1] = v[1:5]
    return inflag, iceflag, liqflag, cldfrac, clddat


def parse_in_aer(text: str, nlayers: int):
    """``readaer`` (1col:1209-1291): naer (3x,i2); per aerosol (2x,i3,4x,i1) then nlay records
    (2x,i3,16f7.4).  A layer named by two aerosols stops the reference; here it raises."""
    lines = text.split("\n")
    k = 0
    naer = _int(lines[k][3:5]); k += 1
    tau = [[0.0] * NBND for _ in range(nlayers)]
    seen = set()
    for _ in range(naer):
        nlay = _int(lines[k][2:5]); k += 1
        for _ in range(nlay):
            ln = lines[k]; k += 1
            lay = _int(ln[2:5])
            if lay in seen:
                raise ValueError("LAYER %d HAS MORE THAN ONE AEROSOL TYPE" % lay)
            aod = _fields(ln, 5, 7, NBND, d=4)
            if max(aod) >= 1.0e-10:
                seen.add(lay)
            tau[lay - 1] = aod
    return tau


def parse_input_rrtm(text: str, cld_text: str | None = None, aer_text: str | None = None, airmwt: float = 0.0) -> Case:
    """``readprof`` for the first ``$`` record set of an INPUT_RRTM file: layer records (iatm = 0) or a sounding that
    ``rrtatm`` layers (iatm = 1; ``airmwt``: see trap T18 in ``rrtatm.py`` -- 0.0 reproduces the reference program)."""
    lines = text.split("\n")
    i = next(k for k, ln in enumerate(lines) if ln[:1] == "$")
    c = Case()
    fl = lines[i + 1].ljust(95)
    c.iaer, c.iatm, c.ixsect = _int(fl[18:20]), _int(fl[49:50]), _int(fl[69:70])          # format 9011
    c.numangs, c.iout, c.idrv, c.imca, c.icld = (_int(fl[83:85]), _int(fl[87:90]), _int(fl[91:92]), _int(fl[93:94]),
                                                 _int(fl[94:95]))
    if c.numangs == -1:
        c.numangs = 0
    if c.iatm not in (0, 1):
        raise ValueError("iatm must be 0 or 1")
    k = i + 2
    rec = lines[k].ljust(95); k += 1                                                         # format 9012
    c.tbound, c.iemiss, c.ireflect = _real(rec[0:10], 3), _int(rec[11:12]), _int(rec[14:15])
    semis = _fields(rec, 15, 5, NBND, d=3)
    if c.idrv == 1:
        c.dtbound = _real(lines[k][0:10], 3); k += 1
    if c.icld in (4, 5):
        c.idcor = _int(lines[k][8:10]); k += 1                                               # 9014
        if c.idcor == 0:
            c.decorr_con = _real(lines[k][0:10], 3); k += 1                                  # 9015
        elif c.idcor == 1:
            c.juldat, c.lat = _int(lines[k][5:10]), _real(lines[k][10:20], 3); k += 1        # 9016
    for b in range(NBND):                                                                    # 1col:934-943
        if c.iemiss == 1 and semis[0] != 0.0:
            c.semiss[b] = semis[0]
        elif c.iemiss == 2 and semis[b] != 0.0:
            c.semiss[b] = semis[b]
    if c.iatm == 1:
        # the sounding is layered by rrtatm (1col:1000-1008); rrtatm.py restates the part of it the example suite uses
        from . import rrtatm
        if c.ixsect == 1:
            raise NotImplementedError("iatm = 1 with cross-sections (XAMNTS)")
        atm, k = rrtatm.layer_atmosphere(lines, k, airmwt=airmwt)
        c.nlayers, c.nmol = atm["nlayers"], atm["nmol"]
        c.pavel, c.tavel, c.pz, c.tz, c.altz = atm["pavel"], atm["tavel"], atm["pz"], atm["tz"], atm["altz"]
        c.wkl, c.wbrodl = atm["wkl"], atm["wbrodl"]
        return _finish_case(c, cld_text, aer_text)
    hdr = lines[k]; k += 1                                                                   # 9013
    iform, c.nlayers, c.nmol = _int(hdr[1:2]), _int(hdr[2:5]), _int(hdr[5:10]) or 7
    if iform == 1:
        w0, w1, o = 15, 10, 41      # g15.7,g10.4,g10.4,a3,i2,1x -> level fields start at column 41
        vw = 15
    else:
        w0, w1, o = 10, 10, 36      # 3f10.4,a3,i2,1x
        vw = 10
    nextra = (max(c.nmol, 7) - 7 + 7) // 8

    def read_amounts(k, n_first=7):
        vals = _fields(lines[k], 0, vw, 8, d=7 if vw == 15 else 3); k += 1
        more = []
        for _ in range(nextra):
            more += _fields(lines[k], 0, vw, 8, d=7 if vw == 15 else 3); k += 1
        return vals, more, k

    for l in range(c.nlayers):
        a = lines[k].ljust(o + 44); k += 1
        c.pavel.append(_real(a[0:w0], 7 if iform else 4))
        c.tavel.append(_real(a[w0:w0 + w1], 4))
        if l == 0:
            c.altz.append(_real(a[o:o + 7], 2)); c.pz.append(_real(a[o + 7:o + 15], 3)); c.tz.append(_real(a[o + 15:o + 22], 2))
        c.altz.append(_real(a[o + 22:o + 29], 2)); c.pz.append(_real(a[o + 29:o + 37], 3)); c.tz.append(_real(a[o + 37:o + 44], 2))
        vals, more, k = read_amounts(k)
        c.wkl.append(vals[:7] + more[:max(c.nmol - 7, 0)])
        c.wbrodl.append(vals[7])
    if c.ixsect == 1:
        nx = _int(lines[k][0:5]); k += 1                                                     # 9300
        names = []
        row = lines[k].ljust(80); k += 1
        names += [row[10 * j:10 * j + 10] for j in range(min(nx, 7))]
        if nx > 7:
            row = lines[k].ljust(80); k += 1
            names += [row[10 * j:10 * j + 10] for j in range(nx - 7)]
        c.xsname = [n.strip() for n in names]                                                # cljust: left-justify
        iformx = _int(lines[k][1:2]); k += 1                                                 # 9301
        xw = 15 if iformx == 1 else 10
        xextra = (max(nx, 7) - 7 + 7) // 8
        for l in range(c.nlayers):
            k += 1                                                                           # dummy layer line
            vals = _fields(lines[k], 0, xw, 8, d=7 if xw == 15 else 3); k += 1
            for _ in range(xextra):
                vals += _fields(lines[k], 0, xw, 8, d=7 if xw == 15 else 3); k += 1
            c.wx0.append((vals[:7] + vals[8:])[:nx])
    return _finish_case(c, cld_text, aer_text)


def _finish_case(c: Case, cld_text, aer_text) -> Case:
    """Tail of ``readprof``: surface temperature default, cloud and aerosol files."""
    if c.tbound < 0:
        c.tbound = c.tz[0]
    if c.icld >= 1:
        if cld_text is None:
            raise ValueError("icld >= 1 needs the IN_CLD_RRTM text")
        c.inflag, c.iceflag, c.liqflag, c.cldfrac, c.clddat = parse_in_cld(cld_text, c.nlayers)
    else:
        c.cldfrac = [0.0] * c.nlayers
        c.clddat = [[0.0] * 4 for _ in range(c.nlayers)]
    if c.iaer == 10:
        if aer_text is None:
            raise ValueError("iaer = 10 needs the IN_AER_RRTM text")
        c.tauaer = parse_in_aer(aer_text, c.nlayers)
    else:
        c.tauaer = [[0.0] * NBND for _ in range(c.nlayers)]
    return c


def xs_slot(name: str) -> int:
    """``xsident``: 1..4 = (ccl4, cfc11, cfc12, cfc22), 0 = not used by RRTMG_LW."""
    n = name.strip().upper()
    for row in _XS_ALIASES:
        for j, alias in enumerate(row):
            if alias is not None and n == alias:
                return j + 1
    return 0


# ----------------------------------------------------------------------------------------------
# Case -> GCM-entry inputs
# ----------------------------------------------------------------------------------------------
def case_to_batch(c: Case, device="cpu"):
    """The arrays ``rrtmg_lw`` / ``mcica_subcol_lw`` take, for one column (``synth.Batch``).
    Mixing ratios follow readprof's imix test (1col:1006-1036): if the first layer's amounts are all
    <= 1 they are mixing ratios relative to dry air already, otherwise molecules/cm2 divided by the
    dry-air column."""
    import torch
    from .synth import Batch
    f64 = torch.float64
    L = c.nlayers
    wkl = np.array([w[:7] for w in c.wkl], dtype=np.float64)            # (L, 7)
    wbrodl = np.array(c.wbrodl, dtype=np.float64)
    imix = 0 if (np.array(c.wkl[0][:c.nmol]) > 1.0).any() else 1
    summol = np.array([sum(w[1:c.nmol]) for w in c.wkl])
    if imix == 1:
        vmr = wkl
        coldry = wbrodl / (1.0 - summol)
    else:
        coldry = wbrodl + summol
        vmr = wkl / coldry[:, None]
    xs = np.zeros((L, 4))
    if c.ixsect == 1 and c.wx0:
        wx0 = np.array(c.wx0, dtype=np.float64)
        imixx = 1 if wx0[0, 0] <= 1.0 else 0
        for ix, name in enumerate(c.xsname):
            s = xs_slot(name)
            if s:
                xs[:, s - 1] = wx0[:, ix] if imixx == 1 else wx0[:, ix] / coldry

    def t(a):
        return torch.tensor(np.ascontiguousarray(a), dtype=f64, device=device)

    def lay(v):
        return t(np.asarray(v, dtype=np.float64).reshape(-1, 1))

    dat = np.array(c.clddat, dtype=np.float64).reshape(L, 4)
    if c.inflag == 0:                                                    # 1col:1093-1112
        tauc = np.repeat(dat[:, 0:1], NBND, axis=1)
        ciwp = clwp = rei = rel = np.zeros(L)
    else:
        tauc = np.zeros((L, NBND))
        fice = dat[:, 1]
        ciwp, clwp, rei, rel = dat[:, 0] * fice, dat[:, 0] * (1.0 - fice), dat[:, 2], dat[:, 3]
    altz = np.array(c.altz, dtype=np.float64)
    b = Batch(ncol=1, nlay=L, play=lay(c.pavel), plev=lay(c.pz), tlay=lay(c.tavel), tlev=lay(c.tz),
              tsfc=t([c.tbound]), h2ovmr=lay(vmr[:, 0]), co2vmr=lay(vmr[:, 1]), o3vmr=lay(vmr[:, 2]),
              n2ovmr=lay(vmr[:, 3]), ch4vmr=lay(vmr[:, 5]), o2vmr=lay(vmr[:, 6]),
              ccl4vmr=lay(xs[:, 0]), cfc11vmr=lay(xs[:, 1]), cfc12vmr=lay(xs[:, 2]), cfc22vmr=lay(xs[:, 3]),
              emis=t(np.array(c.semiss).reshape(NBND, 1)),
              tauaer=t(np.array(c.tauaer, dtype=np.float64).T.reshape(NBND, L, 1)),
              cldfrac=lay(c.cldfrac), ciwp=lay(ciwp), clwp=lay(clwp), rei=lay(rei), rel=lay(rel),
              tauc=t(tauc.reshape(L, 1, NBND)), lat=t([c.lat]), dz=lay((altz[1:] - altz[:-1]) * 1000.0),
              meta=dict(colmode=True, imix=imix))
    return b


# ----------------------------------------------------------------------------------------------
# Driver loop
# ----------------------------------------------------------------------------------------------
class GpuBackend:
    """The CUDA library behind its C ABI (``driver.py`` wrappers); ``device=None`` = host pointers."""

    def __init__(self, device=None):
        self.device = device

    @staticmethod
    def _np(o):
        return {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else v) for k, v in o.items()}

    def heatfac(self):
        """rrlw_con's heatfac as rrtmg_lw_ini set it (src/rrtmg_lw_init.f90:298)."""
        from . import api
        return float(api.get_table("heatfac")[0])

    def band_range(self, istart=1, iend=16):
        """Bands istart..iend of the following runs (the program's istart / iend of a per-band output block,
        src/rrtmg_lw.1col.f90:463-467); band 16 alone takes its own Planck table (setcoef.f90:233-251)."""
        from . import api
        api.set_option("band_lo", 1)
        api.set_option("band_hi", int(iend))
        api.set_option("band_lo", int(istart))

    def clear(self, b, idrv):
        from . import driver
        return self._np(driver.run_clear_or_strict(b, icld=0, idrv=idrv, device=self.device))

    def get_alpha(self, b, icld, idcor, decorr_con, juldat):
        from . import api
        alpha = np.zeros((b.nlay, 1))
        cv = lambda x: np.ascontiguousarray(x.cpu().numpy())  # noqa: E731
        api.get_alpha(1, 1, b.nlay, icld, idcor, decorr_con, cv(b.dz), cv(b.lat), juldat, cv(b.cldfrac), alpha)
        return alpha

    def nomcica(self, b, icld, idrv, flags):
        from . import driver
        return self._np(driver.run_nomcica(b, icld=icld, idrv=idrv, inflglw=flags[0], iceflglw=flags[1],
                                           liqflglw=flags[2], device=self.device))

    def mcica(self, b, icld, idrv, permuteseed, irng, flags, alpha):
        from . import driver
        return self._np(driver.run_fused(b, icld=icld, idrv=idrv, permuteseed=permuteseed, irng=irng, inflglw=flags[0],
                                         iceflglw=flags[1], liqflglw=flags[2], alpha=alpha, device=self.device))


# band limits in cm-1, rrlw_wvn wavenum1 / wavenum2 (src/rrtmg_lw_init.f90:215-220)
WAVENUM1 = (10., 350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600.)
WAVENUM2 = (350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600., 3250.)


def output_blocks(iout: int):
    """(istart, iend, wn1, wn2) of every OUTPUT_RRTM block the program writes for a given ``iout``
    (doc/rrtmg_lw_instructions.txt; src/rrtmg_lw.1col.f90:453-467, 704-711): 0 = the broadband block, n in 1..16 =
    band n alone, 99 = the broadband block followed by the 16 bands, < 0 = nothing."""
    if iout < 0:
        return []
    blocks = []
    if iout == 0 or iout == 99:
        blocks.append((1, 16, WAVENUM1[0], WAVENUM2[15]))
    if 1 <= iout <= 16:
        blocks.append((iout, iout, WAVENUM1[iout - 1], WAVENUM2[iout - 1]))
    if iout == 99:
        blocks += [(b, b, WAVENUM1[b - 1], WAVENUM2[b - 1]) for b in range(1, 17)]
    return blocks


def run_blocks(c: Case, backend, nmca: int | None = None, irng: int = 1):
    """Every output block of the case (17 for ``iout = 99``): list of :func:`run_case` results with ``wn1`` / ``wn2`` /
    ``istart`` / ``iend`` added.  Each block is its own run over bands istart..iend, like the program's
    ``goto 1000`` loop, McICA sampling included."""
    out = []
    try:
        for istart, iend, wn1, wn2 in output_blocks(c.iout):
            backend.band_range(istart, iend)
            r = run_case(c, backend, nmca=nmca, irng=irng)
            r.update(wn1=wn1, wn2=wn2, istart=istart, iend=iend)
            out.append(r)
    finally:
        backend.band_range(1, 16)
    return out


def format_blocks(blocks, iplon: int = 1) -> str:
    return "".join(format_output(r, iplon=iplon, wn1=r["wn1"], wn2=r["wn2"]) for r in blocks)


def run_case(c: Case, backend, nmca: int | None = None, irng: int = 1):
    """The main program's loop for one atmosphere and ONE band range (1col:440-705; the backend's current
    ``band_range``, all 16 bands by default): returns a dict with the level arrays ``uflx, dflx, fnet`` (nlayers+1)
    and ``htr`` (nlayers+1, top entry 0), McICA-averaged over ``nmca`` samples (200 in the reference) and
    ``dtbound``-adjusted when idrv = 1."""
    b = case_to_batch(c)
    L = c.nlayers
    n = 1 if c.imca == 0 else (NMCA if nmca is None else nmca)
    pz = np.array(c.pz)
    acc = {k: np.zeros(L + 1) for k in ("uflx", "dflx", "fnet", "htr")}
    for ims in range(1, n + 1):
        if c.imca == 1 and c.icld >= 1:
            alpha = None
            if c.icld in (4, 5):
                alpha = backend.get_alpha(b, c.icld, c.idcor, c.decorr_con, c.juldat)
            # the column-mode generator seeds every sample with permuteseed = ims * nsubclw
            # (src/mcica_subcol_gen_lw.1col.f90:251, nsubclw = ngptlw = 140)
            o = backend.mcica(b, c.icld, c.idrv, ims * NGPT, irng, (c.inflag, c.iceflag, c.liqflag), alpha)
        elif c.icld >= 1:
            # grid-mean clouds, no sub-columns: cldprop + rtrn (icld = 1) / rtrnmr (1col:498, 557-572)
            o = backend.nomcica(b, c.icld, c.idrv, (c.inflag, c.iceflag, c.liqflag))
        else:
            o = backend.clear(b, c.idrv)
        uflx = np.asarray(o["uflx"], dtype=np.float64).reshape(L + 1).copy()
        dflx = np.asarray(o["dflx"], dtype=np.float64).reshape(L + 1)
        htr = np.concatenate([np.asarray(o["hr"], dtype=np.float64).reshape(L), [0.0]])
        fnet = uflx - dflx
        if c.idrv == 1:                                                  # 1col:587-610
            d = np.asarray(o["duflx_dt"], dtype=np.float64).reshape(L + 1)
            heatfac = backend.heatfac()
            uflx = uflx + d * c.dtbound
            fnet = uflx - dflx
            htr = np.concatenate([heatfac * (fnet[:-1] - fnet[1:]) / (pz[:-1] - pz[1:]), [0.0]])
        for k, v in (("uflx", uflx), ("dflx", dflx), ("fnet", fnet), ("htr", htr)):
            acc[k] += v
    for k in acc:
        acc[k] = acc[k] / n
    acc["pz"] = pz
    acc["nmca"] = n
    return acc


# ----------------------------------------------------------------------------------------------
# OUTPUT_RRTM
# ----------------------------------------------------------------------------------------------
def _f(x: float, w: int, d: int) -> str:
    """Fortran Fw.d: the leading zero goes when the field is too narrow for it."""
    s = "%.*f" % (d, x)
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def format_output(res: dict, iplon: int = 1, wn1: float = 10.0, wn2: float = 3250.0) -> str:
    """One output block with the reference's formats 9899-9958 (1col:737-751): level, pressure,
    upward / downward / net flux, heating rate, top level first."""
    out = [" Wavenumbers: %6.1f - %6.1f cm-1, ATM %6d" % (wn1, wn2, iplon),
           " LEVEL    PRESSURE   UPWARD FLUX   DOWNWARD FLUX    NET FLUX       HEATING RATE",
           "             mb          W/m2          W/m2           W/m2          degree/day"]
    pz = res["pz"]
    for i in range(len(pz) - 1, -1, -1):
        p = pz[i]
        if p < 1.0e-2:
            ps = " " * 9 + _f(p, 7, 6) + " " * 3
        elif p < 1.0e-1:
            ps = " " * 9 + _f(p, 6, 5) + " " * 4
        elif p < 1.0:
            ps = " " * 8 + _f(p, 6, 4) + " " * 5
        elif p < 10.0:
            ps = " " * 7 + _f(p, 6, 3) + " " * 6
        elif p < 100.0:
            ps = " " * 6 + _f(p, 6, 2) + " " * 7
        else:
            ps = " " * 5 + _f(p, 6, 1) + " " * 8
        out.append(" %3d" % i + ps + _f(res["uflx"][i], 8, 4) + " " * 6 + _f(res["dflx"][i], 8, 4) + " " * 6 +
                   _f(res["fnet"][i], 12, 7) + " " * 10 + _f(res["htr"][i], 9, 5))
    out.append("\f")
    return "\n".join(out) + "\n"


def parse_output_rrtm(text: str):
    """All blocks of an OUTPUT_RRTM file: list of dict(wn1, wn2, atm, level, pz, uflx, dflx, fnet, htr),
    arrays ordered surface (level 0) first."""
    blocks = []
    cur = None
    for ln in text.split("\n"):
        s = ln.replace("\f", "").strip()
        if s.startswith("Wavenumbers:"):
            p = s.replace("Wavenumbers:", "").replace("cm-1, ATM", " ").replace(" - ", " ").split()
            cur = dict(wn1=float(p[0]), wn2=float(p[1]), atm=int(p[2]), rows=[])
            blocks.append(cur)
            continue
        p = s.split()
        if cur is not None and len(p) == 6 and p[0].isdigit():
            cur["rows"].append([float(x) for x in p])
    for b in blocks:
        rows = np.array(sorted(b.pop("rows"), key=lambda r: r[0]))
        b["level"] = rows[:, 0].astype(int)
        for j, k in enumerate(("pz", "uflx", "dflx", "fnet", "htr"), start=1):
            b[k] = rows[:, j]
    return blocks
