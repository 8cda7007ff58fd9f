"""Host-side mirror of the reference's GCM interface over the C ABI (ctypes).

Same names and argument meaning as the Fortran entry points the library replaces:

* ``rrtmg_lw_ini(cpdair)``                 -- ``src/rrtmg_lw_init.f90:47``
* ``rrtmg_lw(ncol, nlay, icld, idrv, ...)``  -- ``src/rrtmg_lw_rad.f90:99-108``
* ``mcica_subcol_lw(iplon, ncol, nlay, ...)`` -- ``src/mcica_subcol_gen_lw.f90:183-185``
* ``get_alpha(iplon, ncol, nlayers, ...)``   -- ``src/mcica_subcol_gen_lw.f90:68-69``

plus ``rrtmg_lw_mcica`` (generator + radiation in one device pass).  Arrays are either numpy
float64 (host pointers; the call stages them through the GPU) or CUDA ``torch`` float64 tensors
(device pointers; nothing is copied).  Layout is the reference's Fortran layout seen from C:
``(nlay, ncol)`` row-major == Fortran ``(ncol, nlay)``; McICA arrays ``(nlay, ncol, 140)``;
``emis (16, ncol)``; ``tauaer (16, nlay, ncol)``; ``tauc (nlay, ncol, 16)``.

There is no CPU path: if ``librrtmg_lw_b200.so`` is missing or no CUDA device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RRTMG_LW_B200_LIB selects an alternative build of the same CUDA library (e.g. the FMA-contracted
# variant librrtmg_lw_b200_fma.so); it is never a non-CUDA implementation.
LIB_PATH = os.environ.get("RRTMG_LW_B200_LIB") or os.path.join(_HERE, "lib", "librrtmg_lw_b200.so")

HOST, DEVICE = 0, 1
_dp = C.POINTER(C.c_double)
_lib = None


class RrtmgLwError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        msg = lib().rrtmg_lw_b200_strerror(code).decode()
        super().__init__("rrtmg_lw_b200 error %d%s: %s" % (code, (" in " + where) if where else "", msg))


_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
# argument and result types of every entry point of include/rrtmg_lw_b200.h (tests/test_abi_cpu.py checks the
# argument counts against the header, so a mis-ordered or missing argument fails on the CPU already)
SIGNATURES = {
    "rrtmg_lw_b200_kdata_put": (C.c_int, [C.c_int, C.c_char_p, _dp, C.c_long]),
    "rrtmg_lw_b200_ini": (C.c_int, [C.c_double, _dp, C.c_long, C.c_int]),
    "rrtmg_lw_b200_tables_build": (C.c_long, [C.c_double, _dp, C.c_long, C.c_int, C.c_void_p, C.c_long]),
    "rrtmg_lw_b200_ini_tables": (C.c_int, [C.c_void_p, C.c_long]),
    "rrtmg_lw_b200_rrtmg_lw": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, C.c_int] + [_dp] * 16 + [C.c_int] * 3 +
                               [_dp] * 15 + [C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_rrtmg_lw_nomcica": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, C.c_int] + [_dp] * 16 + [C.c_int] * 3 +
                                       [_dp] * 15 + [C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_rrtmg_lw_mcica": (C.c_int, [C.c_int] * 7 + [_dp] * 16 + [C.c_int] * 3 + [_dp] * 16 +
                                     [C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_mcica_subcol_lw": (C.c_int, [C.c_int] * 6 + [_ip] + [_dp] * 14 + [C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_get_alpha": (C.c_int, [C.c_int] * 6 + [C.c_double, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_sync": (C.c_int, [_ip, _ip]),
    "rrtmg_lw_b200_last_error": (C.c_int, [_ip, _ip]),
    "rrtmg_lw_b200_debug_capture": (C.c_int, [C.c_int]),
    "rrtmg_lw_b200_debug_stages": (C.c_int, [C.c_int, _ip, _dp, _dp, _dp, _ip, _dp, C.POINTER(C.c_ubyte)]),
    "rrtmg_lw_b200_get_table": (C.c_long, [C.c_char_p, _dp, C.c_long]),
    "rrtmg_lw_b200_set_chunk": (C.c_int, [C.c_int]),
    "rrtmg_lw_b200_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "rrtmg_lw_b200_set_device": (C.c_int, [C.c_int]),
    "rrtmg_lw_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rrtmg_lw_b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_gather": (C.c_int, [_dp, C.c_long, _dp, C.c_int, C.c_void_p]),
    "rrtmg_lw_b200_comm_finalize": (C.c_int, []),
    "rrtmg_lw_b200_kiss_jump_selftest": (C.c_int, []),
    "rrtmg_lw_b200_mt_jump_selftest": (C.c_int, []),
    "rrtmg_lw_b200_launch_count": (C.c_long, [C.c_int]),
    "rrtmg_lw_b200_set_timing": (C.c_int, [C.c_int]),
    "rrtmg_lw_b200_last_timing": (C.c_int, [_fp, _fp]),
    "rrtmg_lw_b200_last_band_timing": (C.c_int, [_fp]),
    "rrtmg_lw_b200_last_stage_timing": (C.c_int, [_fp, C.c_int]),
    "rrtmg_lw_b200_fp64_peak": (C.c_int, [C.POINTER(C.c_double)]),
    "rrtmg_lw_b200_strerror": (C.c_char_p, [C.c_int]),
    "rrtmg_lw_b200_version": (C.c_char_p, []),
    "rrtmg_lw_b200_finalize": (None, []),
}


def _declare(L):
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librrtmg_lw_b200.so not built -- run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _is_torch(a):
    return a is not None and hasattr(a, "data_ptr")


class col_offset:
    """Column ``lo`` onwards of a column-fastest (Fortran layout) device array, for running a column sub-range of a
    resident batch: pass it where the array goes and give the call ``ld`` = the array's full column extent.
    ``inner`` = elements per column ahead of the column axis (16 for ``tauc``, 140 for the McICA arrays)."""
    is_cuda = True

    def __init__(self, t, lo: int, inner: int = 1):
        assert t.is_cuda and t.is_contiguous() and t.element_size() == 8 and t.dtype.is_floating_point
        self.base, self.dtype = t, t.dtype
        self._p = t.data_ptr() + int(lo) * int(inner) * 8

    def data_ptr(self):
        return self._p

    def element_size(self):
        return 8

    def is_contiguous(self):
        return True


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        assert a.dtype.is_floating_point and a.element_size() == 8 and a.is_contiguous(), "float64 contiguous tensor"
        return C.cast(C.c_void_p(a.data_ptr()), _dp)
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "float64 C-contiguous array"
    return a.ctypes.data_as(_dp)


def _space(arrs):
    dev = [a for a in arrs if _is_torch(a) and a.is_cuda]
    if dev:
        assert all((a is None) or (_is_torch(a) and a.is_cuda) for a in arrs), "mixing host and device arrays"
        return DEVICE
    assert all((a is None) or isinstance(a, np.ndarray) for a in arrs), "host arrays must be numpy"
    return HOST


def _stream(space):
    if space == DEVICE:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return None


def rrtmg_lw_ini(cpdair: float, kdata: np.ndarray, default_real_kind: int = 4):
    """Replaces ``rrtmg_lw_ini(cpdair)``.  ``kdata`` = unreduced k-distribution blob
    (``rrtmg_lw_b200.kdata``; the reference compiles these numbers in)."""
    if kdata is None:   # arrays were staged one by one with kdata_put
        rc = lib().rrtmg_lw_b200_ini(float(cpdair), None, 0, int(default_real_kind))
    else:
        k = np.ascontiguousarray(kdata, dtype=np.float64)
        rc = lib().rrtmg_lw_b200_ini(float(cpdair), _ptr(k), k.size, int(default_real_kind))
    if rc:
        raise RrtmgLwError(rc, "rrtmg_lw_ini")


def tables_build(cpdair: float, kdata: np.ndarray, default_real_kind: int = 4) -> bytes:
    """Host-only half of ``rrtmg_lw_ini``: the reduced tables as a versioned binary image (no GPU needed)."""
    a = np.ascontiguousarray(kdata, dtype=np.float64)
    n = lib().rrtmg_lw_b200_tables_build(float(cpdair), _ptr(a), a.size, int(default_real_kind), None, 0)
    if n < 0:
        raise RrtmgLwError(-n, "tables_build")
    buf = C.create_string_buffer(n)
    m = lib().rrtmg_lw_b200_tables_build(float(cpdair), _ptr(a), a.size, int(default_real_kind), buf, n)
    if m != n:
        raise RrtmgLwError(7, "tables_build")
    return buf.raw


def rrtmg_lw_ini_tables(blob: bytes):
    """``rrtmg_lw_ini`` from an image made by :func:`tables_build` / ``tools/make_table_blob.py``."""
    rc = lib().rrtmg_lw_b200_ini_tables(blob, len(blob))
    if rc:
        raise RrtmgLwError(rc, "rrtmg_lw_ini_tables")


def kdata_put(band: int, name: str, array: np.ndarray):
    """Hand over one unreduced k-array (Fortran element order) -- what the Fortran shim does after
    the reference's ``lw_kgbNN``; ``rrtmg_lw_ini(cpdair, None)`` then consumes the staged arrays."""
    a = np.ascontiguousarray(np.asarray(array, dtype=np.float64).ravel(order="F"))
    rc = lib().rrtmg_lw_b200_kdata_put(int(band), name.encode(), _ptr(a), a.size)
    if rc:
        raise RrtmgLwError(rc, "kdata_put(%d,%s)" % (band, name))


def rrtmg_lw(ncol, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
             cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, inflglw, iceflglw, liqflglw, cldfmcl, taucmcl, ciwpmcl,
             clwpmcl, reicmcl, relqmcl, tauaer, uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt=None, duflxc_dt=None,
             ld=None):
    """Replaces ``rrtmg_lw``; outputs are written in place.  Returns the (possibly clamped) icld."""
    arrs = [play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr,
            ccl4vmr, emis, cldfmcl, taucmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, tauaer, uflx, dflx, hr, uflxc, dflxc,
            hrc, duflx_dt, duflxc_dt]
    space = _space(arrs)
    ic = C.c_int(int(icld))
    p = [_ptr(a) for a in arrs]
    rc = lib().rrtmg_lw_b200_rrtmg_lw(int(ncol), int(ld or ncol), int(nlay), C.byref(ic), int(idrv), *p[:16],
                                      int(inflglw), int(iceflglw), int(liqflglw), *p[16:], space, _stream(space))
    if rc:
        raise RrtmgLwError(rc, "rrtmg_lw")
    return ic.value


def rrtmg_lw_nomcica(ncol, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
                     cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, inflglw, iceflglw, liqflglw, cldfr, taucld, cicewp,
                     cliqwp, reice, reliq, tauaer, uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt=None, duflxc_dt=None,
                     ld=None):
    """Replaces ``rrtmg_lw`` of ``src/rrtmg_lw_rad.nomcica.f90:99-108`` (grid-mean clouds, no McICA: random overlap
    for icld = 1, maximum-random otherwise).  Returns the (possibly clamped) icld."""
    arrs = [play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr,
            ccl4vmr, emis, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer, uflx, dflx, hr, uflxc, dflxc,
            hrc, duflx_dt, duflxc_dt]
    space = _space(arrs)
    ic = C.c_int(int(icld))
    p = [_ptr(a) for a in arrs]
    rc = lib().rrtmg_lw_b200_rrtmg_lw_nomcica(int(ncol), int(ld or ncol), int(nlay), C.byref(ic), int(idrv), *p[:16],
                                              int(inflglw), int(iceflglw), int(liqflglw), *p[16:], space,
                                              _stream(space))
    if rc:
        raise RrtmgLwError(rc, "rrtmg_lw_nomcica")
    return ic.value


def rrtmg_lw_mcica(ncol, nlay, icld, idrv, permuteseed, irng, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr,
                   ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, inflglw, iceflglw, liqflglw,
                   cldfrac, ciwp, clwp, rei, rel, tauc, alpha, tauaer, uflx, dflx, hr, uflxc, dflxc, hrc,
                   duflx_dt=None, duflxc_dt=None, ld=None):
    """``mcica_subcol_lw`` + ``rrtmg_lw`` fused on the device (sub-columns never materialised)."""
    a1 = [play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr,
          ccl4vmr, emis]
    a2 = [cldfrac, ciwp, clwp, rei, rel, tauc, alpha, tauaer, uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt]
    space = _space(a1 + a2)
    rc = lib().rrtmg_lw_b200_rrtmg_lw_mcica(int(ncol), int(ld or ncol), int(nlay), int(icld), int(idrv),
                                            int(permuteseed), int(irng), *[_ptr(a) for a in a1], int(inflglw),
                                            int(iceflglw), int(liqflglw), *[_ptr(a) for a in a2], space, _stream(space))
    if rc:
        raise RrtmgLwError(rc, "rrtmg_lw_mcica")


def mcica_subcol_lw(iplon, ncol, nlay, icld, permuteseed, irng, play, cldfrac, ciwp, clwp, rei, rel, tauc, alpha,
                    cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl, ld=None):
    """Replaces ``mcica_subcol_lw``; returns the (possibly normalised) irng."""
    arrs = [play, cldfrac, ciwp, clwp, rei, rel, tauc, alpha, cldfmcl, ciwpmcl, clwpmcl, reicmcl, relqmcl, taucmcl]
    space = _space(arrs)
    ir = C.c_int(int(irng))
    rc = lib().rrtmg_lw_b200_mcica_subcol_lw(int(iplon), int(ncol), int(ld or ncol), int(nlay), int(icld),
                                             int(permuteseed), C.byref(ir), *[_ptr(a) for a in arrs], space,
                                             _stream(space))
    if rc:
        raise RrtmgLwError(rc, "mcica_subcol_lw")
    return ir.value


def get_alpha(iplon, ncol, nlayers, icld, idcor, decorr_con, dz, lat, juldat, cldfrac, alpha, ld=None):
    """Replaces ``get_alpha``."""
    arrs = [dz, lat, cldfrac, alpha]
    space = _space(arrs)
    rc = lib().rrtmg_lw_b200_get_alpha(int(iplon), int(ncol), int(ld or ncol), int(nlayers), int(icld), int(idcor),
                                       float(decorr_con), _ptr(dz), _ptr(lat), int(juldat), _ptr(cldfrac),
                                       _ptr(alpha), space, _stream(space))
    if rc:
        raise RrtmgLwError(rc, "get_alpha")


def comm_unique_id() -> bytes:
    """Rank 0: the 128-byte NCCL id every rank passes to :func:`comm_init` (broadcast it with the host's own means)."""
    buf = C.create_string_buffer(128)
    rc = lib().rrtmg_lw_b200_comm_unique_id(buf)
    if rc:
        raise RrtmgLwError(rc, "comm_unique_id")
    return buf.raw


def comm_init(nranks: int, rank: int, unique_id: bytes):
    rc = lib().rrtmg_lw_b200_comm_init(int(nranks), int(rank), unique_id)
    if rc:
        raise RrtmgLwError(rc, "comm_init")


def gather(send, recv, root: int = 0):
    """Gather every rank's ``send`` (CUDA float64 tensor) into ``recv`` (root only: ``(nranks,) + send.shape``) through
    the library's NCCL gather, on the current torch stream."""
    rc = lib().rrtmg_lw_b200_gather(_ptr(send), int(send.numel()), _ptr(recv) if recv is not None else None, int(root),
                                    _stream(DEVICE))
    if rc:
        raise RrtmgLwError(rc, "gather")


def comm_finalize():
    lib().rrtmg_lw_b200_comm_finalize()


def sync():
    col, lay = C.c_int(0), C.c_int(0)
    rc = lib().rrtmg_lw_b200_sync(C.byref(col), C.byref(lay))
    if rc:
        e = RrtmgLwError(rc, "device status (column %d, layer %d)" % (col.value, lay.value))
        e.col, e.lay = col.value, lay.value
        raise e


def set_device(device):
    """Bind the following calls of this process to CUDA device ``device`` (-1: follow the current device)."""
    rc = lib().rrtmg_lw_b200_set_device(int(device))
    if rc:
        raise RrtmgLwError(rc, "set_device")


def set_chunk(n):
    rc = lib().rrtmg_lw_b200_set_chunk(int(n))
    if rc:
        raise RrtmgLwError(rc, "set_chunk")


def set_option(name, value):
    """Tuning switches of the library (``rrtmg_lw_b200_set_option``), e.g. ``("band_kernel", 2)``."""
    rc = lib().rrtmg_lw_b200_set_option(str(name).encode(), int(value))
    if rc:
        raise RrtmgLwError(rc, "set_option(%s)" % name)


def set_mt_block(global_ncol: int, col_offset: int = 0):
    """``irng = 1`` on a sharded batch: the columns of the following calls are columns
    ``[col_offset, col_offset + ncol)`` of a batch of ``global_ncol`` columns, and the library steps the single
    Mersenne-Twister stream of that batch (sub-column, column, layer order,
    ``src/mcica_subcol_gen_lw.f90:523-530``) over the words of the other blocks, so every rank reproduces exactly
    what one unsharded call would draw.  ``set_mt_block(0)`` switches it off."""
    set_option("mt_global_ncol", int(global_ncol))
    set_option("mt_col_offset", int(col_offset) if global_ncol else 0)


def launch_count(reset=False):
    return int(lib().rrtmg_lw_b200_launch_count(1 if reset else 0))


def set_timing(on):
    lib().rrtmg_lw_b200_set_timing(1 if on else 0)


def last_timing():
    a, b = C.c_float(0), C.c_float(0)
    lib().rrtmg_lw_b200_last_timing(C.byref(a), C.byref(b))
    return a.value, b.value


def last_band_timing():
    """CUDA-event milliseconds of each band's kernel in the last call made with timing enabled."""
    a = (C.c_float * 16)()
    lib().rrtmg_lw_b200_last_band_timing(a)
    return [float(x) for x in a]


def last_stage_timing():
    """CUDA-event milliseconds of each timed stage between the prep kernels and the finalize kernel of the last call
    made with timing enabled (two-phase path: taumol, then every sweep kernel in launch order)."""
    a = (C.c_float * 32)()
    n = lib().rrtmg_lw_b200_last_stage_timing(a, 32)
    return [float(a[k]) for k in range(max(0, min(n, 32)))]


def last_error():
    """(column, layer), 1-based, of the input error most recently reported (column of the CALL)."""
    col, lay = C.c_int(0), C.c_int(0)
    lib().rrtmg_lw_b200_last_error(C.byref(col), C.byref(lay))
    return col.value, lay.value


def debug_capture(col):
    """Capture the stage dump of global column ``col`` (0-based) on the next call; None disables."""
    lib().rrtmg_lw_b200_debug_capture(0 if col is None else int(col) + 1)


def debug_stages(col, nlay, want_mask=False):
    ints6 = np.zeros((nlay, 6), dtype=np.int32)
    facs = np.zeros((nlay, 4))
    taug = np.zeros((140, nlay)); fracs = np.zeros((140, nlay))
    itr = np.zeros((140, nlay, 2), dtype=np.int32)
    scal = np.zeros(2)
    mask = np.zeros((140, nlay), dtype=np.uint8) if want_mask else None
    ip = C.POINTER(C.c_int)
    rc = lib().rrtmg_lw_b200_debug_stages(int(col), ints6.ctypes.data_as(ip), _ptr(facs), _ptr(taug), _ptr(fracs),
                                          itr.ctypes.data_as(ip), _ptr(scal),
                                          mask.ctypes.data_as(C.POINTER(C.c_ubyte)) if want_mask else None)
    if rc:
        raise RrtmgLwError(rc, "debug_stages")
    return dict(ints6=ints6, facs=facs, taug=taug, fracs=fracs, itr=itr, laytrop=int(scal[0]), pwvcm=scal[1], mask=mask)


def get_table(name):
    n = lib().rrtmg_lw_b200_get_table(name.encode(), None, 0)
    out = np.zeros(n)
    if n:
        lib().rrtmg_lw_b200_get_table(name.encode(), _ptr(out), n)
    return out


def fp64_peak():
    """Measured FP64 FMA peak of the current device in TFLOP/s."""
    v = C.c_double(0)
    rc = lib().rrtmg_lw_b200_fp64_peak(C.byref(v))
    if rc:
        raise RrtmgLwError(rc, "fp64_peak")
    return v.value


def finalize():
    lib().rrtmg_lw_b200_finalize()
