"""The C-ABI library loads and exports every symbol include/rrtmg_lw_b200.h declares (no
compute calls -- there is no GPU in the CPU test tier), and the host mirror refuses to run
without the CUDA path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "rrtmg_lw_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rrtmg_lw_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rrtmg_lw_b200 import api
    names = _declared()
    assert len(names) >= 14 and "rrtmg_lw_b200_rrtmg_lw" in names and "rrtmg_lw_b200_ini" in names
    lib = ctypes.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_strerror_and_version_do_not_need_a_gpu():
    from rrtmg_lw_b200 import api
    L = api.lib()
    assert b"ICE RADIUS" in L.rrtmg_lw_b200_strerror(2)
    assert b"LIQUID" in L.rrtmg_lw_b200_strerror(3)
    assert b"sm_100a" in L.rrtmg_lw_b200_version()


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_no_cpu_fallback():
    """Without a CUDA device init must fail with ERR_CUDA (5), and compute entries with
    ERR_NOT_INIT (6) -- never a silent CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rrtmg_lw_b200 import api, kdata
    with pytest.raises(api.RrtmgLwError) as e:
        api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob())
    assert e.value.code == 5
    z = np.zeros((4, 2))
    with pytest.raises(api.RrtmgLwError) as e2:
        api.get_alpha(1, 2, 4, 4, 0, 2500.0, z, np.zeros(2), 1, z, z.copy())
    assert e2.value.code == 6


def test_kdata_put_validates_name_and_size():
    """Array-by-array k-data hand-over (the Fortran shim's route): wrong size / unknown name are
    rejected with ERR_ARG (7); ini without the complete set is rejected too."""
    from rrtmg_lw_b200 import api, kdata
    arrs = kdata.synthetic_kdata()
    api.kdata_put(3, "kao", arrs[(3, "kao")])
    with pytest.raises(api.RrtmgLwError) as e:
        api.kdata_put(3, "kao", arrs[(3, "kao")][:10])
    assert e.value.code == 7
    with pytest.raises(api.RrtmgLwError) as e:
        api.kdata_put(3, "nonsense", arrs[(3, "kao")])
    assert e.value.code == 7
    with pytest.raises(api.RrtmgLwError) as e:
        api.rrtmg_lw_ini(1004.0, None)
    assert e.value.code == 7   # incomplete set


def test_set_option_validates_names_and_ranges():
    """rrtmg_lw_b200_set_option needs no device: every documented switch accepts its range and refuses the value next to
    it with ERR_ARG (7); unknown names are refused; the header documents every name the library knows."""
    from rrtmg_lw_b200 import api
    ranges = {"concurrency": (0, 5), "mcica_beside": (0, 4), "tail_group": (0, 1), "kiss_sparse": (0, 1),
              "sweep_groups": (0, 1)}
    defaults = {"concurrency": 1, "mcica_beside": 0, "tail_group": 1, "kiss_sparse": 1, "sweep_groups": 0}
    hdr = open(os.path.join(ROOT, "include", "rrtmg_lw_b200.h")).read()
    for name, (lo, hi) in ranges.items():
        assert '"%s"' % name in hdr, name
        try:
            for v in (lo, hi):
                api.set_option(name, v)
            for v in (lo - 1, hi + 1):
                with pytest.raises(api.RrtmgLwError) as e:
                    api.set_option(name, v)
                assert e.value.code == 7, (name, v)
        finally:
            api.set_option(name, defaults[name])
    with pytest.raises(api.RrtmgLwError) as e:
        api.set_option("no_such_switch", 1)
    assert e.value.code == 7
    with pytest.raises(api.RrtmgLwError):
        api.set_chunk(8)


def test_fortran_shim_binds_only_exported_symbols():
    """Every bind(C, name=...) of fortran/rrtmg_lw_b200_shim.f90 is an export of the library and
    every k-array it hands over exists in the layout header."""
    from rrtmg_lw_b200 import api, kdata
    txt = open(os.path.join(ROOT, "fortran", "rrtmg_lw_b200_shim.f90")).read()
    names = set(re.findall(r"bind\(C, name='([a-z0-9_]+)'\)", txt))
    assert len(names) >= 6
    lib = ctypes.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    puts = set((int(b), n) for b, n in re.findall(r"call put\((\d+), '([a-z0-9_]+)'", txt))
    assert puts == set((a[0], a[1]) for a in kdata.karrays())
    # the two rad files (McICA / non-McICA, one module name like the reference's two sources) call interfaces the
    # common file declares
    for f, c in (("rrtmg_lw_b200_shim_rad.f90", "c_rrtmg_lw"), ("rrtmg_lw_b200_shim_rad_nomcica.f90", "c_rrtmg_lw_nomcica")):
        t = open(os.path.join(ROOT, "fortran", f)).read()
        assert "module rrtmg_lw_rad" in t and "public :: rrtmg_lw, inatm" in t
        assert ("call check(%s(" % c) in t and ("function %s(" % c) in txt


def _dummy_list(text, name):
    """Dummy-argument names of `subroutine name (...)` in free-form Fortran, continuation lines joined."""
    m = re.search(r"subroutine\s+%s\s*&?\s*\(" % name, text, flags=re.I)
    assert m, name
    depth, k = 1, m.end()
    while depth:
        depth += {"(": 1, ")": -1}.get(text[k], 0)
        k += 1
    body = re.sub(r"!.*", "", text[m.end():k - 1])
    return [a.strip().lower() for a in body.replace("&", " ").replace("\n", " ").split(",")]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not mounted")
def test_fortran_shim_keeps_the_reference_dummy_argument_lists():
    """Same public names, same dummy arguments in the same order as the reference's GCM entry points (what makes the
    shim a relink-only replacement)."""
    ref = lambda f: open(os.path.join("/root/reference/src", f)).read()  # noqa: E731
    mine = lambda f: open(os.path.join(ROOT, "fortran", f)).read()  # noqa: E731
    pairs = [("rrtmg_lw_rad.f90", "rrtmg_lw_b200_shim_rad.f90", ("rrtmg_lw", "inatm")),
             ("rrtmg_lw_rad.nomcica.f90", "rrtmg_lw_b200_shim_rad_nomcica.f90", ("rrtmg_lw", "inatm")),
             ("mcica_subcol_gen_lw.f90", "rrtmg_lw_b200_shim.f90", ("get_alpha", "mcica_subcol_lw", "generate_stochastic_clouds")),
             ("rrtmg_lw_init.f90", "rrtmg_lw_b200_shim.f90", ("rrtmg_lw_ini",))]
    for rf, mf, names in pairs:
        for n in names:
            assert _dummy_list(mine(mf), n) == _dummy_list(ref(rf), n), (mf, n)


def test_kdata_header_matches_python_layout():
    from rrtmg_lw_b200 import kdata
    arr = kdata.karrays()
    assert len(arr) == 115 and kdata.KDATA_DOUBLES == 211856
    txt = open(os.path.join(ROOT, "include", "rrtmg_lw_kdata_layout.h")).read()
    assert "#define RRTMG_LW_KDATA_DOUBLES 211856" in txt


def test_product_does_not_reference_the_oracle():
    """Nothing under rrtmg_lw_b200/ may import, link or call the oracle."""
    pkg = os.path.join(ROOT, "rrtmg_lw_b200")
    for dp, _, files in os.walk(pkg):
        if "_generated" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                t = open(os.path.join(dp, f)).read()
                assert "oracle" not in t.lower().replace("oracle/_ref", ""), os.path.join(dp, f)


def test_kiss_jump_ahead_matches_stepping():
    """The closed-form n-step jumps behind the sub-column-parallel KISS kernel equal n single steps of
    kissvec (src/mcica_subcol_gen_lw.f90:731-741) for all four component generators (host-only check)."""
    from rrtmg_lw_b200 import api
    assert api.lib().rrtmg_lw_b200_kiss_jump_selftest() == 0


def test_fp32_variant_is_generated_from_the_same_source():
    """The optional FP32 mode compiles the SAME band-kernel source as float (tools/gen_fp32.py): the generated
    header must hold no double-typed kernel arithmetic outside the accumulator alias, and the one the library was
    built from must be what the generator makes of the current source."""
    import re
    import subprocess
    import sys
    import tempfile
    src = os.path.join(ROOT, "rrtmg_lw_b200", "csrc", "lw_band.cuh")
    gen = os.path.join(ROOT, "rrtmg_lw_b200", "csrc", "_generated", "lw_band32.cuh")
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "lw_band32.cuh")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_fp32.py"), src, out])
        body = open(out).read()
    if os.path.exists(gen):  # built tree: the header the library was compiled from is the current one
        assert body == open(gen).read(), "stale generated header"
    assert "namespace p32" in body and "#define LW_F32 1" in body
    code = "\n".join(l.split("//")[0] for l in body.split("\n"))
    assert not re.search(r"\bdouble2?\b", code), "double-typed code left in the float variant"
    assert "acc_t" in code and "ld.global.nc.v4.f32" in body


def test_mersenne_twister_jump_ahead_matches_stepping():
    """The GF(2) jump-ahead behind the parallel irng = 1 generator (lw_mtjump.cpp: characteristic polynomial of
    MT19937 by Berlekamp-Massey, x^(2^b) mod phi by squaring, one dot-product pass per set bit of the distance) lands on
    the same 624-word state as stepping the generator of src/mcica_random_numbers.f90:134-152 n times, for n up to
    10^6 (host-only check of the table the device kernel uses)."""
    from rrtmg_lw_b200 import api
    assert api.lib().rrtmg_lw_b200_mt_jump_selftest() == 0


def test_table_image_is_built_on_the_host_and_validated(kblob, tmp_path):
    """The host-side half of rrtmg_lw_ini as a tool (packaging item): the versioned table image is built
    without a GPU, is deterministic, depends on the inputs it should depend on, and `rrtmg_lw_b200_ini_tables` refuses a
    corrupt, truncated or foreign image with ERR_ARG before it ever looks for a device (ERR_CUDA on this CPU box for a
    valid one: there is no CPU path)."""
    import subprocess
    import sys
    from rrtmg_lw_b200 import api
    img = api.tables_build(1004.0, kblob, 4)
    assert img[:8] == b"RLWLTBL1" and int.from_bytes(img[8:16], "little") == 1 and len(img) % 8 == 0
    assert img == api.tables_build(1004.0, kblob, 4)
    assert img != api.tables_build(1004.0, kblob, 8)          # default-kind literals are part of the tables
    assert img != api.tables_build(1005.0, kblob, 4)          # heatfac
    with pytest.raises(api.RrtmgLwError) as e:
        api.tables_build(1004.0, kblob[:-1], 4)
    assert e.value.code == 7
    out = tmp_path / "tables.bin"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_table_blob.py"), "--kdata", "synthetic",
                           "--cpdair", "1004.0", "-o", str(out)])
    assert out.read_bytes() == img
    if not _has_gpu():
        with pytest.raises(api.RrtmgLwError) as e:
            api.rrtmg_lw_ini_tables(img)
        assert e.value.code == 5
    bad = bytearray(img); bad[len(img) // 2] ^= 0x10
    for blob in (bytes(bad), img[:-8], img[8:], b"\0" * 64, b""):
        with pytest.raises(api.RrtmgLwError) as e:
            api.rrtmg_lw_ini_tables(blob)
        assert e.value.code == 7
