"""The C ABI driven by a non-Python caller (tests/c_caller/lw_c_caller.c): k-arrays staged one by one with
rrtmg_lw_b200_kdata_put in the shim's order, rrtmg_lw_b200_rrtmg_lw on host arrays with ld > ncol, result checked
against the CPU oracle inside the C program."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(ROOT, "tests", "c_caller")
EXE = os.path.join(DIR, "lw_c_caller")


def _build():
    from oracle import orc
    orc.build()
    subprocess.check_call(["make", "-C", DIR, "-s"])
    assert os.path.exists(EXE)


def test_c_caller_builds_and_fails_loudly_without_a_gpu():
    """gcc links the program against the CUDA library's C ABI; where no CUDA device exists the library must refuse
    (exit code 2 = library error, its message names the missing device) instead of computing anything on the CPU."""
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the gpu-marked test runs the program for real")
    p = subprocess.run([EXE, os.path.join(ROOT, "rrtmg_lw_b200", "data", "lw_static_tables.bin")], capture_output=True,
                       text=True, cwd=ROOT)
    assert p.returncode == 2, (p.returncode, p.stdout, p.stderr)
    assert "no CUDA device" in p.stderr or "CUDA" in p.stderr


@pytest.mark.gpu
def test_c_caller_matches_oracle_on_the_gpu():
    _build()
    p = subprocess.run([EXE, os.path.join(ROOT, "rrtmg_lw_b200", "data", "lw_static_tables.bin")], capture_output=True,
                       text=True, cwd=ROOT, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert "PASS" in p.stdout
