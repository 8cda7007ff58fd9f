import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kblob():
    from rrtmg_lw_b200 import kdata
    return kdata.synthetic_blob()


@pytest.fixture(scope="session")
def oracle(kblob):
    """CPU oracle initialised with the synthetic k-data, default-real-kind 4 (plain gfortran)."""
    from oracle import orc
    orc.build()
    orc.ini(kblob, 1004.0, 4)
    return orc


@pytest.fixture(scope="session")
def gpu(kblob):
    """The CUDA library initialised on cuda:0 with the same k-data; fails loudly without a GPU."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device; the library has no CPU fallback"
    from rrtmg_lw_b200 import api
    api.rrtmg_lw_ini(1004.0, kblob, 4)
    # RRTMG_LW_BAND_KERNEL=2|4 runs the whole suite against the fused per-band kernels / the two-phase path (default)
    bk = os.environ.get("RRTMG_LW_BAND_KERNEL")
    if bk:
        api.set_option("band_kernel", int(bk))
    sg = os.environ.get("RRTMG_LW_SWEEP_GROUPS")
    if sg:
        api.set_option("sweep_groups", int(sg))
    return api


@pytest.fixture(scope="session")
def std_atm():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "std_atm.json")))
