"""Multi-GPU behind the C ABI: device selection (`rrtmg_lw_b200_set_device`) and the NCCL output gather
(`rrtmg_lw_b200_comm_*`, `rrtmg_lw_b200_gather`) driven from a torchrun launch with one rank per GPU; rank 0's gathered
blocks must equal a single-GPU run of the whole batch bit for bit.  Needs two GPUs (skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_abi.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_step_gathered_through_the_c_abi():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    n = 2 if n < 4 else 3   # three ranks when available: uneven blocks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu", "abi_gather_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ABI_GATHER_OK" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])


def test_gather_before_comm_init_is_refused(gpu):
    import torch
    from rrtmg_lw_b200 import api
    t = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    api.comm_finalize()
    with pytest.raises(api.RrtmgLwError) as e:
        api.gather(t, t, root=0)
    assert e.value.code == 6
