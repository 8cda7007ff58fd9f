"""N-GPU check of the sharded path on real hardware: block-partitioned fused McICA run + NCCL gather of the outputs
equals the single-GPU run bit for bit; times the gather of a 1 M-column-per-GPU output set."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from rrtmg_lw_b200 import api, driver, kdata, synth, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
b = synth.make_batch(4099, 72, cloudy=True, aerosol=True)   # not a multiple of the world size
full = driver.run_fused(b, icld=2, idrv=1, device=dev); api.sync()
got = sharding.run_sharded(b, lambda sh: driver.run_fused(sh, icld=2, idrv=1, device=dev), idrv=1)
ok = all(bool(torch.equal(got[k], full[k])) for k in sharding.OUT_KEYS_DRV)
# gather timing at bench size: 6 arrays of (73 | 72) x 1 M columns per rank
n = 1048576
loc = {k: torch.zeros((73 if k in ("uflx", "dflx", "uflxc", "dflxc") else 72, n), dtype=torch.float64, device=dev) for k in sharding.OUT_KEYS}
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sharding.gather_outputs(loc, n * world, dst=0)  # warm-up
torch.cuda.synchronize(); dist.barrier()
e0.record(); res = sharding.gather_outputs(loc, n * world, dst=0); e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = sum(v.numel() * 8 for v in loc.values())
    print(json.dumps({"world": world, "sharded_equals_single_gpu": ok, "gather_to_rank0_ms": float(ms), "bytes_per_rank": nbytes,
                      "GBps_into_rank0": nbytes * (world - 1) / float(ms) / 1e6}))
dist.barrier(); dist.destroy_process_group()
