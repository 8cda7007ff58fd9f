"""Worker of tests/test_multigpu_abi.py (run under torchrun, one rank per GPU): a column-sharded McICA step whose output
blocks are gathered on rank 0 THROUGH THE C ABI (rrtmg_lw_b200_comm_* / rrtmg_lw_b200_gather; torch.distributed is only
the launcher's mailbox for the 128-byte NCCL id, over gloo), in two pipelined column chunks written straight into packed
send buffers (out_ld); rank 0 then runs the whole batch alone and compares bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rrtmg_lw_b200 import api, driver, kdata, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
api.set_device(local)
torch.cuda.set_device(local)
dev = "cuda:%d" % local
dist.init_process_group("gloo")
uid = [api.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
api.comm_init(world, rank, uid[0])

NCOL, NLAY = 3000, 40            # ragged over the ranks' chunks on purpose
per = -(-NCOL // world)
lo, hi = rank * per, min((rank + 1) * per, NCOL)
full = synth.make_batch(NCOL, NLAY, cloudy=True, aerosol=True)   # every rank builds the same batch, takes its block
conv = driver._conv(full, dev)
at = driver.atm_args(full, conv)
cl = [conv(x) for x in (full.cldfrac, full.ciwp, full.clwp, full.rei, full.rel, full.tauc)]
ta = conv(full.tauaer)
ROWS = [("uflx", NLAY + 1), ("dflx", NLAY + 1), ("hr", NLAY), ("uflxc", NLAY + 1), ("dflxc", NLAY + 1), ("hrc", NLAY)]
R = sum(n for _, n in ROWS)
chunks = [(lo, lo + (hi - lo) // 2), (lo + (hi - lo) // 2, hi)]
cw = per - per // 2                # widest chunk of any rank: the fixed send width
send = [torch.zeros((R, cw), dtype=torch.float64, device=dev) for _ in chunks]
recv = [torch.zeros((world, R, cw), dtype=torch.float64, device=dev) if rank == 0 else None for _ in chunks]
for c, (a, b) in enumerate(chunks):
    views, r0 = {}, 0
    for k, n in ROWS:
        views[k] = send[c][r0:r0 + n]
        r0 += n
    api.set_option("out_ld", cw)
    off = lambda t, w=1: api.col_offset(t, a, w)  # noqa: E731
    api.rrtmg_lw_mcica(b - a, NLAY, 2, 0, 1, 0, *[off(t) for t in at[:15]], off(at[15]), 2, 3, 1,
                       *[off(t) for t in cl[:5]], off(cl[5], 16), None, off(ta), views["uflx"], views["dflx"],
                       views["hr"], views["uflxc"], views["dflxc"], views["hrc"], None, None, ld=NCOL)
    api.set_option("out_ld", 0)
    api.gather(send[c], recv[c], root=0)   # NCCL, on this stream, behind the chunk's kernels
torch.cuda.synchronize()
api.sync()
dist.barrier()
ok = True
if rank == 0:
    ref = driver.run_fused(full, icld=2, device=dev)
    torch.cuda.synchronize()
    for r in range(world):
        rlo, rhi = r * per, min((r + 1) * per, NCOL)
        rch = [(rlo, rlo + (rhi - rlo) // 2), (rlo + (rhi - rlo) // 2, rhi)]
        for c, (a, b) in enumerate(rch):
            r0 = 0
            for k, n in ROWS:
                got = recv[c][r, r0:r0 + n, :b - a]
                ok = ok and bool(torch.equal(got, ref[k][:, a:b]))
                r0 += n
    print("ABI_GATHER_OK" if ok else "ABI_GATHER_MISMATCH", flush=True)
api.comm_finalize()
dist.barrier()
sys.exit(0 if ok else 1)
