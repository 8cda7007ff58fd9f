"""N > 1 host logic on CPU: block partition of the column axis and the output gather, with
world_size-2 gloo processes.  The per-rank "compute" is the CPU oracle here (tests may use it as
a checker); on the GPU box the same code path runs driver.run_fused per rank (tests/test_gpu_parity)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrtmg_lw_b200 import sharding, synth


def test_block_range_covers_all_columns():
    for ncol in (1, 5, 64, 1000, 1036800):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.block_range(ncol, world, r)
                assert 0 <= lo <= hi <= ncol
                seen.extend(range(lo, hi) if ncol <= 1000 else [lo, hi])
            if ncol <= 1000:
                assert seen == list(range(ncol))
            else:
                assert seen[0] == 0 and seen[-1] == ncol


def test_shard_batch_slices_every_field():
    b = synth.make_batch(37, 12, cloudy=True, aerosol=True)
    parts = [sharding.shard_batch(b, 3, r) for r in range(3)]
    assert sum(p.ncol for p in parts) == 37
    assert torch.equal(torch.cat([p.play for p in parts], dim=-1), b.play)
    assert torch.equal(torch.cat([p.tsfc for p in parts], dim=0), b.tsfc)
    assert torch.equal(torch.cat([p.tauc for p in parts], dim=1), b.tauc)
    assert torch.equal(torch.cat([p.tauaer for p in parts], dim=-1), b.tauaer)
    assert torch.equal(torch.cat([p.emis for p in parts], dim=-1), b.emis)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ncol, nlay, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import orc
        from rrtmg_lw_b200 import kdata
        orc.ini(kdata.synthetic_blob(), 1004.0, 4)
        b = synth.make_batch(ncol, nlay, cloudy=True)

        def run_local(shard):
            mc = orc.mcica_subcol_lw(shard, 2, permuteseed=1, irng=0)
            r = orc.rrtmg_lw(shard, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=1)
            assert r["rc"] == 0
            return {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in sharding.OUT_KEYS}

        full = sharding.run_sharded(b, run_local)                    # all ranks get everything
        only0 = sharding.run_sharded(b, run_local, gather_to=0)      # only rank 0
        assert (only0 is None) == (rank != 0)
        with pytest.raises(ValueError):
            sharding.run_sharded(b, run_local, irng=1)
        # irng = 1 with the stream-position hook: called with (batch size, first column of the block), then reset;
        # the stand-in draws the WHOLE batch's Mersenne-Twister masks with the checker and keeps its block, which is
        # what the library's stepping over the other blocks' words must reproduce (tests/test_gpu_parity.py)
        seen = []
        lo_, hi_ = sharding.block_range(ncol, world, rank)
        mc_all = orc.mcica_subcol_lw(b, 2, permuteseed=5, irng=1)

        def run_local_mt(shard):
            assert seen == [(ncol, lo_)]
            mc = {k: (np.ascontiguousarray(v[:, lo_:hi_]) if isinstance(v, np.ndarray) else v) for k, v in mc_all.items()}
            r = orc.rrtmg_lw(shard, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=1)
            return {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in sharding.OUT_KEYS}

        full_mt = sharding.run_sharded(b, run_local_mt, irng=1, mt_block=lambda n, o: seen.append((n, o)))
        assert seen == [(ncol, lo_), (0, 0)]
        ref_mt = orc.rrtmg_lw(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc_all, nthreads=1)
        assert all(np.array_equal(full_mt[k].numpy(), ref_mt[k]) for k in sharding.OUT_KEYS)
        # pipelined variant: the rank's block in 3 column chunks, each chunk's gather issued before the next chunk
        # is computed (ragged: 37 columns -> blocks 19 + 18, chunks of 7 -> 7 + 7 + 5 | 7 + 7 + 4)
        shard = sharding.shard_batch(b, world, rank)
        rows = [(k, nlay if k.startswith("hr") else nlay + 1) for k in sharding.OUT_KEYS]
        calls = []

        def run_chunk(lo, hi, outs):
            calls.append((lo, hi))
            sub = sharding.shard_batch(shard, 1, 0)
            for k2, v2 in sub.__dict__.items():
                if torch.is_tensor(v2):
                    d = sharding._COL_DIM.get(k2, v2.dim() - 1)
                    setattr(sub, k2, v2.narrow(d, lo, hi - lo).contiguous())
            sub.ncol = hi - lo
            r = run_local(sub)
            for k2 in outs:
                assert outs[k2].is_contiguous() and outs[k2].shape == r[k2].shape
                outs[k2].copy_(r[k2])

        for dst_ in (None, 0):
            piped = sharding.pipelined_gather(ncol, rows, run_chunk, nchunks=3, dst=dst_)
            assert (piped is None) == (dst_ == 0 and rank != 0)
            if piped is not None:
                assert all(torch.equal(piped[k], full[k]) for k in full), "pipelined gather differs"
        assert calls[:3] == [(0, 7), (7, 14), (14, 19 if rank == 0 else 18)]
        raw = sharding.pipelined_gather(ncol, rows, run_chunk, nchunks=3, dst=None, assemble=False)
        assert raw["per"] == 19 and raw["cw"] == 7 and len(raw["packed"]) == 3 and len(raw["packed"][0]) == world
        if rank == 0:
            q.put({k: v.numpy() for k, v in full.items()})
            assert all(torch.equal(only0[k], full[k]) for k in full)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_run_equals_single_process(oracle):
    """world_size 2, uneven split (37 columns): the gathered result equals the unsharded run bit for
    bit -- the KISS streams are per column, so the partition must not change anything."""
    ncol, nlay = 37, 16
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncol, nlay, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    b = synth.make_batch(ncol, nlay, cloudy=True)
    mc = oracle.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
    ref = oracle.rrtmg_lw(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc)
    for k in sharding.OUT_KEYS:
        assert np.array_equal(got[k], ref[k]), k
