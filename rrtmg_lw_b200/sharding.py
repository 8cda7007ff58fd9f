"""Column sharding of the longwave path over ``torch.distributed`` ranks (one process per GPU).

Columns are independent (``src/rrtmg_lw_rad.f90:486``) and the KISS McICA streams are seeded per
column (``src/mcica_subcol_gen_lw.f90:467-470``), so the column axis is block-partitioned and the
data path needs NO collective; the only exchange is the optional gather of the flux / heating-rate
arrays.  The Mersenne-Twister stream (``irng=1``) is one sequence over (sub-column, column, layer)
whose values depend on how the batch is split: it is refused unless the caller hands in
``mt_block`` (``api.set_mt_block`` on the GPU box), which tells the library where the rank's block sits
in the global stream so that each rank steps over the other ranks' words.

Pure plumbing: nothing in this module computes radiation.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import synth

OUT_KEYS = ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")
OUT_KEYS_DRV = OUT_KEYS + ("duflx_dt", "duflxc_dt")

# how each Batch field is sliced along the column axis: index of the column dimension
_COL_DIM = {"tsfc": 0, "lat": 0, "tauc": 1}


def block_range(ncol: int, world: int, rank: int):
    """Contiguous block partition: rank r owns columns [r*ceil(ncol/world), ...) (may be empty)."""
    per = -(-ncol // world)
    lo = min(rank * per, ncol)
    return lo, min(lo + per, ncol)


def shard_batch(b: "synth.Batch", world: int, rank: int) -> "synth.Batch":
    """The rank's column block of a batch (views made contiguous; layer 1 stays the surface)."""
    lo, hi = block_range(b.ncol, world, rank)
    kw = {}
    for k, v in b.__dict__.items():
        if torch.is_tensor(v):
            d = _COL_DIM.get(k, v.dim() - 1)
            kw[k] = v.narrow(d, lo, hi - lo).contiguous()
        else:
            kw[k] = v
    kw["ncol"] = hi - lo
    return synth.Batch(**kw)


def gather_outputs(local: dict, ncol_total: int, idrv: int = 0, dst: int | None = None, group=None) -> dict | None:
    """Gather the per-rank output blocks ``(nlev, ncol_local)`` into ``(nlev, ncol_total)``.

    ``dst=None``: every rank receives the full arrays (all_gather); otherwise only ``dst`` does.
    Uneven last blocks are padded to the common block width for the collective and trimmed after.
    Works on NCCL (CUDA tensors) and gloo (CPU tensors).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-ncol_total // world)
    keys = OUT_KEYS_DRV if idrv else OUT_KEYS
    # one packed buffer -> one collective: rows of all arrays stacked
    rows = [torch.as_tensor(local[k]) for k in keys]
    nloc = rows[0].shape[-1]
    pack = torch.cat(rows, dim=0)
    if nloc < per:
        pack = torch.cat([pack, pack.new_zeros((pack.shape[0], per - nloc))], dim=1)
    pack = pack.contiguous()
    backend_nccl = pack.is_cuda
    if dst is None or not backend_nccl:
        # gloo has no gather for every build; all_gather is universally available
        bufs = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(bufs, pack, group=group)
    else:
        bufs = [torch.empty_like(pack) for _ in range(world)] if rank == dst else None
        dist.gather(pack, bufs, dst=dst, group=group)
    if dst is not None and rank != dst:
        return None
    full = torch.cat(bufs, dim=1)[:, :ncol_total]
    out, r0 = {}, 0
    for k, t in zip(keys, rows):
        out[k] = full[r0:r0 + t.shape[0]].contiguous()
        r0 += t.shape[0]
    return out


def pipelined_gather(ncol_total: int, rows, run_chunk, nchunks: int = 4, dst: int | None = 0, group=None,
                     device="cpu", assemble: bool = True):
    """The rank's column block is run in ``nchunks`` column chunks; the gather of chunk i travels (NCCL's own
    stream) while chunk i+1 computes on the caller's stream -- SURVEY 8(e): the output gather is comparable to the
    compute time, so it is overlapped instead of appended.

    ``rows``: list of ``(key, nrows)`` of the output arrays (``uflx`` has nlay+1 rows, ``hr`` nlay, ...).
    ``run_chunk(lo, hi, outs)`` computes local columns [lo, hi) into ``outs[key]``, contiguous ``(nrows, hi-lo)``
    tensors on ``device`` -- they are row ranges of the chunk's packed send buffer, so whatever the call writes there is
    what is sent (a C-ABI call whose inputs are a column sub-range of a resident batch shares one ``ld`` with its
    outputs: it writes its own ``(nrows, ncol_local)`` arrays through ``api.col_offset`` and copies the chunk in).
    Returns on ``dst`` (every rank if ``dst`` is None) the dict of full ``(nrows, ncol_total)`` arrays, or with
    ``assemble=False`` the raw ``[chunk][rank]`` packed buffers plus the geometry ``(per, cw)``; None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-ncol_total // world)
    lo_r, hi_r = block_range(ncol_total, world, rank)
    nloc = hi_r - lo_r
    cw = -(-per // max(1, nchunks))
    nchunks = -(-per // cw)
    R = sum(n for _, n in rows)
    want = dst is None or rank == dst
    works, recv, keep = [], [], []
    for c in range(nchunks):
        lo = min(c * cw, nloc)
        hi = min(lo + cw, nloc)
        full_width = hi - lo == cw
        pack = (torch.empty if full_width else torch.zeros)((R, cw), dtype=torch.float64, device=device)
        if hi > lo:
            tgt = pack if full_width else torch.empty((R, hi - lo), dtype=torch.float64, device=device)
            outs, r0 = {}, 0
            for k, n in rows:
                outs[k] = tgt[r0:r0 + n]
                r0 += n
            run_chunk(lo, hi, outs)
            if not full_width:
                pack[:, :hi - lo] = tgt
        if dst is None or not pack.is_cuda:      # gloo: all_gather is universally available
            bufs = [torch.empty_like(pack) for _ in range(world)]
            works.append(dist.all_gather(bufs, pack, group=group, async_op=True))
        else:
            bufs = [torch.empty_like(pack) for _ in range(world)] if rank == dst else None
            works.append(dist.gather(pack, bufs, dst=dst, group=group, async_op=True))
        recv.append(bufs)
        keep.append(pack)
    for w in works:
        w.wait()
    if not want:
        return None
    if not assemble:
        return dict(packed=recv, per=per, cw=cw, rows=list(rows))
    full = torch.empty((R, ncol_total), dtype=torch.float64, device=device)
    for c in range(nchunks):
        for r in range(world):
            g0 = r * per + c * cw
            n = max(0, min(cw, per - c * cw, ncol_total - g0))
            if n > 0:
                full[:, g0:g0 + n] = recv[c][r][:, :n]
    out, r0 = {}, 0
    for k, n in rows:
        out[k] = full[r0:r0 + n]
        r0 += n
    return out


def run_sharded(b: "synth.Batch", run_local, idrv: int = 0, irng: int = 0, gather_to: int | None = None, group=None,
                mt_block=None):
    """Run ``run_local(shard) -> dict`` on this rank's column block and gather the outputs.

    ``run_local`` is the caller's per-rank invocation of the CUDA path (``driver.run_fused`` on the
    GPU box); tests substitute a stand-in to exercise the partition/gather logic on CPU.
    ``mt_block(global_ncol, col_offset)`` (``api.set_mt_block``) is required for ``irng = 1``; it is called with
    the batch size and the block's first column before ``run_local`` and with ``(0, 0)`` after it.
    """
    if irng != 0 and mt_block is None:
        raise ValueError("irng=1 (Mersenne Twister) is one global stream over all columns of a call; "
                         "its values depend on the partition -- use irng=0 (KISS) on the sharded path, "
                         "or pass mt_block=api.set_mt_block")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    shard = shard_batch(b, world, rank)
    if irng != 0:
        mt_block(b.ncol, block_range(b.ncol, world, rank)[0])
    try:
        local = run_local(shard)
    finally:
        if irng != 0:
            mt_block(0, 0)
    return gather_outputs(local, b.ncol, idrv=idrv, dst=gather_to, group=group)
