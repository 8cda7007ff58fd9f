#!/usr/bin/env python3
"""Benchmark of the RRTMG_LW longwave hot path on B200 (BASELINE.json metric: columns/s,
72 layers, 140 g-points, McICA cloudy).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

One "step" = one pass of the hot path (mcica_subcol_lw -> rrtmg_lw, fused device entry) over the
rank's column batch.  N > 1: one process per GPU (torchrun), the column batch is block-partitioned
(weak scaling: 1 M columns per GPU); the step INCLUDES the path's one exchange, the NCCL gather of every
rank's flux / heating-rate arrays into rank 0, overlapped chunk by chunk with the kernels;
value = columns of all ranks / max-over-ranks device time.  The line also carries device-timed numbers
of the other BASELINE.json configurations (C2 clear sky, C4 exponential-random + aerosol + dF/dTs,
C5 the 1440 x 720 x 137-layer grid block-partitioned over the ranks = strong scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLAY = 72
WORKLOAD = "C3: McICA cloudy, icld=2 max-random (KISS), inflglw=2 iceflglw=3 liqflglw=1, 72 layers, 140 g-points"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pin_to_gpu_numa_node(index):
    """Run this rank's host thread (and first-touch its pinned staging buffers) on the CPUs that are local to its
    GPU, so the end-to-end copies of several ranks do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # no NVML / no permission: keep the inherited affinity
        pass


def cpu_sample_run(ncol, nthreads, seed_off=0, block=128):
    """The oracle (CPU restatement) on `ncol` cloudy columns: McICA generation + rrtmg_lw in column
    blocks spread over `nthreads` host threads.  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import orc
    from rrtmg_lw_b200 import synth

    blocks = []
    for i, c0 in enumerate(range(0, ncol, block)):
        blocks.append(synth.make_batch(min(block, ncol - c0), NLAY, cloudy=True, col_offset=1000 + seed_off + i))

    def work(b):
        mc = orc.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
        r = orc.rrtmg_lw(b, icld=2, idrv=0, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=1)
        assert r["rc"] == 0
        return float(r["uflx"][0, 0])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        list(ex.map(work, blocks))
    return time.perf_counter() - t0


def run_reference(args, rank):
    """Reference arm: the reference's own CPU algorithm for the path.  Its Fortran cannot be built
    in this image (no Fortran compiler; k-data missing), so the arm is the C++ restatement
    (oracle, kind "port") on all host cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import orc
    from rrtmg_lw_b200 import kdata
    orc.ini(kdata.synthetic_blob(), 1004.0, 4)
    cores = os.cpu_count() or 1
    ncol = min(args.ref_cols or 2048 * cores, 65536)   # ~1 s per step on all cores; a 5+3-step run stays under a minute
    ncol = max(ncol, 512)
    for w in range(max(args.warmup, 1)):
        cpu_sample_run(min(ncol, 64 * cores), cores, seed_off=w)
    t = 0.0
    for s in range(args.steps):
        t += cpu_sample_run(ncol, cores, seed_off=10 + s)
    val = ncol * args.steps / t
    sample = "%d columns x %d layers per step, blocks of 128 columns over %d threads, oracle C++ port -O3 -ffp-contract=off" % (ncol, NLAY, cores)
    print(json.dumps({
        "impl": "reference", "metric": "columns/sec", "value": val, "unit": "columns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "columns/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def make_device_batch(ncol, device, seed_off, nlay=NLAY, aerosol=False):
    """Device-resident synthetic batch generated in pieces (bounded temporaries)."""
    import torch

    from rrtmg_lw_b200 import synth
    piece = 65536 if nlay <= 80 else 32768
    parts = []
    for i, c0 in enumerate(range(0, ncol, piece)):
        parts.append(synth.make_batch(min(piece, ncol - c0), nlay, cloudy=True, aerosol=aerosol, device=device,
                                      col_offset=seed_off * 4096 + i))
    if len(parts) == 1:
        return parts[0]
    b0 = parts[0]
    kw = {}
    for k, v in b0.__dict__.items():
        if hasattr(v, "shape") and v is not None:
            if k == "tsfc" or k == "lat":
                kw[k] = torch.cat([getattr(p, k) for p in parts], dim=0)
            elif k == "tauc":
                kw[k] = torch.cat([getattr(p, k) for p in parts], dim=1).contiguous()
            else:
                kw[k] = torch.cat([getattr(p, k) for p in parts], dim=-1).contiguous()
            for p in parts:
                setattr(p, k, None)
        else:
            kw[k] = v
    kw["ncol"] = ncol
    return synth.Batch(**kw)


# output arrays of one call, as row ranges of ONE packed (rows, columns) buffer: what a rank computes is what it sends
def out_rows(nlay, idrv=0):
    rows = [("uflx", nlay + 1), ("dflx", nlay + 1), ("hr", nlay), ("uflxc", nlay + 1), ("dflxc", nlay + 1), ("hrc", nlay)]
    if idrv:
        rows += [("duflx_dt", nlay + 1), ("duflxc_dt", nlay + 1)]
    return rows


def packed_views(buf, rows):
    out, r0 = {}, 0
    for k, n in rows:
        out[k] = buf[r0:r0 + n]
        r0 += n
    return out


class Workload:
    """One BASELINE.json configuration resident on this rank's GPU: inputs, a packed output buffer, and step()."""

    def __init__(self, api, driver, dev, ncol, nlay, icld, idrv, aerosol, seed_off, clear=False):
        import torch
        self.api, self.ncol, self.nlay, self.icld, self.idrv, self.clear = api, ncol, nlay, icld, idrv, clear
        b = make_device_batch(ncol, dev, seed_off, nlay=nlay, aerosol=aerosol)
        self.b = b
        self.rows = out_rows(nlay, idrv)
        self.R = sum(n for _, n in self.rows)
        self.pack = torch.zeros((self.R, ncol), dtype=torch.float64, device=dev)
        self.out = packed_views(self.pack, self.rows)
        conv = driver._conv(b, dev)
        self.at = driver.atm_args(b, conv)
        self.cl = [b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc]
        self.alpha = None
        if icld >= 4:
            self.alpha = torch.zeros_like(b.play)
            api.get_alpha(1, ncol, nlay, icld, 1, 2500.0, b.dz, b.lat, 200, b.cldfrac, self.alpha)
        torch.cuda.empty_cache()   # the generator's temporaries go back to the device: the library allocates its own scratch

    def step(self, lo=0, hi=None, out=None, out_ld=None):
        """The hot path on local columns [lo, hi); outputs into `out` (default: the rank's packed buffer)."""
        api, o = self.api, (out or self.out)
        hi = self.ncol if hi is None else hi
        sub = lo != 0 or hi != self.ncol
        co = (lambda x, inner=1: api.col_offset(x, lo, inner)) if sub else (lambda x, inner=1: x)
        du, duc = (o["duflx_dt"], o["duflxc_dt"]) if self.idrv else (None, None)
        if out_ld is not None:
            api.set_option("out_ld", out_ld)
        try:
            if self.clear:
                n = None
                api.rrtmg_lw(hi - lo, self.nlay, 0, self.idrv, *[co(x) for x in self.at], 0, 0, 0, n, n, n, n, n, n,
                             co(self.b.tauaer), o["uflx"], o["dflx"], o["hr"], o["uflxc"], o["dflxc"], o["hrc"], du, duc,
                             ld=self.ncol)
            else:
                cl = [co(x) for x in self.cl[:5]] + [co(self.cl[5], 16)]
                al = co(self.alpha) if self.alpha is not None else None
                api.rrtmg_lw_mcica(hi - lo, self.nlay, self.icld, self.idrv, 1, 0, *[co(x) for x in self.at], 2, 3, 1, *cl,
                                   al, co(self.b.tauaer), o["uflx"], o["dflx"], o["hr"], o["uflxc"], o["dflxc"], o["hrc"],
                                   du, duc, ld=self.ncol)
        finally:
            if out_ld is not None:
                api.set_option("out_ld", 0)


def timed(torch, dist, world, dev, fn, steps, warmup):
    """W untimed + K timed calls of fn, device events, barrier + synchronize on both sides, max over ranks (ms total)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncol", type=int, default=1048576, help="columns per GPU (C3)")
    ap.add_argument("--chunk", type=int, default=524288,
                    help="columns per internal pass of the library (scratch is ~1.3 KB per (column, layer) of the pass)")
    ap.add_argument("--e2e-cols", type=int, default=262144,
                    help="columns per end-to-end step (pinned host buffers, 33 KB per column)")
    ap.add_argument("--ref-cols", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the C2 / C4 / C5 device-timed extras")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32],
                    help="64 = the headline FP64 path; 32 = the optional FP32 mode of the band kernels (secondary line)")
    ap.add_argument("--band-kernel", type=int, default=4, choices=[2, 4],
                    help="4 = two-phase path (taumol per (column, layer) + sweep kernels); 2 = fused per-band kernels")
    ap.add_argument("--gather-transport", default="abi", choices=["abi", "nccl", "p2p"],
                    help="N > 1: how a chunk's outputs reach rank 0: the library's own NCCL gather behind the C ABI "
                         "(rrtmg_lw_b200_gather, default), torch.distributed's NCCL gather, or peer "
                         "copies into IPC-mapped buffers (copy engines; measured 38 GB/s on this pool's boxes against "
                         "660 GB/s for NCCL over NVLink, profiles/r02_p2p_vs_nccl.log -- kept for boxes where peer "
                         "copies take NVLink)")
    ap.add_argument("--gather-chunks", type=int, default=8,
                    help="N > 1: column chunks of a rank's block whose output gathers (NCCL, into rank 0) overlap the next "
                         "chunk's kernels; n = n equal chunks, 0 = tapered schedule (1/4 1/4 1/4 1/8 1/16 1/32 1/32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from rrtmg_lw_b200 import api, driver, kdata

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    all_cpus = os.sched_getaffinity(0)
    pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    # keep stdout to the ONE JSON line: anything a library prints to fd 1 meanwhile (e.g. NCCL's version banner)
    # goes to stderr; fd 1 is restored just before the result line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    blob = kdata.synthetic_blob()
    api.rrtmg_lw_ini(1004.0, blob, 4)
    api.set_chunk(args.chunk)
    api.set_option("band_kernel", args.band_kernel)
    api.set_option("precision", args.precision)
    W = max(args.warmup, 3)
    flops = {}
    try:
        fj = json.load(open(os.path.join(ROOT, "profiles", "flops_per_column.json")))
        flops = {k: v["flops"] for k, v in fj.items() if isinstance(v, dict)}
    except (OSError, ValueError, KeyError):
        pass

    # ------------------------------------------------------------------ headline: C3, weak scaling
    ncol = args.ncol
    wl = Workload(api, driver, dev, ncol, NLAY, 2, 0, False, rank)
    rows = wl.rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the step INCLUDES the gather of every rank's six output arrays into rank 0 (the one exchange of the path,
    # SURVEY 8(e)): the rank's block runs in column chunks whose packed (rows, chunk) output buffers are the send
    # buffers; chunk i's gather (NCCL stream) overlaps chunk i+1's kernels.  N = 1 has nothing to gather.
    # chunk schedule: n > 0 (default 8) = n equal chunks; 0 = tapered, 1/4 1/4 1/4 1/8 1/16 1/32 1/32 of the block (small
    # last chunks shorten the transfer left over at the end of the step; measured equal to 8 equal chunks at N = 8 and
    # 3 % slower at N = 2, so the exposed cost at N = 8 is not that tail -- DESIGN.md section 8)
    if args.gather_chunks > 0:
        cw_u = -(-ncol // args.gather_chunks)
        bounds = [(lo, min(lo + cw_u, ncol)) for lo in range(0, ncol, cw_u)]
    else:
        bounds, lo = [], 0
        for f in (4, 4, 4, 8, 16, 32):
            w_ = max(128, (ncol // f) // 128 * 128)
            if lo + w_ >= ncol:
                break
            bounds.append((lo, lo + w_))
            lo += w_
        bounds.append((lo, ncol))
    nch = len(bounds)
    send = recv = None
    if world > 1:
        send = [torch.zeros((wl.R, hi - lo), dtype=torch.float64, device=dev) for lo, hi in bounds]
        if rank == 0:   # [chunk][rank] packed blocks: the gathered result, no further assembly
            recv = [torch.empty((world, wl.R, hi - lo), dtype=torch.float64, device=dev) for lo, hi in bounds]

    # transport of a chunk into the root.  "abi" (default): the library's NCCL gather behind the C ABI
    # (rrtmg_lw_b200_comm_init / rrtmg_lw_b200_gather: what a Fortran host would call), enqueued on a side stream behind
    # the chunk's kernels.  "nccl": torch.distributed gather (same wire protocol, 660 GB/s per pair over NVLink).
    # "p2p": the root's receive buffers are mapped into every rank (CUDA IPC) and a rank PUSHES its finished chunk with a
    # peer copy on a side stream (copy engines, no SM held by a communication kernel; the root's own chunks are written
    # in place by its kernels) -- on this pool's boxes such copies run at 38 GB/s (PCIe class), so it is not the default.
    transport = args.gather_transport
    peer = None
    copy_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    if world > 1 and transport == "p2p":
        ok = 1
        try:
            from torch.multiprocessing.reductions import reduce_tensor
            obj = [[reduce_tensor(t) for t in recv] if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            if rank != 0:
                peer = [fn(*a) for fn, a in obj[0]]   # the root's (world, R, w) buffers, addressable from this rank
                peer[0][rank, :1, :1].copy_(send[0][:1, :1])   # makes torch enable peer access to the root's device
                torch.cuda.synchronize()
                import ctypes
                cudart = ctypes.CDLL("libcudart.so.12")
                cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                                   ctypes.c_void_p]
                cudart.cudaMemcpyAsync.restype = ctypes.c_int

                def push(c):
                    # a raw peer copy on THIS rank's stream (its copy engine writes over NVLink): nothing is enqueued in a
                    # context of the root's GPU, so the root's kernels are never time-sliced against another process
                    rc_ = cudart.cudaMemcpyAsync(peer[c][rank].data_ptr(), send[c].data_ptr(),
                                                 send[c].numel() * 8, 4, copy_stream.cuda_stream)
                    if rc_ != 0:
                        raise RuntimeError("cudaMemcpyAsync (peer) failed: %d" % rc_)
        except Exception as e:  # noqa: BLE001 -- no peer access / IPC on this box: fall back to NCCL, all ranks together
            sys.stderr.write("rank %d: p2p gather unavailable (%s), using NCCL\n" % (rank, e))
            ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            transport, peer = "nccl", None

    if world > 1 and transport == "abi":
        ok = 1
        try:
            uid = [api.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            api.comm_init(world, rank, uid[0])
        except Exception as e:  # noqa: BLE001 -- e.g. no loadable libnccl.so.2: fall back to torch's, all ranks together
            sys.stderr.write("rank %d: C-ABI gather unavailable (%s), using torch.distributed\n" % (rank, e))
            ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            transport = "nccl"

    def step_gathered():
        if transport == "abi":
            for c, (lo, hi) in enumerate(bounds):
                wl.step(lo, hi, out=packed_views(send[c], rows), out_ld=hi - lo)
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):   # the gather orders behind the chunk's kernels, beside the next chunk's
                    api.gather(send[c], recv[c] if rank == 0 else None, root=0)
            torch.cuda.current_stream().wait_stream(copy_stream)
            return
        if transport == "p2p":
            for c, (lo, hi) in enumerate(bounds):
                if rank == 0:
                    wl.step(lo, hi, out=packed_views(recv[c][0], rows), out_ld=hi - lo)
                else:
                    wl.step(lo, hi, out=packed_views(send[c], rows), out_ld=hi - lo)
                    copy_stream.wait_stream(torch.cuda.current_stream())
                    push(c)
            torch.cuda.current_stream().wait_stream(copy_stream)
            return
        works = []
        for c, (lo, hi) in enumerate(bounds):
            wl.step(lo, hi, out=packed_views(send[c], rows), out_ld=hi - lo)
            works.append(dist.gather(send[c], [recv[c][r] for r in range(world)] if rank == 0 else None, dst=0,
                                     async_op=True))
        for w_ in works:
            w_.wait()

    step = step_gathered if world > 1 else wl.step
    for _ in range(W):
        step()
    barrier()
    api.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.launch_count(reset=True)
    ms = timed(torch, dist, world, dev, step, args.steps, 0)
    launches = api.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    api.sync()
    value = world * ncol * args.steps / (ms * 1e-3)
    gather_res = None
    if world > 1:
        # the same step without the exchange, for reference; and the proof that what arrived is what was computed
        ms_ng = timed(torch, dist, world, dev, wl.step, max(2, min(args.steps, 5)), 1) / max(2, min(args.steps, 5))
        step_gathered()
        barrier()
        if transport == "p2p" and rank == 0:   # the root's kernels wrote its own block straight into recv[c][0]
            for c in range(nch):
                send[c].copy_(recv[c][0])
        wl.step()
        torch.cuda.synchronize()
        mine = torch.stack([send[c].sum() for c in range(nch)]).sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        same = None
        if rank == 0:
            got = [float(sum(recv[c][r].sum() for c in range(nch))) for r in range(world)]
            same = bool(all(abs(got[r] - float(sums[r])) <= 1e-9 * abs(got[r]) for r in range(world)))
            same = same and bool(torch.equal(torch.cat([recv[c][0] for c in range(nch)], dim=1), wl.pack))
        gather_res = {"included_in_value": True, "chunks": nch, "chunk_columns": [hi - lo for lo, hi in bounds], "dst": 0,
                      "transport": ("peer copies (copy engines over NVLink) of the packed (rows, chunk) buffers the kernels write, "
                                    "pushed by each rank into the root's IPC-mapped receive buffers; the root's kernels write "
                                    "its own block in place" if transport == "p2p" else
                                    "the library's NCCL gather behind the C ABI (rrtmg_lw_b200_gather: grouped ncclSend / ncclRecv "
                                    "on a side stream) of packed (rows, chunk) buffers the kernels write directly; no packing, "
                                    "no assembly" if transport == "abi" else
                                    "NCCL gather (torch.distributed) of packed (rows, chunk) buffers the kernels write "
                                    "directly; no packing, no assembly"),
                      "bytes_into_root_per_step": (world - 1) * ncol * wl.R * 8,
                      "ungathered": {"value": world * ncol / (ms_ng * 1e-3), "unit": "columns/s", "ms_per_step": ms_ng},
                      "root_holds_every_ranks_block": same}

    # instrumented extra step: CUDA events (library, launching stream) around every kernel of the pipeline
    api.set_timing(True)
    wl.step()
    torch.cuda.synchronize()
    ms_band, ms_all = api.last_timing()
    stages = api.last_stage_timing()
    api.set_timing(False)

    res = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = api.fp64_peak()
        flops_per_col = flops.get("C3", 606259.0)   # op-counting oracle (tools/count_flops.py)
        bytes_per_col = (37 * NLAY + 19 + 6 * NLAY + 4) * 8.0   # API words in + out, SURVEY 8(d)
        cols_per_s_band = ncol / (ms_band * 1e-3)
        two = args.band_kernel == 4 and args.precision == 64
        roof = {
            "bound": "fp64",
            "kernel": ("2 taumol_kernel<band group> (bands 3 4 5 7 9 / the other eleven) + 13 rtrnmc_kernel<B0,NB,CLD_MASK,0> (bands 1-12 one each, bands 13-16 together) per pass of %d columns (timed one after the other)"
                       if two else "band_kernel<1..16,CLD_MASK,0> (16 launches per pass of %d columns)") % args.chunk,
            "achieved": flops_per_col * cols_per_s_band / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": flops_per_col * cols_per_s_band / 1e12 / fp64_peak,
            "peak_source": "DFMA microbenchmark run in this process (rrtmg_lw_b200_fp64_peak)",
            "flops_per_column": flops_per_col, "band_kernels_ms_per_step": ms_band, "all_kernels_ms_per_step": ms_all,
            "per_stage_ms_per_step": [round(x, 3) for x in stages],
            "stages": ((["taumol"] + ["sweep band %d" % b for b in range(1, 13)] + ["sweep bands 13-16"]) if len(stages) == 14 else
                       (["taumol"] + ["sweep band %d" % b for b in range(1, 17)])) if two else ["band %d" % b for b in range(1, 17)],
            "note": "the reference arithmetic is not FMA-contracted (-fmad=false up to the table indices): 0.5 of the DFMA "
                    "peak is the ceiling of a bit-faithful kernel",
            "traffic": None,
            "hbm": {"achieved": bytes_per_col * cols_per_s_band / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_per_col * cols_per_s_band / 1e9 / hbm_peak, "bytes_per_column": bytes_per_col,
                    "peak_source": hbm_src},
        }
        # DRAM bytes of the same launches from the committed ncu capture (profiles/traffic.json, regenerated from the
        # capture's CSV by tools/traffic_json.py), scaled to this step
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                if tj.get("band_kernel") == args.band_kernel:
                    roof["traffic"] = tj["dram_bytes_per_column"] * ncol
                    roof["traffic_note"] = tj.get("note")
                    roof["hbm"]["traffic_gbs"] = roof["traffic"] / (ms_band * 1e-3) / 1e9
                    roof["hbm"]["traffic_frac"] = roof["hbm"]["traffic_gbs"] / hbm_peak
            except (OSError, ValueError, KeyError):
                pass
        res = {
            "metric": "columns/sec", "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "columns_per_gpu": ncol, "layers": NLAY, "chunk_columns": args.chunk,
                       "entry": "rrtmg_lw_b200_rrtmg_lw_mcica (device pointers)", "l2": "inputs (>=20 GB) larger than L2",
                       "fma": "off (-fmad=false)", "band_kernel": args.band_kernel,
                       "multi_gpu": ("column blocks per rank; value INCLUDES the NCCL gather of all outputs into rank 0"
                                     if world > 1 else "single GPU: nothing to gather"),
                       "precision": "fp64" if args.precision == 64 else
                                    "optional fp32 mode: float band kernels, double flux accumulators (tolerance 0.05 W m-2 / 0.01 K/day)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        }
        if gather_res is not None:
            res["gather"] = gather_res

    # ---- e2e: same metric through the C ABI with HOST (pinned) buffers, H2D + D2H inside the timed region
    b = wl.b
    out = wl.out
    ne = min(args.e2e_cols, ncol)
    names = ["play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
             "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "emis", "cldfrac", "ciwp", "clwp", "rei", "rel", "tauaer"]

    def pinned_buffers(n):
        hb_ = {}
        for k in names:
            sl = getattr(b, k)[..., :n].contiguous()
            pin = torch.empty(sl.shape, dtype=sl.dtype, pin_memory=True)
            pin.copy_(sl)
            hb_[k] = pin.numpy()
        tc_ = torch.zeros((NLAY, n, 16), dtype=torch.float64, pin_memory=True).numpy()
        ho_ = {k: torch.empty((m, n), dtype=torch.float64, pin_memory=True).numpy()
               for k, m in (("uflx", NLAY + 1), ("dflx", NLAY + 1), ("uflxc", NLAY + 1), ("dflxc", NLAY + 1),
                            ("hr", NLAY), ("hrc", NLAY))}
        return hb_, tc_, ho_

    while True:
        try:
            hb, tauc_h, ho = pinned_buffers(ne)
            break
        except RuntimeError:  # not enough page-locked memory on this host: halve the end-to-end batch
            if ne <= 8192:
                raise
            ne //= 2
    h2d = sum(v.nbytes for v in hb.values()) + tauc_h.nbytes
    d2h = sum(v.nbytes for v in ho.values())

    def step_e2e():
        api.rrtmg_lw_mcica(ne, NLAY, 2, 0, 1, 0, *[hb[k] for k in names[:16]], 2, 3, 1, hb["cldfrac"], hb["ciwp"],
                           hb["clwp"], hb["rei"], hb["rel"], tauc_h, None, hb["tauaer"], ho["uflx"], ho["dflx"],
                           ho["hr"], ho["uflxc"], ho["dflxc"], ho["hrc"], None, None)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    ke = max(2, min(args.steps, 5))
    for _ in range(ke):
        step_e2e()
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    if rank == 0:
        res["e2e"] = {"value": world * ne * ke / te, "unit": "columns/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "columns_per_step": ne, "steps": ke,
                      "note": "host pinned buffers -> rrtmg_lw_b200_rrtmg_lw_mcica(HOST) -> host; result fluxes read back"}
        wl.step()
        torch.cuda.synchronize()
        # sanity of the timed outputs
        res["checks"] = {"toa_up_mean": float(out["uflx"][-1].mean()), "sfc_down_mean": float(out["dflx"][0].mean()),
                         "host_equals_device": bool(np.array_equal(ho["uflx"], out["uflx"][:, :ne].cpu().numpy()))}
    del hb, tauc_h, ho

    # ------------------------------------------------------------------ the other BASELINE.json configurations
    # device-timed like the headline (same events, barrier, max over ranks); each with its own op-counted flops/column.
    # C2 and C4 run as per-rank replicas (weak, like C3); C5 is ONE 1440 x 720 grid block-partitioned over the ranks
    # (strong scaling), outputs gathered into rank 0 inside the timed step when N > 1.
    extras = {}
    if not args.no_extra_configs and args.precision == 64:
        del wl, out, b
        send = recv = None
        torch.cuda.empty_cache()

        def run_extra(name, w2, k, note, gather=False):
            f = w2.step
            if gather and world > 1:
                rbuf = torch.empty((world, w2.R, w2.ncol), dtype=torch.float64, device=dev) if rank == 0 else None

                def f():
                    w2.step()
                    dist.gather(w2.pack, [rbuf[r] for r in range(world)] if rank == 0 else None, dst=0)
            msx = timed(torch, dist, world, dev, f, k, 2) / k
            api.set_timing(True)
            w2.step()
            torch.cuda.synchronize()
            mb, ma = api.last_timing()
            api.set_timing(False)
            if rank == 0:
                tot = world * w2.ncol
                fl = flops.get(name)
                extras[name] = {"value": tot / (msx * 1e-3), "unit": "columns/s", "ms_per_step": msx, "columns": tot,
                                "layers": w2.nlay, "steps": k, "config": note, "flops_per_column": fl,
                                "band_kernels_ms_per_step": mb,
                                "fp64_frac_band_kernels": (fl * w2.ncol / (mb * 1e-3) / 1e12 / fp64_peak) if fl else None,
                                "toa_up_mean": float(w2.out["uflx"][-1].mean())}

        api.set_option("release_scratch", 1)
        api.set_chunk(262144)
        w2 = Workload(api, driver, dev, 65536, NLAY, 0, 0, False, rank, clear=True)
        run_extra("C2", w2, 5, "clear sky, rrtmg_lw() strict entry, idrv=0, 65 536 columns x 72 layers per GPU (weak)")
        del w2
        torch.cuda.empty_cache()
        # the reference's literal two-call sequence (README.md:179-183 of the reference): mcica_subcol_lw materialises the
        # four (140, ncol, nlay) sub-column arrays, rrtmg_lw reads them back -- 345 KB per column through the interface
        # instead of 25 KB; what an unmodified GCM gets before it switches to the fused entry
        n2c = 65536
        w2 = Workload(api, driver, dev, n2c, NLAY, 2, 0, False, rank)
        mk3 = lambda: torch.zeros((NLAY, n2c, 140), dtype=torch.float64, device=dev)  # noqa: E731
        mc = {k: mk3() for k in ("cldfmcl", "ciwpmcl", "clwpmcl", "taucmcl")}
        mc["reicmcl"] = torch.zeros((NLAY, n2c), dtype=torch.float64, device=dev)
        mc["relqmcl"] = torch.zeros((NLAY, n2c), dtype=torch.float64, device=dev)

        def two_call(w=w2, m=mc):
            bb, o = w.b, w.out
            api.mcica_subcol_lw(1, n2c, NLAY, 2, 1, 0, bb.play, bb.cldfrac, bb.ciwp, bb.clwp, bb.rei, bb.rel, bb.tauc, None,
                                m["cldfmcl"], m["ciwpmcl"], m["clwpmcl"], m["reicmcl"], m["relqmcl"], m["taucmcl"])
            api.rrtmg_lw(n2c, NLAY, 2, 0, *w.at, 2, 3, 1, m["cldfmcl"], m["taucmcl"], m["ciwpmcl"], m["clwpmcl"],
                         m["reicmcl"], m["relqmcl"], bb.tauaer, o["uflx"], o["dflx"], o["hr"], o["uflxc"], o["dflxc"],
                         o["hrc"], None, None)
        w2.step = two_call
        run_extra("C3_two_call", w2, 3, "the reference's two-call sequence mcica_subcol_lw -> rrtmg_lw with the sub-column "
                  "arrays materialised on the device (345 KB per column through the interface), 65 536 columns x 72 "
                  "layers per GPU (weak); same flops as C3")
        if rank == 0:
            extras["C3_two_call"]["flops_per_column"] = flops.get("C3")
        del w2, mc
        torch.cuda.empty_cache()
        api.set_option("release_scratch", 1)
        w2 = Workload(api, driver, dev, ncol, NLAY, 5, 1, True, rank)
        run_extra("C4", w2, 3, "icld=5 exponential-random overlap, idcor=1 alpha (get_alpha on device), 16-band aerosol "
                  "tau, idrv=1; %d columns x 72 layers per GPU (weak)" % ncol)
        del w2
        torch.cuda.empty_cache()
        api.set_option("release_scratch", 1)
        api.set_chunk(131072)
        n5 = 1440 * 720
        per = -(-n5 // world)
        n5_loc = max(0, min(per, n5 - rank * per))
        w2 = Workload(api, driver, dev, n5_loc, 137, 2, 0, False, 100 + rank)
        run_extra("C5", w2, 2, "global 0.25 degree grid 1440 x 720 = 1 036 800 columns x 137 layers, McICA cloudy, "
                  "block-partitioned over %d GPU(s) (STRONG scaling)%s" % (world, "; outputs gathered into rank 0 inside "
                  "the timed step" if world > 1 else ""), gather=True)
        if rank == 0:
            extras["C5"]["columns"] = n5
            extras["C5"]["value"] = n5 / (extras["C5"]["ms_per_step"] * 1e-3)
            extras["C5"]["scaling"] = "strong"
        del w2
        torch.cuda.empty_cache()
    if rank == 0:
        if extras:
            res["other_configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core again
            from oracle import orc
            orc.ini(blob, 1004.0, 4)
            cores = os.cpu_count() or 1
            nc = max(512, min(2048 * cores, 32768))   # ~20 core-seconds of CPU work
            cpu_sample_run(min(nc, 32 * cores), cores)
            tc = cpu_sample_run(nc, cores, seed_off=3)
            res["cpu_baseline"] = {"value": nc / tc, "unit": "columns/s", "cores": cores, "kind": "port",
                                   "sample": "%d columns x %d layers (McICA generation + rrtmg_lw), blocks of 128 over %d "
                                             "threads; C++ restatement of the reference, g++ -O3 -ffp-contract=off "
                                             "(no Fortran compiler in the image)" % (nc, NLAY, cores)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
