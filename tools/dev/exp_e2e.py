"""Dev experiment: end-to-end (host pointers) throughput vs host_chunk, and the box's raw pinned H2D/D2H bandwidth."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from rrtmg_lw_b200 import api, driver, kdata

torch.cuda.set_device(0)
# raw PCIe
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); print(name, "GB/s", 5 * (1 << 30) / (time.perf_counter() - t0) / 1e9, flush=True)
del h, d
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
ne = 131072; NLAY = 72
b = bench.make_device_batch(ne, "cuda:0", 0)
names = ["play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
         "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "emis", "cldfrac", "ciwp", "clwp", "rei", "rel", "tauaer"]
hb = {}
for k in names:
    sl = getattr(b, k).contiguous(); pin = torch.empty(sl.shape, dtype=sl.dtype, pin_memory=True); pin.copy_(sl); hb[k] = pin.numpy()
tauc_h = torch.zeros((NLAY, ne, 16), dtype=torch.float64, pin_memory=True).numpy()
ho = {k: torch.empty((n, ne), dtype=torch.float64, pin_memory=True).numpy()
      for k, n in (("uflx", NLAY + 1), ("dflx", NLAY + 1), ("uflxc", NLAY + 1), ("dflxc", NLAY + 1), ("hr", NLAY), ("hrc", NLAY))}
h2d = sum(v.nbytes for v in hb.values()) + tauc_h.nbytes
def step():
    api.rrtmg_lw_mcica(ne, NLAY, 2, 0, 1, 0, *[hb[k] for k in names[:16]], 2, 3, 1, hb["cldfrac"], hb["ciwp"],
                       hb["clwp"], hb["rei"], hb["rel"], tauc_h, None, hb["tauaer"], ho["uflx"], ho["dflx"],
                       ho["hr"], ho["uflxc"], ho["dflxc"], ho["hrc"], None, None)
for hc in (8192, 16384, 32768, 65536):
    api.set_option("host_chunk", hc)
    step(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): step()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3
    print(json.dumps({"host_chunk": hc, "cols_per_s": ne / t, "h2d_GBps": h2d / t / 1e9}), flush=True)
