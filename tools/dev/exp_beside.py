"""Dev experiment: timeline of the resident work-queue kernel (McICA masks + cloud optics beside taumol), from the
globaltimer stamps the kernel leaves (LW_KISS_WORK_STATS): how long the resident CTAs ran, how many items they had
taken when taumol ended (= when the mop-up launch began), how long the mop-up took."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from rrtmg_lw_b200 import api, driver, kdata

ncol, chunk = 1048576, 524288
dev = "cuda:0"
torch.cuda.set_device(0)
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
api.set_chunk(chunk)
b = bench.make_device_batch(ncol, dev, 0)
out = driver._outputs(72, ncol, like=b.play)
conv = driver._conv(b, dev)
at = driver.atm_args(b, conv)
cl = [b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc]

def step():
    api.rrtmg_lw_mcica(ncol, 72, 2, 0, 1, 0, *at, 2, 3, 1, *cl, None, b.tauaer, out["uflx"], out["dflx"],
                       out["hr"], out["uflxc"], out["dflxc"], out["hrc"], None, None)

L = api.lib()
L.rrtmg_lw_b200_kiss_work_stats.argtypes = [C.POINTER(C.c_ulonglong)]
L.rrtmg_lw_b200_kiss_work_stats.restype = None
for k in [int(x) for x in os.environ.get("EXP_BESIDE", "0,1,2").split(",")]:
    api.set_option("mcica_beside", k)
    step(); step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record(); torch.cuda.synchronize()
    st = (C.c_ulonglong * 8)()
    L.rrtmg_lw_b200_kiss_work_stats(st)
    t = [int(x) for x in st]
    ms = lambda a, b_: round((t[a] - t[b_]) / 1e6, 3) if t[a] and t[b_] else None
    print(json.dumps({"beside": k, "noprio": os.environ.get("RRTMG_LW_BESIDE_NOPRIO", "0"), "step_ms": round(e0.elapsed_time(e1) / 3, 3),
                      "resident_ms": ms(1, 0), "resident_start_to_mop_start_ms": ms(2, 0), "mop_ms": ms(3, 2),
                      "items_taken_at_mop_start": t[4], "items": chunk // 8}), flush=True)
