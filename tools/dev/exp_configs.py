"""Full-size runs of the other BASELINE.json configs on one GPU (device-resident, CUDA events):
C2 clear-sky 65 536 x 72, C4 icld=5 idcor=1 + aerosol + idrv=1 at 1 M x 72, C5 137 layers at this GPU's 1/8 share
(129 600 columns) of the 0.25-degree grid.  Prints columns/s and basic sanity of the outputs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from rrtmg_lw_b200 import api, driver, kdata, synth

dev = "cuda:0"; torch.cuda.set_device(0)
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)

def batch(ncol, nlay, cloudy, aerosol):
    parts = [synth.make_batch(min(65536, ncol - c0), nlay, cloudy=cloudy, aerosol=aerosol, device=dev, col_offset=7 + i)
             for i, c0 in enumerate(range(0, ncol, 65536))]
    if len(parts) == 1: return parts[0]
    kw = {}
    for k, v in parts[0].__dict__.items():
        if torch.is_tensor(v):
            d = 0 if k in ("tsfc", "lat") else (1 if k == "tauc" else v.dim() - 1)
            kw[k] = torch.cat([getattr(p, k) for p in parts], dim=d).contiguous()
        else: kw[k] = v
    kw["ncol"] = ncol
    return synth.Batch(**kw)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
# C2
b = batch(65536, 72, False, False); out = driver._outputs(72, b.ncol, like=b.play)
ms = timeit(lambda: driver.run_clear_or_strict(b, device=dev, out=out)); api.sync()
res["C2 clear 65536x72 idrv=0"] = {"ms": ms, "cols_per_s": b.ncol / ms * 1e3, "toa_up": float(out["uflx"][-1].mean())}
del b, out
# C4
b = batch(1048576, 72, True, True); out = driver._outputs(72, b.ncol, like=b.play)
al = torch.zeros_like(b.play)
api.get_alpha(1, b.ncol, b.nlay, 5, 1, 2500.0, b.dz.to(dev).contiguous(), b.lat.to(dev).contiguous(), 200, b.cldfrac, al)
ms = timeit(lambda: driver.run_fused(b, icld=5, idrv=1, alpha=al, device=dev, out=out), n=2); api.sync()
res["C4 icld=5 idcor=1 aerosol idrv=1 1Mx72"] = {"ms": ms, "cols_per_s": b.ncol / ms * 1e3, "toa_up": float(out["uflx"][-1].mean()),
                                                "dF_dTs_sfc": float(out["duflx_dt"][0].mean())}
del b, out, al
# C5 share of one of 8 GPUs
b = batch(129600, 137, True, False); out = driver._outputs(137, b.ncol, like=b.play)
ms = timeit(lambda: driver.run_fused(b, icld=2, idrv=0, device=dev, out=out)); api.sync()
res["C5 129600x137 (1/8 of 1440x720) icld=2"] = {"ms": ms, "cols_per_s": b.ncol / ms * 1e3, "toa_up": float(out["uflx"][-1].mean())}
print(json.dumps(res, indent=1))
