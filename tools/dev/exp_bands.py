"""Dev experiment: per-band kernel times of the library selected by RRTMG_LW_B200_LIB on the bench workload;
outputs of the first 8192 columns are saved / compared against a previous variant's (EXP_REF / EXP_SAVE)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from rrtmg_lw_b200 import api, driver, kdata

ncol = int(os.environ.get("EXP_NCOL", 262144))
idrv = int(os.environ.get("EXP_IDRV", 0))
dev = "cuda:0"
torch.cuda.set_device(0)
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
api.set_chunk(int(os.environ.get("EXP_CHUNK", 262144)))
for kv in os.environ.get("EXP_OPTS", "").split(","):
    if kv:
        k, v = kv.split("=")
        api.set_option(k, int(v))
b = bench.make_device_batch(ncol, dev, 0)
if os.environ.get("EXP_SAME_ATM"):
    # upper bound of any table-locality optimisation: every column gets column 0's atmosphere (pressure,
    # temperature, gases), so all lanes of a warp gather the same k-table rows; cloud inputs stay per column
    for k in ("play", "plev", "tlay", "tlev", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
              "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr"):
        t = getattr(b, k)
        t.copy_(t[:, :1].expand_as(t))
    b.tsfc.copy_(b.tsfc[:1].expand_as(b.tsfc))
out = driver._outputs(72, ncol, like=b.play)
conv = driver._conv(b, dev)
at = driver.atm_args(b, conv)
cl = [b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc]
du = torch.zeros_like(out["uflx"]) if idrv else None
duc = torch.zeros_like(out["uflx"]) if idrv else None

icld = int(os.environ.get("EXP_ICLD", 2))

def step():
    if icld == 0:   # clear sky through the strict entry
        api.rrtmg_lw(ncol, 72, 0, idrv, *at, 0, 0, 0, None, None, None, None, None, None, b.tauaer, out["uflx"],
                     out["dflx"], out["hr"], out["uflxc"], out["dflxc"], out["hrc"], du, duc)
        return
    api.rrtmg_lw_mcica(ncol, 72, icld, idrv, 1, 0, *at, 2, 3, 1, *cl, None, b.tauaer, out["uflx"], out["dflx"],
                       out["hr"], out["uflxc"], out["dflxc"], out["hrc"], du, duc)

step(); step()
torch.cuda.synchronize()
api.set_timing(True)
acc = None; tot = 0.0; alln = 0.0; n = 3
for _ in range(n):
    step(); torch.cuda.synchronize()
    st_ = np.array(api.last_stage_timing()); acc = st_ if acc is None else acc + st_; t = api.last_timing(); tot += t[0]; alln += t[1]
api.set_timing(False)
api.sync()
# whole-step time without the per-stage events (the library may then run independent kernels on side streams)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
step(); torch.cuda.synchronize()
e0.record()
for _ in range(n):
    step()
e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / n
cur = {k: out[k][:, :8192].cpu().numpy() for k in ("uflx", "dflx", "uflxc", "dflxc", "hr", "hrc")}
err = None
if os.environ.get("EXP_REF") and os.path.exists(os.environ["EXP_REF"]):
    ref = np.load(os.environ["EXP_REF"])
    err = max(float(np.abs(cur[k] - ref[k]).max()) for k in cur)
if os.environ.get("EXP_SAVE"):
    np.savez(os.environ["EXP_SAVE"], **cur)
print(json.dumps({"lib": os.path.basename(api.LIB_PATH), "opts": os.environ.get("EXP_OPTS", ""), "ncol": ncol, "chunk": int(os.environ.get("EXP_CHUNK", 262144)),
                  "stages_ms": [round(x / n, 3) for x in acc], "bands_total_ms": round(tot / n, 3),
                  "all_ms": round(alln / n, 3), "step_ms": round(step_ms, 3), "maxdiff_vs_ref": err}), flush=True)
