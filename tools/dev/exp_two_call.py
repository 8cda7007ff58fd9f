"""Dev experiment: where the time of the reference's two-call sequence (mcica_subcol_lw -> rrtmg_lw on materialised
sub-column arrays) goes, 65 536 columns x 72 layers on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from rrtmg_lw_b200 import api, driver, kdata

n, L, dev = int(os.environ.get("EXP_NCOL", 65536)), 72, "cuda:0"
torch.cuda.set_device(0)
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
w = bench.Workload(api, driver, dev, n, L, 2, 0, False, 0)
mk3 = lambda: torch.zeros((L, n, 140), dtype=torch.float64, device=dev)
m = {k: mk3() for k in ("cldfmcl", "ciwpmcl", "clwpmcl", "taucmcl")}
m["reicmcl"] = torch.zeros((L, n), dtype=torch.float64, device=dev); m["relqmcl"] = torch.zeros((L, n), dtype=torch.float64, device=dev)
b, o = w.b, w.out
def gen():
    api.mcica_subcol_lw(1, n, L, 2, 1, 0, b.play, b.cldfrac, b.ciwp, b.clwp, b.rei, b.rel, b.tauc, None,
                        m["cldfmcl"], m["ciwpmcl"], m["clwpmcl"], m["reicmcl"], m["relqmcl"], m["taucmcl"])
def rad():
    api.rrtmg_lw(n, L, 2, 0, *w.at, 2, 3, 1, m["cldfmcl"], m["taucmcl"], m["ciwpmcl"], m["clwpmcl"], m["reicmcl"],
                 m["relqmcl"], b.tauaer, o["uflx"], o["dflx"], o["hr"], o["uflxc"], o["dflxc"], o["hrc"], None, None)
def t(f, k=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("mcica_subcol_lw (generator + materialise): %.2f ms" % t(gen))
print("rrtmg_lw strict, cloudy: %.2f ms" % t(rad))
api.set_timing(True); rad(); torch.cuda.synchronize()
print("stages (taumol, sweeps...):", [round(x, 2) for x in api.last_stage_timing()], "band/all:", api.last_timing())
api.set_timing(False)
w.step(); torch.cuda.synchronize()
print("fused entry: %.2f ms" % t(w.step))
