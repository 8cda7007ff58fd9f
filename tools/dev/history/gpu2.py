import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from rrtmg_lw_b200 import kdata, synth, api, driver
from oracle import orc
blob = kdata.synthetic_blob()
orc.ini(blob); api.rrtmg_lw_ini(1004.0, blob)
for n in (128, 256, 384, 1024, 2048):
    b = synth.make_batch(n, 72)
    o = orc.rrtmg_lw(b, nthreads=16)
    for dev in (None, 'cuda:0'):
        g = driver.run_clear_or_strict(b, device=dev)
        api.sync()
        gu = g['uflx'].cpu().numpy() if dev else g['uflx']
        gd = g['dflx'].cpu().numpy() if dev else g['dflx']
        eu = np.abs(gu-o['uflx']); ed = np.abs(gd-o['dflx'])
        badc = np.where(eu.max(0)>1e-6)[0]
        print(n, dev, 'uflx err', eu.max(), 'dflx err', ed.max(), 'bad cols', len(badc), badc[:10], 'bad levels', np.where(eu.max(1)>1e-6)[0][:10])
