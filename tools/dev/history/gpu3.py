import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from rrtmg_lw_b200 import kdata, synth, api, driver
blob = kdata.synthetic_blob()
api.rrtmg_lw_ini(1004.0, blob)
b = synth.make_batch(1024, 72)
g = driver.run_clear_or_strict(b)
api.sync()
print('done', float(g['uflx'].mean()))
