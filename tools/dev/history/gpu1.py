import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from rrtmg_lw_b200 import kdata, synth, api, driver
from oracle import orc
blob = kdata.synthetic_blob()
orc.ini(blob); api.rrtmg_lw_ini(1004.0, blob)
# table parity
for n in ['tau_tbl','exp_tbl','tfn_tbl','rwgt','3:absa','3:absb','3:fracrefa','8:ka_mn2o','5:ccl4','13:ka_mco']:
    a, b_ = orc.get_table(n), api.get_table(n)
    print(n, a.size, b_.size, 'bitwise equal' if a.size==b_.size and np.array_equal(a,b_) else 'DIFF')
b = synth.make_batch(256, 72, cloudy=True)
o = orc.rrtmg_lw(b, nthreads=8)
api.debug_capture(5)
g = driver.run_clear_or_strict(b)
for k in ('uflx','dflx','hr','uflxc','dflxc','hrc'):
    print(k, 'max abs diff', np.abs(o[k]-g[k]).max(), 'max', np.abs(o[k]).max())
s = orc.column_stages(b, 5); d = api.debug_stages(5, 72)
print('ints equal', np.array_equal(s['ints6'], d['ints6']), 'laytrop', s['laytrop'], d['laytrop'], 'pwvcm', s['pwvcm']-d['pwvcm'])
print('facs maxdiff', np.abs(s['facs']-d['facs']).max())
rel = np.abs(s['taug']-d['taug'])/np.maximum(np.abs(s['taug']),1e-300)
print('taug bitwise frac', (s['taug']==d['taug']).mean(), 'max rel', rel.max(), 'fracs bitwise', (s['fracs']==d['fracs']).mean())
print('itr equal', np.array_equal(s['itr'], d['itr']), (s['itr']!=d['itr']).sum())
# cloudy fused vs oracle (mcica + rrtmg_lw)
mc = orc.mcica_subcol_lw(b, 2, permuteseed=1, irng=0)
oc = orc.rrtmg_lw(b, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
print('oracle cloudy rc', oc['rc'])
api.debug_capture(None)
gc = driver.run_fused(b, icld=2)
for k in ('uflx','dflx','hr','uflxc','dflxc','hrc'):
    print('fused', k, 'max abs diff', np.abs(oc[k]-gc[k]).max())
gm = driver.run_mcica_subcol(b, 2)
for k in ('cldfmcl','ciwpmcl','clwpmcl','taucmcl','reicmcl','relqmcl'):
    print('mcica', k, 'equal', np.array_equal(mc[k], gm[k]), 'cloudy frac', (mc['cldfmcl']>0).mean())
gs = driver.run_clear_or_strict(b, icld=2, inflglw=2, iceflglw=3, liqflglw=1, mc=mc)
for k in ('uflx','dflx','hr','uflxc','dflxc','hrc'):
    print('strict', k, 'max abs diff', np.abs(oc[k]-gs[k]).max())
# idrv
oi = orc.rrtmg_lw(b, icld=2, idrv=1, inflglw=2, iceflglw=3, liqflglw=1, mc=mc, nthreads=8)
gi = driver.run_fused(b, icld=2, idrv=1)
for k in ('uflx','duflx_dt','duflxc_dt'):
    print('idrv', k, 'max abs diff', np.abs(oi[k]-gi[k]).max(), np.abs(oi[k]).max())
api.sync()
