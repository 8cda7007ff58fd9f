#!/bin/bash
cd /root/repo
L=$PWD/rrtmg_lw_b200/lib
for v in "" _v4m4 _v4m5 _v4m6; do
  echo "== variant '$v'"
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so python tools/dev/exp_bands.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('taumol', d['stages_ms'][0], 'sweeps', round(d['bands_total_ms']-d['stages_ms'][0],2), 'step', d['step_ms'])"
done
