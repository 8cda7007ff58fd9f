#!/bin/bash
cd /root/repo
L=$PWD/rrtmg_lw_b200/lib
for v in _clr0 "" _clr12 _clr8b14; do
  echo "== variant '$v'"
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so python tools/dev/exp_bands.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages_ms']; print('taumol', s[0], 'b3', s[3], 'b4', s[4], 'b5', s[5], 'sweeps', round(d['bands_total_ms']-s[0],2), 'step', d['step_ms'])"
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
