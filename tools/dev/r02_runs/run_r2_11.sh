#!/bin/bash
set -x
mkdir -p gpurun_out
L=rrtmg_lw_b200/lib
: > gpurun_out/r2_11_exp.log
for v in "" _nst3 _nst4; do
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_11_exp.log 2>&1
done
cat gpurun_out/r2_11_exp.log
