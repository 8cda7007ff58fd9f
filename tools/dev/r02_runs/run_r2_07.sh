#!/bin/bash
set -x
mkdir -p gpurun_out
L=rrtmg_lw_b200/lib
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_07_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_07_tests.log
tail -n 6 gpurun_out/r2_07_tests.log
: > gpurun_out/r2_07_exp.log
for v in "" _s4; do
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_07_exp.log 2>&1
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so EXP_OPTS=band_kernel=4,sweep_groups=1 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_07_exp.log 2>&1
done
cat gpurun_out/r2_07_exp.log
