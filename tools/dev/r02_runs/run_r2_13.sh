#!/bin/bash
set -x
mkdir -p gpurun_out
L=rrtmg_lw_b200/lib
: > gpurun_out/r2_13_exp.log
for v in "" _c2b6 _c4b3 _c2b4; do
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 ncu --metrics gpu__time_duration.sum -k regex:mcica_kiss --launch-skip 2 --launch-count 2 python tools/dev/exp_bands.py 2>&1 | grep -E "gpu__time_duration|lib" >> gpurun_out/r2_13_exp.log
done
cat gpurun_out/r2_13_exp.log
