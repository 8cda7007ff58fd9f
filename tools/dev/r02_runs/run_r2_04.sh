#!/bin/bash
# round 2, GPU run 4: state-word sweeps; parity suite + per-stage times of build variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_04_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_04_tests.log
tail -n 12 gpurun_out/r2_04_tests.log
L=rrtmg_lw_b200/lib
EXP_OPTS=band_kernel=2 EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py > gpurun_out/r2_04_exp.log 2>&1
for v in "" _t3 _t5 _s4; do
  RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200$v.so EXP_OPTS=band_kernel=4 EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_04_exp.log 2>&1
done
EXP_OPTS=band_kernel=4,sweep_groups=0 EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_04_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_04_exp.log 2>&1
cat gpurun_out/r2_04_exp.log
EXP_OPTS=band_kernel=4 timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 8 --launch-count 2 -f -k regex:rtrnmc_kernel -o gpurun_out/r2_04_sweeps python tools/dev/exp_bands.py > gpurun_out/r2_04_ncu.log 2>&1
tail -n 3 gpurun_out/r2_04_ncu.log
