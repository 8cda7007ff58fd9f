#!/bin/bash
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_08_exp.log
EXP_ICLD=0 EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_08_exp.log 2>&1
EXP_ICLD=0 EXP_OPTS=band_kernel=2 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_08_exp.log 2>&1
cat gpurun_out/r2_08_exp.log
EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 18 --launch-count 1 -f -k regex:rtrnmc_kernel -o gpurun_out/r2_08_sweep3 python tools/dev/exp_bands.py > gpurun_out/r2_08_ncu.log 2>&1
tail -n 3 gpurun_out/r2_08_ncu.log
