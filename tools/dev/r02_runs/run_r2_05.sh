#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_05_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_05_tests.log
tail -n 6 gpurun_out/r2_05_tests.log
EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py > gpurun_out/r2_05_exp.log 2>&1
EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_05_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_05_exp.log 2>&1
cat gpurun_out/r2_05_exp.log
EXP_OPTS=band_kernel=4 timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 9 --launch-count 1 -f -k regex:rtrnmc_kernel -o gpurun_out/r2_05_sweep3 python tools/dev/exp_bands.py > gpurun_out/r2_05_ncu.log 2>&1
tail -n 3 gpurun_out/r2_05_ncu.log
