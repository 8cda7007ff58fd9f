#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_10_tests.log
tail -n 4 gpurun_out/r2_10_tests.log
: > gpurun_out/r2_10_exp.log
EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_10_exp.log 2>&1
EXP_OPTS=band_kernel=4,sweep_groups=1 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_10_exp.log 2>&1
EXP_ICLD=0 EXP_OPTS=band_kernel=4,sweep_groups=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_10_exp.log 2>&1
cat gpurun_out/r2_10_exp.log
