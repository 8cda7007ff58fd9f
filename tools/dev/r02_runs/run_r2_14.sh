#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_14_tests.log
tail -n 8 gpurun_out/r2_14_tests.log
: > gpurun_out/r2_14_exp.log
EXP_OPTS=band_kernel=4,concurrency=0 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
EXP_OPTS=band_kernel=4,concurrency=1 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
EXP_OPTS=band_kernel=4,concurrency=1,sweep_groups=1 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
EXP_OPTS=band_kernel=2 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=262144 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_14_exp.log 2>&1
cat gpurun_out/r2_14_exp.log
