#!/bin/bash
# round 2, GPU run 3: full ncu captures of two sweep kernels ({1,2} and {3}) of the second step
set -x
mkdir -p gpurun_out
export EXP_OPTS=band_kernel=4
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 8 --launch-count 2 -f -k regex:rtrnmc_kernel -o gpurun_out/r2_03_sweeps python tools/dev/exp_bands.py > gpurun_out/r2_03_ncu.log 2>&1
tail -n 3 gpurun_out/r2_03_ncu.log
