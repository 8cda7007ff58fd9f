#!/bin/bash
set -x
mkdir -p gpurun_out
EXP_OPTS=band_kernel=4 timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 1 --launch-count 1 -f -k regex:mcica_kiss_par -o gpurun_out/r2_12_kiss python tools/dev/exp_bands.py > gpurun_out/r2_12_ncu.log 2>&1
tail -n 3 gpurun_out/r2_12_ncu.log
