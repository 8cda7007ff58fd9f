#!/bin/bash
# round 2, GPU run 2: parity suite again + full ncu captures of the taumol kernel and two sweep kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_02_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_02_tests.log
tail -25 gpurun_out/r2_02_tests.log
export EXP_OPTS=band_kernel=4
NCU="ncu --set full --import-source on --clock-control none --kernel-name-base demangled --launch-skip 1 --launch-count 1 -f"
timeout 900 $NCU -k regex:'taumol_kernel' -o gpurun_out/r2_02_taumol python tools/dev/exp_bands.py > gpurun_out/r2_02_ncu1.log 2>&1
timeout 900 $NCU -k regex:'rtrnmc_kernel<3, 1, 1, 0' -o gpurun_out/r2_02_sweep3 python tools/dev/exp_bands.py > gpurun_out/r2_02_ncu2.log 2>&1
timeout 900 $NCU -k regex:'rtrnmc_kernel<1, 2, 1, 0' -o gpurun_out/r2_02_sweep12 python tools/dev/exp_bands.py > gpurun_out/r2_02_ncu3.log 2>&1
tail -2 gpurun_out/r2_02_ncu*.log
ls -la gpurun_out/
