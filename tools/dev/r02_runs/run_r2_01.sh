#!/bin/bash
# round 2, GPU run 1: parity of the two-phase path + first timings (per stage), old vs new
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_01_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_01_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_01_tests.log
tail -5 gpurun_out/r2_01_tests.log
# per-stage times, old (2) vs new (4), 262144 columns; outputs of the new path compared with the old one's
EXP_OPTS=band_kernel=2 EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py > gpurun_out/r2_01_exp.log 2>&1
EXP_OPTS=band_kernel=4 EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_01_exp.log 2>&1
EXP_OPTS=band_kernel=4,sweep_groups=0 EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_01_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=1048576 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_01_exp.log 2>&1
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=band_kernel=4 timeout 600 python tools/dev/exp_bands.py >> gpurun_out/r2_01_exp.log 2>&1
cat gpurun_out/r2_01_exp.log
# launch list + DRAM traffic of one 262144-column step of the new path
EXP_OPTS=band_kernel=4 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'taumol|rtrnmc|colprep|kiss|finalize' -c 24 --csv --log-file gpurun_out/r2_01_traffic.csv python tools/dev/exp_bands.py > gpurun_out/r2_01_ncu.log 2>&1
tail -3 gpurun_out/r2_01_ncu.log
