#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 48 -c 3 -o gpurun_out/prof_b123 python bench.py --steps 1 --warmup 1 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
ncu -i gpurun_out/prof_b123.ncu-rep --page raw --csv > gpurun_out/prof_b123_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_b123.ncu-rep --page details > gpurun_out/prof_b123_details.txt 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'band_kernel|prep_kernel|mcica|finalize' -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
