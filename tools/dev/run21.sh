#!/bin/bash
# full ncu capture (with SASS source view) of band 1..3 kernels of the current build
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_kernel -c 3 -o gpurun_out/prof_v6 python tools/dev/exp_bands.py > gpurun_out/ncu_v6.log 2>&1
tail -3 gpurun_out/ncu_v6.log
ncu -i gpurun_out/prof_v6.ncu-rep --page details > gpurun_out/r01_v6_b123_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/prof_v6.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/prof_v6_b1_source.csv 2>/dev/null
ncu -i gpurun_out/prof_v6.ncu-rep --page source --csv --launch-skip 2 --launch-count 1 > gpurun_out/prof_v6_b3_source.csv 2>/dev/null
ncu -i gpurun_out/prof_v6.ncu-rep --page raw --csv > gpurun_out/r01_v6_b123_ncu_raw.csv 2>/dev/null
gzip -f gpurun_out/prof_v6_b1_source.csv gpurun_out/prof_v6_b3_source.csv
rm -f gpurun_out/prof_v6.ncu-rep
ls -la gpurun_out | tail -8
