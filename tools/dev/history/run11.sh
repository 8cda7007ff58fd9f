#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 2 -c 1 -o gpurun_out/prof_k2b3 python bench.py --steps 1 --warmup 0 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 --band-kernel 2 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
for n in k2b3; do
ncu -i gpurun_out/prof_$n.ncu-rep --page details > gpurun_out/prof_${n}_details.txt 2>/dev/null
ncu -i gpurun_out/prof_$n.ncu-rep --page source --csv > gpurun_out/prof_${n}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${n}_source.csv
rm -f gpurun_out/prof_$n.ncu-rep
done
