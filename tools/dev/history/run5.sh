#!/bin/bash
# ncu passes only: launch list of our kernels + full capture of bands 1..3 (three launches), CSV exports kept small
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'band_kernel|prep_kernel|mcica|finalize' -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 48 -c 3 -o gpurun_out/prof_b123 python bench.py --steps 1 --warmup 1 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
ncu -i gpurun_out/prof_b123.ncu-rep --page raw --csv > gpurun_out/prof_b123_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_b123.ncu-rep --page details > gpurun_out/prof_b123_details.txt 2>/dev/null
ncu -i gpurun_out/prof_b123.ncu-rep --page source --csv > gpurun_out/prof_b123_source.csv 2>/dev/null
ls -la gpurun_out
sz=$(stat -c %s gpurun_out/prof_b123.ncu-rep || echo 0); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/prof_b123.ncu-rep; fi
gzip -f gpurun_out/prof_b123_source.csv
du -sh gpurun_out
