#!/bin/bash
# ncu: full capture of gband band 3 and band 14 (kernel-name filter matches the base name only)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gband_kernel -s 2 -c 1 -o gpurun_out/prof_g3 python bench.py --steps 1 --warmup 0 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gband_kernel -s 13 -c 1 -o gpurun_out/prof_g14 python bench.py --steps 1 --warmup 0 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log | cut -c1-300
for n in g3 g14; do
ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$n.ncu-rep --page details > gpurun_out/prof_${n}_details.txt 2>/dev/null
ncu -i gpurun_out/prof_$n.ncu-rep --page source --csv > gpurun_out/prof_${n}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${n}_source.csv
sz=$(stat -c %s gpurun_out/prof_$n.ncu-rep || echo 0); if [ "$sz" -gt 25000000 ]; then rm gpurun_out/prof_$n.ncu-rep; fi
done
du -sh gpurun_out; ls -la gpurun_out
