#!/bin/bash
# round-1 v7 evidence: launch list, per-kernel DRAM traffic, full capture of band 3, bench lines, reference arm
mkdir -p gpurun_out
K='band_kernel|prep_kernel|mcica|finalize'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 19 -c 38 --csv --log-file gpurun_out/r01_v9_launches.csv python bench.py --steps 1 --warmup 1 --ncol 524288 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"$K" -s 19 -c 19 --csv --log-file gpurun_out/r01_v9_traffic.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 2 -c 1 -o gpurun_out/prof_v9b3 python bench.py --steps 1 --warmup 0 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_c.log 2>&1
ncu -i gpurun_out/prof_v9b3.ncu-rep --page details > gpurun_out/r01_v9_band3_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/prof_v9b3.ncu-rep --page raw --csv > gpurun_out/r01_v9_band3_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_v9b3.ncu-rep --page source --csv > gpurun_out/prof_v9b3_source.csv 2>/dev/null
gzip -f gpurun_out/prof_v9b3_source.csv; rm -f gpurun_out/prof_v9b3.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_v9_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r01_v9_bench.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_v9_bench_reference.json 2>> gpurun_out/bench.err; tail -c 300 gpurun_out/r01_v9_bench_reference.json
