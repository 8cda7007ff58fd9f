#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-cols 16384 --band-kernel 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v3.json')); print(d['value'], d['roofline']['band_kernels_ms_per_step'], d['roofline']['all_kernels_ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench_v3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gband|prep_kernel|mcica|finalize|transpose' -c 22 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1
awk -F'","' 'NR>3{print $5, $NF}' gpurun_out/launches_v3.csv | cut -c1-90 | sed -n 1,24p
