#!/bin/bash
mkdir -p gpurun_out
export RRTMG_LW_BAND_KERNEL=2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 --ncol 524288 --no-cpu-baseline --e2e-cols 16384 --band-kernel 2 > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k2.json')); print(d['value'], d['roofline']['band_kernels_ms_per_step'], d['roofline']['all_kernels_ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench_k2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'band_kernel' -c 16 --csv --log-file gpurun_out/launches_k2.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 --band-kernel 2 > gpurun_out/ncu_bench.log 2>&1
awk -F'","' 'NR>3{print $5, $NF}' gpurun_out/launches_k2.csv | cut -c1-70 | tr '\n' ' '; echo
