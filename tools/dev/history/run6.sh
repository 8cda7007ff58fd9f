#!/bin/bash
# first run of the g-lane kernel: parity suite, then bench v3 vs v2, then a launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-cols 16384 --band-kernel 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -c 1500 gpurun_out/bench_v3.json; tail -3 gpurun_out/bench_v3.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-cols 16384 --band-kernel 2 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 600 gpurun_out/bench_v2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1
grep -o '"Kernel Name.*' gpurun_out/launches_v3.csv | head -2; awk -F'","' 'NR>3{print $5, $NF}' gpurun_out/launches_v3.csv | sed -n 1,30p
