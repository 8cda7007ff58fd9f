#!/bin/bash
mkdir -p gpurun_out
EXP_OPTS="precision=32" timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1 > gpurun_out/exp_fp32_final.log; cat gpurun_out/exp_fp32_final.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --precision 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_v7_bench_fp32.json 2> gpurun_out/bench32.err; tail -c 1500 gpurun_out/r01_v7_bench_fp32.json | head -c 700
