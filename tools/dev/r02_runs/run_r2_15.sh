#!/bin/bash
set -x
mkdir -p gpurun_out
nproc
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err; echo "rc=$?"
tail -n 5 gpurun_out/r2_15_bench.err
cat gpurun_out/r2_15_bench.json
