#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k2.json')); print(d['value'], d['roofline']['band_kernels_ms_per_step'], d['roofline']['all_kernels_ms_per_step'], d['e2e'])"; tail -3 gpurun_out/bench_k2.err
