#!/bin/bash
# GPU pass: gpu tests, smoke, bench (with cpu baseline), reference arm, ncu launch list, ncu full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader; nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 32 -c 16 -o gpurun_out/prof_bands python bench.py --steps 1 --warmup 1 --ncol 131072 --chunk 131072 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
ncu -i gpurun_out/prof_bands.ncu-rep --page raw --csv > gpurun_out/prof_bands_raw.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
