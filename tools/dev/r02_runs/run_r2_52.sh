#!/bin/bash
# sweeps over 2 / 3 streams at 524288-column passes; one pass of 1 048 576 columns (131 GB of scratch)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576
for o in 4 5; do
  EXP_CHUNK=524288 EXP_OPTS=concurrency=$o python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee -a gpurun_out/r2_52.jsonl
done
nvidia-smi --query-gpu=memory.total,memory.used --format=csv | tee -a gpurun_out/r2_52.jsonl
EXP_CHUNK=1048576 EXP_OPTS=concurrency=0 timeout 120 python tools/dev/exp_bands.py 2>&1 | tail -n 2 | tee -a gpurun_out/r2_52.jsonl
EXP_CHUNK=1048576 EXP_OPTS=concurrency=1 timeout 120 python tools/dev/exp_bands.py 2>&1 | tail -n 2 | tee -a gpurun_out/r2_52.jsonl
