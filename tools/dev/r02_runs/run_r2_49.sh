#!/bin/bash
# which half of the multi-stream schedule costs time at large passes: concurrency 0 = serial, 1 = masks beside taumol +
# sweeps over four streams, 2 = masks beside taumol only, 3 = sweeps over four streams only
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576
for c in 524288 131072; do
for o in 0 1 2 3; do
  EXP_CHUNK=$c EXP_OPTS=concurrency=$o python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee -a gpurun_out/r2_49_conc.jsonl
done; done
