#!/bin/bash
# serial (concurrency=0) against the multi-stream schedule, after the taumol split
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576
for c in 524288 262144; do
for o in concurrency=1 concurrency=0; do
  EXP_CHUNK=$c EXP_OPTS=$o python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee -a gpurun_out/r2_48_conc.jsonl
done; done
