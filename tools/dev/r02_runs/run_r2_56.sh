#!/bin/bash
# pass size against wave quantisation of the sweep grids (444 / 592 / 740 / 888 resident CTAs)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576
for c in 454656 349568 606208 699136; do
  EXP_CHUNK=$c python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee -a gpurun_out/r2_56_chunks.jsonl
done
