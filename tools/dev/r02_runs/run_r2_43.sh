#!/bin/bash
# timeline of the resident work-queue kernel beside taumol, with and without the high-priority stream
cd /root/repo
mkdir -p gpurun_out
python tools/dev/exp_beside.py 2>&1 | grep '^{' | tee gpurun_out/r2_43_beside.jsonl
RRTMG_LW_BESIDE_NOPRIO=1 EXP_BESIDE=1,2 python tools/dev/exp_beside.py 2>&1 | grep '^{' | tee -a gpurun_out/r2_43_beside.jsonl
