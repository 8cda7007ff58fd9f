#!/bin/bash
# KISS generator at 40 registers (10 CTAs per SM) / 32 registers (12) against 48 (8)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
for v in k40 k12; do xz -d -c gpurun_stage/$v.so.xz > /tmp/librrtmg_lw_b200_$v.so & done
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_50_base.json
wait
for v in k40 k12; do
  RRTMG_LW_B200_LIB=/tmp/librrtmg_lw_b200_$v.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_50_$v.json
done
