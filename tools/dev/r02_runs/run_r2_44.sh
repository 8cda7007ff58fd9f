#!/bin/bash
# experiment: taumol split into two launches -- single-key-species bands (mask 0xFEA3) compiled for more resident CTAs
# (80 / 96 / 64 registers), the binary-species bands 3 4 5 7 9 at 128 (or 168) registers
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_44_base.json
for v in s6 s5 s8 s6m3; do
  RRTMG_LW_B200_LIB=/root/repo/rrtmg_lw_b200/lib/librrtmg_lw_b200_$v.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_44_$v.json
done
