#!/bin/bash
# experiment: band groups of the taumol launches (LW_TAUMOL_VARIANT 0..8, see lw_twophase_inst.cu); the variant
# libraries travel xz-compressed under gpurun_stage/ (the snapshot limit is 512 MiB)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
for v in 1 2 3 4 5 6 7 8; do xz -d -c gpurun_stage/v$v.so.xz > /tmp/librrtmg_lw_b200_v$v.so & done
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_45_v0.json
wait
for v in v1 v2 v3 v4 v5 v6 v7 v8; do
  RRTMG_LW_B200_LIB=/tmp/librrtmg_lw_b200_$v.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_45_$v.json
done
