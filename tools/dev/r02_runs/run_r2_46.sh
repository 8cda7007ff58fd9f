#!/bin/bash
# experiment: resident CTAs of the narrow sweep kernels (LW_SWEEP_MINB_6G / 8G / 10G / 12G); variant libraries travel
# xz-compressed under gpurun_stage/
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
for v in w1 w2 w3 w4; do xz -d -c gpurun_stage/$v.so.xz > /tmp/librrtmg_lw_b200_$v.so & done
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_46_base.json
wait
for v in w1 w2 w3 w4; do
  RRTMG_LW_B200_LIB=/tmp/librrtmg_lw_b200_$v.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_46_$v.json
done
