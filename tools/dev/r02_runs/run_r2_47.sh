#!/bin/bash
# experiment: L2 prefetch of the sweeps' state words LW_SWEEP_L2PF layers beyond the cp.async ring (1 / 2 / 4)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
for v in p1 p2 p4; do xz -d -c gpurun_stage/$v.so.xz > /tmp/librrtmg_lw_b200_$v.so & done
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_47_base.json
wait
for v in p1 p2 p4; do
  RRTMG_LW_B200_LIB=/tmp/librrtmg_lw_b200_$v.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_47_$v.json
done
