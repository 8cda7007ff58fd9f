#!/bin/bash
# taumol lane sort with the eta bin as third key (LW_TAUMOL_SORT_ETA=1) against the (jp, jt) sort
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
xz -d -c gpurun_stage/eta1.so.xz > /tmp/librrtmg_lw_b200_eta1.so &
EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_57_base.json
wait
RRTMG_LW_B200_LIB=/tmp/librrtmg_lw_b200_eta1.so EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_57_eta1.json
