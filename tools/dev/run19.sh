#!/bin/bash
# L1 prefetch of all k-table rows of a layer right after the band-level setup
mkdir -p gpurun_out
L=$PWD/rrtmg_lw_b200/lib
{
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200.so EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -2
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -2
} > gpurun_out/exp_pf.log 2>&1
cat gpurun_out/exp_pf.log
