#!/bin/bash
# forced register budgets: 2 (ptxas default) / 3 / 4 CTAs per SM
mkdir -p gpurun_out
L=$PWD/rrtmg_lw_b200/lib
{
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200.so EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vC.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_minb.log 2>&1
cat gpurun_out/exp_minb.log
