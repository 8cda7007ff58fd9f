#!/bin/bash
# whole-row / half-row load bursts vs one 256-bit block per evaluation
mkdir -p gpurun_out
L=$PWD/rrtmg_lw_b200/lib
{
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200.so EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -2
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -2
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vC.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -2
} > gpurun_out/exp_rows.log 2>&1
cat gpurun_out/exp_rows.log
