#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rrtmg_lw_b200/lib
{
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200.so EXP_SAVE=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so EXP_REF=/tmp/ref.npz timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_tailfma.log 2>&1; cat gpurun_out/exp_tailfma.log
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
