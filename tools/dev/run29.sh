#!/bin/bash
mkdir -p gpurun_out
L=$PWD/rrtmg_lw_b200/lib
{
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vB.so EXP_OPTS="precision=32" timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
RRTMG_LW_B200_LIB=$L/librrtmg_lw_b200_vC.so EXP_OPTS="precision=32" timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_fp32_minb.log 2>&1; cat gpurun_out/exp_fp32_minb.log
