#!/bin/bash
# occupancy sensitivity: 2 CTAs/SM (default) vs 1 CTA/SM (extra dynamic smem)
mkdir -p gpurun_out
{
RRTMG_LW_SMEM_PAD=0 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
RRTMG_LW_SMEM_PAD=100000 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_occ.log 2>&1
cat gpurun_out/exp_occ.log
