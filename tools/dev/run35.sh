#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_OPTS="precision=32" timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_block.log; cat gpurun_out/exp_block.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
