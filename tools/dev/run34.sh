#!/bin/bash
mkdir -p gpurun_out
EXP_OPTS="precision=32" timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1 > gpurun_out/exp_fp32_v8.log; cat gpurun_out/exp_fp32_v8.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
