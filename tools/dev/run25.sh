#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1 > gpurun_out/exp_v7.log; cat gpurun_out/exp_v7.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
