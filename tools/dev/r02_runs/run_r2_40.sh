#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_colmode_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/dev/exp_two_call.py 2>&1 | tail -5
