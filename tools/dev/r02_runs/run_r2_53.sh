#!/bin/bash
# KISS jumps through 4-bit window tables: GPU tests + step time
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_53_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_53_tests.log
tail -n 3 gpurun_out/r2_53_tests.log
EXP_NCOL=1048576 EXP_CHUNK=524288 python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_53_kiss_windows.json
