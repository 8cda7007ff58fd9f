#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_17_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_17_tests.log
tail -n 30 gpurun_out/r2_17_tests.log
