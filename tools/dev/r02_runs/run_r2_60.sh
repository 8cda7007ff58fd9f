#!/bin/bash
# cloud optics with the band loop unrolled (no local-memory arrays): GPU tests + stage times
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_60_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_60_tests.log
tail -n 3 gpurun_out/r2_60_tests.log
EXP_NCOL=1048576 EXP_CHUNK=524288 python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_60_cldoptics_unroll.json
