#!/bin/bash
# racecheck (shared-memory hazards) of the kernels with CTA-level cooperation: taumol sort, KISS lists, MT fill / jump
cd /root/repo
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_nomcica_gpu.py -x -q -m gpu -k "mcica_mask_bit_exact or kiss_stepping or mersenne_twister_stream or ragged or nomcica_error" > gpurun_out/r2_33_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -n 8 gpurun_out/r2_33_racecheck.log
