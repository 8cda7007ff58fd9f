#!/bin/bash
# memcheck of the kernels added this round on small cases + the same-atmosphere bound of the taumol kernel
cd /root/repo
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nomcica_gpu.py tests/test_gpu_parity.py -x -q -m gpu -k "nomcica_overlap or nomcica_error or nomcica_host or mcica_mask_bit_exact or kiss_stepping or mersenne_twister_stream or ragged or empty" > gpurun_out/r2_29_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 6 gpurun_out/r2_29_memcheck.log
echo "== taumol, every column the same atmosphere (upper bound of any table-locality optimisation)"
EXP_SAME_ATM=1 python tools/dev/exp_bands.py 2>&1 | tail -1
python tools/dev/exp_bands.py 2>&1 | tail -1
