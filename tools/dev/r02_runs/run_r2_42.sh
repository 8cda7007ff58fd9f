#!/bin/bash
# experiment: McICA masks + cloud optics as a resident work-queue kernel beside taumol (mcica_beside = CTAs per SM)
cd /root/repo
mkdir -p gpurun_out
export EXP_NCOL=1048576 EXP_CHUNK=524288
EXP_OPTS= EXP_SAVE=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_42_base.json
for k in 1 2; do
  EXP_OPTS=mcica_beside=$k EXP_REF=/tmp/ref.npz python tools/dev/exp_bands.py 2>&1 | tail -n 1 | tee gpurun_out/r2_42_beside$k.json
done
RRTMG_LW_MCICA_BESIDE=2 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_42_tests_beside2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_42_tests_beside2.log
tail -n 4 gpurun_out/r2_42_tests_beside2.log
