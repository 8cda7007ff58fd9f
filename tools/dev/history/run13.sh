#!/bin/bash
mkdir -p gpurun_out
for v in "" _mb3; do
export RRTMG_LW_B200_LIB=$PWD/rrtmg_lw_b200/lib/librrtmg_lw_b200$v.so
echo "== lib $v"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --warmup 3 --ncol 524288 --no-cpu-baseline --e2e-cols 16384 > gpurun_out/bench$v.json 2> gpurun_out/bench$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench$v.json')); print(d['value'], d['roofline']['band_kernels_ms_per_step'], d['roofline']['all_kernels_ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench$v.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'band_kernel' -c 16 --csv --log-file gpurun_out/launches$v.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1
awk -F'","' 'NR>3{print $NF}' gpurun_out/launches$v.csv | tr -d '"' | tr '\n' ' '; echo
done
