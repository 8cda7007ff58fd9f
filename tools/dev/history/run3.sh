#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-cols 16384 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.3e col/s  ms/step %.1f  band ms %.1f all ms %.1f  fp64 frac %.4f  e2e %.3e launches %d clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['band_kernels_ms_per_step'],d['roofline']['all_kernels_ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches'],d['clocks']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'band_kernel|prep_kernel|mcica|finalize' -s 57 -c 19 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --ncol 262144 --no-cpu-baseline --e2e-cols 1024 > gpurun_out/ncu_bench.log 2>&1
