#!/bin/bash
# end-of-round verification on a fresh box: GPU test suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_v9b_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 250 gpurun_out/r01_v9b_bench_reference.json
timeout 900 python bench.py > gpurun_out/r01_v9b_bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/r01_v9b_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, j['roofline']['frac'], j['e2e']['value'], j['cpu_baseline']['value'], j['clocks'])
PY
