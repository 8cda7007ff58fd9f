#!/bin/bash
cd /root/repo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_23_bench_reference.json 2> gpurun_out/r2_23_bench_reference.err
python -c "import json; d=json.loads(open('gpurun_out/r2_23_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
python bench.py > gpurun_out/r2_23_bench.json 2> gpurun_out/r2_23_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2_23_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], {k: round(v['value']) for k, v in d['other_configs'].items()})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'colprep|taumol|rtrnmc|kiss|finalize|cldoptics|fp64_peak|band_kernel|prep_kernel' -c 400 --csv --log-file gpurun_out/r2_23_launches.csv python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r2_23_ncu_bench.log 2>&1
tail -n 2 gpurun_out/r2_23_ncu_bench.log
