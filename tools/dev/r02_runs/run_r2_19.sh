#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python tools/dev/exp_bands.py > gpurun_out/r2_19_stages.txt 2>&1
tail -1 gpurun_out/r2_19_stages.txt
EXP_OPTS=kiss_sparse=0 python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_NCOL=1048576 EXP_CHUNK=524288 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'kiss|cldoptics|colprep|finalize' --launch-skip 16 --launch-count 8 --csv --log-file gpurun_out/r2_19_small.csv python tools/dev/exp_bands.py > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_19_small.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0]!='ID': print(r[4][:40], r[12], r[14])
"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_19_bench.json 2> gpurun_out/r2_19_bench.err
cat gpurun_out/r2_19_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
