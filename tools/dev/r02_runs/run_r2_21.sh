#!/bin/bash
# multi-GPU bench lines (N GPUs of one box): $1 = N
cd /root/repo
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/r2_21_bench_n$N.json 2> gpurun_out/r2_21_bench_n$N.err
tail -2 gpurun_out/r2_21_bench_n$N.err
cat gpurun_out/r2_21_bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d.get('extras'), d.get('gather'))"
