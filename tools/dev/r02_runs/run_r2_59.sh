#!/bin/bash
# N = 2 sanity of the final tree: gathered bench line through the C-ABI gather
cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline > gpurun_out/r2_59_bench_n2.json 2> gpurun_out/r2_59_bench_n2.err
echo "rc=$?"; tail -c 1200 gpurun_out/r2_59_bench_n2.json; tail -n 3 gpurun_out/r2_59_bench_n2.err
