#!/bin/bash
# N-GPU gathered bench with fewer NCCL p2p channels (fewer SMs held by the gather's send / receive kernels)
cd /root/repo
N=$1
for ch in ${CHANNELS:-2 4}; do
  NCCL_MAX_P2P_NCHANNELS=$ch NCCL_MIN_P2P_NCHANNELS=${MINCH:-1} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra-configs > gpurun_out/r2_28_bench_n${N}_ch$ch.json 2> gpurun_out/r2_28_bench_n${N}_ch$ch.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r2_28_bench_n${N}_ch$ch.json').read().strip().splitlines()[-1]); print('channels', $ch, d['value'], d['ms_per_step'], d['gather']['ungathered'])"
done
