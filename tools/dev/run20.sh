#!/bin/bash
# chunk-size sensitivity (TLB reach / DRAM page locality of the plane-strided scratch)
mkdir -p gpurun_out
{
EXP_NCOL=151552 EXP_CHUNK=37888 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_NCOL=151552 EXP_CHUNK=151552 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_NCOL=262144 EXP_CHUNK=65536 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_NCOL=262144 EXP_CHUNK=262144 timeout 600 python tools/dev/exp_bands.py 2>&1 | tail -1
} > gpurun_out/exp_chunk.log 2>&1
cat gpurun_out/exp_chunk.log
