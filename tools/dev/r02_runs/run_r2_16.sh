#!/bin/bash
set -x
mkdir -p gpurun_out
# launch list + DRAM traffic of one step of the default bench workload (1 M columns, 524288-column passes)
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=band_kernel=4,concurrency=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'colprep|taumol|rtrnmc|kiss|finalize|cldoptics' --launch-skip 84 --launch-count 42 --csv --log-file gpurun_out/r2_16_traffic.csv python tools/dev/exp_bands.py > gpurun_out/r2_16_ncu.log 2>&1
tail -n 3 gpurun_out/r2_16_ncu.log
