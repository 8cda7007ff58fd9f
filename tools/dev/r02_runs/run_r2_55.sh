#!/bin/bash
# evidence pass v7 (taumol band groups, pass-size stream schedule, KISS window tables): tests, smoke, bench (+ reference
# arm), launch list of the bench command, per-kernel DRAM traffic of one step, full captures of the dominant kernels
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_55_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_55_tests.log
tail -n 3 gpurun_out/r2_55_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_55_bench.json 2> gpurun_out/r2_55_bench.err
tail -c 1500 gpurun_out/r2_55_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_55_bench_reference.json 2> gpurun_out/r2_55_bench_reference.err
tail -c 400 gpurun_out/r2_55_bench_reference.json
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'colprep|taumol|rtrnmc|kiss|finalize|cldoptics|fp64_peak' -c 400 --csv --log-file gpurun_out/r2_55_launches.csv python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline > gpurun_out/r2_55_ncu_bench.log 2>&1
tail -n 2 gpurun_out/r2_55_ncu_bench.log
# DRAM traffic of one step (1 M columns, 524288-column passes, serial kernels: 19 launches per pass)
EXP_NCOL=1048576 EXP_CHUNK=524288 EXP_OPTS=concurrency=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'colprep|taumol|rtrnmc|kiss|finalize|cldoptics' --launch-skip 76 --launch-count 38 --csv --log-file gpurun_out/r2_55_traffic.csv python tools/dev/exp_bands.py > gpurun_out/r2_55_ncu_traffic.log 2>&1
tail -n 2 gpurun_out/r2_55_ncu_traffic.log
# full captures: both taumol launches of a pass, and the band 3 / 4 / 5 sweeps
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 2 --launch-count 2 -f -k regex:taumol_kernel -o gpurun_out/r2_55_taumol python tools/dev/exp_bands.py > gpurun_out/r2_55_ncu_t.log 2>&1
ls -la gpurun_out/r2_55_*
