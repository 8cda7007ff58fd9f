#!/bin/bash
# final tree (eta sort key on): tests, smoke, bench
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_58_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_58_tests.log
tail -n 3 gpurun_out/r2_58_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_58_bench.json 2> gpurun_out/r2_58_bench.err
tail -c 300 gpurun_out/r2_58_bench.json
