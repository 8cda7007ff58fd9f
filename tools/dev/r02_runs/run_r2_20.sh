#!/bin/bash
cd /root/repo
echo "== base (tail_group=1)"; python tools/dev/exp_bands.py 2>&1 | tail -1
echo "== base tail_group=0"; EXP_OPTS=tail_group=0 python tools/dev/exp_bands.py 2>&1 | tail -1
export RRTMG_LW_B200_LIB=$PWD/rrtmg_lw_b200/lib/librrtmg_lw_b200_tma.so
echo "== TMA bulk-copy variant"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "clear_sky or fused_cloudy or strict or stage_dumps or ragged" 2>&1 | tail -3
python tools/dev/exp_bands.py 2>&1 | tail -1
EXP_ICLD=0 python tools/dev/exp_bands.py 2>&1 | tail -1
unset RRTMG_LW_B200_LIB
echo "== base clear"; EXP_ICLD=0 python tools/dev/exp_bands.py 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_nomcica_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err
cat gpurun_out/r2_20_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
