#!/bin/bash
# KISS sparse stepping + MT jump-ahead: parity, stage timing, bench
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_colmode_gpu.py -x -q -m gpu 2>&1 | tail -5
python tools/dev/exp_bands.py > gpurun_out/r2_18_stages.txt 2>&1
tail -3 gpurun_out/r2_18_stages.txt
python - > gpurun_out/r2_18_mt.txt 2>&1 <<'PY'
import time, torch, sys
sys.path.insert(0, '.')
import bench
from rrtmg_lw_b200 import api, driver, kdata
api.rrtmg_lw_ini(1004.0, kdata.synthetic_blob(), 4)
for ncol in (16384, 262144):
    b = bench.make_device_batch(ncol, "cuda:0", 0)
    for irng in (0, 1):
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            driver.run_fused(b, icld=2, irng=irng, device="cuda:0"); torch.cuda.synchronize()
            print("ncol", ncol, "irng", irng, "rep", rep, "ms", (time.time() - t) * 1e3, flush=True)
PY
cat gpurun_out/r2_18_mt.txt | tail -10
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_18_bench.json 2> gpurun_out/r2_18_bench.err
cat gpurun_out/r2_18_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d.get('extras'))"
