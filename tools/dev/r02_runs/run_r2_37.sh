#!/bin/bash
cd /root/repo
python tools/dev/exp_bands.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages_ms']; print('taumol', s[0], 'sweeps', round(d['bands_total_ms']-s[0],2), s[1:], 'step', d['step_ms'])"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
