#!/bin/bash
# usage: scripts/gpu_check.sh [bench-args...]  -- run on the GPU box: tests then one bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read())
    print('ms_per_step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches_per_step'])
    print(d['stage_ms_last_step'])
    for t in d['top_kernels_ms']: print('  ', t)
    print(d['roofline'])
except Exception as e: print('no json', e)
PY
