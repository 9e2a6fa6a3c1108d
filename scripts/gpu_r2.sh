#!/bin/bash
# usage: scripts/gpu_r2.sh <tag> [pytest-k-expr]   -- on the GPU box: GPU tests, then bench lines of C2 and C3
tag=${1:-x}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 -k "$kexpr" > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?"
else
  timeout 1200 python -m pytest tests -m gpu -x -q --timeout=150 > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?"
fi
tail -30 gpurun_out/tests_$tag.log
for wl in c2_ecoli50x_k31 c3_100mbp30x_k31; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/bench_${tag}_$wl.json 2> gpurun_out/bench_${tag}_$wl.err; echo "bench $wl rc=$?"; tail -c 1200 gpurun_out/bench_${tag}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}_$wl.json').read().strip().splitlines()[-1])
    print('$wl ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches_per_step'])
    print(d['stage_ms_last_step'])
    for t in d['top_kernels_ms']: print('  ', t)
    print('parity', d['parity_checked'])
    print('cpu', d['cpu_baseline'])
except Exception as e: print('no json', e)
PY
done
