#!/bin/bash
# usage: scripts/gpu_profile.sh <tag> [kernel-regex ...] -- tests, bench, ncu launch list, ncu --set full per kernel
mkdir -p gpurun_out
T=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$T.log
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$T.err
python - gpurun_out/bench_$T.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print('ms_per_step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches_per_step'])
print(d['stage_ms_last_step'])
for t in d['top_kernels_ms']: print('  ', t)
r=d['roofline']; print({k:v for k,v in r.items() if k!='other_kernels'})
for t in r.get('other_kernels',[]): print('  ', t)
print(d['cpu_baseline'])
PY
# launch list of one step (after the plain run above exited 0)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$T.log 2>&1; echo "launch list rc=$?"
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/full_${T}_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${T}_$K.log 2>&1; echo "ncu $K rc=$?"
done
