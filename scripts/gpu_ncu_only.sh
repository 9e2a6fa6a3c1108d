#!/bin/bash
# usage: scripts/gpu_ncu_only.sh <tag> <kernel-regex>... -- one plain run, then ncu --set full per kernel
mkdir -p gpurun_out
T=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$T.json 2> gpurun_out/plain_$T.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$T.err; exit 1; }
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/full_${T}_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${T}_$K.log 2>&1; echo "ncu $K rc=$?"
done
