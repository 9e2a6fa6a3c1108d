#!/bin/bash
# usage: scripts/gpu_ncu2.sh <tag> <workload> <skip> <count> <kernel-regex> [more kernel regexes]   -- plain run, then ONE ncu --set full pass
tag=$1; wl=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $wl"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
KR=$(IFS='|'; echo "$*")
ncu --set full --clock-control none --import-source on -k "regex:($KR)" -s $skip -c $cnt -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; grep -E "==PROF==|Error|error" gpurun_out/ncu_$tag.log | head -20
