#!/bin/bash
# usage: scripts/gpu_ncu.sh <kernel-regex> <outname> [skip] -- plain run, then ncu --set full on the kernel
mkdir -p gpurun_out
K=$1; O=$2; S=${3:-3}
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$O.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/$O python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$O.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$O.log | cut -c1-300
