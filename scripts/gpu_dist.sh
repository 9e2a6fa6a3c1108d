#!/bin/bash
# usage: scripts/gpu_dist.sh <tag> <ngpus> [bench-workload]  -- multi-GPU parity tests (IPC and NCCL-only routes), then bench.py at N GPUs
tag=${1:-x}; N=${2:-2}; wl=${3:-c3_100mbp30x_k31}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/dist_${tag}_gpus.txt
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --timeout=600 > gpurun_out/dist_${tag}_tests.log 2>&1; echo "dist tests rc=$?"; tail -25 gpurun_out/dist_${tag}_tests.log
GG_DIST_NO_IPC=1 timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --timeout=600 > gpurun_out/dist_${tag}_tests_noipc.log 2>&1; echo "dist tests (no IPC) rc=$?"; tail -5 gpurun_out/dist_${tag}_tests_noipc.log
if [ "$N" -gt 1 ]; then
  nvidia-smi nvlink -gt d -i 0 > gpurun_out/dist_${tag}_nvlink_before.txt 2>&1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err
  echo "bench N=$N rc=$?"; nvidia-smi nvlink -gt d -i 0 > gpurun_out/dist_${tag}_nvlink_after.txt 2>&1; grep -v OMP_NUM gpurun_out/scale_${tag}_n$N.err | tail -8
  python - gpurun_out/scale_${tag}_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.4g k-mers/s  ms/step %.2f  e2e %.2f ms  stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stage_ms_last_step"].items()}))
    for k,v in list(d["all_kernels_ms"].items())[:14]: print("   ", k, v)
    print("  parity", d["parity_checked"])
except Exception as e: print("  ERR", e)
PY
fi
