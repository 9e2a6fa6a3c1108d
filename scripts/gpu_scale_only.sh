#!/bin/bash
# usage: scripts/gpu_scale_only.sh <tag> <ngpus> [workload] [extra bench args]  -- bench.py at N GPUs (no tests)
tag=$1; N=$2; wl=${3:-c3_100mbp30x_k31}; shift 3
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 --workload $wl "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err
echo "bench N=$N rc=$?"; grep -v OMP_NUM gpurun_out/scale_${tag}_n$N.err | tail -5
python - gpurun_out/scale_${tag}_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.4g k-mers/s  ms/step %.2f  e2e %.2f ms  stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stage_ms_last_step"].items()}))
    for k,v in list(d["all_kernels_ms"].items())[:12]: print("   ", k, v)
    print("  parity", d["parity_checked"])
except Exception as e: print("  ERR", e)
PY
