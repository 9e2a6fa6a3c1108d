#!/bin/bash
# usage: scripts/gpu_scale.sh <tag> N... -- bench.py at each N (torchrun for N > 1), one JSON line per N
mkdir -p gpurun_out
T=$1; shift
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${T}_n$N.json 2> gpurun_out/scale_${T}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${T}_n$N.json 2> gpurun_out/scale_${T}_n$N.err
  fi
  echo "N=$N rc=$?"; tail -c 300 gpurun_out/scale_${T}_n$N.err | grep -v OMP_NUM | tail -3
  python - gpurun_out/scale_${T}_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.4g k-mers/s  ms/step %.2f  e2e %.2f ms  stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stage_ms_last_step"].items()}))
except Exception as e: print("  ERR", e)
PY
done
