#!/bin/bash
# usage: scripts/gpu_variants.sh <tag> <workload> VAR=val[,VAR=val] ...   -- one short bench per environment setting, leaf / stage times printed
tag=$1; wl=$2; shift 2
mkdir -p gpurun_out
for setting in "$@"; do
  envs=$(echo "$setting" | tr ',' ' ')
  name=$(echo "$setting" | tr -c 'A-Za-z0-9\n' '_')
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/var_${tag}_$name.json 2> gpurun_out/var_${tag}_$name.err
  python - gpurun_out/var_${tag}_$name.json "$setting" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ak=d["all_kernels_ms"]
    print("%-40s ms/step %7.2f  %s" % (sys.argv[2], d["ms_per_step"], "  ".join("%s %.2f" % (k.replace("_kernel",""), ak[k][1]) for k in list(ak)[:7])))
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
