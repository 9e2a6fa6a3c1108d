#!/bin/bash
# usage: scripts/gpu_sweep.sh "ENV1=a ENV2=b" "ENV1=c" ... -- one short bench per environment setting
mkdir -p gpurun_out
for e in "$@"; do
  env $e timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "$e FAILED"; tail -5 gpurun_out/sweep.err; continue; }
  python - "$e" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/sweep.json').read())
st=d['stage_ms_last_step']
print(sys.argv[1], '| ms/step %.2f e2e %.2f |' % (d['ms_per_step'], d['e2e']['ms_per_step']), ' '.join('%s=%.2f'%(k,v) for k,v in st.items()), '|', ' '.join('%s:%.2f'%(t['kernel'].replace('_kernel',''),t['ms']) for t in d['top_kernels_ms'][:6]))
PY
done
