#!/bin/bash
# usage: scripts/gpu_final.sh <tag>  -- one GPU: tests, the bench lines of every single-GPU workload, launch list, ncu --set full of the top kernels
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$tag.log
timeout 300 python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "bench reference rc=$?"
for wl in c2_ecoli50x_k31 c2_ecoli50x_k63 c5_100mbp30x_k127; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/bench_${tag}_$wl.json 2> gpurun_out/bench_${tag}_$wl.err; echo "bench $wl rc=$?"; tail -c 300 gpurun_out/bench_${tag}_$wl.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:(skm_leaf_kernel|lookup_kernel|gs_walk_kernel|skm_scan_kernel|skm_l1_kernel|gs_emit_kernel)' -s 0 -c 7 -f -o gpurun_out/full_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for n in ("default","c2_ecoli50x_k31","c2_ecoli50x_k63","c5_100mbp30x_k127"):
    try:
        d=json.loads(open("gpurun_out/bench_${tag}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.2f value %.4g e2e %.2f ms parity %s roofline %s %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["parity_checked"]["ok"], d["roofline"]["kernel"], d["roofline"]["frac"] or -1))
    except Exception as e: print(n, "ERR", e)
PY
