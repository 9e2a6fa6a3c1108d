"""Full-size check and CPU timing of remove_tips!: the oracle builds the graphs of scripts/time_remove_tips.py on the CPU
(same generator, same parameters), the C restatement of the reference's sequential tip removal (oracle/graph_edit.c)
edits them, and the sizes before / after, the passes and the number of tips are compared with the GPU's line
(profiles/r2_g_remove_tips_timing.json).  Prints one JSON line per case with the CPU time next to the GPU time.

    python tests/full_size_tips_check.py [genome_len] [coverage] [gpu_json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import graph_edit as E  # noqa: E402
from oracle import oracle as O      # noqa: E402


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    cov = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    gpu_file = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r2_g_remove_tips_timing.json")
    gpu = [json.loads(x) for x in open(gpu_file).read().strip().splitlines()] if os.path.exists(gpu_file) else []
    rl, k = 150, 31
    nreads = genome * cov // rl // 2 * 2
    threads = max(1, (os.cpu_count() or 1))
    seq, offs = O.synth_reads(2, genome, nreads, rl, True, 500, 10000, threads=threads)
    keys, cnt = O.count_kmers(seq, offs, k, threads=threads)
    del seq
    ok = True
    for min_freq, min_size in ((2, 200), (1, 60)):
        t0 = time.perf_counter()
        og = O.dbg(O.filter_counts(keys, cnt, k, min_freq), k)
        t_build = time.perf_counter() - t0
        c = E.CGraph(og.node_off, og.bases, og.link_row, og.link_tri)
        before = c.sizes()
        del og
        t0 = time.perf_counter()
        passes, removed = c.remove_tips(min_size)
        t_tips = time.perf_counter() - t0
        after = c.sizes()
        tips_left = len(c.find_tip_nodes(min_size))
        c.free()
        wl = "%d bp uniform genome, %dx 150 bp paired, 1%% substitutions, k=31, min-count %d" % (genome, cov, min_freq)
        g = next((x for x in gpu if x["workload"] == wl and x["min_size"] == min_size), None)
        line = {"workload": wl, "min_size": min_size, "cpu": {"nodes_bases_links_before": list(before), "nodes_bases_links_after": list(after),
                                                               "passes": passes, "tips_removed": removed, "tips_left": tips_left,
                                                               "remove_tips_ms": round(1e3 * t_tips, 1), "graph_build_s": round(t_build, 1),
                                                               "what": "oracle/graph_edit.c, 1 thread (the reference is sequential)"}}
        if g is not None:
            same = (list(before) == g["nodes_bases_links_before"] and list(after) == g["nodes_bases_links_after"]
                    and passes == g["passes"] and removed == g["tips_removed"] and tips_left == g["tips_left"])
            line["gpu"] = {"file": os.path.relpath(gpu_file, ROOT), "ms": g["ms"]}
            line["equal"] = same
            ok = ok and same
        print(json.dumps(line), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
