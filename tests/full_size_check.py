#!/usr/bin/env python
"""One-off full-size parity check of a bench workload against the CPU oracle, WITHOUT a GPU.

The oracle builds the WHOLE workload (default: C3, 100 Mbp x 30x, k = 31, min-count 2: 2.4e9 k-mer instances) from the
same deterministic generator as bench.py and reduces the graph to the sizes and the checksum that gg_graph_checksum
returns; bench.py prints the GPU's values for that workload in its `graph` field (1, 2, 4 or 8 GPUs alike).  Equality
of (nodes, bases, link entries, checksum) is the full-size, bit-level parity evidence of the benchmarked result.

The k-mers are counted in 16 slices of the key space so that the collect -> sort -> run-length restatement fits 62 GB of
host memory; slices are disjoint key ranges, so their sorted tables concatenate to the reference's sorted vector.

    python tests/full_size_check.py [workload] [expected-json-line-file] [genome-multiple] [passes]
                                                                       # ~15 min on 8 cores for C3

genome-multiple N builds the job of `bench.py --gpus N --scaling weak` (N x genome, N x reads); with passes > 1 the reads are
generated and scanned `passes` times and every pass keeps only its share of the 64 key-space slices (memory / passes).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table only)
from oracle import oracle as O  # noqa: E402

M64 = (1 << 64) - 1


def sy_mix(z):
    """csrc/synth.cuh sy_mix on uint64 arrays (wrapping arithmetic)"""
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def graph_checksum(g):
    """gg_graph_checksum (csrc/ggb200.cu: checksum_kernel) on the oracle's arrays"""
    acc = 0
    with np.errstate(over="ignore"):
        step = 1 << 24
        for a in range(0, len(g.bases), step):
            i = np.arange(a, min(a + step, len(g.bases)), dtype=np.uint64)
            acc = (acc + int(sy_mix(i * np.uint64(4) + g.bases[a:a + step].astype(np.uint64)).sum(dtype=np.uint64))) & M64
        tri = np.ascontiguousarray(g.link_tri, dtype=np.int64).view(np.uint64)
        for a in range(0, len(tri), step):
            j = np.arange(a, min(a + step, len(tri)), dtype=np.uint64)
            acc = (acc + int((sy_mix(np.uint64(0x5bd1e995) + j) ^ sy_mix(tri[a:a + step])).sum(dtype=np.uint64))) & M64
    return (acc + g.n_nodes * 0x9E3779B97F4A7C15) & M64


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3_100mbp30x_k31"
    w = bench.WORKLOADS[wl]
    k, min_freq, rl = w["k"], w["min_freq"], w["read_len"]
    W = O.words(k)
    mult = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    nreads = bench.nreads_for(w) * mult
    genome_len = w["genome_len"] * mult
    threads = bench.host_threads()
    t0 = time.time()
    bits = 4 if passes == 1 else 6
    nsl, sh = 1 << bits, np.uint64(2 * k - 64 * (W - 1) - bits)   # slices on the top bits of the most significant word
    chunk = 1_000_000
    kept, n_distinct, n_inst = [], 0, 0
    for ps in range(passes):
        q0, q1 = nsl * ps // passes, nsl * (ps + 1) // passes
        slices = [[] for _ in range(nsl)]
        for first in range(0, nreads, chunk):
            cnt = min(chunk, nreads - first)
            seq, offs = O.synth_reads(w["seed"], genome_len, cnt, rl, True, w["frag_len"], w["err_ppm"], first_read=first, threads=threads)
            km = O.extract(seq, offs, k)
            if ps == 0:
                n_inst += len(km)
            s = (km[:, 0] >> sh).astype(np.int64)
            if passes > 1:
                m = (s >= q0) & (s < q1)
                km, s = km[m], s[m]
            order = np.argsort(s, kind="stable")
            km, s = km[order], s[order]
            cuts = np.searchsorted(s, np.arange(nsl + 1))
            for q in range(q0, q1):
                if cuts[q + 1] > cuts[q]:
                    slices[q].append(km[cuts[q]:cuts[q + 1]].copy())
            if (first // chunk) % 8 == 0:
                print("pass %d: extracted reads %d / %d  (%.0f s)" % (ps, first + cnt, nreads, time.time() - t0), flush=True)
        for q in range(q0, q1):
            if not slices[q]:
                continue
            a = np.concatenate(slices[q]).reshape(-1, W)
            slices[q] = None
            keys, cnt = O.sort_count(a, k, threads=threads)
            del a
            n_distinct += len(keys)
            kept.append(O.filter_counts(keys, cnt, k, min_freq, u8=True))
            print("slice %d: %d distinct, %d kept  (%.0f s)" % (q, len(keys), len(kept[-1]), time.time() - t0), flush=True)
    kept = np.concatenate(kept)
    assert np.all(kept[1:, 0] >= kept[:-1, 0])
    print("counting done: %d instances, %d distinct, %d kept  (%.0f s); building the graph (sequential walk) ..." % (n_inst, n_distinct, len(kept), time.time() - t0), flush=True)
    g = O.dbg(kept, k)
    out = {"workload": wl, "genome_multiple": mult, "oracle": {"kmer_instances": int(n_inst), "distinct_kmers": int(n_distinct), "kept_kmers": int(len(kept)),
                                      "nodes": int(g.n_nodes), "bases": int(len(g.bases)), "link_entries": int(len(g.link_tri) // 3),
                                      "checksum": "%016x" % graph_checksum(g)}, "seconds": round(time.time() - t0, 1), "threads": threads}
    if len(sys.argv) > 2 and sys.argv[2] not in ("", "-"):
        d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        gg = d["graph"]
        out["gpu"] = {"file": os.path.relpath(sys.argv[2], ROOT), "n_gpus": d["n_gpus"], **gg}
        o = out["oracle"]
        out["equal"] = (gg["nodes"], gg["bases"], gg["link_entries"], gg["checksum"]) == (o["nodes"], o["bases"], o["link_entries"], o["checksum"])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
