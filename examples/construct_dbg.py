#!/usr/bin/env python
"""De Bruijn graph construction with the B200 path: the recipe of the reference's
examples/construct_dbg.jl / docs/src/man/assembly.md:126-216 through the Python host mirror.

    python examples/construct_dbg.py reads.fastq [more.fastq ...] [-k 31] [--min-freq 2] [--remove-tips 200] [--gfa out]

Without input files a synthetic stand-in for the reference's (absent) ecoli-ref.fasta /
pe-reads.fastq pair is generated: 4,641,652 bp uniform genome, 50x 150 bp paired reads.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reads", nargs="*")
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("--min-freq", type=int, default=2)
    ap.add_argument("--remove-tips", type=int, default=None, metavar="MIN_SIZE",
                    help="remove_tips!(ws, MIN_SIZE) after the build (docs/src/man/assembly.md:234-275)")
    ap.add_argument("--gfa", default=None, help="write <gfa>.gfa + <gfa>.fasta (write_to_gfa1)")
    args = ap.parse_args()
    gg = G.build()
    if args.reads:
        parts = [gg.read_sequences(p) for p in args.reads]
        seq = np.concatenate([p[0] for p in parts])
        offs, base = [np.zeros(1, dtype=np.uint64)], 0
        for s, o in parts:
            offs.append(o[1:] + np.uint64(base))
            base += len(s)
        reads = (seq, np.concatenate(offs))
    else:
        glen = 4641652
        reads = gg.synth_reads(1, glen, 2 * int(round(50 * glen / 300.0)), 150, paired=True, frag_len=700, err_ppm=1000)
    t0 = time.perf_counter()
    counts = gg.serial_mem(args.k, gg.CANONICAL)(reads)            # sm = serial_mem(DNAMer{31}, CANONICAL); sm(ds)
    print("%d k-mer instances, %d distinct" % (counts.n_instances, len(counts)))
    spectrum = gg.spectra(counts)                                  # spectra(counts)
    print("k-mer spectrum, counts 1..10:", spectrum[1:11].tolist())
    graph = gg.dbg_(gg.empty_graph(), counts, args.min_freq)       # dbg!(graph, counts, min_freq)
    print("graph: %d nodes, %d link entries, %.2f s" % (graph.n_nodes(), sum(len(l) for l in graph.links), time.perf_counter() - t0))
    if args.remove_tips is not None:
        t1 = time.perf_counter()
        gg.remove_tips_(graph, args.remove_tips)                   # remove_tips!(ws, 200)
        live = sum(1 for n in graph.nodes if not n.deleted)
        print("after remove_tips!(%d): %d nodes (%d live), %d link entries, %.2f s"
              % (args.remove_tips, graph.n_nodes(), live, sum(len(l) for l in graph.links), time.perf_counter() - t1))
    if args.gfa:
        graph.write_to_gfa1(args.gfa)


if __name__ == "__main__":
    main()
