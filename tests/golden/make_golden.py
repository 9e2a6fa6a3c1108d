#!/usr/bin/env python
"""Generates tests/golden/dbg_golden.json.

The reference (pure Julia) cannot run in this project, so these vectors are NOT reference
outputs: they come from tests/pyref.py, the string-based transliteration of src/dbg.jl /
src/GraphIndexes.jl that shares no code or data representation with the C oracle or the CUDA
path, plus the hand-traced vector G1.  They freeze the expected results
so that the oracle and the CUDA path are both checked against committed data.

    python tests/golden/make_golden.py        # rewrites the fixture (deterministic)
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import pyref  # noqa: E402
from hand_traced import HAND_TRACED  # noqa: E402


def graph_case(name, kmers):
    kmers = sorted(set(kmers))
    nodes, links = pyref.dbg(kmers)
    return {"name": name, "k": len(kmers[0]), "kmers": kmers, "nodes": nodes, "links": [[list(t) for t in ls] for ls in links]}


def reads_case(name, reads, k, min_freq):
    counts = pyref.count(pyref.extract(reads, k))
    kept = [m for m, c in counts if c >= min_freq]
    nodes, links = pyref.dbg(kept)
    uniq, upn, total = pyref.unique_mer_index(nodes, k)
    return {"name": name, "k": k, "min_freq": min_freq, "reads": reads, "counts": [[m, c] for m, c in counts], "nodes": nodes,
            "links": [[list(t) for t in ls] for ls in links],
            "umi": {"unique": [[m, nid, pos] for m, nid, pos in uniq], "unique_per_node": upn, "total_per_node": total}}


def mutate(rnd, s, rate):
    return "".join(rnd.choice([b for b in "ACGT" if b != c]) if rnd.random() < rate else c for c in s)


def main():
    rnd = random.Random(20261019)
    graphs = [graph_case("G1 (docs/src/DeBruijnGraph.md:57-76, canonicalised)", ["AATC", "AATG", "ACGA", "ATTC", "CGAA", "CGTA", "CGTC"])]
    for h in HAND_TRACED[1:]:   # G2..G4: expected values from the hand traces (tests/golden/hand_traced.py), not from pyref
        graphs.append({"name": h["name"] + " (hand-traced through src/dbg.jl)", "k": h["k"], "kmers": h["kmers"], "nodes": h["nodes"],
                       "links": [[list(t) for t in ls] for ls in h["links"]]})
    graphs.append(graph_case("single k-mer self loop", ["AAA"]))
    graphs.append(graph_case("palindromes k=4", ["ACGT", "AATT", "CGCG", "AACG", "ACGA"]))
    graphs.append(graph_case("hairpin k=3", ["AAC", "ACG", "CGT"[::-1].translate(str.maketrans("ACGT", "TGCA")), "ACC"]))
    circ = "ACGGTCATTGCA"
    graphs.append(graph_case("perfect circle k=5", [pyref.canonical((circ + circ)[i:i + 5]) for i in range(len(circ))]))
    for k in (2, 3, 4, 5, 6, 7, 8, 9):
        for dens in (0.05, 0.3, 0.7):
            n = max(2, min(150, int(4 ** k * dens)))
            ks = {pyref.canonical("".join(rnd.choice("ACGT") for _ in range(k))) for _ in range(n)}
            graphs.append(graph_case("random k=%d density %.2f" % (k, dens), list(ks)))
    reads_cases = []
    for k, mf, glen, nr, rl, err in ((5, 1, 60, 30, 20, 0.0), (11, 2, 300, 120, 40, 0.01), (31, 2, 300, 100, 70, 0.01), (33, 2, 300, 100, 80, 0.01),
                                     (63, 1, 300, 40, 100, 0.0), (65, 2, 300, 100, 110, 0.005), (127, 1, 400, 24, 150, 0.0)):
        genome = "".join(rnd.choice("ACGT") for _ in range(glen))
        reads = []
        for _ in range(nr):
            p = rnd.randrange(0, glen - rl + 1)
            r = mutate(rnd, genome[p:p + rl], err)
            if rnd.random() < 0.5:
                r = pyref.rc(r)
            if rnd.random() < 0.1:
                q = rnd.randrange(len(r))
                r = r[:q] + "N" + r[q + 1:]
            if rnd.random() < 0.1:
                r = r.lower()
            reads.append(r)
        reads.append("ACGT"[: min(4, k - 1)])  # shorter than k: yields nothing
        reads.append("")
        reads_cases.append(reads_case("reads k=%d min_freq=%d" % (k, mf), reads, k, mf))
    out = {"generator": "tests/golden/make_golden.py (tests/pyref.py)", "graphs": graphs, "reads": reads_cases}
    with open(os.path.join(HERE, "dbg_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)
        f.write("\n")
    print("wrote %d graph cases, %d read cases" % (len(graphs), len(reads_cases)))


if __name__ == "__main__":
    main()
