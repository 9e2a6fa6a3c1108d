"""Committed golden vectors (tests/golden/dbg_golden.json, made by tests/golden/make_golden.py from
the independent Python transliteration) against the C oracle (CPU) and the CUDA path (gpu)."""
import json
import os
import sys

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "dbg_golden.json")))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from hand_traced import HAND_TRACED  # noqa: E402  (G1..G4: src/dbg.jl traced by hand, independent of pyref and of the oracle)


def _links(ls):
    return [[tuple(t) for t in row] for row in ls]


@pytest.mark.parametrize("case", GOLD["graphs"], ids=lambda c: c["name"])
def test_oracle_graph_golden(case):
    g = O.dbg(O.kmers_from_strings(case["kmers"]), case["k"])
    assert g.nodes() == case["nodes"]
    assert g.links() == _links(case["links"])


@pytest.mark.parametrize("case", HAND_TRACED, ids=lambda c: c["name"])
def test_oracle_hand_traced(case):
    g = O.dbg(O.kmers_from_strings(case["kmers"]), case["k"])
    assert g.nodes() == case["nodes"]
    assert g.links() == case["links"]


@pytest.mark.parametrize("case", GOLD["reads"], ids=lambda c: c["name"])
def test_oracle_reads_golden(case):
    k = case["k"]
    seq = np.frombuffer("".join(case["reads"]).encode(), dtype=np.uint8)
    offs = np.cumsum([0] + [len(r) for r in case["reads"]]).astype(np.uint64)
    keys, cnt = O.count_kmers(seq, offs, k)
    assert [O.words_to_kmer(w, k) for w in keys] == [m for m, _ in case["counts"]]
    assert cnt.tolist() == [c for _, c in case["counts"]]
    g = O.dbg(O.filter_counts(keys, cnt, k, case["min_freq"]), k)
    assert g.nodes() == case["nodes"] and g.links() == _links(case["links"])


@pytest.fixture(scope="module")
def gpu():
    pkg = G.build()
    ctx = pkg.Context(0)
    yield pkg, ctx
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["graphs"], ids=lambda c: c["name"])
def test_gpu_graph_golden(gpu, case):
    pkg, ctx = gpu
    ga = pkg.dbg_arrays(O.kmers_from_strings(case["kmers"]), k=case["k"], ctx=ctx)
    assert ga.node_strings() == case["nodes"]
    assert ga.link_lists() == _links(case["links"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", HAND_TRACED, ids=lambda c: c["name"])
def test_gpu_hand_traced(gpu, case):
    pkg, ctx = gpu
    ga = pkg.dbg_arrays(O.kmers_from_strings(case["kmers"]), k=case["k"], ctx=ctx)
    assert ga.node_strings() == case["nodes"]
    assert ga.link_lists() == case["links"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["reads"], ids=lambda c: c["name"])
def test_gpu_reads_golden(gpu, case):
    pkg, ctx = gpu
    k = case["k"]
    counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)(case["reads"])
    keys, c32, c8 = counts.export()
    assert [O.words_to_kmer(w, k) for w in keys] == [m for m, _ in case["counts"]]
    assert c32.tolist() == [c for _, c in case["counts"]]
    ga = pkg.dbg_arrays(counts, min_freq=case["min_freq"], ctx=ctx)
    assert ga.node_strings() == case["nodes"] and ga.link_lists() == _links(case["links"])
    ga2 = pkg.dbg_from_reads(case["reads"], k, case["min_freq"], ctx=ctx)   # fused reads -> graph pipeline
    assert ga2.node_strings() == case["nodes"] and ga2.link_lists() == _links(case["links"])
    umi = pkg.UniqueMerIndex(case["nodes"], k, ctx=ctx)
    exp = case["umi"]
    assert [O.words_to_kmer(w, k) for w in umi.keys] == [m for m, _, _ in exp["unique"]]
    assert umi.node.tolist() == [n for _, n, _ in exp["unique"]] and umi.pos.tolist() == [p for _, _, p in exp["unique"]]
    assert umi.unique_mers_per_node.tolist() == exp["unique_per_node"] and umi.total_mers_per_node.tolist() == exp["total_per_node"]
    counts.free()
