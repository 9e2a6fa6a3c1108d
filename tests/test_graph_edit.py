"""Tip removal (SURVEY section 8 f, rank 3): remove_node!, collapse_all_unitigs!, remove_tips!.

CPU: the sequential restatement of the reference (oracle/graph_edit.py) against hand-checked cases and against the
executable model of the data-parallel formulation (tests/collapse_model.py) on fuzzed graphs.
GPU: the CUDA entry points (gg_graph_remove_nodes, gg_graph_collapse_all_unitigs, gg_graph_remove_tips) against the
sequential restatement through the C ABI, bit-exact: node sequences, deleted nodes, per-node link order.
"""
import random

import numpy as np
import pytest

import __graft_entry__ as G
import collapse_model as M
from oracle import graph_edit as E
from oracle import oracle as O


import json
import os

TIPS_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tips_golden.json")))["cases"]


def _tl(ls):
    return [[tuple(t) for t in row] for row in ls]


def _dbg_graph(seed, k=15, glen=3000, nreads=900, err=8000, min_freq=1, read_len=60):
    seq, offs = O.synth_reads(seed, glen, nreads, read_len, True, 150, err)
    keys, cnt = O.count_kmers(seq, offs, k)
    og = O.dbg(O.filter_counts(keys, cnt, k, min_freq), k)
    return og.nodes(), og.links()


def _random_graph(rnd, nn, nl, dup=0.1):
    """random sequences, random symmetric links made with add_link! (distances >= 0: no overlap to check), with
    duplicate links, links between the two ends of one node, self links and deleted nodes"""
    g = E.Graph(["".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 12))) for _ in range(nn)], [[] for _ in range(nn)])
    for _ in range(nl):
        a = rnd.randint(1, nn) * rnd.choice((1, -1))
        b = rnd.randint(1, nn) * rnd.choice((1, -1))
        g.add_link(a, b, rnd.randint(0, 5))
        if rnd.random() < dup:
            g.add_link(a, b, rnd.randint(0, 5))
    for _ in range(rnd.randint(0, 2)):
        g.remove_node(rnd.randint(1, nn))
    return g.nodes, g.links


def _chain_graph(rnd, k=5):
    """sequences cut into overlapping pieces (paths and circles, pieces stored in either orientation, node ids
    shuffled), linked by (k-1)-overlaps pushed in random order, plus a few branch links without overlap"""
    pieces, joins = [], []          # joins: (piece index a, piece index b): end of a -> start of b
    for _ in range(rnd.randint(1, 6)):
        L = rnd.randint(k + 2, 60)
        s = "".join(rnd.choice("ACGT") for _ in range(L))
        circle = rnd.random() < 0.3
        if circle:
            s = s + s[:k - 1]
        nk = len(s) - k + 1
        cuts = sorted(rnd.sample(range(1, nk), min(rnd.randint(1, 5), nk - 1)))
        first = len(pieces)
        a = 0
        for c in cuts + [nk]:
            pieces.append(s[a:c + k - 1])       # k-mers a .. c-1
            a = c
        joins += [(i, i + 1) for i in range(first, len(pieces) - 1)]
        if circle:
            joins.append((len(pieces) - 1, first))
    perm = list(range(len(pieces)))
    rnd.shuffle(perm)                # piece i becomes node perm[i] + 1
    sign = [rnd.choice((1, -1)) for _ in pieces]
    nodes = [None] * len(pieces)
    for i, p in enumerate(pieces):
        nodes[perm[i]] = p if sign[i] > 0 else E.revcomp(p)
    g = E.Graph(nodes, [[] for _ in nodes])
    ident = lambda i: (perm[i] + 1) * sign[i]
    rnd.shuffle(joins)
    for a, b in joins:
        if rnd.random() < 0.5:
            g.add_link(-ident(a), ident(b), -(k - 1))
        else:
            g.add_link(ident(b), -ident(a), -(k - 1))
    for _ in range(rnd.randint(0, 3)):
        g.add_link(rnd.randint(1, len(nodes)) * rnd.choice((1, -1)), rnd.randint(1, len(nodes)) * rnd.choice((1, -1)), 3)
    return g.nodes, g.links


def _seq_collapse(nodes, links, min_nodes=2):
    g = E.Graph(nodes, links)
    new = g.collapse_all_unitigs(min_nodes, True)
    return g.nodes, g.links, len(new)


def test_restatement_small_cases():
    # three nodes in a row, (k-1) = 2 overlaps: ACGT -> GTA -> TAC ; all collapse into one canonical node
    g = E.Graph(["ACGT", "GTA", "TAC"], [[], [], []])
    g.add_link(-1, 2, -2)
    g.add_link(-2, 3, -2)
    assert g.find_all_unitigs(2) == [[1, 2, 3]]
    new = g.collapse_all_unitigs(2, True)
    assert new == [4] and g.nodes == ["", "", "", "ACGTAC"] and g.links == [[], [], [], []]


def test_restatement_tip_rules():
    # src/Graphs.jl:778-802 on a fork: only a dead end next to a branching neighbour (or an isolated node) is a tip
    g = E.Graph(["AAAC", "AACGGG", "AACT", "CGGGT", "TTTTT"], [[] for _ in range(5)])
    g.add_link(-1, 2, -3)
    g.add_link(-1, 3, -3)
    g.add_link(-2, 4, -4)
    # node 3: one backward link, no forward link, and node 1 has two forward links -> tip; node 5 isolated -> tip;
    # node 4: its only neighbour (2) has one forward link -> not a tip; node 1: two forward links -> not a tip
    assert g.find_tip_nodes(5) == [3, 5]
    assert g.find_tip_nodes(4) == [3]        # node 5 is longer than min_size
    passes, removed = g.remove_tips(4)
    # after removing 3: 1 -> 2 -> 4 is one unitig AAACGGGT; canonical of AAACGGGT vs ACCCGTTT: AAAC.. < ACCC.. keep
    assert (passes, removed) == (2, 1)
    assert g.nodes == ["", "", "", "", "TTTTT", "AAACGGGT"] and all(not ls for ls in g.links)


def test_model_equals_restatement_fuzz_random_links():
    rnd = random.Random(5)
    hits = 0
    for trial in range(400):
        nodes, links = _random_graph(rnd, rnd.randint(1, 14), rnd.randint(0, 14))
        exp = _seq_collapse(nodes, links)
        got = M.collapse_all_unitigs(nodes, links, 2)
        assert got[0] == exp[0] and got[1] == exp[1] and got[2] == exp[2], (trial, nodes, links)
        hits += exp[2] > 0
    assert hits > 100


def test_model_equals_restatement_fuzz_chains():
    rnd = random.Random(9)
    circles = 0
    for trial in range(300):
        nodes, links = _chain_graph(rnd)
        for mn in (2, 3):
            exp = _seq_collapse(nodes, links, mn)
            got = M.collapse_all_unitigs(nodes, links, mn)
            assert got[0] == exp[0] and got[1] == exp[1] and got[2] == exp[2], (trial, mn, nodes, links)
        circles += any(len(ls) == 0 for ls in exp[1][len(nodes):])
    assert circles > 10


@pytest.mark.parametrize("seed,err,min_size", [(3, 8000, 45), (4, 20000, 30), (5, 3000, 10000), (6, 12000, 20)])
def test_model_equals_restatement_remove_tips_dbg(seed, err, min_size):
    nodes, links = _dbg_graph(seed, err=err)
    g = E.Graph(nodes, links)
    assert M.find_tip_nodes(nodes, links, min_size) == g.find_tip_nodes(min_size)
    passes, removed = g.remove_tips(min_size)
    n2, l2, p2, r2 = M.remove_tips(nodes, links, min_size)
    assert (p2, r2) == (passes, removed) and n2 == g.nodes and l2 == g.links
    if min_size >= 30:
        assert removed > 0 and len(g.nodes) > len(nodes)


def test_remove_nodes_model():
    rnd = random.Random(2)
    for _ in range(100):
        nodes, links = _random_graph(rnd, rnd.randint(2, 10), rnd.randint(0, 12))
        ids = [rnd.randint(1, len(nodes)) * rnd.choice((1, -1)) for _ in range(rnd.randint(0, 4))]
        g = E.Graph(nodes, links)
        for n in ids:
            g.remove_node(n)
        n2, l2 = M.remove_nodes(nodes, links, ids)
        assert n2 == g.nodes and l2 == g.links


@pytest.mark.parametrize("case", TIPS_GOLD, ids=lambda c: c["name"])
def test_tips_golden_restatement_and_model(case):
    """committed vectors (tests/golden/make_tips_golden.py; oracle/make_ref_golden.jl regenerates them with the real
    package) against the sequential restatement and the model of the device formulation"""
    nodes, links = case["nodes"], _tl(case["links"])
    g = E.Graph(nodes, links)
    assert g.find_tip_nodes(case["min_size"]) == case["tips"] == M.find_tip_nodes(nodes, links, case["min_size"])
    c = case["collapse"]
    new = g.collapse_all_unitigs(2, True)
    assert new == c["new_nodes"] and g.nodes == c["nodes"] and g.links == _tl(c["links"])
    assert M.collapse_all_unitigs(nodes, links, 2) == (c["nodes"], _tl(c["links"]), len(c["new_nodes"]))
    r = case["remove_tips"]
    g = E.Graph(nodes, links)
    assert g.remove_tips(case["min_size"]) == (r["passes"], r["removed"]) and g.nodes == r["nodes"] and g.links == _tl(r["links"])
    assert M.remove_tips(nodes, links, case["min_size"]) == (r["nodes"], _tl(r["links"]), r["passes"], r["removed"])


import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from hand_traced import TIPS_HAND_TRACED  # noqa: E402  (src/Graphs.jl + src/remove_tips.jl traced by hand)


@pytest.mark.parametrize("case", TIPS_HAND_TRACED, ids=lambda c: c["name"])
def test_tips_hand_traced_restatement_and_model(case):
    nodes, links = case["nodes"], case["links"]
    assert E.Graph(nodes, links).find_tip_nodes(case["min_size"]) == case["tips"]
    if "collapse" in case:
        c = case["collapse"]
        g = E.Graph(nodes, links)
        g.collapse_all_unitigs(2, True)
        assert (g.nodes, g.links) == (c["nodes"], c["links"])
        assert M.collapse_all_unitigs(nodes, links, 2)[:2] == (c["nodes"], c["links"])
    if "remove_tips" in case:
        r = case["remove_tips"]
        g = E.Graph(nodes, links)
        assert g.remove_tips(case["min_size"]) == (r["passes"], r["removed"]) and (g.nodes, g.links) == (r["nodes"], r["links"])
        assert M.remove_tips(nodes, links, case["min_size"]) == (r["nodes"], r["links"], r["passes"], r["removed"])


# ------------------------------------------------ CUDA path through the C ABI -------------------------------------------
@pytest.fixture(scope="module")
def pkg():
    return G.build()


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _device_graph(pkg, ctx, nodes, links, k=0):
    node_off = np.zeros(len(nodes) + 1, dtype=np.uint64)
    node_off[1:] = np.cumsum([len(x) for x in nodes], dtype=np.uint64)
    bases = np.frombuffer("".join(nodes).encode(), dtype=np.uint8).copy()
    link_row = np.zeros(len(nodes) + 1, dtype=np.uint64)
    link_row[1:] = np.cumsum([len(ls) for ls in links], dtype=np.uint64)
    tri = np.array([t for ls in links for t in ls], dtype=np.int64).reshape(-1)
    return pkg.DeviceGraph.from_arrays(pkg.GraphArrays(k, node_off, bases, link_row, tri), ctx=ctx)


def _back(dg):
    ga = dg.arrays()
    return ga.node_strings(), ga.link_lists()


@pytest.mark.gpu
def test_gpu_collapse_fuzz(pkg, ctx):
    """gg_graph_collapse_all_unitigs against join_path! one path after the other (random links, chains, circles)"""
    rnd = random.Random(31)
    hits = 0
    for trial in range(120):
        nodes, links = _chain_graph(rnd) if trial % 2 else _random_graph(rnd, rnd.randint(1, 16), rnd.randint(0, 18), dup=0.2)
        for mn in (2, 3):
            try:
                exp = _seq_collapse(nodes, links, mn)
            except ValueError:
                exp = None
            dg = _device_graph(pkg, ctx, nodes, links)
            try:
                if exp is None:
                    with pytest.raises(ValueError):
                        dg.collapse_all_unitigs(mn)
                    continue
                out, m = dg.collapse_all_unitigs(mn)
                try:
                    got = _back(out)
                    assert m == exp[2] and got[0] == exp[0] and got[1] == exp[1], (trial, mn, nodes, links)
                    hits += m > 0
                finally:
                    out.free()
            finally:
                dg.free()
    assert hits > 60


@pytest.mark.gpu
def test_gpu_remove_nodes(pkg, ctx):
    rnd = random.Random(32)
    for _ in range(40):
        nodes, links = _random_graph(rnd, rnd.randint(2, 12), rnd.randint(0, 14))
        ids = [rnd.randint(1, len(nodes)) * rnd.choice((1, -1)) for _ in range(rnd.randint(0, 4))]
        g = E.Graph(nodes, links)
        for n in ids:
            g.remove_node(n)
        dg = _device_graph(pkg, ctx, nodes, links)
        out = dg.remove_nodes(ids)
        assert _back(out) == (g.nodes, g.links)
        out.free()
        with pytest.raises(ValueError):
            dg.remove_nodes([len(nodes) + 1])
        dg.free()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,err,min_size,k", [(3, 8000, 45, 15), (4, 20000, 30, 15), (5, 3000, 10000, 15), (6, 12000, 20, 15), (7, 10000, 70, 31)])
def test_gpu_remove_tips_dbg(pkg, ctx, seed, err, min_size, k):
    """reads -> graph on the device -> remove_tips! on the device, against the sequential restatement on the oracle's graph"""
    seq, offs = O.synth_reads(seed, 3000, 900, 60, True, 150, err)
    keys, cnt = O.count_kmers(seq, offs, k)
    og = O.dbg(O.filter_counts(keys, cnt, k, 1), k)
    g = E.Graph(og.nodes(), og.links())
    passes, removed = g.remove_tips(min_size)
    ga = pkg.dbg_from_reads((seq, offs), k, 1, ctx=ctx)
    assert ga.node_strings() == og.nodes() and ga.link_lists() == og.links()
    dg = pkg.DeviceGraph.from_arrays(ga, ctx=ctx)
    out, p2, r2 = dg.remove_tips(min_size)
    got = _back(out)
    assert (p2, r2) == (passes, removed)
    assert got[0] == g.nodes and got[1] == g.links
    if min_size >= 30:
        assert removed > 0 and len(g.nodes) > og.n_nodes
    # the edited graph is a graph like any other: no tip of that size is left, the writer accepts it
    assert out.find_tip_nodes(min_size).size == 0
    out.free()
    dg.free()
    # the reference-named host entry mutates the Python graph
    sdg = pkg.empty_graph()
    for s in og.nodes():
        sdg.add_node_(s)
    for i, ls in enumerate(og.links()):
        sdg.links[i].extend(pkg.SDGLink(*t) for t in ls)
    assert pkg.remove_tips_(sdg, min_size, ctx=ctx) is sdg
    assert [n.seq for n in sdg.nodes] == g.nodes and sdg.link_tuples() == g.links
    assert all(n.deleted == (n.seq == "") for n in sdg.nodes)


@pytest.mark.gpu
def test_gpu_collapse_rejects_asymmetric_links(pkg, ctx):
    dg = _device_graph(pkg, ctx, ["ACGT", "GTAA"], [[(-1, 2, -2)], []])
    with pytest.raises(ValueError):
        dg.collapse_all_unitigs(2)
    dg.free()
    dg = _device_graph(pkg, ctx, ["ACGT", "TTAA"], [[(-1, 2, -2)], [(2, -1, -2)]])
    with pytest.raises(ValueError, match="do not overlap"):
        dg.collapse_all_unitigs(2)
    dg.free()


@pytest.mark.gpu
def test_gpu_remove_tips_large(pkg, ctx):
    """a graph beyond the single-launch sizes of the scan and sort primitives: 151,278 link entries, and 36,992 entries that
    touch a collapsed path in the first pass (> 32,768: the multi-pass radix sort of the 128-bit keys, not the rank sort)"""
    k = 21
    seq, offs = O.synth_reads(8, 150000, 30000, 100, True, 300, 10000)
    ga = pkg.dbg_from_reads((seq, offs), k, 1, ctx=ctx)
    nodes, links = ga.node_strings(), ga.link_lists()
    assert len(ga.link_triples) // 3 > 100000
    g = E.Graph(nodes, links)
    passes, removed = g.remove_tips(60)
    dg = pkg.DeviceGraph.from_arrays(ga, ctx=ctx)
    import time
    t0 = time.perf_counter()
    out, p2, r2 = dg.remove_tips(60)
    dt = time.perf_counter() - t0
    got = _back(out)
    print("remove_tips on the device: %d nodes, %d link entries -> %d nodes, %d passes, %d tips, %.1f ms (first call)"
          % (len(nodes), len(ga.link_triples) // 3, len(got[0]), p2, r2, 1e3 * dt))
    assert (p2, r2) == (passes, removed) and removed > 1000
    assert got[0] == g.nodes and got[1] == g.links
    c1, m = dg.collapse_all_unitigs(2)       # the few paths dbg! leaves open (its walk stops at used k-mers, src/dbg.jl:147-150)
    exp = _seq_collapse(nodes, links)
    assert m == exp[2] and _back(c1) == (exp[0], exp[1])
    c1.free(); out.free(); dg.free()


@pytest.mark.gpu
@pytest.mark.parametrize("case", TIPS_GOLD, ids=lambda c: c["name"])
def test_gpu_tips_golden(pkg, ctx, case):
    nodes, links = case["nodes"], _tl(case["links"])
    dg = _device_graph(pkg, ctx, nodes, links)
    assert dg.find_tip_nodes(case["min_size"]).tolist() == case["tips"]
    c = case["collapse"]
    out, m = dg.collapse_all_unitigs(2)
    assert m == len(c["new_nodes"]) and _back(out) == (c["nodes"], _tl(c["links"]))
    out.free()
    r = case["remove_tips"]
    out, passes, removed = dg.remove_tips(case["min_size"])
    assert (passes, removed) == (r["passes"], r["removed"]) and _back(out) == (r["nodes"], _tl(r["links"]))
    out.free()
    dg.free()


@pytest.mark.gpu
@pytest.mark.parametrize("case", TIPS_HAND_TRACED, ids=lambda c: c["name"])
def test_gpu_tips_hand_traced(pkg, ctx, case):
    dg = _device_graph(pkg, ctx, case["nodes"], case["links"])
    assert dg.find_tip_nodes(case["min_size"]).tolist() == case["tips"]
    if "collapse" in case:
        out, _ = dg.collapse_all_unitigs(2)
        assert _back(out) == (case["collapse"]["nodes"], case["collapse"]["links"])
        out.free()
    if "remove_tips" in case:
        r = case["remove_tips"]
        out, passes, removed = dg.remove_tips(case["min_size"])
        assert (passes, removed) == (r["passes"], r["removed"]) and _back(out) == (r["nodes"], r["links"])
        out.free()
    dg.free()


def test_gfa1_of_an_edited_graph_skips_the_emptied_nodes(tmp_path):
    """write_to_gfa1 (src/Graphs.jl:892-926) skips deleted nodes: the array writer of libggb200 (host code, no GPU) on a
    tip-removed graph against the Python transliteration of the reference writer"""
    pkg = G.build()
    nodes, links = _dbg_graph(3)
    g = E.Graph(nodes, links)
    g.remove_tips(45)
    assert "" in g.nodes
    sdg = pkg.empty_graph()
    for s in g.nodes:
        sdg.add_node_(pkg.SDGNode(s, s == ""))
    for i, ls in enumerate(g.links):
        sdg.links[i].extend(pkg.SDGLink(*t) for t in ls)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    from genomegraphs_b200.host import _sdg_arrays
    pkg.write_gfa1_arrays(a, _sdg_arrays(sdg, 15))
    sdg.write_to_gfa1(b)
    assert open(a + ".fasta").read() == open(b + ".fasta").read()
    ga, gb = open(a + ".gfa").read().replace(a + ".fasta", "F"), open(b + ".gfa").read().replace(b + ".fasta", "F")
    assert ga == gb and ga.count("\nS\t") == sum(1 for s in g.nodes if s)


# ------------------------------------------- the C restatement (oracle/graph_edit.c) ------------------------------------
def test_c_restatement_equals_python_restatement_fuzz():
    rnd = random.Random(41)
    errors = 0
    for trial in range(500):
        nodes, links = _chain_graph(rnd) if trial % 2 else _random_graph(rnd, rnd.randint(1, 16), rnd.randint(0, 18), dup=0.2)
        assert E.CGraph.from_lists(nodes, links).lists() == (list(nodes), [list(ls) for ls in links])
        assert E.CGraph.from_lists(nodes, links).find_tip_nodes(7) == E.Graph(nodes, links).find_tip_nodes(7)
        ids = [rnd.randint(1, len(nodes)) * rnd.choice((1, -1)) for _ in range(rnd.randint(0, 3))]
        g, c = E.Graph(nodes, links), E.CGraph.from_lists(nodes, links)
        for n in ids:
            g.remove_node(n)
        c.remove_nodes(ids)
        assert c.lists() == (g.nodes, g.links)
        for mn in (2, 3):
            g, c = E.Graph(nodes, links), E.CGraph.from_lists(nodes, links)
            try:
                new = g.collapse_all_unitigs(mn, True)
            except ValueError:
                errors += 1
                with pytest.raises(ValueError):
                    c.collapse_all_unitigs(mn)
                continue
            assert c.collapse_all_unitigs(mn) == len(new) and c.lists() == (g.nodes, g.links), (trial, mn)
        g, c = E.Graph(nodes, links), E.CGraph.from_lists(nodes, links)
        try:
            exp = g.remove_tips(8)
        except ValueError:
            continue
        assert c.remove_tips(8) == exp and c.lists() == (g.nodes, g.links), trial


@pytest.mark.parametrize("case", TIPS_GOLD, ids=lambda c: c["name"])
def test_c_restatement_golden(case):
    nodes, links = case["nodes"], _tl(case["links"])
    assert E.CGraph.from_lists(nodes, links).find_tip_nodes(case["min_size"]) == case["tips"]
    c = E.CGraph.from_lists(nodes, links)
    assert c.collapse_all_unitigs(2) == len(case["collapse"]["new_nodes"]) and c.lists() == (case["collapse"]["nodes"], _tl(case["collapse"]["links"]))
    c = E.CGraph.from_lists(nodes, links)
    r = case["remove_tips"]
    assert c.remove_tips(case["min_size"]) == (r["passes"], r["removed"]) and c.lists() == (r["nodes"], _tl(r["links"]))


def test_c_restatement_dbg_graph_arrays():
    """array interface (the layout of the C ABI) on a graph of a few thousand nodes"""
    seq, offs = O.synth_reads(8, 60000, 12000, 100, True, 300, 10000)
    keys, cnt = O.count_kmers(seq, offs, 21)
    og = O.dbg(O.filter_counts(keys, cnt, 21, 1), 21)
    c = E.CGraph(og.node_off, og.bases, og.link_row, og.link_tri)
    g = E.Graph(og.nodes(), og.links())
    assert c.remove_tips(60) == g.remove_tips(60)
    assert c.lists() == (g.nodes, g.links) and c.sizes()[0] == len(g.nodes)


def test_overlap_errors_agree_in_all_statements():
    """a base changed inside an overlap: "Sequences do not overlap." (src/Graphs.jl:1023) exactly when the pair sits in a
    collapsed path -- the Python restatement, the C restatement and the model raise together or agree on the result"""
    rnd = random.Random(77)
    raised = fine = 0
    for trial in range(300):
        nodes, links = _chain_graph(rnd)
        nodes = list(nodes)
        i = rnd.randrange(len(nodes))
        s = nodes[i]
        p = rnd.choice([rnd.randrange(0, min(4, len(s))), len(s) - 1 - rnd.randrange(0, min(4, len(s)))])
        nodes[i] = s[:p] + rnd.choice([c for c in "ACGT" if c != s[p]]) + s[p + 1:]
        try:
            exp = _seq_collapse(nodes, links)
        except ValueError:
            raised += 1
            with pytest.raises(ValueError):
                M.collapse_all_unitigs(nodes, links, 2)
            with pytest.raises(ValueError):
                E.CGraph.from_lists(nodes, links).collapse_all_unitigs(2)
            continue
        fine += 1
        assert M.collapse_all_unitigs(nodes, links, 2) == exp
        c = E.CGraph.from_lists(nodes, links)
        assert c.collapse_all_unitigs(2) == exp[2] and c.lists() == (exp[0], exp[1])
    assert raised > 50 and fine > 50
