"""SURVEY section 8(f) rows next to the path: GFA1 + FASTA persistence from the exported arrays (and a loader), spectra-cn,
find_tip_nodes on the device-resident graph, 4-bit packed read input; handles that outlive their context."""
import os

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import oracle as O


@pytest.fixture(scope="module")
def pkg():
    return G.build()


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _oracle_graph(pkg, seed=3, k=15, glen=3000, nreads=900, err=8000, min_freq=1):
    seq, offs = pkg.synth_reads(seed, glen, nreads, 60, True, 150, err)
    keys, cnt = O.count_kmers(seq, offs, k)
    return seq, offs, keys, cnt, O.dbg(O.filter_counts(keys, cnt, k, min_freq), k)


def _sdg_from(pkg, og):
    sdg = pkg.empty_graph()
    for s in og.nodes():
        sdg.add_node_(s)
    for i, ls in enumerate(og.links()):
        sdg.links[i].extend(pkg.SDGLink(*t) for t in ls)
    return sdg


def test_gfa1_writer_from_arrays_matches_object_writer(pkg, tmp_path):
    """gg_gfa1_write (host code, arrays) against the Python transliteration of write_to_gfa1 (src/Graphs.jl:892-926)"""
    _, _, _, _, og = _oracle_graph(pkg, err=600)
    ga = pkg.GraphArrays(15, og.node_off.astype(np.uint64), og.bases, og.link_row.astype(np.uint64), og.link_tri)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    pkg.write_gfa1_arrays(a, ga)
    _sdg_from(pkg, og).write_to_gfa1(b)
    assert open(a + ".fasta").read() == open(b + ".fasta").read()
    assert open(a + ".gfa").read().replace(a + ".fasta", "F") == open(b + ".gfa").read().replace(b + ".fasta", "F")
    assert any(len(line) == 71 for line in open(a + ".fasta")) and og.n_nodes > 3   # wrapped at 70 bases


def _find_tips_ref(nodes, links, min_size):
    """find_tip_nodes! (src/Graphs.jl:778-802) over node strings + per-node (source, destination, dist) lists"""
    def fwd(n):
        return [l for l in links[abs(n) - 1] if l[0] == -n]
    out = set()
    for n in range(1, len(nodes) + 1):
        if len(nodes[n - 1]) == 0 or len(nodes[n - 1]) > min_size:
            continue
        fwl, bwl = fwd(n), fwd(-n)
        if len(fwl) == 1 and len(bwl) == 0 and len(fwd(-fwl[0][1])) > 1:
            out.add(n)
        if len(fwl) == 0 and len(bwl) == 1 and len(fwd(-bwl[0][1])) > 1:
            out.add(n)
        if not fwl and not bwl:
            out.add(n)
    return sorted(out)


@pytest.mark.gpu
@pytest.mark.parametrize("min_size", [20, 45, 10000])
def test_find_tip_nodes_and_gfa_roundtrip(pkg, ctx, tmp_path, min_size):
    seq, offs, keys, cnt, og = _oracle_graph(pkg)
    ga = pkg.dbg_arrays(O.filter_counts(keys, cnt, 15, 1), k=15, ctx=ctx)
    assert ga.node_strings() == og.nodes()
    dg = pkg.DeviceGraph.from_arrays(ga, ctx=ctx)
    try:
        tips = dg.find_tip_nodes(min_size).tolist()
        assert tips == _find_tips_ref(og.nodes(), og.links(), min_size)
        if min_size == 45:
            assert 0 < len(tips) < og.n_nodes
        f = str(tmp_path / "g")
        dg.write_to_gfa1(f)
        back = pkg.DeviceGraph.load_from_gfa1(f + ".gfa", f + ".fasta", ctx=ctx)
        try:
            gb = back.arrays()
            assert gb.node_strings() == og.nodes() and gb.k == 15
            # one add_link! per L line: the same links at every node (the order inside a node follows the file)
            assert [sorted(x) for x in gb.link_lists()] == [sorted(x) for x in og.links()]
        finally:
            back.free()
    finally:
        dg.free()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [15, 31, 40])
def test_spectra_cn(pkg, ctx, k):
    """spectra(reads_kcount, graph_kcount) (docs/src/man/assembly.md:221-228) against a dictionary join"""
    seq, offs = pkg.synth_reads(70 + k, 4000, 2500, 80, True, 200, 6000)
    reads = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
    ga = pkg.dbg_from_reads((seq, offs), k, 3, ctx=ctx)
    graph = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)(ga.node_strings())
    got = pkg.spectra_cn(reads, graph)
    rk, rc, _ = reads.export()
    gk, gc, _ = graph.export()
    a = {tuple(r): min(int(c), 255) for r, c in zip(rk.tolist(), rc.tolist())}
    b = {tuple(r): min(int(c), 255) for r, c in zip(gk.tolist(), gc.tolist())}
    exp = np.zeros((256, 256), dtype=np.uint64)
    for key in set(a) | set(b):
        exp[a.get(key, 0), b.get(key, 0)] += 1
    assert np.array_equal(got, exp)
    assert got[:, 1:].sum() == len(b) and got[3:, 0].sum() == 0   # every k-mer kept at min-count 3 is in the graph, once per unitig set
    reads.free(); graph.free()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [21, 31, 63])
def test_4bit_input_equals_ascii_input(pkg, ctx, k):
    seq, offs = pkg.synth_reads(90 + k, 20000, 3000, 101, True, 300, 4000, n_ppm=500)
    data4, boffs = pkg.pack_4bit((seq, offs))
    assert data4.nbytes * 2 <= seq.nbytes + 16
    c4 = pkg.count_kmers_4bit(data4, boffs, k, ctx=ctx)
    k4, n4, _ = c4.export()
    okeys, ocnt = O.count_kmers(seq, offs, k)
    assert np.array_equal(k4, okeys) and np.array_equal(n4, ocnt)
    g4 = pkg.dbg_from_reads_4bit(data4, boffs, k, 2, ctx=ctx)
    og = O.dbg(O.filter_counts(okeys, ocnt, k, 2), k)
    assert g4.node_strings() == og.nodes() and g4.link_lists() == og.links()
    c4.free()


@pytest.mark.gpu
def test_handles_outlive_their_context(pkg):
    """gg_ctx_destroy with live handles only marks the context; the last handle's release tears it down
    (garbage-collected hosts free objects in no particular order)"""
    c = pkg.Context(0)
    seq, offs = pkg.synth_reads(5, 5000, 500, 80, False, 0, 0)
    counts = pkg.serial_mem(21, pkg.CANONICAL, ctx=c)((seq, offs))
    L = pkg._native.lib()
    h_ctx, h_counts = c.h, counts.h
    L.gg_ctx_destroy(h_ctx)          # handle still alive
    c.h = None
    # the table is still readable
    nu = pkg._native.C.c_uint64()
    assert L.gg_counts_sizes(h_counts, pkg._native.C.byref(nu), None, None, None) == 0 and nu.value > 0
    L.gg_counts_free(h_counts)       # last handle: the context goes with it
    counts.h = None
