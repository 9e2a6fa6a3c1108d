"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs.  Bit-exact for every k-mer word, count, node sequence, node id and link entry
(including per-node link order).  All tests need a B200 (`-m gpu`)."""
import ctypes as C
import random

import numpy as np
import pytest

import __graft_entry__ as G
import pyref
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return G.build()


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _random_reads(rnd, n, lo, hi, alphabet="ACGT"):
    return ["".join(rnd.choice(alphabet) for _ in range(rnd.randint(lo, hi))) for _ in range(n)]


def _graph_equal(ga, og, tag=""):
    assert ga.node_strings() == og.nodes(), tag
    assert ga.link_lists() == og.links(), tag


# ------------------------------ K1/K2: pack + extract -------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 95, 96, 97, 127, 128])
def test_extract_read_order(pkg, ctx, k):
    rnd = random.Random(k)
    alphabet = "ACGTacgtNRY-" if k < 40 else "ACGT" * 30 + "Nn"
    reads = _random_reads(rnd, 300, 0, 3 * k + 40, alphabet)
    reads[5] = ""
    reads[17] = "A" * (k - 1) if k > 1 else ""
    reads[18] = "ACGT" * 100
    seq, offs = pkg.pack_reads(reads)
    for mode in (pkg.CANONICAL, pkg.NONCANONICAL):
        got = pkg.extract_kmers((seq, offs), k, mode, ctx=ctx)
        want = O.extract(seq, offs, k, bool(mode))
        assert got.shape == want.shape
        assert np.array_equal(got, want)


def test_extract_edge_cases(pkg, ctx):
    assert pkg.extract_kmers([], 31, ctx=ctx).shape == (0, 1)
    assert pkg.extract_kmers(["", "", ""], 31, ctx=ctx).shape == (0, 1)
    assert pkg.extract_kmers(["ACGT"], 31, ctx=ctx).shape == (0, 1)
    assert pkg.extract_kmers(["N" * 100], 5, ctx=ctx).shape == (0, 1)
    # read boundaries must break windows: two reads of k-1 bases give nothing
    assert pkg.extract_kmers(["ACGTA", "CGTAC"], 6, ctx=ctx).shape == (0, 1)
    got = pkg.extract_kmers(["ACGTAC"], 6, ctx=ctx)
    assert got.shape == (1, 1) and O.words_to_kmer(got[0], 6) == pyref.canonical("ACGTAC")
    with pytest.raises(ValueError):
        pkg.extract_kmers(["ACGT"], 129, ctx=ctx)
    with pytest.raises(ValueError):
        pkg.serial_mem(0)


def test_extract_unaligned_lengths(pkg, ctx):
    rnd = random.Random(99)
    for total in (1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 511, 512, 513, 4097):
        reads = ["".join(rnd.choice("ACGT") for _ in range(total))]
        seq, offs = pkg.pack_reads(reads)
        for k in (1, 4, 31):
            assert np.array_equal(pkg.extract_kmers((seq, offs), k, ctx=ctx), O.extract(seq, offs, k, True))


# ------------------------------ K3/K4: count + filter --------------------------------------
@pytest.mark.parametrize("k", [1, 3, 4, 11, 16, 31, 32, 33, 63, 64, 65, 96, 127, 128])
def test_count_filter_vs_oracle(pkg, ctx, k):
    rnd = random.Random(1000 + k)
    base = _random_reads(rnd, 60, max(k, 20), 3 * k + 60, "ACGT" * 40 + "N")
    reads = [rnd.choice(base)[rnd.randint(0, 5):] for _ in range(2500)]  # heavy duplication -> counts > 1
    seq, offs = pkg.pack_reads(reads)
    counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
    keys, c32, c8 = counts.export()
    okeys, ocounts = O.count_kmers(seq, offs, k)
    assert counts.n_instances == len(O.extract(seq, offs, k))
    assert np.array_equal(keys, okeys)
    assert np.array_equal(c32, ocounts)
    assert np.array_equal(c8, np.minimum(ocounts, 255).astype(np.uint8))
    spec = counts.spectrum()
    assert np.array_equal(spec, np.bincount(np.minimum(ocounts, 255), minlength=256).astype(np.uint64))
    for mf in (2, 7, 300):
        c2 = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs)).filter_(mf)
        assert np.array_equal(c2.export()[0], O.filter_counts(okeys, ocounts, k, mf))
    c3 = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs)).filter_(256, u8_saturated=True)
    assert len(c3) == 0
    nc = pkg.serial_mem(k, pkg.NONCANONICAL, ctx=ctx)((seq, offs)).export()
    onk, onc = O.count_kmers(seq, offs, k, canonical=False)
    assert np.array_equal(nc[0], onk) and np.array_equal(nc[1], onc)


def test_count_empty(pkg, ctx):
    c = pkg.serial_mem(31, ctx=ctx)([])
    assert len(c) == 0 and c.export()[0].shape == (0, 1)
    c = pkg.serial_mem(31, ctx=ctx)(["ACGT", "NNNN"])
    assert len(c) == 0
    assert len(c.filter_(2)) == 0


# ------------------------------ K5-K8: unitigs + links --------------------------------------
def test_g1_golden(pkg, ctx):
    kmers = ["AATC", "AATG", "ACGA", "ATTC", "CGAA", "CGTA", "CGTC"]
    ga = pkg.dbg_arrays(O.kmers_from_strings(kmers), k=4, ctx=ctx)
    assert ga.node_strings() == ["AATC", "AATG", "ACGAAT", "CGTA", "CGTC"]
    assert ga.link_lists() == [[(1, -3, -3)], [(2, -3, -3)], [(-3, 1, -3), (-3, 2, -3), (3, 4, -3), (3, 5, -3)],
                               [(4, 3, -3)], [(5, 3, -3)]]
    sdg = pkg.dbg_(pkg.empty_graph(), O.kmers_from_strings(list(reversed(kmers))), k=4, ctx=ctx)  # unsorted input
    assert [n.seq for n in sdg.nodes] == ga.node_strings()
    assert sdg.link_tuples() == ga.link_lists()
    assert sdg.sequence(-3) == pyref.rc("ACGAAT")


def test_dbg_trivial(pkg, ctx):
    ga = pkg.dbg_arrays(np.zeros((0, 1), dtype=np.uint64), k=5, ctx=ctx)
    assert ga.n_nodes == 0
    ga = pkg.dbg_arrays(O.kmers_from_strings(["AAA"]), k=3, ctx=ctx)   # self-loop, single node: no links (src/dbg.jl:267)
    assert ga.node_strings() == ["AAA"] and ga.link_lists() == [[]]
    with pytest.raises(ValueError):
        pkg.dbg_(pkg.empty_graph(), [[1, 2, 3]])


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8, 9])
def test_dbg_fuzz_small_k(pkg, ctx, k):
    """hairpins, palindromes, circles, Moebius circles and self-loops only occur at small k."""
    rnd = random.Random(7000 + k)
    space = 4 ** k
    for trial in range(80):
        dens = rnd.choice([0.02, 0.1, 0.3, 0.6, 0.9])
        n = max(1, min(400, int(space * dens)))
        ks = sorted({pyref.canonical("".join(rnd.choice("ACGT") for _ in range(k))) for _ in range(n)})
        km = O.kmers_from_strings(ks)
        _graph_equal(pkg.dbg_arrays(km, k=k, ctx=ctx), O.dbg(km, k), (k, trial))


@pytest.mark.parametrize("k", [5, 9, 15, 16, 31, 32, 33, 63, 64, 65, 127, 128])
def test_dbg_genome_like(pkg, ctx, k):
    rnd = random.Random(9000 + k)
    base = "".join(rnd.choice("ACGT") for _ in range(2500))
    rep = base[100:100 + 2 * k]
    genome = base[:1300] + rep + base[1300:] + pyref.rc(base[-(2 * k):])   # repeat + hairpin
    circ = "".join(rnd.choice("ACGT") for _ in range(3 * k + 17))
    circ_lin = circ + circ[:k - 1]                                           # perfect circle
    circ2 = "".join(rnd.choice("ACGT") for _ in range(64))
    pieces = [genome, circ_lin, circ2 + circ2[:k - 1] if k <= 60 else circ2]
    seq, offs = pkg.pack_reads(pieces)
    km = np.unique(O.extract(seq, offs, k), axis=0) if O.words(k) == 1 else O.sort_count(O.extract(seq, offs, k), k)[0]
    og = O.dbg(km, k)
    ga = pkg.dbg_arrays(km, k=k, ctx=ctx)
    _graph_equal(ga, og, k)
    assert og.n_nodes >= 3


@pytest.mark.parametrize("k,min_freq", [(31, 2), (63, 2), (127, 2), (21, 1), (32, 3)])
def test_reads_to_graph_pipeline(pkg, ctx, k, min_freq):
    """config-2 shaped (scaled down): 50x paired 150 bp reads with errors -> count -> filter -> graph."""
    seq, offs = pkg.synth_reads(seed=3 + k, genome_len=60000, nreads=20000, read_len=150, paired=True, frag_len=500,
                                err_ppm=5000, n_ppm=100)
    ga = pkg.dbg_from_reads((seq, offs), k, min_freq, ctx=ctx)
    okeys, ocounts = O.count_kmers(seq, offs, k, threads=4)
    og = O.dbg(O.filter_counts(okeys, ocounts, k, min_freq), k)
    _graph_equal(ga, og, (k, min_freq))
    ms, launches = ctx.last_timings()
    assert launches > 0 and ms["total"] > 0


def test_perfect_circles_only(pkg, ctx):
    rnd = random.Random(5)
    for k in (4, 7, 31, 33):
        pieces = []
        for _ in range(6):
            c = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(k + 3, 5 * k)))
            pieces.append(c + c[:k - 1])
        seq, offs = pkg.pack_reads(pieces)
        km = O.sort_count(O.extract(seq, offs, k), k)[0]
        _graph_equal(pkg.dbg_arrays(km, k=k, ctx=ctx), O.dbg(km, k), k)


# ------------------------------ UniqueMerIndex ------------------------------------------------
@pytest.mark.parametrize("k", [3, 4, 7, 31, 33, 64, 127])
def test_unique_mer_index(pkg, ctx, k):
    rnd = random.Random(500 + k)
    nodes = _random_reads(rnd, 40, 0, 4 * k + 9, "ACGT" if k > 4 else "AC")
    nodes.append(nodes[0])
    nodes.append("")
    idx = pkg.UniqueMerIndex(nodes, k, ctx=ctx)
    keys, node, pos, upn, tpn = O.unique_mer_index(nodes, k)
    assert np.array_equal(idx.keys, keys)
    assert np.array_equal(idx.node, node)
    assert np.array_equal(idx.pos, pos)
    assert np.array_equal(idx.unique_mers_per_node, upn)
    assert np.array_equal(idx.total_mers_per_node, tpn)
    if len(keys):
        ok, p = idx.find_unique_mer(keys[0])
        assert ok and p.node == node[0] and p.position == pos[0]
    assert idx.find_unique_mer(np.full(O.words(k), 2**64 - 1, dtype=np.uint64))[0] is False
    assert idx.isunmappable(len(nodes)) and idx.isunmappable(-len(nodes))


# ------------------------------ synthetic generator -------------------------------------------
def test_synth_device_matches_host(pkg, ctx):
    import ctypes as C
    N = pkg._native
    nreads, L = 5000, 150
    host, _ = pkg.synth_reads(42, 100000, nreads, L, paired=True, frag_len=600, err_ppm=10000, n_ppm=300)
    d = C.c_void_p()
    N.check(ctx.h, N.lib().gg_dev_alloc(ctx.h, nreads * L, C.byref(d)))
    N.check(ctx.h, N.lib().gg_synth_reads_dev(ctx.h, 42, 100000, nreads, L, 1, 600, 10000, 300, d))
    back = np.zeros(nreads * L, dtype=np.uint8)
    N.check(ctx.h, N.lib().gg_memcpy_d2h(ctx.h, back.ctypes.data_as(C.c_void_p), d, nreads * L))
    N.lib().gg_dev_free(ctx.h, d)
    assert np.array_equal(host, back)


# ------------------------------ overflow paths of the counting leaves ------------------------
@pytest.mark.parametrize("k", [15, 31, 32, 33, 63, 65, 127])
@pytest.mark.parametrize("mod", [1, 3])
def test_count_overflow_paths(pkg, ctx, k, mod, monkeypatch):
    """GG_MSD_TEST_OVERFLOW=m sends every m-th child through the overflow path (k <= 32: in-CTA
    re-partition; k > 32: host-driven radix sort + run-length count of the child's segment)."""
    monkeypatch.setenv("GG_MSD_TEST_OVERFLOW", str(mod))
    monkeypatch.setenv("GG_MSD_DT", "64")   # many small children even on a small input
    monkeypatch.setenv("GG_FORCE_MSD", "1")  # 17 <= k <= 32 would take the super-k-mer route otherwise
    seq, offs = pkg.synth_reads(seed=300 + k, genome_len=6000, nreads=1600, read_len=150, paired=True, frag_len=400, err_ppm=5000)
    for min_freq in (1, 2):
        okeys, ocnt = O.count_kmers(seq, offs, k)
        counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
        keys, c32, _ = counts.export()
        assert np.array_equal(keys, okeys) and np.array_equal(c32, ocnt), (k, mod)
        ga = pkg.dbg_from_reads((seq, offs), k, min_freq, ctx=ctx)   # fused min-count filter inside the leaves
        og = O.dbg(O.filter_counts(okeys, ocnt, k, min_freq), k)
        assert ga.node_strings() == og.nodes() and ga.link_lists() == og.links(), (k, mod, min_freq)
        counts.free()


# ------------------------------ super-k-mer route (17 <= k <= 32) -----------------------------
@pytest.mark.parametrize("k", [17, 18, 20, 24, 25, 27, 31, 32])
def test_skm_count_and_graph(pkg, ctx, k):
    """minimizer-partitioned counting: every k of both window geometries (WIN = 9 for k <= 24, 17 above), reads with N,
    lower case, reads shorter than k, heavy duplication; counts, saturated view, fused min-count filter, graph."""
    rnd = random.Random(4000 + k)
    base = _random_reads(rnd, 80, 10, 4 * k + 70, "ACGT" * 40 + "Nacgt")
    reads = [rnd.choice(base)[rnd.randint(0, 7):] for _ in range(4000)] + ["A" * 200, "ACGT" * 60, "", "AC"]
    seq, offs = pkg.pack_reads(reads)
    okeys, ocnt = O.count_kmers(seq, offs, k)
    for mode in (pkg.CANONICAL, pkg.NONCANONICAL):
        if mode == pkg.NONCANONICAL and k == 32:
            continue
        counts = pkg.serial_mem(k, mode, ctx=ctx)((seq, offs))
        keys, c32, c8 = counts.export()
        wk, wc = (okeys, ocnt) if mode == pkg.CANONICAL else O.count_kmers(seq, offs, k, canonical=False)
        assert counts.n_instances == len(O.extract(seq, offs, k))
        assert np.array_equal(keys, wk) and np.array_equal(c32, wc), (k, mode)
        assert np.array_equal(c8, np.minimum(wc, 255).astype(np.uint8))
        counts.free()
    for mf in (1, 2, 5, 256):
        ga = pkg.dbg_from_reads((seq, offs), k, mf, ctx=ctx)
        og = O.dbg(O.filter_counts(okeys, ocnt, k, mf, u8=True), k)   # freq(x) is UInt8-saturated: nothing passes 256
        _graph_equal(ga, og, (k, mf))


@pytest.mark.parametrize("k", [21, 31, 32])
@pytest.mark.parametrize("variant", ["leaf_overflow_all", "leaf_overflow_some", "tiny_bins", "huge_bins_cap_overflow", "leaf_variants"])
def test_skm_stress_paths(pkg, ctx, k, variant, monkeypatch):
    """GG_SKM_TEST_OVERFLOW=m: every m-th leaf bin goes to the overflow list (generic sort + run-length + filter);
    GG_SKM_BIN: target instances per leaf bin (64: thousands of near-empty bins; 200000: few bins whose tables really
    overflow); GG_SKM_TEST_CAP: level-0 capacities too small -> exact repeat of the scatter."""
    env = {"leaf_overflow_all": {"GG_SKM_TEST_OVERFLOW": "1"}, "leaf_overflow_some": {"GG_SKM_TEST_OVERFLOW": "3", "GG_SKM_BIN": "512"},
           "tiny_bins": {"GG_SKM_BIN": "64"}, "huge_bins_cap_overflow": {"GG_SKM_BIN": "200000", "GG_SKM_TEST_CAP": "1"},
           "leaf_variants": {"GG_SKM_VARIANT": "1"}}[variant]
    for a, b in env.items():
        monkeypatch.setenv(a, b)
    seq, offs = pkg.synth_reads(seed=600 + k, genome_len=40000, nreads=8000, read_len=150, paired=True, frag_len=400, err_ppm=5000, n_ppm=200)
    okeys, ocnt = O.count_kmers(seq, offs, k, threads=4)
    counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
    keys, c32, _ = counts.export()
    assert np.array_equal(keys, okeys) and np.array_equal(c32, ocnt), (k, variant)
    counts.free()
    for mf in (2, 3):
        ga = pkg.dbg_from_reads((seq, offs), k, mf, ctx=ctx)
        og = O.dbg(O.filter_counts(okeys, ocnt, k, mf), k)
        _graph_equal(ga, og, (k, variant, mf))
    if variant == "leaf_variants":
        monkeypatch.setenv("GG_SKM_VARIANT", "2")
        c2 = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
        assert np.array_equal(c2.export()[0], okeys)
        c2.free()


def test_skm_larger_input_sampled_capacities(pkg, ctx):
    """>= 256 tiles (8.4 M windows): capacities come from a 1/16 sample of the tiles; kept-pair sort takes the bucket route."""
    k = 31
    seq, offs = pkg.synth_reads(seed=91, genome_len=1500000, nreads=80000, read_len=150, paired=True, frag_len=500, err_ppm=10000)
    okeys, ocnt = O.count_kmers(seq, offs, k, threads=8)
    counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
    keys, c32, _ = counts.export()
    assert np.array_equal(keys, okeys) and np.array_equal(c32, ocnt)
    counts.free()
    ga = pkg.dbg_from_reads((seq, offs), k, 1, ctx=ctx)
    og = O.dbg(okeys, k)
    assert np.array_equal(ga.bases, og.bases) and np.array_equal(ga.link_triples, og.link_tri)


# ------------------------------ error behaviour at the boundary -------------------------------
def test_error_behaviour(pkg, ctx):
    """Bad arguments come back as status codes (ValueError = the reference's ArgumentError), never as crashes,
    and the context stays usable afterwards."""
    N, L = pkg._native, pkg._native.lib()
    seq = np.frombuffer(b"ACGTACGTACGT", dtype=np.uint8)
    h = C.c_void_p()
    bad_offs = np.array([0, 8, 4, 12], dtype=np.uint64)  # decreasing
    assert L.gg_count_kmers(ctx.h, seq.ctypes.data_as(C.c_void_p), bad_offs.ctypes.data_as(C.c_void_p), 3, 4, 1, C.byref(h)) == N.GG_ERR_ARG
    assert b"non-decreasing" in L.gg_last_error(ctx.h)
    off0 = np.array([4, 8, 12], dtype=np.uint64)        # bytes before the first read: rejected (the reference has no such thing)
    assert L.gg_count_kmers(ctx.h, seq.ctypes.data_as(C.c_void_p), off0.ctypes.data_as(C.c_void_p), 2, 4, 1, C.byref(h)) == N.GG_ERR_ARG
    d_seq, d_off = C.c_void_p(), C.c_void_p()
    N.check(ctx.h, L.gg_dev_alloc(ctx.h, 64, C.byref(d_seq)))
    N.check(ctx.h, L.gg_dev_alloc(ctx.h, 64, C.byref(d_off)))
    N.check(ctx.h, L.gg_memcpy_h2d(ctx.h, d_seq, seq.ctypes.data_as(C.c_void_p), 12))
    for bad in (bad_offs, off0, np.array([0, 4, 8, 11], dtype=np.uint64)):   # decreasing / gap at the start / not ending at nbytes
        N.check(ctx.h, L.gg_memcpy_h2d(ctx.h, d_off, bad.ctypes.data_as(C.c_void_p), bad.nbytes))
        assert L.gg_count_kmers_dev(ctx.h, d_seq, d_off, len(bad) - 1, 12, 4, 1, C.byref(h)) == N.GG_ERR_ARG
        assert L.gg_dbg_from_reads_dev(ctx.h, d_seq, d_off, len(bad) - 1, 12, 4, 1, C.byref(h)) == N.GG_ERR_ARG
    L.gg_dev_free(ctx.h, d_seq); L.gg_dev_free(ctx.h, d_off)
    offs = np.array([0, 12], dtype=np.uint64)
    for k in (0, -3, 129):
        assert L.gg_count_kmers(ctx.h, seq.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), 1, k, 1, C.byref(h)) == N.GG_ERR_ARG
    assert L.gg_count_kmers(ctx.h, None, offs.ctypes.data_as(C.c_void_p), 1, 4, 1, C.byref(h)) == N.GG_ERR_ARG      # null bytes, non-empty reads
    assert L.gg_count_kmers(ctx.h, seq.ctypes.data_as(C.c_void_p), None, 1, 4, 1, C.byref(h)) == N.GG_ERR_ARG       # null offsets
    assert L.gg_count_kmers(ctx.h, seq.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), 1, 4, 1, None) == N.GG_ERR_ARG
    assert L.gg_dbg_build(ctx.h, None, C.byref(h)) == N.GG_ERR_ARG
    with pytest.raises(ValueError):
        pkg.serial_mem(200)
    with pytest.raises(ValueError):   # counts given with unsorted k-mers
        km = O.kmers_from_strings(["CCCC", "AAAA"])
        cnt = np.array([1, 1], dtype=np.uint32)
        N.check(ctx.h, L.gg_counts_from_kmers(ctx.h, km.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 2, 4, C.byref(h)))
    # still healthy
    c = pkg.serial_mem(4, ctx=ctx)(["ACGTACGTACGT"])
    assert c.export()[1].tolist() == [3, 4, 2]   # ACGT, CGTA (= TACG), GTAC


# ------------------------------ pipelined host path --------------------------------------------
@pytest.mark.parametrize("case", ["random", "sorted", "forced"])
def test_host_pipeline_chunks(pkg, ctx, case, monkeypatch):
    """gg_count_kmers / gg_dbg_from_reads copy the reads in chunks and run pack, masks, pass A and a speculative
    pass B behind the transfer.  8 KB chunks make a small input take that route; reads sorted by content make the
    first chunk unrepresentative (bucket overflow -> exact pass B); 'forced' discards the speculation."""
    monkeypatch.setenv("GG_H2D_CHUNK_KB", "8")
    if case == "forced":
        monkeypatch.setenv("GG_PIPE_TEST_OVERFLOW", "1")
    seq, offs = pkg.synth_reads(seed=77, genome_len=30000, nreads=6000, read_len=150, paired=True, frag_len=400, err_ppm=3000, n_ppm=200)
    if case == "sorted":
        reads = sorted(bytes(seq[i * 150:(i + 1) * 150]) for i in range(6000))
        seq = np.frombuffer(b"".join(reads), dtype=np.uint8)
    for k in (31, 15):
        okeys, ocnt = O.count_kmers(seq, offs, k)
        counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
        keys, c32, _ = counts.export()
        assert np.array_equal(keys, okeys) and np.array_equal(c32, ocnt), (case, k)
        counts.free()
        ga = pkg.dbg_from_reads((seq, offs), k, 2, ctx=ctx)
        og = O.dbg(O.filter_counts(okeys, ocnt, k, 2), k)
        assert ga.node_strings() == og.nodes() and ga.link_lists() == og.links(), (case, k)


def test_workspace(pkg, ctx):
    """WorkSpace entry points (names of src/workspace/WorkSpace.jl): counts stored by name, dbg! from a store."""
    seq, offs = pkg.synth_reads(seed=5, genome_len=4000, nreads=800, read_len=100, paired=True, frag_len=300, err_ppm=2000)
    ws = pkg.WorkSpace(ctx=ctx)
    ws.add_mer_counts_(21, pkg.CANONICAL, "reads", (seq, offs))
    okeys, ocnt = O.count_kmers(seq, offs, 21)
    assert np.array_equal(ws.mer_counts("reads").export()[0], okeys)
    with pytest.raises(KeyError):
        ws.mer_counts("nope")
    ws.dbg_("reads", 2)
    og = O.dbg(O.filter_counts(okeys, ocnt, 21, 2), 21)
    assert [n.seq for n in ws.graph().nodes] == og.nodes() and ws.graph().link_tuples() == og.links()
    assert "1 " not in repr(ws)[:5] and "workspace" in repr(ws)
    from oracle import graph_edit as E
    eg = E.Graph(og.nodes(), og.links())
    eg.remove_tips(40)
    ws.remove_tips_(40)
    assert [n.seq for n in ws.graph().nodes] == eg.nodes and ws.graph().link_tuples() == eg.links
