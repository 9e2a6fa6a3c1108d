"""Pins the CPU oracle: hand-traced golden vector G1, the documented
example of /root/reference/docs/src/DeBruijnGraph.md:57-76, and differential checks against
the independent string-based transliteration tests/pyref.py."""
import random

import numpy as np
import pytest

import pyref
from oracle import oracle as O


def _strs(keys, k):
    return [O.words_to_kmer(r, k) for r in keys]


def test_words_and_roundtrip():
    assert [O.words(k) for k in (1, 31, 32, 33, 63, 64, 65, 127, 128)] == [1, 1, 1, 2, 2, 2, 4, 4, 4]
    rnd = random.Random(1)
    for k in (3, 31, 32, 33, 64, 65, 96, 127, 128):
        s = "".join(rnd.choice("ACGT") for _ in range(k))
        assert O.words_to_kmer(O.kmer_to_words(s), k) == s


def test_g1_golden():
    # canonicalised k-mers of docs/src/DeBruijnGraph.md:57-58, k=4
    kmers = ["AATC", "AATG", "ACGA", "ATTC", "CGAA", "CGTA", "CGTC"]
    g = O.dbg(O.kmers_from_strings(kmers), 4)
    assert g.nodes() == ["AATC", "AATG", "ACGAAT", "CGTA", "CGTC"]
    assert g.links() == [
        [(1, -3, -3)],
        [(2, -3, -3)],
        [(-3, 1, -3), (-3, 2, -3), (3, 4, -3), (3, 5, -3)],
        [(4, 3, -3)],
        [(5, 3, -3)],
    ]
    nodes, links = pyref.dbg(kmers)
    assert nodes == g.nodes() and links == g.links()


def test_doc_example_merged_node():
    # docs/src/DeBruijnGraph.md:57-76: nodes ATTC/TTCG(=CGAA)/TCGT(=ACGA) collapse into ACGAAT
    raw = ["ATTC", "TTCG", "TCGT", "AATC", "AATG", "CGTA", "CGTC"]
    canon = sorted({pyref.canonical(x) for x in raw})
    g = O.dbg(O.kmers_from_strings(canon), 4)
    assert "ACGAAT" in g.nodes() and g.n_nodes == 5


def test_single_node_has_no_links():
    # src/dbg.jl:267 -- connect only if n_nodes > 1 (a self-overlapping single node gets no link)
    g = O.dbg(O.kmers_from_strings(["AAA"]), 3)
    assert g.nodes() == ["AAA"] and g.links() == [[]]


def test_empty():
    g = O.dbg(np.zeros((0, 1), dtype=np.uint64), 5)
    assert g.n_nodes == 0


def _random_reads(rnd, n, lo, hi, alphabet="ACGT"):
    return ["".join(rnd.choice(alphabet) for _ in range(rnd.randint(lo, hi))) for _ in range(n)]


def _pack(reads):
    seq = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return seq, offs


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 13, 31, 32, 33, 47, 63, 64, 65, 100, 127, 128])
def test_extract_count_vs_pyref(k):
    rnd = random.Random(100 + k)
    reads = _random_reads(rnd, 40, 0, 3 * k + 20, "ACGTacgtN" if k < 60 else "ACGTN" + "ACGT" * 20)
    seq, offs = _pack(reads)
    for canon in (True, False):
        got = _strs(O.extract(seq, offs, k, canon), k)
        assert got == pyref.extract(reads, k, canon)
    keys, counts = O.count_kmers(seq, offs, k)
    want = pyref.count(pyref.extract(reads, k))
    assert _strs(keys, k) == [a for a, _ in want]
    assert counts.tolist() == [b for _, b in want]
    keys2, counts2 = O.count_kmers(seq, offs, k, threads=4)
    assert np.array_equal(keys, keys2) and np.array_equal(counts, counts2)
    for mf in (1, 2, 3):
        f = O.filter_counts(keys, counts, k, mf)
        assert _strs(f, k) == [a for a, b in want if b >= mf]


def test_filter_u8_saturation():
    keys = O.kmers_from_strings(["AAA", "AAC", "AAG"])
    counts = np.array([254, 255, 1000], dtype=np.uint32)
    assert len(O.filter_counts(keys, counts, 3, 256, u8=True)) == 0
    assert len(O.filter_counts(keys, counts, 3, 255, u8=True)) == 2
    assert len(O.filter_counts(keys, counts, 3, 256, u8=False)) == 1


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8, 9])
def test_dbg_fuzz_vs_pyref(k):
    """Random canonical k-mer sets at small k: exercises hairpins, palindromes, circles, self-loops."""
    rnd = random.Random(7000 + k)
    space = 4 ** k
    for trial in range(60):
        dens = rnd.choice([0.02, 0.1, 0.3, 0.6, 0.9])
        n = max(1, min(300, int(space * dens)))
        ks = set()
        for _ in range(n):
            s = "".join(rnd.choice("ACGT") for _ in range(k))
            ks.add(pyref.canonical(s))
        ks = sorted(ks)
        g = O.dbg(O.kmers_from_strings(ks), k)
        nodes, links = pyref.dbg(ks)
        assert g.nodes() == nodes, (k, trial)
        assert g.links() == links, (k, trial)


@pytest.mark.parametrize("k", [5, 9, 15, 31, 33, 63, 65, 127])
def test_dbg_genome_like_vs_pyref(k):
    """k-mers of a random genome with repeats + a circular piece: long unitigs and perfect circles."""
    rnd = random.Random(9000 + k)
    base = "".join(rnd.choice("ACGT") for _ in range(600))
    rep = base[100:100 + 2 * k]
    genome = base[:300] + rep + base[300:] + rc_tail(base, k)
    circ = "".join(rnd.choice("ACGT") for _ in range(3 * k + 17))
    circ_lin = circ + circ[:k - 1]          # every k-mer of the circle exactly once -> perfect circle
    ks = sorted(set(pyref.extract([genome, circ_lin], k)))
    g = O.dbg(O.kmers_from_strings(ks), k)
    nodes, links = pyref.dbg(ks)
    assert g.nodes() == nodes
    assert g.links() == links
    assert sum(len(s) - k + 1 for s in nodes) >= len(ks)


def rc_tail(base, k):
    # a hairpin: sequence followed by its own reverse complement
    return pyref.rc(base[-(2 * k):])


@pytest.mark.parametrize("k", [3, 4, 7, 31, 33, 127])
def test_unique_mer_index_vs_pyref(k):
    rnd = random.Random(500 + k)
    nodes = _random_reads(rnd, 12, 0, 4 * k + 9, "ACGT" if k > 4 else "AC")
    nodes.append(nodes[0])
    keys, node, pos, upn, tpn = O.unique_mer_index(nodes, k)
    uniq, upn_r, tot_r = pyref.unique_mer_index(nodes, k)
    assert _strs(keys, k) == [u[0] for u in uniq]
    assert node.tolist() == [u[1] for u in uniq]
    assert pos.tolist() == [u[2] for u in uniq]
    assert upn.tolist() == upn_r and tpn.tolist() == tot_r


def test_oracle_synth_matches_product_generator():
    """The oracle's C restatement of the synthetic read generator gives the product generator's bytes
    (the CPU legs of bench.py use the oracle's, so that they never load libggb200.so)."""
    import __graft_entry__ as G
    pkg = G.build()
    for args in [(7, 5000, 400, 60, True, 200, 0, 0, 0), (3, 20000, 501, 150, False, 0, 20000, 500, 0),
                 (1, 4641652, 2000, 150, True, 700, 1000, 0, 1000), (2, 100000, 64, 150, True, 700, 10000, 100, 4000)]:
        seed, glen, n, rl, paired, fl, err, nppm, first = args
        a, ao = pkg.synth_reads(seed, glen, n, rl, paired, fl, err, nppm, first_read=first)
        b, bo = O.synth_reads(seed, glen, n, rl, paired, fl, err, nppm, first_read=first, threads=3)
        assert np.array_equal(a, b) and np.array_equal(ao, bo)
    with pytest.raises(ValueError):
        O.synth_reads(7, 10, 4, 60)
