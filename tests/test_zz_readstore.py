"""read_datastore (src/GenomeGraphs.jl:79-83, docs/src/man/assembly.md:77-100) in the host layer: FASTQ pair -> filtered,
4-bit encoded paired-read store on disk -> memory-mapped -> the device counters.  CPU: filter, container round trip.
GPU (this file sorts last on purpose: its GPU case was written after the round's GPU lease was spent and has not run on a
B200 yet; it only composes entry points the other tests cover): counting and graph from the store = from the filtered reads."""
import random

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import oracle as O


def _fastq_pair(tmp_path, seed=4, n=300):
    rnd = random.Random(seed)
    genome = "".join(rnd.choice("ACGT") for _ in range(3000))
    comp = str.maketrans("ACGT", "TGCA")
    r1, r2 = [], []
    for _ in range(n):
        p = rnd.randint(0, len(genome) - 400)
        frag = genome[p:p + 400]
        a = frag[:rnd.randint(40, 130)]
        b = frag[::-1].translate(comp)[:rnd.randint(40, 130)]
        if rnd.random() < 0.05:
            a = a[:10] + "N" + a[11:]
        r1.append(a)
        r2.append(b)
    for name, rs in (("R1.fastq", r1), ("R2.fastq", r2)):
        with open(tmp_path / name, "w") as f:
            for i, s in enumerate(rs):
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))
    return r1, r2


def _expected(r1, r2, minlen, maxlen):
    out = []
    for a, b in zip(r1, r2):
        if len(a) >= minlen and len(b) >= minlen:
            out += [a[:maxlen], b[:maxlen]]
    return out


def test_read_datastore_filter_and_container(tmp_path):
    pkg = G.build()
    r1, r2 = _fastq_pair(tmp_path)
    ds = pkg.read_datastore(str(tmp_path / "R1.fastq"), str(tmp_path / "R2.fastq"), str(tmp_path / "ecoli-pe"), 60, 100, 0, pkg.FwRv)
    exp = _expected(r1, r2, 60, 100)
    assert 0 < len(exp) < 2 * len(r1) and any(len(x) == 100 for x in exp)
    assert len(ds) == len(exp) and ds.n_pairs() == len(exp) // 2
    assert [ds[i] for i in range(len(ds))] == exp
    assert (tmp_path / "ecoli-pe.prseq").exists() and ds.name == "ecoli-pe" and (ds.minlen, ds.maxlen, ds.orientation) == (60, 100, "FwRv")
    again = pkg.PairedReads.open(str(tmp_path / "ecoli-pe.prseq"))
    assert np.array_equal(again.base_offsets, ds.base_offsets) and np.array_equal(again.data4, ds.data4)
    assert ds.data4.nbytes * 2 <= sum(len(x) for x in exp) + 16          # 4 bits per base
    with open(tmp_path / "bad.prseq", "wb") as f:
        f.write(b"not a datastore")
    with pytest.raises(ValueError):
        pkg.PairedReads.open(str(tmp_path / "bad.prseq"))
    with pytest.raises(IndexError):
        ds[len(ds)]


@pytest.mark.gpu
def test_gpu_counts_and_graph_from_the_store(tmp_path):
    pkg = G.build()
    ctx = pkg.Context(0)
    r1, r2 = _fastq_pair(tmp_path, seed=6, n=500)
    ds = pkg.read_datastore(str(tmp_path / "R1.fastq"), str(tmp_path / "R2.fastq"), str(tmp_path / "pe"), 50, 120)
    exp = _expected(r1, r2, 50, 120)
    seq = np.frombuffer("".join(exp).encode(), dtype=np.uint8)
    offs = np.cumsum([0] + [len(x) for x in exp]).astype(np.uint64)
    for k in (21, 31):
        okeys, ocnt = O.count_kmers(seq, offs, k)
        counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)(ds)
        keys, c32, _ = counts.export()
        assert np.array_equal(keys, okeys) and np.array_equal(c32, ocnt)
        counts.free()
        ga = pkg.dbg_from_reads(ds, k, 2, ctx=ctx)
        og = O.dbg(O.filter_counts(okeys, ocnt, k, 2), k)
        assert ga.node_strings() == og.nodes() and ga.link_lists() == og.links()
    ctx.close()
