"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/ggb200.h declares; no compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    return G.build()


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "ggb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(pkg._native.SO_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table mirrors the header one to one
    assert declared == set(pkg._native.SIGNATURES)


def test_kmer_words(pkg):
    L = pkg._native.lib()
    assert [L.gg_kmer_words(k) for k in (0, 1, 32, 33, 64, 65, 128, 129)] == [0, 1, 1, 2, 2, 4, 4, 0]


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg._native.GGError):
        pkg.Context(0)


def test_product_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, "genomegraphs.jl_b200")
    for dp, _, fs in os.walk(pkgdir):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "ggoracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_synth_reads_host_semantics(pkg):
    """reads are substrings of the synthetic genome (either strand); paired mates face each other."""
    G_, L, F = 5000, 60, 200
    genome = pkg.synth_genome(7, G_).tobytes().decode()
    comp = str.maketrans("ACGT", "TGCA")
    rcg = genome[::-1].translate(comp)
    seq, offs = pkg.synth_reads(7, G_, 400, L, paired=True, frag_len=F)
    assert offs.tolist() == [i * L for i in range(401)]
    reads = [seq[i * L:(i + 1) * L].tobytes().decode() for i in range(400)]
    for j in range(200):
        r1, r2 = reads[2 * j], reads[2 * j + 1]
        p = genome.find(r1)
        if p >= 0:                       # forward fragment: mate 2 is the rc of the fragment end
            assert genome[p + F - L:p + F] == r2[::-1].translate(comp)
        else:
            q = rcg.find(r1)
            assert q >= 0
            assert rcg[q + F - L:q + F] == r2[::-1].translate(comp)
    # determinism + error / N injection rates
    seq2, _ = pkg.synth_reads(7, G_, 400, L, paired=True, frag_len=F)
    assert np.array_equal(seq, seq2)
    e, _ = pkg.synth_reads(7, G_, 400, L, paired=True, frag_len=F, err_ppm=100000)
    frac = float(np.mean(e != seq))
    assert 0.07 < frac < 0.13
    n, _ = pkg.synth_reads(7, G_, 400, L, paired=True, frag_len=F, n_ppm=50000)
    assert 0.03 < float(np.mean(n == ord("N"))) < 0.07
    with pytest.raises(ValueError):
        pkg.synth_reads(7, 10, 4, 60)


def test_read_sequences(pkg, tmp_path):
    """FASTQ (4-line records) and multi-line FASTA -> concatenated sequence bytes + offsets."""
    fq = tmp_path / "a.fastq"
    fq.write_text("@r1\nACGTN\n+\nIIIII\n@r2\nGGCC\n+\nIIII\n")
    fa = tmp_path / "b.fasta"
    fa.write_text(">x desc\nACGT\nAC\n>y\nTTTT\n")
    s, o = pkg.read_sequences(str(fq))
    assert bytes(s) == b"ACGTNGGCC" and o.tolist() == [0, 5, 9]
    s, o = pkg.read_sequences(str(fa))
    assert bytes(s) == b"ACGTACTTTT" and o.tolist() == [0, 6, 10]


def test_gfa_writer(pkg, tmp_path):
    """write_to_gfa1 (src/Graphs.jl:892-926): one S line per node, each link once."""
    g = pkg.empty_graph()
    for s in ("AATC", "AATG", "ACGAAT"):
        g.add_node_(s)
    g.add_link_(1, -3, -3)
    g.add_link_(2, -3, -3)
    g.write_to_gfa1(str(tmp_path / "g"))
    lines = (tmp_path / "g.gfa").read_text().splitlines()
    assert lines[0] == "H\tVN:Z:1.0" and sum(l.startswith("S\t") for l in lines) == 3
    assert sum(l.startswith("L\t") for l in lines) == 2
    assert (tmp_path / "g.fasta").read_text() == ">seq1\nAATC\n>seq2\nAATG\n>seq3\nACGAAT\n"
    assert g.sequence(-1) == "GATT"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU port of the reference algorithm) prints one JSON line with the contract keys, with the
    arm's own thread count stated, the same `config` as the GPU arm would print, and WITHOUT mapping the product library."""
    import json
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'tiny', '--steps', '1', '--warmup', '0']; "
            "import bench; bench.main(); print('MAPS_HAS_PRODUCT', 'libggb200' in open('/proc/self/maps').read())" % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torch.distributed.run exports: the arm must not inherit it
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "MAPS_HAS_PRODUCT False"
    d = json.loads(lines[-2])
    assert d["impl"] == "reference" and d["unit"] == "k-mers/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and "sample" in d["cpu_baseline"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"] == "tiny" and "sample" not in d["config"]
    import bench
    args = type("A", (), {"workload": "tiny", "scaling": "strong", "gpus": 1})()
    assert d["config"] == bench.base_config(args, dict(bench.WORKLOADS["tiny"]), 1)
    assert bench.DEFAULT_WORKLOAD == "c3_100mbp30x_k31"


def test_bench_shard_range_matches_package(pkg):
    import bench
    for n, world in [(101, 3), (20000000, 8), (6, 4), (0, 2)]:
        for r in range(world):
            assert bench.shard_range(n, r, world, 2) == pkg.dist.shard_range(n, r, world, 2)


def test_graph_entry_points_reject_a_null_handle(pkg):
    """no device is needed to be told that there is no graph: every gg_graph_* entry returns GG_ERR_ARG for NULL"""
    import ctypes as C
    L, N = pkg._native.lib(), pkg._native
    out, a, b = C.c_void_p(), C.c_uint64(), C.c_uint64()
    assert L.gg_graph_remove_nodes(None, None, 0, C.byref(out)) == N.GG_ERR_ARG
    assert L.gg_graph_collapse_all_unitigs(None, 2, C.byref(out), C.byref(a)) == N.GG_ERR_ARG
    assert L.gg_graph_remove_tips(None, 200, C.byref(out), C.byref(a), C.byref(b)) == N.GG_ERR_ARG
    assert L.gg_graph_find_tip_nodes(None, 200, None, C.byref(a)) == N.GG_ERR_ARG
    assert L.gg_graph_checksum(None, C.byref(a)) == N.GG_ERR_ARG
    assert out.value is None
