"""ctypes wrapper of the CPU oracle (oracle/libggoracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.
PARITY UNPINNED -- see the header of oracle/gg_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libggoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("gg_oracle.c", "oracle_impl.h", "graph_edit.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libggoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u32p, u64p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64))
        L.ggo_words.restype = C.c_int
        L.ggo_words.argtypes = [C.c_int]
        L.ggo_extract.restype = C.c_uint64
        L.ggo_extract.argtypes = [u8p, u64p, C.c_uint64, C.c_int, C.c_int, u64p]
        L.ggo_sort_count.restype = C.c_uint64
        L.ggo_sort_count.argtypes = [u64p, C.c_uint64, C.c_int, u32p, C.c_int]
        L.ggo_filter.restype = C.c_uint64
        L.ggo_filter.argtypes = [u64p, u32p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, u64p]
        L.ggo_dbg_build.restype = C.c_void_p
        L.ggo_dbg_build.argtypes = [u64p, C.c_uint64, C.c_int]
        L.ggo_graph_sizes.restype = None
        L.ggo_graph_sizes.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.ggo_graph_export.restype = None
        L.ggo_graph_export.argtypes = [C.c_void_p, u64p, u8p, u64p, i64p]
        L.ggo_graph_free.restype = None
        L.ggo_graph_free.argtypes = [C.c_void_p]
        L.ggo_unique_mer_index.restype = C.c_uint64
        L.ggo_unique_mer_index.argtypes = [u8p, u64p, C.c_uint64, C.c_int, u64p, i64p, u32p, u64p, u64p]
        L.ggo_max_threads.restype = C.c_int
        L.ggo_synth_reads_range.restype = C.c_int
        L.ggo_synth_reads_range.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                            C.c_uint32, u8p, C.c_int]
        # graph_edit.c: sequential restatement of remove_tips! and what it calls
        L.ggo_edit_graph_new.restype = C.c_void_p
        L.ggo_edit_graph_new.argtypes = [C.c_uint64, u64p, u8p, u64p, i64p]
        L.ggo_edit_graph_free.restype = None
        L.ggo_edit_graph_free.argtypes = [C.c_void_p]
        L.ggo_edit_graph_sizes.restype = None
        L.ggo_edit_graph_sizes.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.ggo_edit_graph_export.restype = None
        L.ggo_edit_graph_export.argtypes = [C.c_void_p, u64p, u8p, u64p, i64p]
        L.ggo_edit_find_tips.restype = C.c_uint64
        L.ggo_edit_find_tips.argtypes = [C.c_void_p, C.c_uint64, i64p]
        L.ggo_edit_remove_nodes.restype = None
        L.ggo_edit_remove_nodes.argtypes = [C.c_void_p, i64p, C.c_uint64]
        L.ggo_edit_collapse.restype = C.c_int64
        L.ggo_edit_collapse.argtypes = [C.c_void_p, C.c_uint64]
        L.ggo_edit_remove_tips.restype = C.c_int64
        L.ggo_edit_remove_tips.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def words(k):
    w = lib().ggo_words(int(k))
    if w == 0:
        raise ValueError("k must be in 1..128")
    return w


def _seq_offs(seq, offs):
    seq = np.ascontiguousarray(np.frombuffer(bytes(seq), dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else seq, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    if seq.size == 0:
        seq = np.zeros(1, dtype=np.uint8)
    return seq, offs


def extract(seq, offs, k, canonical=True):
    """All k-mer windows in read order -> (n, W) uint64 array (most significant word first)."""
    seq, offs = _seq_offs(seq, offs)
    w = words(k)
    nreads = len(offs) - 1
    n = lib().ggo_extract(_p(seq, C.c_uint8), _p(offs, C.c_uint64), nreads, k, int(canonical), None)
    out = np.zeros((max(n, 1), w), dtype=np.uint64)
    lib().ggo_extract(_p(seq, C.c_uint8), _p(offs, C.c_uint64), nreads, k, int(canonical), _p(out, C.c_uint64))
    return out[:n]


def sort_count(kmers, k, threads=1):
    """(n, W) k-mers -> sorted unique (u, W) + u32 counts."""
    w = words(k)
    a = np.array(kmers, dtype=np.uint64, copy=True).reshape(-1, w)
    n = a.shape[0]
    counts = np.zeros(max(n, 1), dtype=np.uint32)
    if n == 0:
        return a, counts[:0]
    u = lib().ggo_sort_count(_p(a, C.c_uint64), n, k, _p(counts, C.c_uint32), int(threads))
    return a[:u].copy(), counts[:u].copy()


def count_kmers(seq, offs, k, canonical=True, threads=1):
    return sort_count(extract(seq, offs, k, canonical), k, threads)


def filter_counts(keys, counts, k, min_freq, u8=False):
    w = words(k)
    keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, w)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    out = np.zeros((max(len(keys), 1), w), dtype=np.uint64)
    n = lib().ggo_filter(_p(keys, C.c_uint64), _p(counts, C.c_uint32), len(keys), k, int(min_freq), int(u8), _p(out, C.c_uint64))
    return out[:n].copy()


class Graph:
    """nodes: list of ASCII strings (creation order, id = index + 1);
    links: list (per node) of (src, dst, dist) tuples in push order."""

    def __init__(self, node_off, bases, link_row, link_tri):
        self.node_off, self.bases, self.link_row, self.link_tri = node_off, bases, link_row, link_tri

    @property
    def n_nodes(self):
        return len(self.node_off) - 1

    def nodes(self):
        b = self.bases.tobytes().decode()
        return [b[self.node_off[i]:self.node_off[i + 1]] for i in range(self.n_nodes)]

    def links(self):
        t = self.link_tri.reshape(-1, 3)
        return [[tuple(int(x) for x in t[j]) for j in range(self.link_row[i], self.link_row[i + 1])] for i in range(self.n_nodes)]


def dbg(kmers, k):
    """dbg!(graph, kmers): kmers (n, W) canonical, duplicate-free (sorted here if needed)."""
    w = words(k)
    a = np.array(kmers, dtype=np.uint64, copy=True).reshape(-1, w)
    h = lib().ggo_dbg_build(_p(a, C.c_uint64) if len(a) else None, len(a), k)
    try:
        nn, nb, nl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().ggo_graph_sizes(h, C.byref(nn), C.byref(nb), C.byref(nl))
        node_off = np.zeros(nn.value + 1, dtype=np.uint64)
        bases = np.zeros(max(nb.value, 1), dtype=np.uint8)
        link_row = np.zeros(nn.value + 1, dtype=np.uint64)
        link_tri = np.zeros(max(nl.value, 1) * 3, dtype=np.int64)
        lib().ggo_graph_export(h, _p(node_off, C.c_uint64), _p(bases, C.c_uint8), _p(link_row, C.c_uint64), _p(link_tri, C.c_int64))
        return Graph(node_off.astype(np.int64), bases[:nb.value], link_row.astype(np.int64), link_tri[:nl.value * 3])
    finally:
        lib().ggo_graph_free(h)


def unique_mer_index(node_seqs, k):
    """node_seqs: list of ASCII strings -> (keys (u,W), node i64, pos u32, unique_per_node, total_per_node)."""
    w = words(k)
    seq = np.frombuffer("".join(node_seqs).encode(), dtype=np.uint8)
    offs = np.zeros(len(node_seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in node_seqs])
    seq, offs = _seq_offs(seq, offs)
    nn = len(node_seqs)
    t = int(sum(max(0, len(s) - k + 1) for s in node_seqs))
    keys = np.zeros((max(t, 1), w), dtype=np.uint64)
    node = np.zeros(max(t, 1), dtype=np.int64)
    pos = np.zeros(max(t, 1), dtype=np.uint32)
    upn = np.zeros(max(nn, 1), dtype=np.uint64)
    tpn = np.zeros(max(nn, 1), dtype=np.uint64)
    u = lib().ggo_unique_mer_index(_p(seq, C.c_uint8), _p(offs, C.c_uint64), nn, k, _p(keys, C.c_uint64),
                                   _p(node, C.c_int64), _p(pos, C.c_uint32), _p(upn, C.c_uint64), _p(tpn, C.c_uint64))
    return keys[:u], node[:u], pos[:u], upn[:nn], tpn[:nn]


def synth_reads(seed, genome_len, nreads, read_len, paired=False, frag_len=0, err_ppm=0, n_ppm=0, first_read=0, threads=1):
    """Deterministic synthetic reads -> (seq_bytes, offsets); the same bytes as the product's gg_synth_reads_range_host / _dev
    (C restatement of csrc/synth.cuh inside the oracle library, so the CPU legs of bench.py need not load the product)."""
    out = np.zeros(int(nreads) * int(read_len), dtype=np.uint8)
    rc = lib().ggo_synth_reads_range(int(seed), int(genome_len), int(first_read), int(nreads), int(read_len), int(bool(paired)), int(frag_len),
                                     int(err_ppm), int(n_ppm), _p(out, C.c_uint8), int(threads))
    if rc != 0:
        raise ValueError("bad synthetic read parameters")
    return out, np.arange(int(nreads) + 1, dtype=np.uint64) * np.uint64(read_len)


# ---- helpers shared by tests (string <-> word conversion) -------------------------------
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def kmer_to_words(s):
    k = len(s)
    v = 0
    for ch in s:
        v = (v << 2) | _CODE[ch]
    w = words(k)
    return [(v >> (64 * (w - 1 - i))) & 0xFFFFFFFFFFFFFFFF for i in range(w)]


def words_to_kmer(ws, k):
    v = 0
    for x in ws:
        v = (v << 64) | int(x)
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def kmers_from_strings(strs):
    k = len(strs[0])
    return np.array([kmer_to_words(s) for s in strs], dtype=np.uint64).reshape(-1, words(k))
