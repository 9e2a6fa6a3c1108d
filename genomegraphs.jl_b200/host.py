"""Python host mirror of the GenomeGraphs.jl call surface for the DBG construction path.

Same names, argument meaning and error behaviour as the reference (Julia `!` suffix is
spelled `_`): `serial_mem(k, CANONICAL)(reads)`, `mer`/`freq`, `empty_graph`, `dbg_`
(`dbg!`, src/dbg.jl:261-278), `SequenceDistanceGraph` / `SDGNode` / `SDGLink`
(src/Graphs.jl:65-68,163-167,232-243), `UniqueMerIndex` (src/GraphIndexes.jl:27-102).
Everything computational is a call into libggb200.so (CUDA, sm_100a); there is no CPU path.
The Julia host files that `ccall` the same ABI are under genomegraphs.jl_b200/julia/.
"""
import ctypes as C

import os

import numpy as np

from . import _native as N

CANONICAL = 1
NONCANONICAL = 0


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU, one private stream (gg_ctx).  A process-wide default is created on demand."""

    _default = None

    def __init__(self, device=0):
        h = C.c_void_p()
        code = N.lib().gg_ctx_create(int(device), C.byref(h))
        if code != N.GG_OK:
            raise N.GGError(code, "no usable CUDA device %d (libggb200 has no CPU fallback)" % device)
        self.h = h
        self.device = device

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls(0)
        return cls._default

    def close(self):
        if self.h:
            N.lib().gg_ctx_destroy(self.h)
            self.h = None

    def last_timings(self):
        ms = (C.c_float * 9)()
        n = C.c_uint64()
        N.lib().gg_ctx_last_timings(self.h, ms, C.byref(n))
        names = ["pack", "extract", "sort", "count", "filter", "lookup", "unitigs", "links", "total"]
        return dict(zip(names, [float(x) for x in ms])), int(n.value)


def pack_reads(reads):
    """list of str/bytes -> (seq_bytes uint8[n], read_offsets uint64[nreads+1])."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    seq = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return seq, offs


def read_sequences(path):
    """Sequence lines of a FASTA or FASTQ file -> (seq_bytes, read_offsets): the host side of
    `read_datastore` / `FASTQ.Reader` (src/GenomeGraphs.jl:79-83, examples/construct_dbg.jl:15-21).
    Only the sequence lines matter for this path; headers and qualities are dropped on the host."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    seqs = []
    with op(path, "rb") as f:
        first = f.read(1)
        f.seek(0)
        if first == b"@":     # FASTQ: 4-line records
            for i, line in enumerate(f):
                if i % 4 == 1:
                    seqs.append(line.strip())
        else:                 # FASTA: a record's sequence may span lines
            cur = []
            for line in f:
                if line.startswith(b">"):
                    if cur:
                        seqs.append(b"".join(cur))
                    cur = []
                else:
                    cur.append(line.strip())
            if cur:
                seqs.append(b"".join(cur))
    return pack_reads(seqs)


def _as_reads(reads):
    if isinstance(reads, tuple) and len(reads) == 2:
        seq = np.ascontiguousarray(reads[0], dtype=np.uint8)
        offs = np.ascontiguousarray(reads[1], dtype=np.uint64)
        return seq, offs
    return pack_reads(reads)


class MerCounts:
    """`Vector{MerCount{M}}` held on the device (gg_counts): sorted unique k-mers + counts."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def _sizes(self):
        nu, ni, k, w = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
        N.check(self.ctx.h, N.lib().gg_counts_sizes(self.h, C.byref(nu), C.byref(ni), C.byref(k), C.byref(w)))
        return nu.value, ni.value, k.value, w.value

    def __len__(self):
        return self._sizes()[0]

    @property
    def k(self):
        return self._sizes()[2]

    @property
    def n_instances(self):
        return self._sizes()[1]

    def export(self):
        """-> (kmer words (U, W) uint64, counts uint32, freq uint8 saturated)."""
        nu, _, _, w = self._sizes()
        keys = np.zeros((nu, w), dtype=np.uint64)
        c32 = np.zeros(nu, dtype=np.uint32)
        c8 = np.zeros(nu, dtype=np.uint8)
        N.check(self.ctx.h, N.lib().gg_counts_export(self.h, _ptr(keys), _ptr(c8), _ptr(c32)))
        return keys, c32, c8

    def filter_(self, min_freq, u8_saturated=False):
        """filter!(x -> freq(x) >= min_freq, counts) -- mutates, like src/dbg.jl:275."""
        N.check(self.ctx.h, N.lib().gg_counts_filter(self.h, int(min_freq), int(bool(u8_saturated))))
        return self

    def spectrum(self):
        h = np.zeros(256, dtype=np.uint64)
        N.check(self.ctx.h, N.lib().gg_counts_spectrum(self.h, h.ctypes.data_as(N.u64p)))
        return h

    def free(self):
        if self.h:
            if self.ctx.h:  # a closed context has already released the device
                N.lib().gg_counts_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def mer(counts):
    """mer.(counts): the k-mer words."""
    return counts.export()[0]


def freq(counts):
    """freq.(counts): UInt8-saturated counts (KmerAnalysis MerCount)."""
    return counts.export()[2]


def spectra(counts):
    return counts.spectrum()


def serial_mem(k, mode=CANONICAL, ctx=None):
    """`serial_mem(DNAMer{k}, CANONICAL)` -> counter; `counter(reads)` -> MerCounts
    (docs/src/man/assembly.md:126-128).  reads: list of str/bytes or (seq_bytes, offsets)."""
    if N.lib().gg_kmer_words(int(k)) == 0:
        raise ValueError("k must be in 1..128")

    def counter(reads):
        c = ctx or Context.default()
        if isinstance(reads, PairedReads):     # the read store's own 4-bit encoding goes to the device as it is
            return count_kmers_4bit(reads.data4, reads.base_offsets, k, mode, ctx=c)
        seq, offs = _as_reads(reads)
        h = C.c_void_p()
        N.check(c.h, N.lib().gg_count_kmers(c.h, _ptr(seq) if seq.size else None, _ptr(offs), len(offs) - 1, int(k), int(mode), C.byref(h)))
        return MerCounts(c, h)

    return counter


def extract_kmers(reads, k, mode=CANONICAL, ctx=None):
    """each(M, seq) + canonical over all reads, in read order -> (n, W) uint64."""
    c = ctx or Context.default()
    seq, offs = _as_reads(reads)
    w = N.lib().gg_kmer_words(int(k))
    if w == 0:
        raise ValueError("k must be in 1..128")
    n = C.c_uint64(0)
    sp = _ptr(seq) if seq.size else None
    N.check(c.h, N.lib().gg_extract_kmers(c.h, sp, _ptr(offs), len(offs) - 1, int(k), int(mode), None, C.byref(n)))
    out = np.zeros((n.value, w), dtype=np.uint64)
    if n.value:
        N.check(c.h, N.lib().gg_extract_kmers(c.h, sp, _ptr(offs), len(offs) - 1, int(k), int(mode), _ptr(out), C.byref(n)))
    return out


class SDGNode:
    """src/Graphs.jl:65-68"""
    __slots__ = ("seq", "deleted")

    def __init__(self, seq, deleted=False):
        self.seq, self.deleted = seq, deleted

    def __eq__(self, o):
        return isinstance(o, SDGNode) and self.seq == o.seq and self.deleted == o.deleted

    def __repr__(self):
        return "SDGNode(%r, %r)" % (self.seq, self.deleted)


class SDGLink:
    """src/Graphs.jl:163-167; `==` ignores dist (:183-185)."""
    __slots__ = ("source", "destination", "dist")

    def __init__(self, source, destination, dist=0):
        self.source, self.destination, self.dist = int(source), int(destination), int(dist)

    def __eq__(self, o):
        return isinstance(o, SDGLink) and self.source == o.source and self.destination == o.destination

    def astuple(self):
        return (self.source, self.destination, self.dist)

    def __repr__(self):
        return "SDGLink(%d, %d, %d)" % self.astuple()


class SequenceDistanceGraph:
    """src/Graphs.jl:232-235: nodes::Vector{SDGNode}, links::Vector{Vector{SDGLink}}."""

    def __init__(self):
        self.nodes = []
        self.links = []

    def n_nodes(self):
        return len(self.nodes)

    def add_node_(self, seq):
        """add_node! src/Graphs.jl:418-422,442-444"""
        self.nodes.append(seq if isinstance(seq, SDGNode) else SDGNode(seq, False))
        self.links.append([])
        return len(self.nodes)

    def _check_node_id(self, i):
        if not (0 < abs(i) <= len(self.nodes)):
            raise IndexError("Sequence graph has no node with ID of %d" % i)

    def add_link_(self, source, dest, dist):
        """add_link! src/Graphs.jl:473-476"""
        self._check_node_id(source)
        self._check_node_id(dest)
        self.links[abs(source) - 1].append(SDGLink(source, dest, dist))
        self.links[abs(dest) - 1].append(SDGLink(dest, source, dist))

    def sequence(self, n):
        """src/Graphs.jl:592-600"""
        self._check_node_id(n)
        s = self.nodes[abs(n) - 1].seq
        if n < 0:
            s = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
        return s

    def link_tuples(self):
        return [[l.astuple() for l in ls] for ls in self.links]

    def write_to_gfa1(self, filename):
        """src/Graphs.jl:892-926: <filename>.gfa + <filename>.fasta"""
        fasta_filename = filename + ".fasta"
        with open(filename + ".gfa", "w") as gfa, open(fasta_filename, "w") as fa:
            gfa.write("H\tVN:Z:1.0\n")
            for nid, n in enumerate(self.nodes, start=1):
                if n.deleted:
                    continue
                gfa.write("S\tseq%d\t*\tLN:i:%d\tUR:Z:%s\n" % (nid, len(n.seq), fasta_filename))
                fa.write(">seq%d\n" % nid)   # FASTX FASTA.Writer: sequence lines of 70 bases
                for p in range(0, len(n.seq), 70):
                    fa.write(n.seq[p:p + 70] + "\n")
            for ls in self.links:
                for l in ls:
                    if l.source <= l.destination:
                        a = "seq%d\t-" % l.source if l.source > 0 else "seq%d\t+" % -l.source
                        b = "seq%d\t+" % l.destination if l.destination > 0 else "seq%d\t-" % -l.destination
                        gfa.write("L\t%s\t%s\t%dM\n" % (a, b, -l.dist if l.dist < 0 else 0))


def empty_graph(_seqtype=None):
    """empty_graph(LongSequence{DNAAlphabet{4}}) src/GenomeGraphs.jl:85"""
    return SequenceDistanceGraph()


class GraphArrays:
    """Raw export of a gg_graph (CSR arrays), for large graphs where Python objects are too slow."""

    def __init__(self, k, node_offsets, bases, link_row_offsets, link_triples):
        self.k = k
        self.node_offsets, self.bases = node_offsets, bases
        self.link_row_offsets, self.link_triples = link_row_offsets, link_triples

    @property
    def n_nodes(self):
        return len(self.node_offsets) - 1

    def node_strings(self):
        b = self.bases.tobytes().decode()
        o = self.node_offsets
        return [b[int(o[i]):int(o[i + 1])] for i in range(self.n_nodes)]

    def link_lists(self):
        t = self.link_triples.reshape(-1, 3)
        r = self.link_row_offsets
        return [[tuple(int(x) for x in t[j]) for j in range(int(r[i]), int(r[i + 1]))] for i in range(self.n_nodes)]


def write_gfa1_arrays(filename, ga):
    """write_to_gfa1 (src/Graphs.jl:892-926) straight from exported arrays (host code in libggb200, no GPU needed):
    <filename>.gfa + <filename>.fasta"""
    nl = len(ga.link_triples) // 3
    code = N.lib().gg_gfa1_write(str(filename).encode(), ga.n_nodes, _ptr(ga.node_offsets), _ptr(ga.bases) if len(ga.bases) else None,
                                 _ptr(ga.link_row_offsets), _ptr(ga.link_triples) if nl else None)
    if code != N.GG_OK:
        raise OSError("cannot write %s.gfa / %s.fasta" % (filename, filename))


class DeviceGraph:
    """gg_graph handle: the graph stays on the device (find_tip_nodes, GFA1 persistence, export)."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    @classmethod
    def from_arrays(cls, ga, ctx=None):
        c = ctx or Context.default()
        h = C.c_void_p()
        nl = len(ga.link_triples) // 3
        N.check(c.h, N.lib().gg_graph_from_arrays(c.h, int(ga.k), ga.n_nodes, _ptr(np.ascontiguousarray(ga.node_offsets, dtype=np.uint64)),
                                                  _ptr(ga.bases) if len(ga.bases) else None, _ptr(np.ascontiguousarray(ga.link_row_offsets, dtype=np.uint64)),
                                                  _ptr(ga.link_triples) if nl else None, C.byref(h)))
        return cls(c, h)

    @classmethod
    def load_from_gfa1(cls, gfa_file, fasta_file, ctx=None):
        """load_from_gfa1!(sg, gfafile, fafile) (unfinished upstream, src/Graphs.jl:943-958)"""
        c = ctx or Context.default()
        h = C.c_void_p()
        N.check(c.h, N.lib().gg_graph_load_gfa1(c.h, str(gfa_file).encode(), str(fasta_file).encode(), C.byref(h)))
        return cls(c, h)

    def arrays(self):
        return _export_graph(self.ctx, self.h)

    def write_to_gfa1(self, filename):
        N.check(self.ctx.h, N.lib().gg_graph_write_gfa1(self.h, str(filename).encode()))

    def find_tip_nodes(self, min_size):
        """find_tip_nodes(sg, min_size) (src/Graphs.jl:778-816): ascending node ids"""
        n = C.c_uint64(0)
        N.check(self.ctx.h, N.lib().gg_graph_find_tip_nodes(self.h, int(min_size), None, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.int64)
        n2 = C.c_uint64(len(out))
        N.check(self.ctx.h, N.lib().gg_graph_find_tip_nodes(self.h, int(min_size), _ptr(out), C.byref(n2)))
        return out[:n2.value]

    def remove_nodes(self, ids):
        """remove_node! (src/Graphs.jl:459-466) for every id: a new DeviceGraph"""
        a = np.ascontiguousarray(ids, dtype=np.int64)
        h = C.c_void_p()
        N.check(self.ctx.h, N.lib().gg_graph_remove_nodes(self.h, _ptr(a) if a.size else None, a.size, C.byref(h)))
        return DeviceGraph(self.ctx, h)

    def collapse_all_unitigs(self, min_nodes=2, consume=True):
        """collapse_all_unitigs!(sg, min_nodes, consume) (src/Graphs.jl:528-539): (new DeviceGraph, number of new nodes)"""
        if not consume:
            raise ValueError("only consume = true (what remove_tips! uses) runs on the device")
        h = C.c_void_p()
        m = C.c_uint64(0)
        N.check(self.ctx.h, N.lib().gg_graph_collapse_all_unitigs(self.h, int(min_nodes), C.byref(h), C.byref(m)))
        return DeviceGraph(self.ctx, h), m.value

    def remove_tips(self, min_size):
        """remove_tips!(sdg, min_size) (src/remove_tips.jl:2-30): (new DeviceGraph, passes, tips removed)"""
        h = C.c_void_p()
        p, r = C.c_uint64(0), C.c_uint64(0)
        N.check(self.ctx.h, N.lib().gg_graph_remove_tips(self.h, int(min_size), C.byref(h), C.byref(p), C.byref(r)))
        return DeviceGraph(self.ctx, h), p.value, r.value

    def free(self):
        if self.h:
            N.lib().gg_graph_free(self.h)
            self.h = None


def pack_4bit(reads):
    """reads -> (uint64 words, base offsets) in the 4-bit encoding of LongSequence{DNAAlphabet{4}} (the reference's read
    store, src/GenomeGraphs.jl:79-83): 16 bases per word, base i in bits 4i .. 4i+3, A 1, C 2, G 4, T 8, other letters 15"""
    seq, offs = _as_reads(reads)
    lut = np.full(256, 15, dtype=np.uint64)
    for ch, v in (("A", 1), ("C", 2), ("G", 4), ("T", 8)):
        lut[ord(ch)] = v
        lut[ord(ch.lower())] = v
    nib = lut[seq]
    pad = (-len(nib)) % 16
    nib = np.concatenate([nib, np.zeros(pad, dtype=np.uint64)]).reshape(-1, 16)
    words = (nib << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(axis=1, dtype=np.uint64) if len(nib) else np.zeros(0, dtype=np.uint64)
    return np.ascontiguousarray(words, dtype=np.uint64), offs


def count_kmers_4bit(data4, offs, k, mode=1, ctx=None):
    c = ctx or Context.default()
    h = C.c_void_p()
    N.check(c.h, N.lib().gg_count_kmers_4bit(c.h, _ptr(data4) if data4.size else None, _ptr(offs), len(offs) - 1, int(k), int(mode), C.byref(h)))
    return MerCounts(c, h)


def dbg_from_reads_4bit(data4, offs, k, min_freq, ctx=None):
    c = ctx or Context.default()
    gh = C.c_void_p()
    N.check(c.h, N.lib().gg_dbg_from_reads_4bit(c.h, _ptr(data4) if data4.size else None, _ptr(offs), len(offs) - 1, int(k), int(min_freq), C.byref(gh)))
    try:
        return _export_graph(c, gh)
    finally:
        N.lib().gg_graph_free(gh)


def spectra_cn(reads_counts, graph_counts):
    """`spectra(reads_kcount, graph_kcount)` (docs/src/man/assembly.md:221-228): (256, 256) joint histogram of the
    UInt8-saturated counts, [count in reads, count in graph]"""
    h = np.zeros(65536, dtype=np.uint64)
    N.check(reads_counts.ctx.h, N.lib().gg_counts_spectrum_cn(reads_counts.h, graph_counts.h, h.ctypes.data_as(N.u64p)))
    return h.reshape(256, 256)


def _export_graph(ctx, gh):
    nn, nb, nl, k = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
    N.check(ctx.h, N.lib().gg_graph_sizes(gh, C.byref(nn), C.byref(nb), C.byref(nl), C.byref(k)))
    node_off = np.zeros(nn.value + 1, dtype=np.uint64)
    bases = np.zeros(nb.value, dtype=np.uint8)
    link_row = np.zeros(nn.value + 1, dtype=np.uint64)
    tri = np.zeros(3 * nl.value, dtype=np.int64)
    N.check(ctx.h, N.lib().gg_graph_export(gh, _ptr(node_off), _ptr(bases) if nb.value else None, _ptr(link_row), _ptr(tri) if nl.value else None))
    return GraphArrays(k.value, node_off, bases, link_row, tri)


def _kmers_to_counts(ctx, kmers, k):
    a = np.ascontiguousarray(kmers, dtype=np.uint64)
    w = N.lib().gg_kmer_words(int(k))
    if w == 0:
        raise ValueError("k must be in 1..128")
    a = a.reshape(-1, w)
    h = C.c_void_p()
    N.check(ctx.h, N.lib().gg_counts_from_kmers(ctx.h, _ptr(a) if a.size else None, None, a.shape[0], int(k), C.byref(h)))
    return MerCounts(ctx, h)


def dbg_arrays(kmers_or_counts, k=None, min_freq=None, ctx=None):
    """Device DBG build returning raw arrays (GraphArrays)."""
    c = ctx or Context.default()
    if isinstance(kmers_or_counts, MerCounts):
        counts, own = kmers_or_counts, False
        if min_freq is not None:
            counts.filter_(min_freq, u8_saturated=True)   # freq(x) is the UInt8-saturated count (src/dbg.jl:275)
    else:
        if k is None:
            raise ValueError("Didn't provide a collection of kmers for the `kmers` parameter did ya?")  # src/dbg.jl:263
        counts, own = _kmers_to_counts(c, kmers_or_counts, k), True
    gh = C.c_void_p()
    try:
        N.check(c.h, N.lib().gg_dbg_build(c.h, counts.h, C.byref(gh)))
        return _export_graph(c, gh)
    finally:
        if gh:
            N.lib().gg_graph_free(gh)
        if own:
            counts.free()


def dbg_(sdg, kmers_or_counts, min_freq=None, k=None, ctx=None):
    """dbg!(sdg, kmers) (src/dbg.jl:261-272) and dbg!(sg, counted_kmers, min_freq) (:274-278).
    kmers: (n, W) uint64 words with k given, or a MerCounts (filtered in place when min_freq
    is given, exactly like the reference mutates the caller's vector)."""
    if not isinstance(kmers_or_counts, MerCounts) and k is None:
        raise ValueError("Didn't provide a collection of kmers for the `kmers` parameter did ya?")
    ga = dbg_arrays(kmers_or_counts, k=k, min_freq=min_freq, ctx=ctx)
    base = len(sdg.nodes)
    if base != 0:
        raise ValueError("dbg! expects an empty graph")  # node ids in links are absolute
    for s in ga.node_strings():
        sdg.add_node_(s)
    for i, ls in enumerate(ga.link_lists()):
        sdg.links[i].extend(SDGLink(*t) for t in ls)
    return sdg


FwRv, RvFw = "FwRv", "RvFw"     # PairedReadOrientation of ReadDatastores (docs/src/man/assembly.md:96-100)
_PRSEQ_MAGIC = b"GGB200PRSEQ\x00\x00\x00\x00\x01"     # 16 bytes: everything behind it stays 8-byte aligned


class PairedReads:
    """The paired-read datastore `read_datastore(...)` makes (src/GenomeGraphs.jl:79-83: `PairedReads{DNAAlphabet{4}}(fwq, rvq,
    name, name, minlen, maxlen, insertlen, mode)`), held the way this path consumes it: the reads' 4-bit encoding (the
    encoding of LongSequence{DNAAlphabet{4}}, 16 bases per uint64) + base offsets, reads 2i / 2i+1 = R1 / R2 of pair i.
    The file `<name>.prseq` is this repository's own container for that payload (header, offsets, 4-bit words; little
    endian) -- ReadDatastores.jl is not vendored in the reference, so its on-disk layout is not restated; what is kept is
    the content and the documented filter: a pair is dropped when either read is shorter than minlen, longer reads are
    cut to maxlen (docs/src/man/assembly.md:86-88).  Opened stores are memory-mapped: the words go from the page cache
    straight into the pinned staging of gg_count_kmers_4bit / gg_dbg_from_reads_4bit."""

    def __init__(self, data4, base_offsets, name, minlen, maxlen, insertlen, orientation, path=None):
        self.data4, self.base_offsets = data4, base_offsets
        self.name, self.minlen, self.maxlen, self.insertlen, self.orientation, self.path = name, minlen, maxlen, insertlen, orientation, path

    def __len__(self):
        return len(self.base_offsets) - 1

    def n_pairs(self):
        return len(self) // 2

    def __getitem__(self, i):
        """read i as a string (N for every ambiguous code)"""
        if not 0 <= i < len(self):
            raise IndexError(i)
        lo, hi = int(self.base_offsets[i]), int(self.base_offsets[i + 1])
        idx = np.arange(lo, hi, dtype=np.uint64)
        nib = (np.asarray(self.data4)[(idx >> np.uint64(4)).astype(np.int64)] >> ((idx & np.uint64(15)) * np.uint64(4))) & np.uint64(15)
        lut = np.full(16, ord("N"), dtype=np.uint8)
        lut[[1, 2, 4, 8]] = [ord(x) for x in "ACGT"]
        return lut[nib.astype(np.int64)].tobytes().decode()

    def write(self, path):
        hdr = np.array([len(self), int(self.base_offsets[-1]), len(self.data4), self.minlen, self.maxlen, self.insertlen,
                        0 if self.orientation == FwRv else 1], dtype="<u8")
        nm = self.name.encode()
        with open(path, "wb") as f:
            f.write(_PRSEQ_MAGIC)
            f.write(hdr.tobytes())
            f.write(np.array([len(nm)], dtype="<u8").tobytes())
            f.write(nm + b"\x00" * ((-len(nm)) % 8))
            f.write(np.ascontiguousarray(self.base_offsets, dtype="<u8").tobytes())
            f.write(np.ascontiguousarray(self.data4, dtype="<u8").tobytes())
        self.path = str(path)
        return self

    @classmethod
    def open(cls, path):
        """open(PairedReads, "name.prseq"): header parsed, offsets and 4-bit words memory-mapped"""
        with open(path, "rb") as f:
            if f.read(len(_PRSEQ_MAGIC)) != _PRSEQ_MAGIC:
                raise ValueError("%s is not a paired read datastore of this package" % path)
            nreads, nbases, nwords, minlen, maxlen, insertlen, orient = (int(x) for x in np.frombuffer(f.read(56), dtype="<u8"))
            ln = int(np.frombuffer(f.read(8), dtype="<u8")[0])
            name = f.read(ln + ((-ln) % 8))[:ln].decode()
            pos = f.tell()
        offs = np.memmap(path, dtype="<u8", mode="r", offset=pos, shape=(nreads + 1,))
        data = np.memmap(path, dtype="<u8", mode="r", offset=pos + 8 * (nreads + 1), shape=(nwords,)) if nwords else np.zeros(0, dtype=np.uint64)
        if int(offs[-1]) != nbases or nwords != (nbases + 15) // 16:
            raise ValueError("%s: header and payload disagree" % path)
        return cls(data, offs, name, minlen, maxlen, insertlen, FwRv if orient == 0 else RvFw, str(path))


def read_datastore(r1file, r2file, name, minlen, maxlen, insertlen=0, mode=FwRv):
    """read_datastore(R1file, R2file, name, minlen, maxlen, insertlen, mode) (src/GenomeGraphs.jl:79-83): the two FASTQ
    files are read pair by pair, a pair is dropped when either read is shorter than `minlen`, reads longer than `maxlen`
    are cut (docs/src/man/assembly.md:86-88); the store is written to `<name>.prseq` and returned memory-mapped."""
    (s1, o1), (s2, o2) = read_sequences(r1file), read_sequences(r2file)
    if len(o1) != len(o2):
        raise ValueError("the two FASTQ files hold different numbers of reads")
    l1, l2 = np.diff(o1.astype(np.int64)), np.diff(o2.astype(np.int64))
    keep = (l1 >= minlen) & (l2 >= minlen)
    reads = []
    for i in np.nonzero(keep)[0]:
        reads.append(s1[int(o1[i]):int(o1[i]) + min(int(l1[i]), maxlen)].tobytes())
        reads.append(s2[int(o2[i]):int(o2[i]) + min(int(l2[i]), maxlen)].tobytes())
    data4, offs = pack_4bit(pack_reads(reads))
    path = name if str(name).endswith(".prseq") else str(name) + ".prseq"
    PairedReads(data4, offs, os.path.basename(str(name)), int(minlen), int(maxlen), int(insertlen), mode).write(path)
    return PairedReads.open(path)


def _sdg_arrays(sdg, k=0):
    """SequenceDistanceGraph (Python objects) -> GraphArrays; deleted nodes are empty"""
    seqs = ["" if n.deleted else n.seq for n in sdg.nodes]
    node_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    node_off[1:] = np.cumsum([len(x) for x in seqs], dtype=np.uint64)
    bases = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    link_row = np.zeros(len(seqs) + 1, dtype=np.uint64)
    link_row[1:] = np.cumsum([len(ls) for ls in sdg.links], dtype=np.uint64)
    tri = np.array([l.astuple() for ls in sdg.links for l in ls], dtype=np.int64).reshape(-1)
    return GraphArrays(k, node_off, bases, link_row, tri)


def _sdg_assign(sdg, ga):
    seqs = ga.node_strings()
    sdg.nodes[:] = [SDGNode(x, len(x) == 0) for x in seqs]     # empty_node: deleted, no sequence (src/Graphs.jl:86-88)
    sdg.links[:] = [[SDGLink(*t) for t in ls] for ls in ga.link_lists()]
    return sdg


def remove_tips_(sdg, min_size, ctx=None):
    """remove_tips!(sdg, min_size) (src/remove_tips.jl:2-30) on the device; mutates and returns `sdg` like the reference.
    Accepts a SequenceDistanceGraph (uploaded, edited, written back) or a DeviceGraph (a new DeviceGraph is returned)."""
    if isinstance(sdg, DeviceGraph):
        return sdg.remove_tips(min_size)[0]
    dg = DeviceGraph.from_arrays(_sdg_arrays(sdg), ctx=ctx)
    try:
        out, _, _ = dg.remove_tips(min_size)
        try:
            return _sdg_assign(sdg, out.arrays())
        finally:
            out.free()
    finally:
        dg.free()


def dbg_from_reads(reads, k, min_freq, ctx=None):
    """reads -> graph arrays in one device pipeline (count, filter, build)."""
    c = ctx or Context.default()
    if isinstance(reads, PairedReads):
        return dbg_from_reads_4bit(reads.data4, reads.base_offsets, k, min_freq, ctx=c)
    seq, offs = _as_reads(reads)
    gh = C.c_void_p()
    N.check(c.h, N.lib().gg_dbg_from_reads(c.h, _ptr(seq) if seq.size else None, _ptr(offs), len(offs) - 1, int(k), int(min_freq), C.byref(gh)))
    try:
        return _export_graph(c, gh)
    finally:
        N.lib().gg_graph_free(gh)


class WorkSpace:
    """A live WorkSpace for the construction path: a graph plus named, device-resident k-mer count stores.
    In the reference v0.2.0 the type is dead code (src/workspace/WorkSpace.jl is not included and names types
    that exist nowhere); only its field / accessor names are kept (sdg, mer_count_stores, add_mer_counts!,
    mer_counts)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or Context.default()
        self.sdg = SequenceDistanceGraph()
        self.mer_count_stores = {}

    def graph(self):
        return self.sdg

    def add_mer_counts_(self, k, mode, name, reads):
        """add_mer_counts!(ws, DNAMer{k}, mode, name): count the k-mers of `reads` on the GPU, keep the table"""
        self.mer_count_stores[name] = serial_mem(k, mode, ctx=self.ctx)(reads)
        return self

    def mer_counts(self, name):
        if name not in self.mer_count_stores:
            raise KeyError("WorkSpace has no mer count datastore datastore called %s" % name)
        return self.mer_count_stores[name]

    def dbg_(self, name, min_freq):
        """dbg!(ws, counted_kmers, min_freq): build the workspace graph from a stored count table"""
        self.sdg = dbg_(SequenceDistanceGraph(), self.mer_counts(name), min_freq, ctx=self.ctx)
        return self

    def remove_tips_(self, min_size):
        """remove_tips!(ws, min_size) (the commented-out WorkSpace form at src/remove_tips.jl:1-5, :28-29): edits the
        workspace graph on the device"""
        remove_tips_(self.sdg, min_size, ctx=self.ctx)
        return self

    def __repr__(self):
        lines = ["Graph Genome workspace", " Sequence distance graph (%d nodes)" % self.sdg.n_nodes(),
                 " Mer count stores: %d" % len(self.mer_count_stores)]
        lines += ["  %s: %d distinct %d-mers" % (n, len(c), c.k) for n, c in self.mer_count_stores.items()]
        return "\n".join(lines)


class GraphStrandPosition:
    """src/GraphIndexes.jl:14-17"""
    __slots__ = ("node", "position")

    def __init__(self, node, position):
        self.node, self.position = int(node), int(position)

    def __eq__(self, o):
        return (self.node, self.position) == (o.node, o.position)

    def __repr__(self):
        return "(Node: %d, Base position: %d)" % (self.node, self.position)


class UniqueMerIndex:
    """UniqueMerIndex{M}(graph) src/GraphIndexes.jl:62-102 (intended semantics)."""

    def __init__(self, graph, k, ctx=None):
        c = ctx or Context.default()
        seqs = [n.seq for n in graph.nodes] if isinstance(graph, SequenceDistanceGraph) else list(graph)
        seq, offs = pack_reads(seqs)
        h = C.c_void_p()
        N.check(c.h, N.lib().gg_unique_mer_index(c.h, _ptr(seq) if seq.size else None, _ptr(offs), len(seqs), int(k), C.byref(h)))
        try:
            nu, nn, kk, w = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
            N.check(c.h, N.lib().gg_umi_sizes(h, C.byref(nu), C.byref(nn), C.byref(kk), C.byref(w)))
            self.k, self.words = kk.value, w.value
            self.keys = np.zeros((nu.value, w.value), dtype=np.uint64)
            self.node = np.zeros(nu.value, dtype=np.int64)
            self.pos = np.zeros(nu.value, dtype=np.uint32)
            self.unique_mers_per_node = np.zeros(nn.value, dtype=np.uint64)
            self.total_mers_per_node = np.zeros(nn.value, dtype=np.uint64)
            N.check(c.h, N.lib().gg_umi_export(h, _ptr(self.keys) if nu.value else None, _ptr(self.node) if nu.value else None,
                                               _ptr(self.pos) if nu.value else None,
                                               _ptr(self.unique_mers_per_node) if nn.value else None,
                                               _ptr(self.total_mers_per_node) if nn.value else None))
        finally:
            N.lib().gg_umi_free(h)
        self._dict = None

    @property
    def mer_to_graphposition(self):
        if self._dict is None:
            self._dict = {tuple(int(x) for x in self.keys[i]): GraphStrandPosition(self.node[i], self.pos[i]) for i in range(len(self.keys))}
        return self._dict

    def find_unique_mer(self, mer_words):
        """src/GraphIndexes.jl:37-41 -> (exists, GraphStrandPosition)"""
        key = tuple(int(x) for x in np.atleast_1d(np.asarray(mer_words, dtype=np.uint64)))
        pos = self.mer_to_graphposition.get(key, GraphStrandPosition(0, 0))
        return pos != GraphStrandPosition(0, 0), pos

    def isunmappable(self, nodeid):
        """src/GraphIndexes.jl:33-35"""
        return self.unique_mers_per_node[abs(int(nodeid)) - 1] == 0


def synth_reads(seed, genome_len, nreads, read_len, paired=False, frag_len=0, err_ppm=0, n_ppm=0, first_read=0):
    """Deterministic synthetic reads (host generator) -> (seq_bytes, offsets).  `first_read` selects
    reads first_read .. first_read + nreads of the same job (a rank's shard)."""
    out = np.zeros(int(nreads) * int(read_len), dtype=np.uint8)
    code = N.lib().gg_synth_reads_range_host(int(seed), int(genome_len), int(first_read), int(nreads), int(read_len), int(bool(paired)),
                                             int(frag_len), int(err_ppm), int(n_ppm), _ptr(out))
    if code != N.GG_OK:
        raise ValueError("bad synthetic read parameters")
    offs = np.arange(int(nreads) + 1, dtype=np.uint64) * np.uint64(read_len)
    return out, offs


def synth_genome(seed, genome_len):
    out = np.zeros(int(genome_len), dtype=np.uint8)
    N.check(None, N.lib().gg_synth_genome_host(int(seed), int(genome_len), _ptr(out)))
    return out
