/*
 * gg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the de Bruijn graph construction path of
 * BioJulia/GenomeGraphs.jl v0.2.0: reads -> canonical k-mers -> sort + run-length
 * counts -> min-count filter -> unitig nodes -> (k-1)-overlap links, plus the
 * UniqueMerIndex core.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (genomegraphs.jl_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED (see oracle_impl.h header): no reference tests or goldens exist and
 * the Julia reference cannot run here.  Half of the path (k-mer types, counting) lives
 * in un-vendored dependencies: BioSequences.jl compat "2", KmerAnalysis.jl compat
 * "0.4.4", ReadDatastores.jl "0.4", FASTX.jl "1.1" (Project.toml:12-17, no Manifest);
 * their behaviour is restated from the published semantics and anchored on the
 * reference's own call sites (src/dbg.jl, src/GraphIndexes.jl, docs/src/man/assembly.md).
 *
 * Assumptions written down:
 *  (1) 2-bit code A0 C1 G2 T3, first base most significant; order = unsigned order.
 *  (2) canonical = min(x, rc(x)); iscanonical = x <= rc(x).
 *  (3) Mer(::Integer) masks to 2k bits (needed by src/dbg.jl:32-33).
 *  (4) windows with a non-ACGT symbol are skipped; lower-case acgt == upper case.
 *  (5) counts are kept as u32 here; the UInt8-saturating MerCount view is min(c, 255).
 *  (6) canonical!(::LongSequence) keeps the sequence if seq <= revcomp(seq) lexicographically.
 *  (7) sort! on (mer, nid) tuples orders by mer then signed nid (src/dbg.jl:219-220).
 *  (8) reads shorter than k yield nothing.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int or_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

static inline uint64_t or_rev2(uint64_t x) { /* reverse the order of the 32 2-bit groups */
    x = __builtin_bswap64(x);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    return x;
}

/* growable byte vector (a LongSequence stand-in holding codes 0..3) */
typedef struct { uint8_t *d; uint64_t n, cap; } or_bytes;
static void or_bytes_init(or_bytes *b) { b->d = NULL; b->n = 0; b->cap = 0; }
static void or_bytes_free(or_bytes *b) { free(b->d); b->d = NULL; b->n = b->cap = 0; }
static void or_bytes_push(or_bytes *b, uint8_t v) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 64; b->d = (uint8_t *)realloc(b->d, b->cap); }
    b->d[b->n++] = v;
}
/* canonical!(s): BioSequences iscanonical compares s with its reverse complement
 * position by position (recalled); ties keep s.  Use site src/dbg.jl:156,179. */
static void or_seq_canonical(or_bytes *s) {
    uint64_t n = s->n;
    int keep = 1;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t f = s->d[i], r = (uint8_t)(3 - s->d[n - 1 - i]);
        if (f < r) { keep = 1; break; }
        if (r < f) { keep = 0; break; }
    }
    if (keep) return;
    for (uint64_t i = 0, j = n - 1; i < j; ++i, --j) {
        uint8_t a = s->d[i], b = s->d[j];
        s->d[i] = (uint8_t)(3 - b); s->d[j] = (uint8_t)(3 - a);
    }
    if (n & 1) s->d[n / 2] = (uint8_t)(3 - s->d[n / 2]);
}

/* SequenceDistanceGraph stand-in: src/Graphs.jl:232-235 (nodes, links) */
typedef struct { int64_t owner, src, dst, dist; } or_push;
typedef struct {
    uint64_t n_nodes;
    uint64_t *node_off; uint64_t node_cap;   /* n_nodes + 1 offsets into bases */
    uint8_t *bases; uint64_t nbases, bases_cap;
    or_push *push; uint64_t npush, push_cap;  /* link entries in push order */
    uint64_t *link_row;                       /* CSR row offsets, n_nodes + 1 (after finalize) */
    int64_t *link_tri;                        /* 3 * npush : src, dst, dist grouped by owner */
} or_graph;

static or_graph *or_graph_new(void) {
    or_graph *g = (or_graph *)calloc(1, sizeof(or_graph));
    g->node_cap = 16; g->node_off = (uint64_t *)malloc(g->node_cap * sizeof(uint64_t));
    g->node_off[0] = 0;
    return g;
}
/* add_node! src/Graphs.jl:418-422,442-444: push node, IDs are 1-based and dense */
static void or_graph_add_node(or_graph *g, const or_bytes *s) {
    if (g->n_nodes + 2 > g->node_cap) {
        g->node_cap *= 2;
        g->node_off = (uint64_t *)realloc(g->node_off, g->node_cap * sizeof(uint64_t));
    }
    if (g->nbases + s->n > g->bases_cap) {
        g->bases_cap = (g->nbases + s->n) * 2 + 64;
        g->bases = (uint8_t *)realloc(g->bases, g->bases_cap);
    }
    memcpy(g->bases + g->nbases, s->d, s->n);
    g->nbases += s->n;
    g->n_nodes += 1;
    g->node_off[g->n_nodes] = g->nbases;
}
/* add_link! src/Graphs.jl:473-476: push (src,dst,d) to links[|src|] and (dst,src,d) to links[|dst|] */
static void or_graph_add_link(or_graph *g, int64_t src, int64_t dst, int64_t dist) {
    if (g->npush + 2 > g->push_cap) {
        g->push_cap = g->push_cap ? g->push_cap * 2 : 64;
        g->push = (or_push *)realloc(g->push, g->push_cap * sizeof(or_push));
    }
    or_push a = { src < 0 ? -src : src, src, dst, dist };
    or_push b = { dst < 0 ? -dst : dst, dst, src, dist };
    g->push[g->npush++] = a;
    g->push[g->npush++] = b;
}
static void or_graph_finalize_links(or_graph *g) {
    uint64_t nn = g->n_nodes;
    g->link_row = (uint64_t *)calloc(nn + 2, sizeof(uint64_t));
    g->link_tri = (int64_t *)malloc((g->npush ? g->npush : 1) * 3 * sizeof(int64_t));
    for (uint64_t i = 0; i < g->npush; ++i) g->link_row[g->push[i].owner]++;
    /* link_row[o] currently = count of owner o (1-based); exclusive scan into row starts */
    uint64_t sum = 0;
    for (uint64_t o = 0; o <= nn; ++o) { uint64_t c = g->link_row[o]; g->link_row[o] = sum; sum += c; }
    /* now link_row[o] = start of owner o (for o >= 1), link_row[0] = 0 with zero entries */
    uint64_t *cur = (uint64_t *)malloc((nn + 2) * sizeof(uint64_t));
    memcpy(cur, g->link_row, (nn + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < g->npush; ++i) {
        uint64_t p = cur[g->push[i].owner]++;
        g->link_tri[3 * p] = g->push[i].src;
        g->link_tri[3 * p + 1] = g->push[i].dst;
        g->link_tri[3 * p + 2] = g->push[i].dist;
    }
    free(cur);
    /* shift so that row offsets are indexed by node-1: row[i] = start of node i+1 */
    for (uint64_t o = 0; o < nn; ++o) g->link_row[o] = g->link_row[o + 1];
    g->link_row[nn] = g->npush;
}

#define W 1
#define SUF _w1
#include "oracle_impl.h"
#undef W
#undef SUF
#define W 2
#define SUF _w2
#include "oracle_impl.h"
#undef W
#undef SUF
#define W 4
#define SUF _w4
#include "oracle_impl.h"
#undef W
#undef SUF

/* ------------------------------- exported C API (ctypes) ----------------------------- */
#define EXPORT __attribute__((visibility("default")))

/* words per k-mer: 1 (k<=32), 2 (k<=64), 4 (k<=128); 0 if unsupported */
EXPORT int ggo_words(int k) { return k < 1 ? 0 : k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 0; }

#define DISPATCH(k, call1, call2, call4, bad) \
    switch (ggo_words(k)) { case 1: call1; break; case 2: call2; break; case 4: call4; break; default: bad; }

EXPORT uint64_t ggo_extract(const uint8_t *seq, const uint64_t *offs, uint64_t nreads, int k,
                            int canonical, uint64_t *out_words) {
    DISPATCH(k,
        return or_extract_w1(seq, offs, nreads, k, canonical, out_words, NULL, NULL, NULL),
        return or_extract_w2(seq, offs, nreads, k, canonical, out_words, NULL, NULL, NULL),
        return or_extract_w4(seq, offs, nreads, k, canonical, out_words, NULL, NULL, NULL),
        return 0)
    return 0;
}

/* sorts words[n*W] in place, unique keys to the front, counts[u]; returns u */
EXPORT uint64_t ggo_sort_count(uint64_t *words, uint64_t n, int k, uint32_t *counts, int threads) {
    DISPATCH(k,
        return or_sort_count_w1(words, n, k, counts, threads),
        return or_sort_count_w2(words, n, k, counts, threads),
        return or_sort_count_w4(words, n, k, counts, threads),
        return 0)
    return 0;
}

/* a6: filter!(x -> freq(x) >= min_freq) src/dbg.jl:275 then [mer(x) ...] :276.
 * u8 != 0 applies the UInt8-saturated MerCount view: freq = min(count, 255). */
EXPORT uint64_t ggo_filter(const uint64_t *words, const uint32_t *counts, uint64_t u, int k,
                           uint64_t min_freq, int u8, uint64_t *out_words) {
    int w = ggo_words(k);
    uint64_t n = 0;
    for (uint64_t i = 0; i < u; ++i) {
        uint64_t f = counts[i];
        if (u8 && f > 255) f = 255;
        if (f >= min_freq) {
            for (int j = 0; j < w; ++j) out_words[n * (uint64_t)w + (uint64_t)j] = words[i * (uint64_t)w + (uint64_t)j];
            ++n;
        }
    }
    return n;
}

EXPORT void *ggo_dbg_build(uint64_t *words, uint64_t n, int k) {
    or_graph *g = or_graph_new();
    DISPATCH(k, or_dbg_w1(g, words, n, k), or_dbg_w2(g, words, n, k), or_dbg_w4(g, words, n, k), return NULL)
    return g;
}
EXPORT void ggo_graph_sizes(void *h, uint64_t *n_nodes, uint64_t *total_bases, uint64_t *n_link_entries) {
    or_graph *g = (or_graph *)h;
    *n_nodes = g->n_nodes; *total_bases = g->nbases; *n_link_entries = g->npush;
}
/* bases_ascii: 'A','C','G','T'; node_offsets: n_nodes+1; link_row_offsets: n_nodes+1;
 * link_triples: 3 * n_link_entries (src, dst, dist) in per-node push order. */
EXPORT void ggo_graph_export(void *h, uint64_t *node_offsets, uint8_t *bases_ascii,
                             uint64_t *link_row_offsets, int64_t *link_triples) {
    or_graph *g = (or_graph *)h;
    static const char A[4] = { 'A', 'C', 'G', 'T' };
    memcpy(node_offsets, g->node_off, (g->n_nodes + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < g->nbases; ++i) bases_ascii[i] = (uint8_t)A[g->bases[i]];
    memcpy(link_row_offsets, g->link_row, (g->n_nodes + 1) * sizeof(uint64_t));
    memcpy(link_triples, g->link_tri, g->npush * 3 * sizeof(int64_t));
}
EXPORT void ggo_graph_free(void *h) {
    or_graph *g = (or_graph *)h;
    if (!g) return;
    free(g->node_off); free(g->bases); free(g->push); free(g->link_row); free(g->link_tri); free(g);
}

EXPORT uint64_t ggo_unique_mer_index(const uint8_t *seq, const uint64_t *offs, uint64_t n_nodes, int k,
                                     uint64_t *keys_out, int64_t *node_out, uint32_t *pos_out,
                                     uint64_t *unique_per_node, uint64_t *total_per_node) {
    DISPATCH(k,
        return or_unique_mer_index_w1(seq, offs, n_nodes, k, keys_out, node_out, pos_out, unique_per_node, total_per_node),
        return or_unique_mer_index_w2(seq, offs, n_nodes, k, keys_out, node_out, pos_out, unique_per_node, total_per_node),
        return or_unique_mer_index_w4(seq, offs, n_nodes, k, keys_out, node_out, pos_out, unique_per_node, total_per_node),
        return 0)
    return 0;
}

EXPORT int ggo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- deterministic synthetic reads ---------------------------------
 * C restatement of the counter-based generator in genomegraphs.jl_b200/csrc/synth.cuh, so that
 * the CPU legs of bench.py (cpu_baseline, --impl reference) build their input without loading
 * the product library.  Same bytes as gg_synth_reads_range_host / _dev (tests/test_oracle.py). */
static inline uint64_t osy_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t osy_key(uint64_t seed, uint64_t stream) { return osy_mix(seed + stream * 0x9E3779B97F4A7C15ULL); }
static inline uint64_t osy_rnd(uint64_t key, uint64_t i) { return osy_mix(key + i * 0xD6E8FEB86659FD93ULL); }

EXPORT int ggo_synth_reads_range(uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t nreads, uint32_t read_len,
                                 int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *out, int threads) {
    const uint64_t F = paired ? frag_len : read_len;
    if (!out || read_len == 0 || F < read_len || genome_len < F || (paired && (first_read & 1))) return -1;
    const uint64_t gkey = osy_key(seed, 1), skey = osy_key(seed, 2), dkey = osy_key(seed, 3), ekey = osy_key(seed, 4), nkey = osy_key(seed, 5);
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long rr = 0; rr < (long long)nreads; ++rr) {
        const uint64_t r = first_read + (uint64_t)rr;
        const uint64_t j = paired ? (r >> 1) : r;
        const uint32_t mate = paired ? (uint32_t)(r & 1) : 0u;
        const uint64_t start = (uint64_t)(((unsigned __int128)osy_rnd(skey, j) * (genome_len - F + 1)) >> 64);
        const uint32_t strand = (uint32_t)(osy_rnd(dkey, j) & 1u);
        for (uint32_t p = 0; p < read_len; ++p) {
            const uint64_t t = mate ? (F - 1 - p) : p;
            const uint64_t g = strand ? (start + F - 1 - t) : (start + t);
            uint32_t b = (uint32_t)(osy_rnd(gkey, g >> 5) >> (2 * (g & 31))) & 3u;
            if (mate ^ strand) b = 3u - b;
            const uint64_t idx = r * (uint64_t)read_len + p;
            if (err_ppm) {
                const uint64_t e = osy_rnd(ekey, idx);
                if ((uint32_t)(((uint64_t)(uint32_t)e * 1000000ULL) >> 32) < err_ppm) b = (b + 1u + (uint32_t)((e >> 32) % 3u)) & 3u;
            }
            uint8_t c = (uint8_t)(0x54474341u >> (8 * b));
            if (n_ppm) {
                const uint64_t e = osy_rnd(nkey, idx);
                if ((uint32_t)(((uint64_t)(uint32_t)e * 1000000ULL) >> 32) < n_ppm) c = (uint8_t)'N';
            }
            out[(uint64_t)rr * read_len + p] = c;
        }
    }
    return 0;
}
