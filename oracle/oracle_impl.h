/*
 * oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * W-templated body of the CPU oracle (included three times by gg_oracle.c with
 * W = 1, 2, 4 and SUF = _w1, _w2, _w4).  It restates, as plainly as possible,
 * the reference algorithm of BioJulia/GenomeGraphs.jl v0.2.0 for the de Bruijn
 * graph construction path.  Each function cites the reference file:line it
 * follows (paths relative to /root/reference).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path (test/runtests.jl:1-7 is an empty module) and Julia is not available
 * in this environment, so the oracle is pinned only against the hand-traced
 * vector G1, the documented example of
 * docs/src/DeBruijnGraph.md:57-76 and an independent string-based Python
 * transliteration (tests/pyref.py).
 *
 * k-mer value layout (BioSequences.jl 2.x Mer / BigMer, recalled -- not in repo):
 *   2 bits per base, A=0 C=1 G=2 T=3, first base in the MOST significant of the
 *   2k used bits, value right-aligned in a W*64-bit unsigned integer.  Word 0 of
 *   a KM is the most significant word.  Ordering = unsigned integer order.
 */

#define OR_CAT_(a, b) a##b
#define OR_CAT(a, b) OR_CAT_(a, b)
#define FN(name) OR_CAT(name, SUF)
#define KM FN(km)

typedef struct { uint64_t w[W]; } KM;

static inline int FN(km_cmp)(const KM *a, const KM *b) {
    for (int i = 0; i < W; ++i)
        if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
static inline int FN(km_eq)(const KM *a, const KM *b) { return FN(km_cmp)(a, b) == 0; }

static inline KM FN(km_zero)(void) { KM r; for (int i = 0; i < W; ++i) r.w[i] = 0; return r; }

/* low 2k bits set */
static inline KM FN(km_mask)(int k) {
    KM r;
    int bits = 2 * k;
    for (int i = W - 1; i >= 0; --i) {
        if (bits >= 64) { r.w[i] = ~(uint64_t)0; bits -= 64; }
        else if (bits > 0) { r.w[i] = (((uint64_t)1) << bits) - 1; bits = 0; }
        else r.w[i] = 0;
    }
    return r;
}
static inline KM FN(km_and)(KM a, KM b) { for (int i = 0; i < W; ++i) a.w[i] &= b.w[i]; return a; }

static inline KM FN(km_shl2)(KM x) {
    KM r;
    for (int i = 0; i < W; ++i)
        r.w[i] = (x.w[i] << 2) | ((i + 1 < W) ? (x.w[i + 1] >> 62) : 0);
    return r;
}
static inline KM FN(km_shr2)(KM x) {
    KM r;
    for (int i = W - 1; i >= 0; --i)
        r.w[i] = (x.w[i] >> 2) | ((i > 0) ? (x.w[i - 1] << 62) : 0);
    return r;
}
/* OR the 2-bit code b in at (even) bit position pos counted from the LSB of the whole value */
static inline KM FN(km_or2)(KM x, unsigned b, int pos) {
    x.w[W - 1 - pos / 64] |= ((uint64_t)b) << (pos % 64);
    return x;
}
static inline unsigned FN(km_get2)(const KM *x, int pos) {
    return (unsigned)((x->w[W - 1 - pos / 64] >> (pos % 64)) & 3u);
}
/* logical right shift of the whole W*64-bit value by s bits, 0 <= s < 64*W */
static inline KM FN(km_shr)(KM x, int s) {
    KM r;
    int ws = s / 64, bs = s % 64;
    for (int i = W - 1; i >= 0; --i) {
        int src = i - ws;
        uint64_t lo = (src >= 0) ? x.w[src] : 0;
        uint64_t hi = (src - 1 >= 0) ? x.w[src - 1] : 0;
        r.w[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
    return r;
}
/* BioSequences reverse_complement(::Mer) (recalled): complement = 3 - code */
static inline KM FN(km_rc)(KM x, int k) {
    KM t;
    for (int i = 0; i < W; ++i) t.w[W - 1 - i] = or_rev2(~x.w[i]);
    return FN(km_shr)(t, 64 * W - 2 * k);
}
/* BioSequences canonical(x) = min(x, rc(x)) (recalled); use sites src/dbg.jl:73,84 */
static inline KM FN(km_canonical)(KM x, int k) {
    KM r = FN(km_rc)(x, k);
    return FN(km_cmp)(&x, &r) <= 0 ? x : r;
}
/* iscanonical(x) = x <= rc(x); use sites src/dbg.jl:203,211 and src/GraphIndexes.jl:80 */
static inline int FN(km_iscanonical)(KM x, int k) {
    KM r = FN(km_rc)(x, k);
    return FN(km_cmp)(&x, &r) <= 0;
}

/* ---- a2/a3: window iteration + canonicalisation ------------------------------------
 * each(M, seq) of BioSequences (recalled) -- in-repo statement src/GraphIndexes.jl:78-82.
 * Windows containing a non-ACGT symbol are skipped (the run restarts after it); reads
 * shorter than k yield nothing.  If out == NULL only counts.  Returns #windows emitted.
 * pos_out (optional): 1-based window start inside its read; read_out: 0-based read idx;
 * strand_out: 1 if fw <= rc (forward is canonical) else 0.
 */
static uint64_t FN(or_extract)(const uint8_t *seq, const uint64_t *offs, uint64_t nreads, int k,
                               int canonical, uint64_t *out, uint32_t *pos_out,
                               uint64_t *read_out, uint8_t *strand_out) {
    uint64_t n = 0;
    KM mask = FN(km_mask)(k);
    for (uint64_t r = 0; r < nreads; ++r) {
        KM fw = FN(km_zero)(), rc = FN(km_zero)();
        int run = 0;
        for (uint64_t p = offs[r]; p < offs[r + 1]; ++p) {
            int c = or_code(seq[p]);
            if (c < 0) { run = 0; fw = FN(km_zero)(); rc = FN(km_zero)(); continue; }
            fw = FN(km_and)(FN(km_or2)(FN(km_shl2)(fw), (unsigned)c, 0), mask);
            rc = FN(km_or2)(FN(km_shr2)(rc), (unsigned)(3 - c), 2 * (k - 1));
            if (++run >= k) {
                if (out) {
                    int fwcanon = FN(km_cmp)(&fw, &rc) <= 0;
                    const KM *e = (canonical && !fwcanon) ? &rc : &fw;
                    for (int i = 0; i < W; ++i) out[n * W + i] = e->w[i];
                    if (pos_out) pos_out[n] = (uint32_t)(p - offs[r] + 1 - (uint64_t)k + 1);
                    if (read_out) read_out[n] = r;
                    if (strand_out) strand_out[n] = (uint8_t)fwcanon;
                }
                ++n;
            }
        }
    }
    return n;
}

/* ---- a4: collect -> sort! -> run-length collapse (KmerAnalysis serial_mem, recalled;
 * same pattern in-repo at src/GraphIndexes.jl:88-90,43-60).  LSD byte radix sort (any
 * correct sort gives the same multiset order); tmp has the same size as keys.
 */
static void FN(or_radix_sort)(KM *keys, KM *tmp, uint64_t n, int k) {
    int nbytes = (2 * k + 7) / 8;
    KM *src = keys, *dst = tmp;
    uint64_t *hist = (uint64_t *)malloc(256 * sizeof(uint64_t));
    for (int b = 0; b < nbytes; ++b) {
        int word = W - 1 - b / 8, sh = (b % 8) * 8;
        memset(hist, 0, 256 * sizeof(uint64_t));
        for (uint64_t i = 0; i < n; ++i) hist[(src[i].w[word] >> sh) & 0xFF]++;
        int trivial = 0;
        for (int d = 0; d < 256; ++d) if (hist[d] == n) trivial = 1;
        if (trivial) continue;
        uint64_t sum = 0;
        for (int d = 0; d < 256; ++d) { uint64_t c = hist[d]; hist[d] = sum; sum += c; }
        for (uint64_t i = 0; i < n; ++i) dst[hist[(src[i].w[word] >> sh) & 0xFF]++] = src[i];
        KM *t = src; src = dst; dst = t;
    }
    if (src != keys) memcpy(keys, src, n * sizeof(KM));
    free(hist);
}

/* sort keys (in place) and collapse runs: unique keys to the front of keys[], u32 counts */
static uint64_t FN(or_sort_count)(uint64_t *words, uint64_t n, int k, uint32_t *counts, int threads) {
    KM *keys = (KM *)words;
    if (n == 0) return 0;
    KM *tmp = (KM *)malloc(n * sizeof(KM));
    if (threads <= 1) {
        FN(or_radix_sort)(keys, tmp, n, k);
    } else {
        /* stronger CPU bar: split on the top byte of the 2k-bit key, sort buckets in parallel */
        int topbits = 2 * k < 8 ? 2 * k : 8;
        int nb = 1 << topbits, sh = 2 * k - topbits;
        uint64_t *start = (uint64_t *)calloc((size_t)nb + 1, sizeof(uint64_t));
        uint64_t *cur = (uint64_t *)malloc((size_t)nb * sizeof(uint64_t));
        for (uint64_t i = 0; i < n; ++i) {
            KM s = FN(km_shr)(keys[i], sh);
            start[(s.w[W - 1] & (uint64_t)(nb - 1)) + 1]++;
        }
        for (int d = 0; d < nb; ++d) start[d + 1] += start[d];
        for (int d = 0; d < nb; ++d) cur[d] = start[d];
        for (uint64_t i = 0; i < n; ++i) {
            KM s = FN(km_shr)(keys[i], sh);
            tmp[cur[s.w[W - 1] & (uint64_t)(nb - 1)]++] = keys[i];
        }
        memcpy(keys, tmp, n * sizeof(KM));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int d = 0; d < nb; ++d)
            FN(or_radix_sort)(keys + start[d], tmp + start[d], start[d + 1] - start[d], k);
        free(start); free(cur);
    }
    free(tmp);
    uint64_t u = 0, i = 0;
    while (i < n) {
        uint64_t j = i + 1;
        while (j < n && FN(km_eq)(&keys[j], &keys[i])) ++j;
        keys[u] = keys[i];
        counts[u] = (j - i > 0xFFFFFFFFull) ? 0xFFFFFFFFu : (uint32_t)(j - i);
        ++u;
        i = j;
    }
    return u;
}

/* ---- a7: src/dbg.jl:30-34 and :36-42 ------------------------------------------------ */
static inline void FN(or_fw_neighbours)(KM x, int k, KM out[4]) {
    KM base = FN(km_and)(FN(km_shl2)(x), FN(km_mask)(k)); /* M(::UInt) masks to 2k bits */
    for (unsigned b = 0; b < 4; ++b) out[b] = FN(km_or2)(base, b, 0);
}
static inline void FN(or_bw_neighbours)(KM x, int k, KM out[4]) {
    KM base = FN(km_shr2)(x);
    for (unsigned b = 0; b < 4; ++b) out[b] = FN(km_or2)(base, b, 2 * (k - 1));
}

typedef struct { KM kmer; uint64_t idx; } FN(kidx); /* src/dbg.jl:23-26 */

/* searchsortedfirst (0-based lower bound) */
static inline uint64_t FN(or_lower_bound)(const KM *list, uint64_t n, const KM *c) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (FN(km_cmp)(&list[mid], c) < 0) lo = mid + 1; else hi = mid;
    }
    return lo;
}
/* a8: get_fw_idxs! src/dbg.jl:81-90 / get_bw_idxs! :70-79 (dir 0 = fw, 1 = bw) */
static int FN(or_get_idxs)(FN(kidx) out[4], KM x, const KM *list, uint64_t n, int k, int dir) {
    KM nb[4];
    int cnt = 0;
    if (dir == 0) FN(or_fw_neighbours)(x, k, nb); else FN(or_bw_neighbours)(x, k, nb);
    for (int b = 0; b < 4; ++b) {
        KM c = FN(km_canonical)(nb[b], k);
        uint64_t i = FN(or_lower_bound)(list, n, &c);
        if (i >= n) i = n - 1; /* min(searchsortedfirst(..), length(..)) */
        if (FN(km_eq)(&list[i], &c)) { out[cnt].kmer = nb[b]; out[cnt].idx = i; ++cnt; }
    }
    return cnt;
}
/* a9: is_end_fw src/dbg.jl:57-68, is_end_bw :44-55 */
static int FN(or_is_end)(KM x, const KM *list, uint64_t n, int k, int dir) {
    FN(kidx) next[4];
    if (FN(or_get_idxs)(next, x, list, n, k, dir) != 1) return 1;
    KM p = next[0].kmer;
    if (FN(or_get_idxs)(next, p, list, n, k, !dir) != 1) return 1;
    return 0;
}

/* LongSequence(kmer): append the k bases (codes 0..3) of x to the byte vector */
static void FN(or_seq_from_kmer)(or_bytes *s, KM x, int k) {
    for (int j = k - 1; j >= 0; --j) or_bytes_push(s, (uint8_t)FN(km_get2)(&x, 2 * j));
}
static KM FN(or_kmer_from_codes)(const uint8_t *codes, int k) {
    KM x = FN(km_zero)();
    for (int j = 0; j < k; ++j) x = FN(km_or2)(FN(km_shl2)(x), codes[j], 0);
    return x;
}

/* a10: build_unitigs! src/dbg.jl:94-185 (list must be sorted, canonical, duplicate-free) */
static void FN(or_build_unitigs)(or_graph *g, const KM *list, uint64_t n, int k) {
    uint8_t *used = (uint8_t *)calloc(n ? n : 1, 1);
    FN(kidx) fwn[4];
    for (uint64_t si = 0; si < n; ++si) {
        if (used[si]) continue;                                       /* :109-112 */
        KM start = list[si];
        int end_bw = FN(or_is_end)(start, list, n, k, 1);             /* :116 */
        int end_fw = FN(or_is_end)(start, list, n, k, 0);             /* :117 */
        if (!end_bw && !end_fw) continue;                             /* :119-122 */
        or_bytes s; or_bytes_init(&s);
        if (end_bw && end_fw) {                                       /* :124-128 */
            FN(or_seq_from_kmer)(&s, start, k);
            used[si] = 1;
        } else {
            KM cur = start;                                           /* :132-137 */
            used[si] = 1;
            if (end_fw) { cur = FN(km_rc)(start, k); end_fw = end_bw; }
            FN(or_seq_from_kmer)(&s, cur, k);                         /* :140 */
            while (!end_fw) {                                         /* :142-154 */
                FN(or_get_idxs)(fwn, cur, list, n, k, 0);
                cur = fwn[0].kmer;
                if (used[fwn[0].idx]) break;                          /* :147-150 */
                used[fwn[0].idx] = 1;
                or_bytes_push(&s, (uint8_t)FN(km_get2)(&cur, 0));     /* push!(s, last(cur)) */
                end_fw = FN(or_is_end)(cur, list, n, k, 0);
            }
        }
        or_seq_canonical(&s);                                         /* canonical!(s) :156 */
        or_graph_add_node(g, &s);
        or_bytes_free(&s);
    }
    /* perfect circles src/dbg.jl:160-181 */
    for (uint64_t ni = 0; ni < n; ++ni) {
        if (used[ni]) continue;
        KM start = list[ni], cur = start;
        used[ni] = 1;
        or_bytes s; or_bytes_init(&s);
        FN(or_seq_from_kmer)(&s, cur, k);
        for (;;) {
            FN(or_get_idxs)(fwn, cur, list, n, k, 0);
            cur = fwn[0].kmer;
            if (FN(km_eq)(&cur, &start)) break;
            used[fwn[0].idx] = 1;
            or_bytes_push(&s, (uint8_t)FN(km_get2)(&cur, 0));
        }
        or_seq_canonical(&s);
        or_graph_add_node(g, &s);
        or_bytes_free(&s);
    }
    free(used);
}

typedef struct { KM mer; int64_t nid; } FN(ovl);
static int FN(or_ovl_cmp)(const void *a, const void *b) { /* tuple order (mer, nid) :219-220 */
    const FN(ovl) *x = (const FN(ovl) *)a, *y = (const FN(ovl) *)b;
    int c = FN(km_cmp)(&x->mer, &y->mer);
    if (c) return c;
    return x->nid < y->nid ? -1 : (x->nid > y->nid ? 1 : 0);
}

/* a12 + a13: find_unitig_overlaps src/dbg.jl:190-222, connect_unitigs_by_overlaps! :224-240 */
static void FN(or_connect)(or_graph *g, int k) {
    uint64_t nn = g->n_nodes;
    int k1 = k - 1;
    FN(ovl) *in = (FN(ovl) *)malloc((2 * nn + 1) * sizeof(FN(ovl)));
    FN(ovl) *out = (FN(ovl) *)malloc((2 * nn + 1) * sizeof(FN(ovl)));
    uint64_t ni = 0, no = 0;
    for (uint64_t id = 1; id <= nn; ++id) {
        const uint8_t *seq = g->bases + g->node_off[id - 1];
        uint64_t len = g->node_off[id] - g->node_off[id - 1];
        KM first = FN(or_kmer_from_codes)(seq, k1);
        if (FN(km_iscanonical)(first, k1)) { in[ni].mer = first; in[ni].nid = (int64_t)id; ++ni; }
        else { out[no].mer = FN(km_rc)(first, k1); out[no].nid = (int64_t)id; ++no; }
        KM last = FN(or_kmer_from_codes)(seq + len - (uint64_t)k1, k1);
        if (FN(km_iscanonical)(last, k1)) { out[no].mer = last; out[no].nid = -(int64_t)id; ++no; }
        else { in[ni].mer = FN(km_rc)(last, k1); in[ni].nid = -(int64_t)id; ++ni; }
    }
    qsort(in, ni, sizeof(FN(ovl)), FN(or_ovl_cmp));
    qsort(out, no, sizeof(FN(ovl)), FN(or_ovl_cmp));
    uint64_t next_out = 0;
    for (uint64_t i = 0; i < ni; ++i) {
        while (next_out < no && FN(km_cmp)(&out[next_out].mer, &in[i].mer) < 0) ++next_out;
        uint64_t o = next_out;
        while (o < no && FN(km_eq)(&out[o].mer, &in[i].mer)) {
            or_graph_add_link(g, in[i].nid, out[o].nid, -(int64_t)k + 1);
            ++o;
        }
    }
    free(in); free(out);
}

/* a15: dbg!(sdg, kmers) src/dbg.jl:261-272.  Sorts the list if needed (:98-100). */
static void FN(or_dbg)(or_graph *g, uint64_t *words, uint64_t n, int k) {
    KM *list = (KM *)words;
    int sorted = 1;
    for (uint64_t i = 1; i < n; ++i) if (FN(km_cmp)(&list[i - 1], &list[i]) > 0) { sorted = 0; break; }
    if (!sorted) {
        KM *tmp = (KM *)malloc(n * sizeof(KM));
        FN(or_radix_sort)(list, tmp, n, k);
        free(tmp);
    }
    FN(or_build_unitigs)(g, list, n, k);
    if (g->n_nodes > 1) FN(or_connect)(g, k);                         /* :267-269 */
    or_graph_finalize_links(g);
}

/* a16: UniqueMerIndex{M}(graph) src/GraphIndexes.jl:62-102 with the F11 fixes (intended
 * semantics): unique_mers_per_node has n_nodes entries.  Node sequences are ASCII.
 * Outputs sorted unique keys + (signed node, 1-based pos).  Returns #unique. */
typedef struct { KM mer; int64_t node; uint32_t pos; } FN(umi);
static int FN(or_umi_cmp)(const void *a, const void *b) {
    return FN(km_cmp)(&((const FN(umi) *)a)->mer, &((const FN(umi) *)b)->mer);
}
static uint64_t FN(or_unique_mer_index)(const uint8_t *seq, const uint64_t *offs, uint64_t nn, int k,
                                        uint64_t *keys_out, int64_t *node_out, uint32_t *pos_out,
                                        uint64_t *unique_per_node, uint64_t *total_per_node) {
    uint64_t t = FN(or_extract)(seq, offs, nn, k, 1, NULL, NULL, NULL, NULL);
    for (uint64_t i = 0; i < nn; ++i) {                               /* :63-72 */
        uint64_t len = offs[i + 1] - offs[i];
        total_per_node[i] = len >= (uint64_t)k ? len + 1 - (uint64_t)k : 0;
        unique_per_node[i] = 0;
    }
    if (t == 0) return 0;
    uint64_t *kw = (uint64_t *)malloc(t * W * sizeof(uint64_t));
    uint32_t *pp = (uint32_t *)malloc(t * sizeof(uint32_t));
    uint64_t *rr = (uint64_t *)malloc(t * sizeof(uint64_t));
    uint8_t *ss = (uint8_t *)malloc(t);
    FN(or_extract)(seq, offs, nn, k, 1, kw, pp, rr, ss);              /* :75-86 */
    FN(umi) *v = (FN(umi) *)malloc(t * sizeof(FN(umi)));
    for (uint64_t i = 0; i < t; ++i) {
        for (int j = 0; j < W; ++j) v[i].mer.w[j] = kw[i * W + j];
        v[i].node = ss[i] ? (int64_t)(rr[i] + 1) : -(int64_t)(rr[i] + 1);
        v[i].pos = pp[i];
    }
    qsort(v, t, sizeof(FN(umi)), FN(or_umi_cmp));                     /* :88 */
    uint64_t u = 0, i = 0;
    while (i < t) {                                                   /* :43-60 */
        uint64_t j = i + 1;
        while (j < t && FN(km_eq)(&v[j].mer, &v[i].mer)) ++j;
        if (j - i == 1) {
            for (int q = 0; q < W; ++q) keys_out[u * W + q] = v[i].mer.w[q];
            node_out[u] = v[i].node;
            pos_out[u] = v[i].pos;
            int64_t a = v[i].node < 0 ? -v[i].node : v[i].node;
            unique_per_node[a - 1] += 1;                              /* :98 */
            ++u;
        }
        i = j;
    }
    free(kw); free(pp); free(rr); free(ss); free(v);
    return u;
}

#undef KM
#undef FN
#undef OR_CAT
#undef OR_CAT_
