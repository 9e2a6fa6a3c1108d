// graph.cuh -- K5..K8: sorted unique canonical k-mers -> unitig nodes + (k-1)-overlap links.
//
// Replaces (reference) build_unitigs! src/dbg.jl:94-185 and connect_unitigs_by_overlaps!
// src/dbg.jl:224-240 with a data-parallel formulation that yields the same nodes, node ids,
// orientations and per-node link order (derivation in DESIGN.md, "Unitigs without a walk"):
//
//   vertex 2i   = stored (canonical) k-mer i read forwards, vertex 2i+1 = its reverse complement
//   succ[v]     = the sole forward neighbour of v when !is_end_fw(v)   (src/dbg.jl:57-68)
//   The relation is symmetric under mirror(v) = v^1, so the vertices form disjoint simple
//   paths and cycles; the reference's sequential walk emits exactly one node per mirror pair
//   of paths (first half of a self-mirror path), then one node per mirror pair of cycles
//   started at the smallest stored k-mer.  Node creation order = ascending smallest stored
//   index of the two path ends (paths first, then cycles by smallest stored index).
#pragma once
#include <cooperative_groups.h>
#include "sort.cuh"

#define V_NONE 0xFFFFFFFFu  // no vertex / dead vertex
#define V_TERM 0x80000000u  // Wyllie record flag: ptr is the tail of the path

// ------------------------------- prefix index over sorted keys ----------------------------
template <int W>
__device__ __forceinline__ u32 key_prefix(const Kmer<W> &x, int k, int pbits) {
    if (pbits == 0) return 0;
    Kmer<W> s = km_shr<W>(x, 2 * k - pbits);
    return (u32)s.w[W - 1];
}
template <int W>
__global__ void prefix_index_kernel(const u64 *__restrict__ keys, u64 n, int k, int pbits, u32 *__restrict__ start) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    u32 t1 = i < n ? key_prefix<W>(km_load<W>(keys, i), k, pbits) : (1u << pbits);
    u32 t0 = i > 0 ? key_prefix<W>(km_load<W>(keys, i - 1), k, pbits) + 1 : 0;
    for (u32 t = t0; t <= t1; ++t) start[t] = (u32)i;
}
// index of c in keys or V_NONE
template <int W>
__device__ __forceinline__ u32 find_key(const u64 *__restrict__ keys, const u32 *__restrict__ start, int k, int pbits,
                                        const Kmer<W> &c) {
    u32 t = key_prefix<W>(c, k, pbits);
    u32 lo = start[t], hi = start[t + 1];
    while (lo < hi) {
        u32 mid = lo + ((hi - lo) >> 1);
        int cmp = km_cmp<W>(km_load<W>(keys, mid), c);
        if (cmp == 0) return mid;
        if (cmp < 0) lo = mid + 1; else hi = mid;
    }
    return V_NONE;
}

// --------------------------------- K5: neighbour lookup -----------------------------------
// 8 lanes per stored k-mer: lanes 0-3 the forward neighbours (src/dbg.jl:30-34, 81-90),
// lanes 4-7 the backward ones (:36-42, :70-79).  flags = cf | cb << 3 | palindrome << 6.
template <int W>
// Stored k-mers [i0, i1) are looked up (the multi-GPU graph stage gives every rank one index range).
__global__ void __launch_bounds__(256) lookup_kernel(const u64 *__restrict__ keys, u64 i0, u64 i1, int k, int pbits,
                                                     const u32 *__restrict__ start, u8 *__restrict__ flags,
                                                     u32 *__restrict__ nbF, u32 *__restrict__ nbB) {
    u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 i = i0 + (gid >> 3);
    u32 sub = (u32)gid & 7u, b = sub & 3u;
    bool live = i < i1;
    u32 v = V_NONE;
    bool pal = false;
    if (live) {
        Kmer<W> x = km_load<W>(keys, i);
        Kmer<W> nb = (sub < 4) ? km_fw_nb<W>(x, b, k) : km_bw_nb<W>(x, b, k);
        Kmer<W> r = km_rc<W>(nb, k);
        bool flip = km_lt<W>(r, nb);
        u32 idx = find_key<W>(keys, start, k, pbits, flip ? r : nb);
        if (idx != V_NONE) v = (idx << 1) | (flip ? 1u : 0u);
        if (sub == 0) pal = km_eq<W>(x, km_rc<W>(x, k));
    }
    u32 m = __ballot_sync(FULL_MASK, v != V_NONE);
    u32 gbase = lane_id() & ~7u;
    u32 bits = (m >> gbase) & 0xFFu;
    u32 fwb = bits & 0xFu, bwb = bits >> 4;
    int srcF = (int)gbase + (fwb ? __ffs(fwb) - 1 : 0);
    int srcB = (int)gbase + 4 + (bwb ? __ffs(bwb) - 1 : 0);
    u32 vF = __shfl_sync(FULL_MASK, v, srcF);
    u32 vB = __shfl_sync(FULL_MASK, v, srcB);
    if (live && sub == 0) {
        u32 cf = __popc(fwb), cb = __popc(bwb);
        flags[i] = (u8)(cf | (cb << 3) | (pal ? 64u : 0u));
        nbF[i] = cf == 1 ? vF : V_NONE;
        nbB[i] = cb == 1 ? vB : V_NONE;
    }
}

// ------------------------- K6: end flags -> successor pointers ----------------------------
// is_end_fw / is_end_bw (src/dbg.jl:44-68) from the per-k-mer neighbour counts.
// rec[v] = dist << 32 | ptr : Wyllie record; ptr has V_TERM when it already names the tail.
__global__ void succ_kernel(const u8 *__restrict__ flags, const u32 *__restrict__ nbF, const u32 *__restrict__ nbB, u64 n,
                            u32 *__restrict__ succ, u64 *__restrict__ rec) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 f = flags[i];
    u32 cf = f & 7u, cb = (f >> 3) & 7u;
    bool pal = (f >> 6) & 1u;
    u32 sF = V_NONE, sB = V_NONE;
    if (cf == 1) {
        u32 v = nbF[i];
        u32 fj = flags[v >> 1];
        u32 cnt = (v & 1u) ? (fj & 7u) : ((fj >> 3) & 7u);  // #bw neighbours of the oriented neighbour
        if (cnt == 1) sF = v;
    }
    if (cb == 1) {
        u32 v = nbB[i];
        u32 fj = flags[v >> 1];
        u32 cnt = (v & 1u) ? ((fj >> 3) & 7u) : (fj & 7u);  // #fw neighbours of the oriented neighbour
        if (cnt == 1) {
            bool palj = (fj >> 6) & 1u;
            sB = palj ? (v & ~1u) : (v ^ 1u);  // successor of rc(x) is rc(q)
        }
    }
    u32 v0 = (u32)(2 * i), v1 = v0 + 1;
    succ[v0] = sF;
    rec[v0] = sF == V_NONE ? (u64)(v0 | V_TERM) : ((1ULL << 32) | sF);
    if (pal) {
        succ[v1] = V_NONE;
        rec[v1] = (u64)V_NONE;  // dead vertex: a palindromic k-mer has a single orientation
    } else {
        succ[v1] = sB;
        rec[v1] = sB == V_NONE ? (u64)(v1 | V_TERM) : ((1ULL << 32) | sB);
    }
}

// one pointer-jumping round (Wyllie list ranking towards the tail)
__global__ void wyllie_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 nv, u32 *__restrict__ active) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u64 r = in[v];
    u32 ptr = (u32)r;
    if (ptr == V_NONE || (ptr & V_TERM)) { out[v] = r; return; }
    u64 q = in[ptr];
    u32 d = (u32)(r >> 32) + (u32)(q >> 32);
    u32 np = (u32)q;
    out[v] = ((u64)d << 32) | np;
    if (!(np & V_TERM)) *active = 1;
}
// all pointer-jumping rounds of a list in ONE cooperative launch (grid-wide barrier between rounds): the
// reduced splitter list is small, so separate launches would cost more than the rounds themselves
__global__ void __launch_bounds__(256) wyllie_fused_kernel(u64 *a, u64 *b, u64 nv, int rounds) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    u64 *in = a, *out = b;
    for (int r = 0; r < rounds; ++r) {
        for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (u64)gridDim.x * blockDim.x) {
            u64 rec = in[v];
            const u32 ptr = (u32)rec;
            if (ptr != V_NONE && !(ptr & V_TERM)) {
                const u64 q = in[ptr];
                rec = ((u64)((u32)(rec >> 32) + (u32)(q >> 32)) << 32) | (u32)q;
            }
            out[v] = rec;
        }
        grid.sync();
        u64 *t = in; in = out; out = t;
    }
}
// ---- splitter-based list ranking (work-efficient front end of the Wyllie rounds) -----------
// Splitters = path heads + a 2^-sl hash sample.  Every splitter walks to the next splitter (or
// to the tail), labelling the vertices it passes; the 2^sl times smaller list of splitters is ranked by
// pointer jumping; every vertex then derives (tail, distance to tail) from its splitter.
// Vertices on cycles keep a non-terminal record and go through the cycle path below.
__device__ __forceinline__ bool hash_splitter(u32 v, int sl) { return ((v * 0x9E3779B1u) >> (32 - sl)) == 0u; }
__global__ void split_flag_kernel(const u32 *__restrict__ succ, const u8 *__restrict__ flags, const u64 *__restrict__ rec, u64 nv,
                                  int sl, u32 *__restrict__ sflag) {
    u64 vv = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= nv) return;
    u32 v = (u32)vv;
    if ((u32)rec[v] == V_NONE) { sflag[v] = 0; return; }
    u32 mv = ((flags[v >> 1] >> 6) & 1u) ? v : (v ^ 1u);
    bool head = succ[mv] == V_NONE;  // pred(v) does not exist  <=>  the mirror vertex has no successor
    sflag[v] = (head || hash_splitter(v, sl)) ? 1u : 0u;
}
__global__ void split_list_kernel(const u32 *__restrict__ sflag, const u32 *__restrict__ sidx, u64 nv, u32 *__restrict__ red_vertex) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv && sflag[v]) red_vertex[sidx[v]] = (u32)v;
}
// one thread per splitter (dense: every lane of a warp walks).  A walk can only run into a
// hash-sampled splitter: heads have no predecessor, so no memory lookup is needed for the test.
__global__ void split_walk_kernel(const u32 *__restrict__ succ, const u32 *__restrict__ red_vertex, const u32 *__restrict__ sidx, u64 ns,
                                  int sl, u64 *__restrict__ ownloc, u64 *__restrict__ red_rec,
                                  u32 *__restrict__ red_tail) {
    u64 ii = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= ns) return;
    const u32 i = (u32)ii;
    u32 u = red_vertex[i], d = 0;
    for (;;) {
        u32 n = succ[u];
        if (n == V_NONE) { red_rec[i] = ((u64)d << 32) | (i | V_TERM); red_tail[i] = u; break; }
        if (hash_splitter(n, sl)) { red_rec[i] = ((u64)(d + 1) << 32) | sidx[n]; red_tail[i] = V_NONE; break; }
        ++d;
        ownloc[n] = ((u64)i << 32) | d;  // owner splitter and distance from it, one store
        u = n;
    }
}
__global__ void split_expand_kernel(const u64 *__restrict__ red_rec, const u32 *__restrict__ red_tail, const u32 *__restrict__ sflag,
                                    const u32 *__restrict__ sidx, const u64 *__restrict__ ownloc, u64 nv, u64 *__restrict__ rec) {
    u64 vv = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= nv) return;
    u32 v = (u32)vv;
    if ((u32)rec[v] == V_NONE) return;  // dead
    u32 i, back = 0;
    if (sflag[v]) i = sidx[v];
    else {
        const u64 ol = ownloc[v];
        i = (u32)(ol >> 32);
        if (i == V_NONE) return;  // on a cycle that holds no splitter: record stays non-terminal
        back = (u32)ol;
    }
    u64 r = red_rec[i];
    u32 ptr = (u32)r;
    if (!(ptr & V_TERM)) return;  // cycle
    u32 tail = red_tail[ptr & ~V_TERM];
    rec[v] = ((u64)((u32)(r >> 32) - back) << 32) | (tail | V_TERM);
}

__global__ void count_cyclic_kernel(const u64 *__restrict__ rec, u64 nv, u32 *__restrict__ flag) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 ptr = (u32)rec[v];
    if (ptr != V_NONE && !(ptr & V_TERM)) *flag = 1;
}
// cycles: cm[v] = min << 32 | ptr over the vertices reachable in 2^r steps
__global__ void cyc_init_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ succ, u64 nv, u64 *__restrict__ cm) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 ptr = (u32)rec[v];
    bool cyc = ptr != V_NONE && !(ptr & V_TERM);
    cm[v] = cyc ? (((u64)v << 32) | succ[v]) : ~0ULL;
}
__global__ void cyc_round_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 nv) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u64 r = in[v];
    if (r == ~0ULL) { out[v] = r; return; }
    u64 q = in[(u32)r];
    u32 m = min((u32)(r >> 32), (u32)(q >> 32));
    out[v] = ((u64)m << 32) | (u32)q;
}
// cut every emitting cycle (smallest vertex id even) just before its leader; drop the others
__global__ void cyc_cut_kernel(const u64 *__restrict__ cm, const u32 *__restrict__ succ, u64 nv, u64 *__restrict__ rec,
                               u32 *__restrict__ cyc_head) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u64 r = cm[v];
    if (r == ~0ULL) { cyc_head[v] = V_NONE; return; }
    u32 leader = (u32)(r >> 32);
    if (leader & 1u) {  // the mirror cycle holds the canonical orientation of the smallest k-mer
        rec[v] = (u64)V_NONE;
        cyc_head[v] = V_NONE;
        return;
    }
    cyc_head[v] = leader;
    rec[v] = (succ[v] == leader) ? (u64)((u32)v | V_TERM) : ((1ULL << 32) | succ[v]);
}

__host__ __device__ __forceinline__ u8 code_to_ascii(u32 code) { return (u8)(0x54474341u >> (8 * code)); }  // "ACGT"
template <int W>
__device__ __forceinline__ Kmer<W> vertex_kmer(const u64 *__restrict__ keys, u32 v, int k) {
    Kmer<W> x = km_load<W>(keys, v >> 1);
    return (v & 1u) ? km_rc<W>(x, k) : x;
}
__device__ __forceinline__ u32 mirror_of(const u8 *__restrict__ flags, u32 v) {
    return ((flags[v >> 1] >> 6) & 1u) ? v : (v ^ 1u);
}

// per vertex: head and rank inside its path; path heads / deciding vertices publish the
// node record keyed by the node's creation key (smallest stored index of the two path ends).
template <int W>
__global__ void __launch_bounds__(256) resolve_kernel(const u64 *__restrict__ keys, const u8 *__restrict__ flags,
                                                      const u64 *__restrict__ rec, const u32 *__restrict__ cyc_head, u64 nv, int k,
                                                      u32 *__restrict__ vhead, u32 *__restrict__ vrank,
                                                      uint4 *__restrict__ prec, u32 *__restrict__ pflag, u32 *__restrict__ cflag) {
    u64 vv = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= nv) return;
    u32 v = (u32)vv;
    u64 r = rec[v];
    u32 ptr = (u32)r;
    if (ptr == V_NONE) { vhead[v] = V_NONE; vrank[v] = 0; return; }
    u32 tail = ptr & ~V_TERM, dist = (u32)(r >> 32);
    u32 ch = cyc_head ? cyc_head[v] : V_NONE;
    if (ch != V_NONE) {  // vertex of an emitting cycle, cut at its leader
        u32 n = (u32)(rec[ch] >> 32) + 1;
        u32 rank = n - 1 - dist;
        vhead[v] = ch;
        vrank[v] = rank;
        if (rank == 0) {
            u32 key = v >> 1;
            prec[key] = make_uint4(v, 0u, n, V_NONE); cflag[key] = 1;
        }
        return;
    }
    u32 mv = mirror_of(flags, v);
    u64 rm = rec[mv];
    u32 head = mirror_of(flags, ((u32)rm) & ~V_TERM);
    u32 rank = (u32)(rm >> 32);
    u32 n = rank + dist + 1;
    vhead[v] = head;
    vrank[v] = rank;
    bool sm = mirror_of(flags, tail) == head;  // the path is its own mirror image
    if (!sm) {
        if (rank == 0) {
            // canonical!(s): keep the orientation whose first k-mer is smaller (src/dbg.jl:156)
            Kmer<W> a = vertex_kmer<W>(keys, v, k), b = vertex_kmer<W>(keys, mirror_of(flags, tail), k);
            if (km_lt<W>(a, b)) {
                u32 key = min(v >> 1, tail >> 1);
                prec[key] = make_uint4(v, 0u, n, V_NONE); pflag[key] = 1;
            }
        }
    } else {
        // the walk stops when it meets its own reverse strand (src/dbg.jl:147-150): the node is
        // the first ceil(n/2) vertices, stored in whichever orientation is canonical
        u32 h = (n + 1) / 2;
        if (rank == n - h) {
            Kmer<W> a = vertex_kmer<W>(keys, head, k), b = vertex_kmer<W>(keys, v, k);
            u32 key = head >> 1;
            prec[key] = make_uint4(head, km_cmp<W>(a, b) <= 0 ? 0u : (n - h), h, V_NONE);
            pflag[key] = 1;
        }
    }
}
// node ids from the scanned flags; node lengths
__global__ void node_len_kernel(const u32 *__restrict__ pflag, const u32 *__restrict__ pscan, const u32 *__restrict__ cflag,
                                const u32 *__restrict__ cscan, u64 n, u32 n_path, int k, uint4 *__restrict__ prec,
                                u32 *__restrict__ node_len) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 id = V_NONE;
    if (pflag[i]) id = pscan[i];
    else if (cflag && cflag[i]) id = n_path + cscan[i];
    if (id != V_NONE) {
        prec[i].w = id;
        node_len[id] = (u32)k + prec[i].z - 1;
    }
}
// every vertex of an emitted range writes its base(s)
template <int W>
__global__ void __launch_bounds__(256) emit_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ rec,
                                                   const u32 *__restrict__ cyc_head, const u32 *__restrict__ vhead,
                                                   const u32 *__restrict__ vrank, u64 nv, int k, const uint4 *__restrict__ prec,
                                                   const u32 *__restrict__ node_off,
                                                   u8 *__restrict__ bases, u32 *__restrict__ node_first, u32 *__restrict__ node_last) {
    u64 vv = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= nv) return;
    u32 v = (u32)vv;
    u32 head = vhead[v];
    if (head == V_NONE) return;
    u32 tail = ((u32)rec[v]) & ~V_TERM;
    bool cyc = cyc_head && cyc_head[v] != V_NONE;
    u32 key = cyc ? (head >> 1) : min(head >> 1, tail >> 1);
    const uint4 pr = prec[key];  // (head, first emitted rank, emitted count, node id) of the path keyed here
    if (pr.x != head) return;    // the mirror orientation is the emitted one
    u32 id = pr.w;
    u32 rank = vrank[v], re = pr.y, cnt = pr.z;
    if (rank < re || rank - re >= cnt) return;
    u32 j = rank - re;
    u64 off = node_off[id];
    Kmer<W> x = vertex_kmer<W>(keys, v, k);
    if (j == 0) {
        for (int q = 0; q < k; ++q) bases[off + q] = code_to_ascii(km_base<W>(x, q, k));
        node_first[id] = v;
    } else {
        bases[off + k - 1 + j] = code_to_ascii((u32)(x.w[W - 1] & 3ULL));
    }
    if (j == cnt - 1) node_last[id] = v;
}

// ------------------------------------ K8: links ---------------------------------------------
// find_unitig_overlaps (src/dbg.jl:190-222): record 2*id+0 = first (k-1)-mer, 2*id+1 = last.
// Composite sort key: [ (k-1)-mer words | nid + n_nodes ]  (tuple order mer, signed nid).
template <int W>
struct OverlapF {
    const u64 *keys;
    const u32 *node_first, *node_last;
    u64 *out;  // (W+1) words per record
    u64 nn;
    int k, want_in;
    __device__ void make(u64 r, Kmer<W> &mer, i64 &nid, bool &is_in) const {
        u64 id = r >> 1;
        int k1 = k - 1;
        Kmer<W> m;
        if (!(r & 1)) {
            m = km_shr<W>(vertex_kmer<W>(keys, node_first[id], k), 2);  // first k-1 bases
            nid = (i64)(id + 1);
        } else {
            m = km_mask<W>(vertex_kmer<W>(keys, node_last[id], k), k1);  // last k-1 bases
            nid = -(i64)(id + 1);
        }
        Kmer<W> rcm = km_rc<W>(m, k1);
        bool canon = km_cmp<W>(m, rcm) <= 0;
        mer = canon ? m : rcm;
        is_in = (!(r & 1)) ? canon : !canon;  // src/dbg.jl:203-217
    }
    __device__ bool pred(u64 r) const {
        Kmer<W> mer; i64 nid; bool is_in;
        make(r, mer, nid, is_in);
        return is_in == (want_in != 0);
    }
    __device__ void emit(u64 r, u64 dst) const {
        Kmer<W> mer; i64 nid; bool is_in;
        make(r, mer, nid, is_in);
#pragma unroll
        for (int j = 0; j < W; ++j) out[dst * (W + 1) + j] = mer.w[j];
        out[dst * (W + 1) + W] = (u64)(nid + (i64)nn);
    }
};
// for each `in` record: range of `out` records with the same (k-1)-mer (src/dbg.jl:230-239)
template <int W>
__global__ void join_count_kernel(const u64 *__restrict__ in, u64 ni, const u64 *__restrict__ out, u64 no,
                                  u32 *__restrict__ lb, u32 *__restrict__ cnt) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    Kmer<W> m;
#pragma unroll
    for (int j = 0; j < W; ++j) m.w[j] = in[i * (W + 1) + j];
    u64 lo = 0, hi = no;
    while (lo < hi) {
        u64 mid = lo + ((hi - lo) >> 1);
        Kmer<W> o;
#pragma unroll
        for (int j = 0; j < W; ++j) o.w[j] = out[mid * (W + 1) + j];
        if (km_cmp<W>(o, m) < 0) lo = mid + 1; else hi = mid;
    }
    u64 e = lo;
    for (;;) {
        if (e >= no) break;
        Kmer<W> o;
#pragma unroll
        for (int j = 0; j < W; ++j) o.w[j] = out[e * (W + 1) + j];
        if (!km_eq<W>(o, m)) break;
        ++e;
    }
    lb[i] = (u32)lo;
    cnt[i] = (u32)(e - lo);
}
// add_link! calls in reference order; each call makes two pushes (src/Graphs.jl:473-476)
template <int W>
__global__ void join_emit_kernel(const u64 *__restrict__ in, u64 ni, const u64 *__restrict__ out, const u32 *__restrict__ lb,
                                 const u32 *__restrict__ cnt, const u32 *__restrict__ off, u64 nn, i64 *__restrict__ call_src,
                                 i64 *__restrict__ call_dst, u64 *__restrict__ push_owner, u64 *__restrict__ push_idx) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    i64 src = (i64)in[i * (W + 1) + W] - (i64)nn;
    u32 c = cnt[i], l = lb[i], o = off[i];
    for (u32 q = 0; q < c; ++q) {
        i64 dst = (i64)out[(u64)(l + q) * (W + 1) + W] - (i64)nn;
        u64 call = (u64)o + q;
        call_src[call] = src;
        call_dst[call] = dst;
        push_owner[2 * call] = (u64)(src < 0 ? -src : src);
        push_owner[2 * call + 1] = (u64)(dst < 0 ? -dst : dst);
        push_idx[2 * call] = 2 * call;
        push_idx[2 * call + 1] = 2 * call + 1;
    }
}
__global__ void link_rows_kernel(const u64 *__restrict__ owner, const u64 *__restrict__ pidx, u64 np, u64 nn,
                                 const i64 *__restrict__ call_src, const i64 *__restrict__ call_dst, i64 dist,
                                 u64 *__restrict__ row, i64 *__restrict__ tri) {
    u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > np) return;
    u64 o1 = q < np ? owner[q] : nn + 1;
    u64 o0 = q > 0 ? owner[q - 1] + 1 : 1;
    for (u64 o = o0; o <= o1; ++o) row[o - 1] = q;
    if (q < np) {
        u64 p = pidx[q], call = p >> 1;
        i64 s = call_src[call], d = call_dst[call];
        tri[3 * q] = (p & 1) ? d : s;
        tri[3 * q + 1] = (p & 1) ? s : d;
        tri[3 * q + 2] = dist;
    }
}
__global__ void widen_offsets_kernel(const u32 *__restrict__ in, u64 n, u64 total, u64 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[n] = total;
}
__global__ void fill_u64_kernel(u64 *p, u64 n, u64 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct gg_graph {
    CtxRef ctx;
    int k = 0;
    u64 n_nodes = 0, total_bases = 0, n_link_entries = 0;
    u64 *d_node_off = nullptr;  // n_nodes + 1
    u8 *d_bases = nullptr;      // ASCII
    u64 *d_link_row = nullptr;  // n_nodes + 1
    i64 *d_link_tri = nullptr;  // 3 * n_link_entries
};

static inline int ilog2_ceil(u64 x) { int b = 0; while ((1ULL << b) < x) ++b; return b; }


// pointer jumping over the reduced (splitter) list: a[i] = dist << 32 | next (V_TERM: list end).  Returns the buffer
// that holds the ranked records.  One cooperative launch when the device supports it.
static u64 *rank_reduced_list(gg_ctx *ctx, u64 *ra, u64 *rb, u64 ns) {
    if (ns == 0) return ra;
    int rounds = ilog2_ceil(ns) + 2;  // fixed count: no host round trips
    if (ctx->coop_blocks < 0) {       // co-resident CTAs per SM of the fused kernel on THIS device (0: not supported)
        int supported = 0, nb = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, ctx->device);
        if (supported) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wyllie_fused_kernel, 256, 0);
        ctx->coop_blocks = supported ? nb : 0;
    }
    if (ctx->coop_blocks > 0 && !getenv("GG_NO_COOP")) {
        u64 grid = ceil_div(ns, 256);
        if (grid > (u64)ctx->coop_blocks * ctx->num_sms) grid = (u64)ctx->coop_blocks * ctx->num_sms;
        u64 nsv = ns;
        void *args[] = {&ra, &rb, &nsv, &rounds};
        cudaEvent_t pa = nullptr, pb = nullptr;
        if (ctx->profile) { pa = prof_event(ctx); pb = prof_event(ctx); cudaEventRecord(pa, ctx->stream); }
        CK(cudaLaunchCooperativeKernel((void *)wyllie_fused_kernel, dim3((unsigned)grid), dim3(256), args, 0, ctx->stream));
        ctx->launches++;
        if (ctx->profile) { cudaEventRecord(pb, ctx->stream); ctx->prof.push_back({std::string("wyllie_fused_kernel"), pa, pb}); }
        return (rounds & 1) ? rb : ra;
    }
    DBuf<u32> flag(ctx, 1);
    for (int r = 0; r < rounds; ++r) {
        LAUNCH(ctx, wyllie_kernel, (unsigned)ceil_div(ns, 256), 256, 0, ra, rb, ns, flag.p);
        std::swap(ra, rb);
    }
    return ra;
}

template <int W>
static void build_links(gg_ctx *ctx, gg_graph *g, const u64 *keys, const u32 *node_first, const u32 *node_last) {
    const int k = g->k;
    const u64 nn = g->n_nodes;
    g->d_link_row = dalloc<u64>(ctx, nn + 1);
    if (nn <= 1 || k < 2) {  // src/dbg.jl:267: no linking for a single node
        LAUNCH(ctx, fill_u64_kernel, (unsigned)ceil_div(nn + 1, 256), 256, 0, g->d_link_row, nn + 1, 0ULL);
        g->d_link_tri = dalloc<i64>(ctx, 3);
        g->n_link_entries = 0;
        return;
    }
    constexpr int KW = W + 1;
    OverlapF<W> f{keys, node_first, node_last, nullptr, nn, k, 1};
    Compactor<OverlapF<W>> cin(ctx, 2 * nn, "overlap"), cout_(ctx, 2 * nn, "overlap");
    u64 ni = cin.count(f);
    u64 no = 2 * nn - ni;
    DBuf<u64> in(ctx, (ni + 1) * KW), in_alt(ctx, (ni + 1) * KW), out(ctx, (no + 1) * KW), out_alt(ctx, (no + 1) * KW);
    f.out = in.p;
    cin.emit(f);
    f.want_in = 0;
    f.out = out.p;
    cout_.count(f);
    cout_.emit(f);
    std::vector<BitRange> ranges = {{0, ilog2_ceil(2 * nn + 1)}, {64, 64 + 2 * (k - 1)}};
    u64 *a = in.p, *b = in_alt.p;
    radix_sort<KW>(ctx, &a, &b, nullptr, nullptr, ni, ranges);
    u64 *c = out.p, *d = out_alt.p;
    radix_sort<KW>(ctx, &c, &d, nullptr, nullptr, no, ranges);
    DBuf<u32> lb(ctx, ni + 1), cnt(ctx, ni + 1), off(ctx, ni + 1);
    DBuf<u64> tot(ctx, 1);
    u64 ncalls = 0;
    if (ni) {
        LAUNCH(ctx, join_count_kernel<W>, (unsigned)ceil_div(ni, 128), 128, 0, a, ni, c, no, lb.p, cnt.p);
        exclusive_scan_u32(ctx, cnt.p, off.p, ni, tot.p);
        ncalls = read_scalar<u64>(ctx, tot.p);
    }
    u64 np = 2 * ncalls;
    g->n_link_entries = np;
    g->d_link_tri = dalloc<i64>(ctx, 3 * (np ? np : 1));
    if (np == 0) {
        LAUNCH(ctx, fill_u64_kernel, (unsigned)ceil_div(nn + 1, 256), 256, 0, g->d_link_row, nn + 1, 0ULL);
        return;
    }
    if (np >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "too many link entries (%llu)", (unsigned long long)np);
    DBuf<i64> call_src(ctx, ncalls), call_dst(ctx, ncalls);
    DBuf<u64> owner(ctx, np), owner_alt(ctx, np), pidx(ctx, np), pidx_alt(ctx, np);
    LAUNCH(ctx, join_emit_kernel<W>, (unsigned)ceil_div(ni, 128), 128, 0, a, ni, c, lb.p, cnt.p, off.p, nn, call_src.p, call_dst.p, owner.p, pidx.p);
    u64 *ow = owner.p, *ow2 = owner_alt.p, *pi = pidx.p, *pi2 = pidx_alt.p;
    std::vector<BitRange> oranges = {{0, ilog2_ceil(nn + 2)}};
    radix_sort<1>(ctx, &ow, &ow2, &pi, &pi2, np, oranges);  // stable: keeps push order per node
    LAUNCH(ctx, link_rows_kernel, (unsigned)ceil_div(np + 1, 256), 256, 0, ow, pi, np, nn, call_src.p, call_dst.p, -(i64)(k - 1), g->d_link_row, g->d_link_tri);
}

// keys: U sorted unique canonical k-mers (device). Builds the graph handle.
// cm (optional): communicator of a multi-GPU job in which every rank holds the same table; the neighbour
// lookups are then split over the ranks by index range and their results all-gathered.
template <int W>
static gg_graph *build_graph(gg_ctx *ctx, const u64 *keys, u64 U, int k, gg_comm *cm = nullptr) {
    gg_graph *g = new gg_graph();
    g->ctx = ctx;
    g->k = k;
    try {
        if (U == 0) {
            g->d_node_off = dalloc<u64>(ctx, 1);
            CK(cudaMemsetAsync(g->d_node_off, 0, sizeof(u64), ctx->stream));
            g->d_link_row = dalloc<u64>(ctx, 1);
            CK(cudaMemsetAsync(g->d_link_row, 0, sizeof(u64), ctx->stream));
            g->d_bases = dalloc<u8>(ctx, 1);
            g->d_link_tri = dalloc<i64>(ctx, 3);
            return g;
        }
        if (U >= (1ULL << 30)) GG_THROW(GG_ERR_LIMIT, "%llu k-mers exceed the 2^30 per-GPU graph limit", (unsigned long long)U);
        const u64 nv = 2 * U;
        // ---- K5 ----
        stage_begin(ctx, ST_LOOKUP);
        int pbits = ilog2_ceil(U) - 1;
        if (pbits > 24) pbits = 24;
        if (pbits > 2 * k) pbits = 2 * k;
        if (pbits < 0) pbits = 0;
        DBuf<u32> start(ctx, (1ULL << pbits) + 1);
        LAUNCH(ctx, prefix_index_kernel<W>, (unsigned)ceil_div(U + 1, 256), 256, 0, keys, U, k, pbits, start.p);
        const int P = (cm && cm->world > 1 && !getenv("GG_DIST_NO_SHARDED_LOOKUP")) ? cm->world : 1;
        const u64 chunk = (ceil_div(U, (u64)P) + 15) & ~15ULL;  // keys per rank (16-aligned pieces for the all-gathers)
        DBuf<u8> flags(ctx, chunk * P);
        DBuf<u32> nbF(ctx, chunk * P), nbB(ctx, chunk * P);
        if (P == 1) {
            LAUNCH(ctx, lookup_kernel<W>, (unsigned)ceil_div(8 * U, 256), 256, 0, keys, 0ULL, U, k, pbits, start.p, flags.p, nbF.p, nbB.p);
        } else {
            NcclApi *nc = nccl_api();
            const u64 i0 = chunk * cm->rank < U ? chunk * cm->rank : U, i1 = i0 + chunk < U ? i0 + chunk : U;
            if (i1 > i0) LAUNCH(ctx, lookup_kernel<W>, (unsigned)ceil_div(8 * (i1 - i0), 256), 256, 0, keys, i0, i1, k, pbits, start.p, flags.p, nbF.p, nbB.p);
            NCK(nc->AllGather(flags.p + chunk * cm->rank, flags.p, chunk, ncclUint8, cm->comm, ctx->stream));
            NCK(nc->AllGather(nbF.p + chunk * cm->rank, nbF.p, chunk, ncclUint32, cm->comm, ctx->stream));
            NCK(nc->AllGather(nbB.p + chunk * cm->rank, nbB.p, chunk, ncclUint32, cm->comm, ctx->stream));
        }
        stage_end(ctx, ST_LOOKUP);
        // ---- K6/K7 ----
        stage_begin(ctx, ST_UNITIG);
        DBuf<u32> succ(ctx, nv);
        DBuf<u64> recA(ctx, nv), recB(ctx, nv);
        LAUNCH(ctx, succ_kernel, (unsigned)ceil_div(U, 256), 256, 0, flags.p, nbF.p, nbB.p, U, succ.p, recA.p);
        nbF.release();
        nbB.release();
        start.release();
        DBuf<u32> d_flag(ctx, 1);
        u64 *rin = recA.p, *rout = recB.p;
        const int max_rounds = ilog2_ceil(nv) + 2;
        auto run_wyllie = [&]() {
            for (int round = 0; round < max_rounds; round += 2) {
                CK(cudaMemsetAsync(d_flag.p, 0, sizeof(u32), ctx->stream));
                for (int s = 0; s < 2; ++s) {
                    LAUNCH(ctx, wyllie_kernel, (unsigned)ceil_div(nv, 256), 256, 0, rin, rout, nv, d_flag.p);
                    std::swap(rin, rout);
                }
                if (read_scalar<u32>(ctx, d_flag.p) == 0) break;
            }
        };
        if (getenv("GG_PLAIN_WYLLIE")) {
            run_wyllie();
        } else {
            DBuf<u32> sflag(ctx, nv), sidx(ctx, nv);
            DBuf<u64> ownloc(ctx, nv);
            DBuf<u64> stot(ctx, 1);
            const int sl = getenv("GG_SPLIT_LOG") ? atoi(getenv("GG_SPLIT_LOG")) : 4;  // splitters: heads + a 2^-sl hash sample
            LAUNCH(ctx, split_flag_kernel, (unsigned)ceil_div(nv, 256), 256, 0, succ.p, flags.p, rin, nv, sl, sflag.p);
            exclusive_scan_u32(ctx, sflag.p, sidx.p, nv, stot.p);
            const u64 ns = read_scalar<u64>(ctx, stot.p);
            CK(cudaMemsetAsync(ownloc.p, 0xFF, nv * sizeof(u64), ctx->stream));
            DBuf<u64> redA(ctx, ns + 1), redB(ctx, ns + 1);
            DBuf<u32> red_tail(ctx, ns + 1), red_vertex(ctx, ns + 1);
            LAUNCH(ctx, split_list_kernel, (unsigned)ceil_div(nv, 256), 256, 0, sflag.p, sidx.p, nv, red_vertex.p);
            if (ns) LAUNCH(ctx, split_walk_kernel, (unsigned)ceil_div(ns, 128), 128, 0, succ.p, red_vertex.p, sidx.p, ns, sl, ownloc.p, redA.p, red_tail.p);
            u64 *ra = rank_reduced_list(ctx, redA.p, redB.p, ns);
            LAUNCH(ctx, split_expand_kernel, (unsigned)ceil_div(nv, 256), 256, 0, ra, red_tail.p, sflag.p, sidx.p, ownloc.p, nv, rin);
        }
        CK(cudaMemsetAsync(d_flag.p, 0, sizeof(u32), ctx->stream));
        LAUNCH(ctx, count_cyclic_kernel, (unsigned)ceil_div(nv, 256), 256, 0, rin, nv, d_flag.p);
        bool has_cycles = read_scalar<u32>(ctx, d_flag.p) != 0;
        DBuf<u32> cyc_head;
        if (has_cycles) {  // perfect circles (src/dbg.jl:160-181)
            DBuf<u64> cmA(ctx, nv), cmB(ctx, nv);
            LAUNCH(ctx, cyc_init_kernel, (unsigned)ceil_div(nv, 256), 256, 0, rin, succ.p, nv, cmA.p);
            u64 *ci = cmA.p, *co = cmB.p;
            for (int round = 0; round < max_rounds; ++round) {
                LAUNCH(ctx, cyc_round_kernel, (unsigned)ceil_div(nv, 256), 256, 0, ci, co, nv);
                std::swap(ci, co);
            }
            cyc_head.alloc(ctx, nv);
            LAUNCH(ctx, cyc_cut_kernel, (unsigned)ceil_div(nv, 256), 256, 0, ci, succ.p, nv, rin, cyc_head.p);
            cmA.release();
            cmB.release();
            run_wyllie();
        }
        const u64 *rec = rin;
        DBuf<u32> vhead(ctx, nv), vrank(ctx, nv);
        DBuf<uint4> prec(ctx, U);
        DBuf<u32> pflag(ctx, U), cflag;
        CK(cudaMemsetAsync(prec.p, 0xFF, U * sizeof(uint4), ctx->stream));
        CK(cudaMemsetAsync(pflag.p, 0, U * sizeof(u32), ctx->stream));
        if (has_cycles) {
            cflag.alloc(ctx, U);
            CK(cudaMemsetAsync(cflag.p, 0, U * sizeof(u32), ctx->stream));
        }
        LAUNCH(ctx, resolve_kernel<W>, (unsigned)ceil_div(nv, 256), 256, 0, keys, flags.p, rec, cyc_head.p, nv, k, vhead.p, vrank.p,
               prec.p, pflag.p, cflag.p);
        DBuf<u32> pscan(ctx, U), cscan;
        DBuf<u64> tot(ctx, 2);
        exclusive_scan_u32(ctx, pflag.p, pscan.p, U, tot.p);
        u64 n_path = 0, n_cyc = 0;
        if (has_cycles) {
            cscan.alloc(ctx, U);
            exclusive_scan_u32(ctx, cflag.p, cscan.p, U, tot.p + 1);
            CK(cudaMemcpyAsync(ctx->h_scalar, tot.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            n_path = ctx->h_scalar[0];
            n_cyc = ctx->h_scalar[1];
        } else {
            n_path = read_scalar<u64>(ctx, tot.p);
        }
        const u64 nn = n_path + n_cyc;
        g->n_nodes = nn;
        DBuf<u32> node_len(ctx, nn + 1), node_off32(ctx, nn + 1), node_first(ctx, nn + 1), node_last(ctx, nn + 1);
        LAUNCH(ctx, node_len_kernel, (unsigned)ceil_div(U, 256), 256, 0, pflag.p, pscan.p, cflag.p, cscan.p, U, (u32)n_path, k,
               prec.p, node_len.p);
        exclusive_scan_u32(ctx, node_len.p, node_off32.p, nn, tot.p);
        u64 total_bases = read_scalar<u64>(ctx, tot.p);
        if (total_bases >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "node sequences exceed 2^32 bases (%llu)", (unsigned long long)total_bases);
        g->total_bases = total_bases;
        g->d_bases = dalloc<u8>(ctx, total_bases ? total_bases : 1);
        g->d_node_off = dalloc<u64>(ctx, nn + 1);
        LAUNCH(ctx, widen_offsets_kernel, (unsigned)ceil_div(nn + 1, 256), 256, 0, node_off32.p, nn, total_bases, g->d_node_off);
        LAUNCH(ctx, emit_kernel<W>, (unsigned)ceil_div(nv, 256), 256, 0, keys, rec, cyc_head.p, vhead.p, vrank.p, nv, k, prec.p,
               node_off32.p, g->d_bases, node_first.p, node_last.p);
        stage_end(ctx, ST_UNITIG);
        // ---- K8 ----
        stage_begin(ctx, ST_LINKS);
        build_links<W>(ctx, g, keys, node_first.p, node_last.p);
        stage_end(ctx, ST_LINKS);
        CK(cudaStreamSynchronize(ctx->stream));
        return g;
    } catch (...) {
        dfree(ctx, g->d_node_off); dfree(ctx, g->d_bases); dfree(ctx, g->d_link_row); dfree(ctx, g->d_link_tri);
        delete g;
        throw;
    }
}
