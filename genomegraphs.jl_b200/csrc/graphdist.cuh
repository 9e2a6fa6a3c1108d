// graphdist.cuh -- K5..K7 with the per-vertex work SHARDED over the ranks of a communicator (and the default
// single-GPU formulation: one rank, no collectives).
//
// Same result as graph.cuh's replicated formulation (reference: build_unitigs! src/dbg.jl:94-185): nodes, node ids,
// orientations.  Every rank holds the whole sorted table of kept k-mers (all-gathered: the binary searches, the
// canonical! comparisons and the link stage read arbitrary k-mers) and OWNS the k-mers [i0, i1) = vertices
// [2 i0, 2 i1); everything that costs one random access per vertex runs on the owner only:
//
//   lookup (8 searches per k-mer)  -> flags all-gathered (1 B per k-mer)
//   succ                           -> all-gathered (4 B per vertex): the walks follow chains anywhere
//   splitter flags                 -> all-gathered as a BITMAP; rank-in-bitmap (one 8-byte entry per 32 vertices:
//                                     prefix << 32 | bits) replaces the per-vertex flag / index arrays
//   walks from the own splitters   -> (owner splitter, distance) of every passed vertex is stored into the shard of
//                                     the rank that owns THAT vertex: peer stores over NVLink (CUDA IPC mapping of
//                                     the exchange buffer), or a full-size array + ncclAllReduce(min) without it
//   reduced list (one record per splitter) -> exchanged, ranked by pointer jumping on every rank (2^-sl of the vertices)
//   expand, resolve                -> own vertices; a deciding vertex appends its node record to a local list and
//                                     marks the node's creation key in a bitmap over the k-mers
//   node ids                       -> bitmap summed over the ranks (disjoint bits), rank-in-bitmap = creation order
//   node table (head, first rank, count, length) -> each rank fills the ids it found, summed over the ranks
//   emit                           -> own vertices write their bases; bases / first / last vertex summed over the ranks
//
// Perfect circles (src/dbg.jl:160-181) are rare; when any rank sees one, all ranks fall back to graph.cuh.
#pragma once
#include <memory>
#include "graph.cuh"

struct GsPeers {
    u64 *own[MSD_MAX_PEERS];   // ownloc shard of every rank
    u32 vchunk;                // vertices per rank
};
__device__ __forceinline__ u32 sel_rank(const u64 *__restrict__ sel, u32 x) {
    const u64 e = sel[x >> 5];
    return (u32)(e >> 32) + __popc((u32)e & ((1u << (x & 31u)) - 1u));
}
__device__ __forceinline__ bool sel_bit(const u64 *__restrict__ sel, u32 x) { return ((u32)sel[x >> 5] >> (x & 31u)) & 1u; }
// pf: the flags when k is even, null when k is odd -- no k-mer of odd length equals its own reverse complement, so the
// palindrome bit needs no (random) look-up and the mirror vertex is v ^ 1
__device__ __forceinline__ bool gs_dead(const u8 *__restrict__ pf, u32 v, u64 U) {
    return (u64)(v >> 1) >= U || (pf && (v & 1u) && ((pf[v >> 1] >> 6) & 1u));
}
__device__ __forceinline__ u32 gs_mirror(const u8 *__restrict__ pf, u32 v) { return (pf && ((pf[v >> 1] >> 6) & 1u)) ? v : (v ^ 1u); }

// is_end_fw / is_end_bw (src/dbg.jl:44-68) of the own k-mers -> successor pointers of their two vertices
__global__ void gs_succ_kernel(const u8 *__restrict__ flags, const u32 *__restrict__ nbF, const u32 *__restrict__ nbB, u64 i0, u64 chunk, u64 U,
                               u32 *__restrict__ succ) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunk) return;
    const u64 i = i0 + t;
    u32 sF = V_NONE, sB = V_NONE;
    if (i < U) {
        const u32 f = flags[i];
        const u32 cf = f & 7u, cb = (f >> 3) & 7u;
        const bool pal = (f >> 6) & 1u;
        if (cf == 1) {
            const u32 v = nbF[t];
            const u32 fj = flags[v >> 1];
            const u32 cnt = (v & 1u) ? (fj & 7u) : ((fj >> 3) & 7u);  // #bw neighbours of the oriented neighbour
            if (cnt == 1) sF = v;
        }
        if (cb == 1 && !pal) {
            const u32 v = nbB[t];
            const u32 fj = flags[v >> 1];
            const u32 cnt = (v & 1u) ? ((fj >> 3) & 7u) : (fj & 7u);  // #fw neighbours of the oriented neighbour
            if (cnt == 1) sB = ((fj >> 6) & 1u) ? (v & ~1u) : (v ^ 1u);  // successor of rc(x) is rc(q)
        }
    }
    succ[2 * i] = sF;
    succ[2 * i + 1] = sB;   // a palindromic k-mer has a single orientation: its odd vertex is dead
}
// splitters = path heads + a 2^-sl hash sample, as one bit per own vertex (v0 is a multiple of 32)
__global__ void gs_flag_kernel(const u32 *__restrict__ succ, const u8 *__restrict__ flags, u64 U, u32 v0, u32 nvl, int sl, u32 *__restrict__ bits) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    bool f = false;
    if (t < nvl) {
        const u32 v = v0 + t;
        if (!gs_dead(flags, v, U)) {
            const u32 mv = gs_mirror(flags, v);
            f = succ[mv] == V_NONE || hash_splitter(v, sl);  // no predecessor  <=>  the mirror vertex has no successor
        }
    }
    const u32 w = __ballot_sync(FULL_MASK, f);
    if (lane_id() == 0 && t < nvl) bits[(v0 + t) >> 5] = w;
}
__global__ void gs_popc_kernel(const u32 *__restrict__ bits, u64 nw, u32 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) out[i] = __popc(bits[i]);
}
__global__ void gs_sel_kernel(const u32 *__restrict__ bits, const u32 *__restrict__ pref, u64 nw, u64 *__restrict__ sel) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) sel[i] = ((u64)pref[i] << 32) | bits[i];
}
__global__ void gs_pick_kernel(const u32 *__restrict__ pref, u64 stride, int n, u32 *__restrict__ out) {
    if ((int)threadIdx.x < n) out[threadIdx.x] = pref[(u64)threadIdx.x * stride];
}
__global__ void gs_list_kernel(const u64 *__restrict__ sel, u32 v0, u32 nvl, u32 soff, u32 *__restrict__ red_vertex) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nvl) return;
    const u32 v = v0 + t;
    if (sel_bit(sel, v)) red_vertex[sel_rank(sel, v) - soff] = v;
}
// (Measured on one GPU: keeping succ[v] and the label of v in ONE 16-byte entry, so that a step would touch a single
// sector, made this kernel 5-6x SLOWER, 14 -> 75-93 ms on C3, whichever of the load / store came first: stores into
// sectors whose fill is still in flight do not pipeline.  Successors and labels therefore stay in separate arrays.)
// one thread per own splitter: walk to the next splitter (only a hash-sampled one can be met: heads have no
// predecessor) or to the tail; every passed vertex learns (splitter, distance) in its owner's shard
__global__ void gs_walk_kernel(const u32 *__restrict__ succ, const u32 *__restrict__ red_vertex, const u64 *__restrict__ sel, u32 ns_own, u32 soff,
                               int sl, GsPeers pt, u64 *__restrict__ red_rec, u32 *__restrict__ red_tail) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns_own) return;
    const u32 i = soff + j;
    u32 u = red_vertex[j], d = 0;
    for (;;) {
        const u32 n = succ[u];
        if (n == V_NONE) { red_rec[i] = ((u64)d << 32) | (i | V_TERM); red_tail[i] = u; break; }
        if (hash_splitter(n, sl)) { red_rec[i] = ((u64)(d + 1) << 32) | sel_rank(sel, n); red_tail[i] = V_NONE; break; }
        ++d;
        const u32 q = n / pt.vchunk;
        pt.own[q][n - q * pt.vchunk] = ((u64)i << 32) | d;
        u = n;
    }
}
// own vertices: rec = dist-to-tail << 32 | tail | V_TERM; V_NONE for dead vertices; vertices of cycles raise *cyc
__global__ void gs_expand_kernel(const u64 *__restrict__ red_rec, const u32 *__restrict__ red_tail, const u64 *__restrict__ sel,
                                 const u64 *__restrict__ ownloc, const u8 *__restrict__ flags, u64 U, u32 v0, u32 nvl, u64 *__restrict__ rec,
                                 u32 *__restrict__ cyc) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nvl) return;
    const u32 v = v0 + t;
    if (gs_dead(flags, v, U)) { rec[t] = (u64)V_NONE; return; }
    u32 i, back = 0;
    if (sel_bit(sel, v)) i = sel_rank(sel, v);
    else {
        const u64 ol = __ldcs(ownloc + t);
        i = (u32)(ol >> 32);
        back = (u32)ol;
    }
    u64 out = 0;   // non-terminal: on a cycle
    if (i != V_NONE) {
        const u64 r = red_rec[i];
        const u32 ptr = (u32)r;
        if (ptr & V_TERM) out = ((u64)((u32)(r >> 32) - back) << 32) | (red_tail[ptr & ~V_TERM] | V_TERM);
    }
    if (out == 0) *cyc = 1u;
    rec[t] = out;
}
// own vertices: head and rank inside the path; the deciding vertex of a node appends (creation key, head, first
// emitted rank, emitted count) to the rank's list and marks the key (see resolve_kernel in graph.cuh for the rules)
template <int W>
__global__ void __launch_bounds__(256) gs_resolve_kernel(const u64 *__restrict__ keys, const u8 *__restrict__ flags, const u64 *__restrict__ rec,
                                                         u32 v0, u32 nvl, int k, u32 *__restrict__ vhead, u32 *__restrict__ vrank,
                                                         uint4 *__restrict__ list, u32 *__restrict__ list_n, u32 *__restrict__ nodemap) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nvl) return;
    const u32 v = v0 + t;
    const u64 r = rec[t];
    const u32 ptr = (u32)r;
    if (ptr == V_NONE) { vhead[t] = V_NONE; vrank[t] = 0; return; }
    const u32 tail = ptr & ~V_TERM, dist = (u32)(r >> 32);
    const bool pal = flags && ((flags[v >> 1] >> 6) & 1u);
    const u64 rm = rec[pal ? t : (t ^ 1u)];
    const u32 head = gs_mirror(flags, ((u32)rm) & ~V_TERM);
    const u32 rank = (u32)(rm >> 32);
    const u32 n = rank + dist + 1;
    vhead[t] = head;
    vrank[t] = rank;
    const bool sm = gs_mirror(flags, tail) == head;  // the path is its own mirror image
    uint4 out;
    bool emit = false;
    if (!sm) {
        if (rank == 0) {
            // canonical!(s): keep the orientation whose first k-mer is smaller (src/dbg.jl:156)
            const Kmer<W> a = vertex_kmer<W>(keys, v, k), b = vertex_kmer<W>(keys, gs_mirror(flags, tail), k);
            if (km_lt<W>(a, b)) { out = make_uint4(min(v >> 1, tail >> 1), v, 0u, n); emit = true; }
        }
    } else {
        // the walk stops when it meets its own reverse strand (src/dbg.jl:147-150): the node is the first ceil(n/2)
        // vertices, stored in whichever orientation is canonical
        const u32 h = (n + 1) / 2;
        if (rank == n - h) {
            const Kmer<W> a = vertex_kmer<W>(keys, head, k), b = vertex_kmer<W>(keys, v, k);
            out = make_uint4(head >> 1, head, km_cmp<W>(a, b) <= 0 ? 0u : (n - h), h);
            emit = true;
        }
    }
    if (emit) {
        list[atomicAdd(list_n, 1u)] = out;
        atomicOr(&nodemap[out.x >> 5], 1u << (out.x & 31u));
    }
}
__global__ void gs_nodefill_kernel(const uint4 *__restrict__ list, u32 n, const u64 *__restrict__ nsel, int k, uint4 *__restrict__ ntab) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 e = list[i];
    ntab[sel_rank(nsel, e.x)] = make_uint4(e.y, e.z, e.w, (u32)k + e.w - 1u);   // head, first emitted rank, count, bases
}
__global__ void gs_nodelen_kernel(const uint4 *__restrict__ ntab, u64 nn, u32 *__restrict__ len) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) len[i] = ntab[i].w;
}
// the node's first-base offset takes the place of its length: emission then needs ONE random access per vertex
__global__ void gs_nodeoff_kernel(const u32 *__restrict__ off, u64 nn, uint4 *__restrict__ ntab) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) ntab[i].w = off[i];
}
// every own vertex of an emitted range writes its base(s)
template <int W>
__global__ void __launch_bounds__(256) gs_emit_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ rec, const u32 *__restrict__ vhead,
                                                      const u32 *__restrict__ vrank, u32 v0, u32 nvl, int k, const u64 *__restrict__ nsel,
                                                      const uint4 *__restrict__ ntab, u32 *__restrict__ packed,
                                                      u32 *__restrict__ node_first, u32 *__restrict__ node_last) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nvl) return;
    // (the per-vertex arrays are read once: streaming loads keep L2 for the node table, the rank structure and the packed bases)
    const u32 head = __ldcs(vhead + t);
    if (head == V_NONE) return;
    const u32 v = v0 + t;
    const u32 tail = ((u32)__ldcs(rec + t)) & ~V_TERM;
    const u32 id = sel_rank(nsel, min(head >> 1, tail >> 1));
    const uint4 pr = ntab[id];   // (head, first emitted rank, emitted count, offset of the first base) of the node keyed here
    if (pr.x != head) return;    // the mirror orientation is the emitted one
    const u32 rank = __ldcs(vrank + t), re = pr.y, cnt = pr.z;
    if (rank < re || rank - re >= cnt) return;
    const u32 j = rank - re;
    const u64 off = pr.w;
    const Kmer<W> x = vertex_kmer<W>(keys, v, k);
    // 2-bit codes OR-ed into a zeroed packed array (16 bases per word): the vertices of a node sit at random table
    // positions, and the packed array of all nodes is small enough to stay in L2, where the atomics are served
    if (j == 0) {
        for (int q = 0; q < k; ++q) {
            const u32 c = km_base<W>(x, q, k);
            if (c) atomicOr(&packed[(off + q) >> 4], c << (2 * ((off + q) & 15)));
        }
        node_first[id] = v;
    } else {
        const u32 c = (u32)(x.w[W - 1] & 3ULL);
        const u64 p = off + k - 1 + j;
        if (c) atomicOr(&packed[p >> 4], c << (2 * (p & 15)));
    }
    if (j == cnt - 1) node_last[id] = v;
}
// packed 2-bit codes -> ASCII, 16 bases per thread
__global__ void gs_unpack_kernel(const u32 *__restrict__ packed, u64 nbases, u8 *__restrict__ bases) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 16 >= nbases) return;
    const u32 x = packed[w];
    u32 o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        u32 v = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) v |= (u32)code_to_ascii((x >> (2 * (4 * q + e))) & 3u) << (8 * e);
        o[q] = v;
    }
    if (w * 16 + 16 <= nbases) reinterpret_cast<uint4 *>(bases)[w] = make_uint4(o[0], o[1], o[2], o[3]);
    else for (u64 e = 0; w * 16 + e < nbases; ++e) bases[w * 16 + e] = (u8)(o[e >> 2] >> (8 * (e & 3)));
}

// walk labels a rank wrote for vertices of ANOTHER rank's range: compacted to (vertex index inside the range, label)
struct LabelF {
    const u64 *lab;   // the range of the full-size label array
    u32 *oidx;
    u64 *olab;
    __device__ bool pred(u64 i) const { return lab[i] != ~0ULL; }
    __device__ void emit(u64 i, u64 dst) const { oidx[dst] = (u32)i; olab[dst] = lab[i]; }
};
__global__ void gs_apply_labels_kernel(const u32 *__restrict__ idx, const u64 *__restrict__ lab, u64 n, u64 *__restrict__ shard) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) shard[idx[i]] = lab[i];
}
// rank p's piece buf[off[p] .. + cnt[p]) (elements of esz bytes) -> every rank, in place
static void dist_allgatherv_inplace(gg_comm *cm, void *buf, const u64 *off, const u64 *cnt, size_t esz) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int me = cm->rank;
    NCK(nc->GroupStart());
    for (int p = 0; p < cm->world; ++p) {
        if (p == me) continue;
        if (cnt[me]) NCK(nc->Send((const u8 *)buf + off[me] * esz, cnt[me] * esz, ncclUint8, p, cm->comm, ctx->stream));
        if (cnt[p]) NCK(nc->Recv((u8 *)buf + off[p] * esz, cnt[p] * esz, ncclUint8, p, cm->comm, ctx->stream));
    }
    NCK(nc->GroupEnd());
}
// bitmap (u32 words) -> rank-in-bitmap entries; returns the number of set bits; optional: prefix at P word positions
static u64 gs_build_sel(gg_ctx *ctx, const u32 *bits, u64 nw, DBuf<u64> &sel, u64 pick_stride = 0, int npick = 0, u32 *h_pick = nullptr) {
    DBuf<u32> pref(ctx, nw + 1), picked(ctx, 16);
    DBuf<u64> tot(ctx, 1);
    LAUNCH(ctx, gs_popc_kernel, (unsigned)ceil_div(nw, 256), 256, 0, bits, nw, pref.p);
    exclusive_scan_u32(ctx, pref.p, pref.p, nw, tot.p);
    sel.alloc(ctx, nw + 1);
    LAUNCH(ctx, gs_sel_kernel, (unsigned)ceil_div(nw, 256), 256, 0, bits, pref.p, nw, sel.p);
    if (npick) {
        LAUNCH(ctx, gs_pick_kernel, 1, 32, 0, pref.p, pick_stride, npick, picked.p);
        CK(cudaMemcpyAsync(ctx->h_scalar + 8, picked.p, npick * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    }
    const u64 total = read_scalar<u64>(ctx, tot.p);
    if (npick) memcpy(h_pick, ctx->h_scalar + 8, npick * sizeof(u32));
    return total;
}

// keys: the whole table of U sorted unique canonical k-mers on every rank.  Returns nullptr (on all ranks alike) when
// the input holds perfect circles: the caller then takes graph.cuh's formulation.
template <int W>
static gg_graph *build_graph_sharded(gg_ctx *ctx, const u64 *keys, u64 U, int k, gg_comm *cm) {
    NcclApi *nc = (cm && cm->world > 1) ? nccl_api() : nullptr;
    const int P = nc ? cm->world : 1, me = nc ? cm->rank : 0;
    if (U == 0 || U >= (1ULL << 30)) return build_graph<W>(ctx, keys, U, k, cm);   // (empty graph / the limit error)
    gg_graph *g = new gg_graph();
    g->ctx = ctx;
    g->k = k;
    try {
        const u64 chunk = (ceil_div(U, (u64)P) + 15) & ~15ULL;   // k-mers per rank: 2 * chunk vertices, a multiple of 32
        const u64 i0 = chunk * me < U ? chunk * me : U, i1 = i0 + chunk < U ? i0 + chunk : U;
        const u32 nvl = (u32)(2 * chunk), v0 = (u32)(2 * chunk * me);
        const u64 nvp = 2 * chunk * P;   // padded vertex count
        // splitters: heads + a 2^-sl hash sample.  Fewer splitters shrink the reduced list (ranked on every rank) and make its
        // records cache-resident for the expansion, as long as every rank keeps >= ~2^18 walks in flight (measured on C2 / C3)
        int sl = 4;
        while (sl < 7 && (nvp / (u64)P) >> (sl + 1) >= (1ULL << (P > 1 ? 17 : 18))) ++sl;   // (the list is ranked on EVERY rank: keep it short)
        if (getenv("GG_SPLIT_LOG")) sl = atoi(getenv("GG_SPLIT_LOG"));
        // ---- K5: neighbour lookup of the own k-mers ----
        stage_begin(ctx, ST_LOOKUP);
        int pbits = ilog2_ceil(U) - 1;
        if (pbits > 24) pbits = 24;
        if (pbits > 2 * k) pbits = 2 * k;
        if (pbits < 0) pbits = 0;
        DBuf<u8> flags;
        const bool even_k = (k & 1) == 0;
        DBuf<u32> nbF, nbB, succ;
        auto local1 = [&] {
            DBuf<u32> start(ctx, (1ULL << pbits) + 1);
            LAUNCH(ctx, prefix_index_kernel<W>, (unsigned)ceil_div(U + 1, 256), 256, 0, keys, U, k, pbits, start.p);
            flags.alloc(ctx, chunk * P);
            nbF.alloc(ctx, chunk); nbB.alloc(ctx, chunk);
            dfill(ctx, flags.p + chunk * me, chunk, 0u);
            // (lookup_kernel indexes its outputs by k-mer: the neighbour arrays are handed over shifted by i0)
            if (i1 > i0) LAUNCH(ctx, lookup_kernel<W>, (unsigned)ceil_div(8 * (i1 - i0), 256), 256, 0, keys, i0, i1, k, pbits, start.p, flags.p, nbF.p - i0, nbB.p - i0);
            CK(cudaStreamSynchronize(ctx->stream));   // `start` goes back to the allocator
        };
        if (nc) dist_guard(cm, "neighbour lookup", local1); else local1();
        if (nc) NCK(nc->AllGather(flags.p + chunk * me, flags.p, chunk, ncclUint8, cm->comm, ctx->stream));
        stage_end(ctx, ST_LOOKUP);
        // ---- K6: successor pointers ----
        stage_begin(ctx, ST_UNITIG);
        DBuf<u32> bits;
        // walk labels: 0 = every rank labels into a full-size local array and the labels that belong to other ranks are
        // compacted and exchanged in bulk (default); 1 = peer stores into the owners' shards (CUDA IPC); 2 = full-size
        // arrays merged with ncclAllReduce(min).  Measured on 4 x B200 (C3): the 8-byte peer stores, random over the
        // owner's shard, reach 0.4 G stores/s per GPU -- 110 ms for the walks against 4 ms here.
        const int label_route = getenv("GG_GRAPH_LABELS") ? atoi(getenv("GG_GRAPH_LABELS")) : 0;
        bool ipc = false;
        DBuf<u64> own_full;
        GsPeers pt;
        memset(&pt, 0, sizeof(pt));
        pt.vchunk = nvl;
        if (nc && label_route == 1) ipc = dist_ensure_recv(cm, nvl);   // collective: the exchange buffer doubles as the label shard
        auto local2 = [&] {
            succ.alloc(ctx, nvp);
            LAUNCH(ctx, gs_succ_kernel, (unsigned)ceil_div(chunk, 256), 256, 0, flags.p, nbF.p, nbB.p, chunk * me, chunk, U, succ.p);
            nbF.release(); nbB.release();
            bits.alloc(ctx, nvp / 32 + 1);
            if (ipc) {
                for (int q = 0; q < P; ++q) pt.own[q] = cm->peerR[q];
                CK(cudaMemsetAsync(cm->R, 0xFF, (u64)nvl * sizeof(u64), ctx->stream));
            } else {
                own_full.alloc(ctx, nvp);
                for (int q = 0; q < P; ++q) pt.own[q] = own_full.p + (u64)q * nvl;
                CK(cudaMemsetAsync(own_full.p, 0xFF, nvp * sizeof(u64), ctx->stream));
            }
        };
        if (nc) dist_guard(cm, "successor pointers", local2); else local2();
        if (nc) NCK(nc->AllGather(succ.p + (u64)nvl * me, succ.p, nvl, ncclUint32, cm->comm, ctx->stream));
        // ---- K7: splitters, walks, reduced list ----
        LAUNCH(ctx, gs_flag_kernel, (unsigned)ceil_div(nvl, 256), 256, 0, succ.p, even_k ? flags.p : (const u8 *)nullptr, U, v0, nvl, sl, bits.p);
        if (nc) NCK(nc->AllGather(bits.p + (u64)(nvl / 32) * me, bits.p, nvl / 32, ncclUint32, cm->comm, ctx->stream));
        DBuf<u64> sel, redA, redB;
        DBuf<u32> red_tail, red_vertex;
        u32 soffs[MSD_MAX_PEERS + 1] = {0};
        u64 ns = 0;
        u32 *cyc = nullptr;
        DBuf<u32> cycb;
        DBuf<u64> rec;
        auto local3 = [&] {
            ns = gs_build_sel(ctx, bits.p, nvp / 32, sel, nvl / 32, P, soffs);
            soffs[P] = (u32)ns;
            if (ns >= 0x7FFFFFFFull) GG_THROW(GG_ERR_LIMIT, "too many splitters (%llu)", (unsigned long long)ns);
            const u32 ns_own = soffs[me + 1] - soffs[me];
            redA.alloc(ctx, ns + 1); redB.alloc(ctx, ns + 1); red_tail.alloc(ctx, ns + 1); red_vertex.alloc(ctx, (u64)ns_own + 1);
            LAUNCH(ctx, gs_list_kernel, (unsigned)ceil_div(nvl, 256), 256, 0, sel.p, v0, nvl, soffs[me], red_vertex.p);
            if (ns_own) LAUNCH(ctx, gs_walk_kernel, (unsigned)ceil_div(ns_own, 128), 128, 0, succ.p, red_vertex.p, sel.p, ns_own, soffs[me], sl, pt, redA.p, red_tail.p);
            cycb.alloc(ctx, 4);
            dfill(ctx, cycb.p, 16, 0u);
            rec.alloc(ctx, nvl);
        };
        if (nc) dist_guard(cm, "splitter walks", local3); else local3();
        if (nc) {
            u64 off[MSD_MAX_PEERS], cnt[MSD_MAX_PEERS];
            for (int q = 0; q < P; ++q) { off[q] = soffs[q]; cnt[q] = soffs[q + 1] - soffs[q]; }
            dist_allgatherv_inplace(cm, redA.p, off, cnt, sizeof(u64));      // (also: every rank's peer stores have landed)
            dist_allgatherv_inplace(cm, red_tail.p, off, cnt, sizeof(u32));
            if (!ipc && label_route == 2) NCK(nc->AllReduce(own_full.p, own_full.p, nvp, ncclUint64, ncclMin, cm->comm, ctx->stream));
            if (!ipc && label_route != 2) {
                // ---- labels of foreign vertices: compacted per owner, counts all-gathered, one exchange, applied by the owner ----
                DBuf<u32> sidx_, ridx_;
                DBuf<u64> slab_, rlab_, cntd;
                std::vector<u64> scnt(P, 0), soff(P, 0), rcnt(P, 0), roff(P, 0), allc((size_t)P * P, 0);
                dist_guard(cm, "label compaction", [&] {
                    cntd.alloc(ctx, (u64)P * P + P);
                    std::vector<std::unique_ptr<Compactor<LabelF>>> cps(P);
                    u64 run = 0;
                    for (int q = 0; q < P; ++q) {
                        soff[q] = run;
                        if (q == me) continue;
                        cps[q].reset(new Compactor<LabelF>(ctx, nvl, "label"));
                        scnt[q] = cps[q]->count(LabelF{own_full.p + (u64)q * nvl, nullptr, nullptr});
                        run += scnt[q];
                    }
                    sidx_.alloc(ctx, run + 1); slab_.alloc(ctx, run + 1);
                    for (int q = 0; q < P; ++q)
                        if (q != me) cps[q]->emit(LabelF{own_full.p + (u64)q * nvl, sidx_.p + soff[q], slab_.p + soff[q]});
                    CK(cudaStreamSynchronize(ctx->stream));   // the compactors' tile counts go back to the allocator
                    CK(cudaMemcpyAsync(cntd.p + (u64)P * P, scnt.data(), P * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
                });
                NCK(nc->AllGather(cntd.p + (u64)P * P, cntd.p, P, ncclUint64, cm->comm, ctx->stream));
                CK(cudaMemcpyAsync(allc.data(), cntd.p, (size_t)P * P * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                u64 nrecv = 0;
                for (int p = 0; p < P; ++p) { roff[p] = nrecv; rcnt[p] = allc[(size_t)p * P + me]; nrecv += rcnt[p]; }
                dist_guard(cm, "label receive buffers", [&] { ridx_.alloc(ctx, nrecv + 1); rlab_.alloc(ctx, nrecv + 1); });
                dist_alltoallv(cm, sidx_.p, soff.data(), scnt.data(), ridx_.p, roff.data(), rcnt.data(), sizeof(u32));
                dist_alltoallv(cm, slab_.p, soff.data(), scnt.data(), rlab_.p, roff.data(), rcnt.data(), sizeof(u64));
                if (nrecv) LAUNCH(ctx, gs_apply_labels_kernel, (unsigned)ceil_div(nrecv, 256), 256, 0, ridx_.p, rlab_.p, nrecv, own_full.p + (u64)nvl * me);
                CK(cudaStreamSynchronize(ctx->stream));   // the exchange buffers go back to the allocator
            }
        }
        const u64 *ownloc = ipc ? cm->R : (own_full.p + (u64)nvl * me);
        if (!nc) ownloc = own_full.p;
        cyc = cycb.p;
        const u64 *ra = rank_reduced_list(ctx, redA.p, redB.p, ns);
        LAUNCH(ctx, gs_expand_kernel, (unsigned)ceil_div(nvl, 256), 256, 0, ra, red_tail.p, sel.p, ownloc, even_k ? flags.p : (const u8 *)nullptr, U, v0, nvl, rec.p, cyc);
        if (nc) NCK(nc->AllReduce(cyc, cyc, 1, ncclUint32, ncclMax, cm->comm, ctx->stream));
        if (read_scalar<u32>(ctx, cyc) != 0) {   // perfect circles somewhere: the replicated formulation handles them
            delete g;
            return nullptr;
        }
        succ.release(); redA.release(); redB.release(); red_tail.release(); red_vertex.release(); sel.release(); bits.release(); own_full.release();
        // ---- nodes: creation keys, ids, table ----
        DBuf<u32> vhead, vrank, nodemap, list_n;
        DBuf<uint4> list;
        const u64 nmw = ceil_div(U, 32);
        auto local4 = [&] {
            vhead.alloc(ctx, nvl); vrank.alloc(ctx, nvl); nodemap.alloc(ctx, nmw + 1); list_n.alloc(ctx, 4); list.alloc(ctx, chunk + 1);
            dfill(ctx, nodemap.p, (nmw + 1) * sizeof(u32), 0u);
            dfill(ctx, list_n.p, 16, 0u);
            LAUNCH(ctx, gs_resolve_kernel<W>, (unsigned)ceil_div(nvl, 256), 256, 0, keys, even_k ? flags.p : (const u8 *)nullptr, rec.p, v0, nvl, k, vhead.p, vrank.p, list.p, list_n.p, nodemap.p);
        };
        if (nc) dist_guard(cm, "node discovery", local4); else local4();
        if (nc) NCK(nc->AllReduce(nodemap.p, nodemap.p, nmw, ncclUint32, ncclSum, cm->comm, ctx->stream));   // disjoint bits: sum = or
        DBuf<u64> nsel;
        DBuf<uint4> ntab;
        u64 nn = 0;
        auto local5 = [&] {
            nn = gs_build_sel(ctx, nodemap.p, nmw, nsel);
            const u32 nl = read_scalar<u32>(ctx, list_n.p);
            ntab.alloc(ctx, nn + 1);
            dfill(ctx, ntab.p, (nn + 1) * sizeof(uint4), 0u);
            if (nl) LAUNCH(ctx, gs_nodefill_kernel, (unsigned)ceil_div(nl, 256), 256, 0, list.p, nl, nsel.p, k, ntab.p);
        };
        if (nc) dist_guard(cm, "node table", local5); else local5();
        if (nc && nn) NCK(nc->AllReduce(ntab.p, ntab.p, 4 * nn, ncclUint32, ncclSum, cm->comm, ctx->stream));
        g->n_nodes = nn;
        DBuf<u32> node_off32, node_first, node_last, packed;
        u64 total_bases = 0;
        auto local6 = [&] {
            DBuf<u32> node_len(ctx, nn + 1);
            DBuf<u64> tot(ctx, 1);
            node_off32.alloc(ctx, nn + 1); node_first.alloc(ctx, nn + 1); node_last.alloc(ctx, nn + 1);
            if (nn) LAUNCH(ctx, gs_nodelen_kernel, (unsigned)ceil_div(nn, 256), 256, 0, ntab.p, nn, node_len.p);
            exclusive_scan_u32(ctx, node_len.p, node_off32.p, nn, tot.p);
            total_bases = read_scalar<u64>(ctx, tot.p);
            if (total_bases >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "node sequences exceed 2^32 bases (%llu)", (unsigned long long)total_bases);
            g->total_bases = total_bases;
            g->d_bases = dalloc<u8>(ctx, total_bases + 16);
            g->d_node_off = dalloc<u64>(ctx, nn + 1);
            LAUNCH(ctx, widen_offsets_kernel, (unsigned)ceil_div(nn + 1, 256), 256, 0, node_off32.p, nn, total_bases, g->d_node_off);
            packed.alloc(ctx, total_bases / 16 + 4);
            dfill(ctx, packed.p, (total_bases / 16 + 1) * sizeof(u32), 0u);
            if (nc) {   // every rank contributes the codes / end vertices of its own vertices: zero elsewhere, summed below
                dfill(ctx, node_first.p, (nn + 1) * sizeof(u32), 0u);
                dfill(ctx, node_last.p, (nn + 1) * sizeof(u32), 0u);
            }
            if (nn) LAUNCH(ctx, gs_nodeoff_kernel, (unsigned)ceil_div(nn, 256), 256, 0, node_off32.p, nn, ntab.p);
            LAUNCH(ctx, gs_emit_kernel<W>, (unsigned)ceil_div(nvl, 256), 256, 0, keys, rec.p, vhead.p, vrank.p, v0, nvl, k, nsel.p, ntab.p,
                   packed.p, node_first.p, node_last.p);
        };
        if (nc) dist_guard(cm, "node emission", local6); else local6();
        if (nc) {
            if (total_bases) NCK(nc->AllReduce(packed.p, packed.p, total_bases / 16 + 1, ncclUint32, ncclSum, cm->comm, ctx->stream));   // disjoint bits
            if (nn) {
                NCK(nc->AllReduce(node_first.p, node_first.p, nn, ncclUint32, ncclSum, cm->comm, ctx->stream));
                NCK(nc->AllReduce(node_last.p, node_last.p, nn, ncclUint32, ncclSum, cm->comm, ctx->stream));
            }
        }
        if (total_bases) LAUNCH(ctx, gs_unpack_kernel, (unsigned)ceil_div(ceil_div(total_bases, 16), 256), 256, 0, packed.p, total_bases, g->d_bases);
        stage_end(ctx, ST_UNITIG);
        // ---- K8: links (on every rank: two records per node) ----
        stage_begin(ctx, ST_LINKS);
        auto local7 = [&] { build_links<W>(ctx, g, keys, node_first.p, node_last.p); CK(cudaStreamSynchronize(ctx->stream)); };
        if (nc) dist_guard(cm, "links", local7); else local7();
        stage_end(ctx, ST_LINKS);
        return g;
    } catch (...) {
        dfree(ctx, g->d_node_off); dfree(ctx, g->d_bases); dfree(ctx, g->d_link_row); dfree(ctx, g->d_link_tri);
        delete g;
        throw;
    }
}

// every rank's sorted slice (slices in key order by rank) -> the whole table on every rank.  The slices are balanced, so
// they are padded to the largest one and moved with ONE ncclAllGather (ring / NVLS inside NCCL: a mesh of sends costs
// several times more at 8 ranks), then compacted.
__global__ void keys_compact_kernel(const u64 *__restrict__ pk, u64 maxw, int P, const u64 *__restrict__ nw, const u64 *__restrict__ offw, u64 *__restrict__ keys) {
    for (int p = 0; p < P; ++p) {
        const u64 np = nw[p], o = offw[p];
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (u64)gridDim.x * blockDim.x) keys[o + i] = pk[(u64)p * maxw + i];
    }
}
template <int W>
static u64 keys_allgatherv(gg_comm *cm, const u64 *slice, u64 n_local, DBuf<u64> &keys) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world;
    DBuf<u64> sz;
    dist_guard(cm, "table sizes", [&] {
        sz.alloc(ctx, 3 * (u64)P + 1);
        CK(cudaMemcpyAsync(sz.p + 3 * P, &n_local, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    });
    NCK(nc->AllGather(sz.p + 3 * P, sz.p, 1, ncclUint64, cm->comm, ctx->stream));
    u64 n[MSD_MAX_PEERS], nwo[2 * MSD_MAX_PEERS], total = 0, maxn = 1;
    CK(cudaMemcpyAsync(n, sz.p, P * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < P; ++p) { nwo[p] = n[p] * W; nwo[P + p] = total * W; total += n[p]; maxn = n[p] > maxn ? n[p] : maxn; }
    DBuf<u64> mk, pk;
    dist_guard(cm, "gathered table", [&] {
        keys.alloc(ctx, (total ? total : 1) * W);
        mk.alloc(ctx, maxn * W); pk.alloc(ctx, (u64)P * maxn * W);
        CK(cudaMemcpyAsync(sz.p + P, nwo, 2 * P * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        if (n_local) CK(cudaMemcpyAsync(mk.p, slice, n_local * W * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    });
    NCK(nc->AllGather(mk.p, pk.p, maxn * W, ncclUint64, cm->comm, ctx->stream));
    LAUNCH(ctx, keys_compact_kernel, (unsigned)(ctx->num_sms * 8), 256, 0, pk.p, maxn * W, P, sz.p + P, sz.p + 2 * P, keys.p);
    CK(cudaStreamSynchronize(ctx->stream));   // nwo / the padded buffers
    return total;
}
