// prims.cuh -- device-wide building blocks: exclusive scan, order-preserving compaction.
#pragma once
#include "common.cuh"

#define FULL_MASK 0xFFFFFFFFu

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, v, d);
        if ((int)lane_id() >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}

// block-wide exclusive scan of one u32 per thread; blockDim.x <= 1024, multiple of 32.
// `wsum` is a shared array of >= 33 u32. Returns exclusive prefix; *total gets the block sum.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *wsum, u32 *total) {
    u32 inc = warp_incl_scan(v);
    u32 wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();  // protect wsum reuse
    if (lane_id() == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u32 s = lane_id() < nw ? wsum[lane_id()] : 0;
        u32 si = warp_incl_scan(s);
        wsum[lane_id()] = si - s;
        if (lane_id() == 31) wsum[32] = si;
    }
    __syncthreads();
    u32 r = wsum[wid] + inc - v;
    *total = wsum[32];
    return r;
}

// ------------------------------- exclusive scan (u32) ------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const u32 *__restrict__ in, u64 n, u32 *__restrict__ partial) {
    __shared__ u32 ws[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE;
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        u64 idx = base + (u64)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    u32 tot;
    block_excl_scan(s, ws, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// single block: exclusive scan of partial[0..nb) in place; total (u64) to *total
__global__ void __launch_bounds__(1024) scan_partials_kernel(u32 *partial, u64 nb, u64 *total) {
    __shared__ u32 ws[33];
    u64 run = 0;
    for (u64 base = 0; base < nb; base += 1024) {
        u64 idx = base + threadIdx.x;
        u32 v = idx < nb ? partial[idx] : 0;
        u32 tot;
        u32 ex = block_excl_scan(v, ws, &tot);
        if (idx < nb) partial[idx] = (u32)(run + ex);
        run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = run;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(const u32 *__restrict__ in, u32 *__restrict__ out, u64 n,
                                                                  const u32 *__restrict__ partial) {
    __shared__ u32 ws[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;  // blocked: thread owns 8 consecutive
    u32 v[SCAN_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    u32 tot;
    u32 ex = block_excl_scan(s, ws, &tot) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}
// small arrays (bucket / child / tile tables): one CTA, one launch
#define SCAN_SMALL_N 65536
__global__ void __launch_bounds__(1024) scan_small_kernel(const u32 *__restrict__ in, u32 *__restrict__ out, u32 n, u64 *total) {
    __shared__ u32 ws[33];
    u64 run = 0;
    for (u32 base = 0; base < n; base += 4096) {  // 4 consecutive items per thread
        const u32 i0 = base + threadIdx.x * 4;
        u32 v[4], s = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = (i0 + j < n) ? in[i0 + j] : 0; s += v[j]; }
        u32 tot;
        u32 ex = block_excl_scan(s, ws, &tot) + (u32)run;
#pragma unroll
        for (int j = 0; j < 4; ++j) { if (i0 + j < n) out[i0 + j] = ex; ex += v[j]; }
        run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = run;
}
// out may alias in. d_total (device u64, optional) receives the grand total (exact in u64 as
// long as each tile sum fits u32; offsets themselves are u32 -> callers keep totals < 2^32).
static void exclusive_scan_u32(gg_ctx *ctx, const u32 *in, u32 *out, u64 n, u64 *d_total) {
    if (n == 0) {
        if (d_total) CK(cudaMemsetAsync(d_total, 0, sizeof(u64), ctx->stream));
        return;
    }
    if (n <= SCAN_SMALL_N) {
        LAUNCH(ctx, scan_small_kernel, 1, 1024, 0, in, out, (u32)n, d_total);
        return;
    }
    u64 nb = ceil_div(n, SCAN_TILE);
    DBuf<u32> partial(ctx, nb);
    LAUNCH(ctx, scan_reduce_kernel, (unsigned)nb, SCAN_THREADS, 0, in, n, partial.p);
    LAUNCH(ctx, scan_partials_kernel, 1, 1024, 0, partial.p, nb, d_total);
    LAUNCH(ctx, scan_final_kernel, (unsigned)nb, SCAN_THREADS, 0, in, out, n, partial.p);
}

// ------------------------- order-preserving stream compaction -----------------------------
// F provides:  __device__ bool pred(u64 i) const;  __device__ void emit(u64 i, u64 dst) const;
// Tile = 256 threads x 8 items; a warp owns 256 consecutive indices (8 coalesced rounds).
#define CP_THREADS 256
#define CP_ITEMS 8
#define CP_TILE (CP_THREADS * CP_ITEMS)

template <class F>
__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(F f, u64 n, u32 *__restrict__ tile_counts) {
    __shared__ u32 ws[33];
    u32 wid = threadIdx.x >> 5;
    u64 wbase = (u64)blockIdx.x * CP_TILE + (u64)wid * (32 * CP_ITEMS);
    u32 c = 0;
#pragma unroll
    for (int it = 0; it < CP_ITEMS; ++it) {
        u64 i = wbase + (u64)it * 32 + lane_id();
        if (i < n && f.pred(i)) ++c;
    }
    u32 tot;
    block_excl_scan(c, ws, &tot);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}
template <class F>
__global__ void __launch_bounds__(CP_THREADS) compact_emit_kernel(F f, u64 n, const u32 *__restrict__ tile_offsets) {
    __shared__ u32 ws[33];
    u32 wid = threadIdx.x >> 5;
    u64 wbase = (u64)blockIdx.x * CP_TILE + (u64)wid * (32 * CP_ITEMS);
    u32 flags = 0, c = 0;
#pragma unroll
    for (int it = 0; it < CP_ITEMS; ++it) {
        u64 i = wbase + (u64)it * 32 + lane_id();
        bool p = (i < n) && f.pred(i);
        flags |= (p ? 1u : 0u) << it;
        c += p;
    }
    // per-warp totals -> block exclusive scan over warps
    u32 wtot = warp_sum(c);
    u32 tot;
    u32 wex = block_excl_scan(lane_id() == 0 ? wtot : 0, ws, &tot);
    wex = __shfl_sync(FULL_MASK, wex, 0);
    u64 dst = (u64)tile_offsets[blockIdx.x] + wex;
#pragma unroll
    for (int it = 0; it < CP_ITEMS; ++it) {
        bool p = (flags >> it) & 1u;
        u32 b = __ballot_sync(FULL_MASK, p);
        if (p) f.emit(wbase + (u64)it * 32 + lane_id(), dst + __popc(b & lanemask_lt()));
        dst += __popc(b);
    }
}
// returns (via *d_total, device u64) the number of selected items; emits them in index order.
// The caller must make sure the output is large enough (two-phase: count first if unknown).
template <class F>
struct Compactor {
    gg_ctx *ctx;
    u64 n, ntiles;
    DBuf<u32> tiles;
    DBuf<u64> total;
    std::string tag;
    Compactor(gg_ctx *c, u64 n_, const char *tag_ = "compact")
        : ctx(c), n(n_), ntiles(ceil_div(n_ ? n_ : 1, CP_TILE)), tiles(c, ntiles), total(c, 1), tag(tag_) {}
    // phase 1: count + scan; returns the total on the host (one stream sync)
    u64 count(const F &f) {
        if (n == 0) return 0;
        LAUNCH_NAMED(ctx, tag + "_count_kernel", compact_count_kernel<F>, (unsigned)ntiles, CP_THREADS, 0, f, n, tiles.p);
        exclusive_scan_u32(ctx, tiles.p, tiles.p, ntiles, total.p);
        return read_scalar<u64>(ctx, total.p);
    }
    // phase 2
    void emit(const F &f) {
        if (n == 0) return;
        LAUNCH_NAMED(ctx, tag + "_emit_kernel", compact_emit_kernel<F>, (unsigned)ntiles, CP_THREADS, 0, f, n, tiles.p);
    }
};
