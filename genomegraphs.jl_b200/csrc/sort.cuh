// sort.cuh -- K3/K4: stable LSD radix sort (keys of KW u64 words, optional u64 payload),
// run-length counting and min-count filtering.
//
// Replaces (reference): KmerAnalysis serial_mem collect -> sort! -> collapse
// (docs/src/man/assembly.md:126-128), the same sort + run-length pattern at
// src/GraphIndexes.jl:88-90,43-60, sort!(in)/sort!(out) at src/dbg.jl:219-220 and
// filter!(freq >= min_freq) at src/dbg.jl:275.
#pragma once
#include "prims.cuh"

struct BitRange {
    int lo, hi;  // key bits [lo, hi), counted from the LSB of the whole KW*64-bit key
};

template <int KW>
__device__ __forceinline__ u32 key_digit(const Kmer<KW> &x, int bit, int nbits) {
    int wi = KW - 1 - (bit >> 6), sh = bit & 63;
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        if (i == wi) lo = x.w[i];
        if (i == wi - 1) hi = x.w[i];
    }
    u64 v = lo >> sh;
    if (sh > 56) v |= hi << (64 - sh);
    return (u32)(v & ((1u << nbits) - 1u));
}

template <int KW>
struct RadixCfg {
    static constexpr int ITEMS = (16 / KW) < 3 ? 3 : (16 / KW);
    static constexpr int TILE = 256 * ITEMS;
};

template <int KW>
__global__ void __launch_bounds__(256) radix_hist_kernel(const u64 *__restrict__ keys, u64 n, int bit, int nbits,
                                                         u32 *__restrict__ block_hist, u64 ntiles) {
    constexpr int TILE = RadixCfg<KW>::TILE;
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u64 tbase = (u64)blockIdx.x * TILE;
#pragma unroll 4
    for (int i = threadIdx.x; i < TILE; i += 256) {
        u64 idx = tbase + i;
        if (idx < n) atomicAdd(&h[key_digit<KW>(km_load<KW>(keys, idx), bit, nbits)], 1u);
    }
    __syncthreads();
    block_hist[(u64)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <int KW, bool PAY>
__global__ void __launch_bounds__(256) radix_scatter_kernel(const u64 *__restrict__ kin, u64 *__restrict__ kout,
                                                            const u64 *__restrict__ pin, u64 *__restrict__ pout, u64 n, int bit,
                                                            int nbits, const u32 *__restrict__ scanned, u64 ntiles) {
    constexpr int ITEMS = RadixCfg<KW>::ITEMS;
    constexpr int TILE = RadixCfg<KW>::TILE;
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);
    u64 *spay = skeys + (size_t)TILE * KW;
    __shared__ u32 whist[8][256];
    __shared__ u32 lbase[256];
    __shared__ u32 gbase[256];
    __shared__ u32 ws[33];
    const u32 tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const u64 tile = blockIdx.x, tbase = tile * TILE, wbase = tbase + (u64)wid * (32 * ITEMS);
    for (int i = tid; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0;
    __syncthreads();
    Kmer<KW> key[ITEMS];
    u64 pay[ITEMS];
    u32 dg[ITEMS], rk[ITEMS];
    u32 inmask = 0;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        u64 idx = wbase + (u64)it * 32 + lane;
        bool in = idx < n;
        dg[it] = 0; rk[it] = 0;
        if (in) {
            key[it] = km_load<KW>(kin, idx);
            if (PAY) pay[it] = pin[idx];
        }
        u32 act = __ballot_sync(FULL_MASK, in);
        if (in) {
            u32 d = key_digit<KW>(key[it], bit, nbits);
            u32 peers = __match_any_sync(act, d);
            int leader = __ffs(peers) - 1;
            u32 old = 0;
            if ((int)lane == leader) {
                old = whist[wid][d];
                whist[wid][d] = old + __popc(peers);
            }
            old = __shfl_sync(peers, old, leader);
            rk[it] = old + __popc(peers & lanemask_lt());
            dg[it] = d;
            inmask |= 1u << it;
        }
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive scan over warps of digit d, then over digits
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { u32 t = whist[w][tid]; whist[w][tid] = run; run += t; }
        u32 tot;
        u32 ex = block_excl_scan(run, ws, &tot);
        lbase[tid] = ex;
        gbase[tid] = scanned[(u64)tid * ntiles + tile];
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if ((inmask >> it) & 1u) {
            u32 pos = lbase[dg[it]] + whist[wid][dg[it]] + rk[it];
#pragma unroll
            for (int j = 0; j < KW; ++j) skeys[(size_t)pos * KW + j] = key[it].w[j];
            if (PAY) spay[pos] = pay[it];
        }
    }
    __syncthreads();
    u64 rem = n - tbase;
    u32 cnt = rem < (u64)TILE ? (u32)rem : (u32)TILE;
    for (u32 i = tid; i < cnt; i += 256) {
        Kmer<KW> x = km_load<KW>(skeys, i);
        u32 d = key_digit<KW>(x, bit, nbits);
        u64 dst = (u64)gbase[d] + (i - lbase[d]);
        km_store<KW>(kout, dst, x);
        if (PAY) pout[dst] = spay[i];
    }
}

// Small inputs (graph link records, tiny k-mer sets): ONE launch, one CTA, all digit passes inside
// (a pass of the tiled path costs five launches, which dominates below ~64 K keys).
#define RADIX_SMALL_N 32768
#define RADIX_MAX_PASSES 40
struct RadixPasses {
    int n;
    short bit[RADIX_MAX_PASSES];
    signed char nbits[RADIX_MAX_PASSES];
};
template <int KW, bool PAY>
__global__ void __launch_bounds__(256) radix_small_kernel(u64 *k0, u64 *k1, u64 *p0, u64 *p1, u32 n, RadixPasses pl) {
    constexpr int ITEMS = RadixCfg<KW>::ITEMS;
    constexpr int TILE = RadixCfg<KW>::TILE;
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);
    u64 *spay = skeys + (size_t)TILE * KW;
    __shared__ u32 whist[8][256];
    __shared__ u32 lbase[256];
    __shared__ u32 gbase[256];
    __shared__ u32 ws[33];
    const u32 tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    u64 *kin = k0, *kout = k1, *pin = p0, *pout = p1;
    for (int pass = 0; pass < pl.n; ++pass) {
        const int bit = pl.bit[pass], nbits = pl.nbits[pass];
        gbase[tid] = 0;
        __syncthreads();
        for (u32 i = tid; i < n; i += 256) atomicAdd(&gbase[key_digit<KW>(km_load<KW>(kin, i), bit, nbits)], 1u);
        __syncthreads();
        {
            u32 c = gbase[tid], tot;
            u32 ex = block_excl_scan(c, ws, &tot);
            gbase[tid] = ex;
        }
        __syncthreads();
        for (u32 tbase = 0; tbase < n; tbase += TILE) {
            const u32 wbase = tbase + wid * (32 * ITEMS);
            for (int i = tid; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0;
            __syncthreads();
            Kmer<KW> key[ITEMS];
            u64 pay[ITEMS];
            u32 dg[ITEMS], rk[ITEMS];
            u32 inmask = 0;
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                u32 idx = wbase + it * 32 + lane;
                bool in = idx < n;
                dg[it] = 0; rk[it] = 0;
                if (in) {
                    key[it] = km_load<KW>(kin, idx);
                    if (PAY) pay[it] = pin[idx];
                }
                u32 act = __ballot_sync(FULL_MASK, in);
                if (in) {
                    u32 d = key_digit<KW>(key[it], bit, nbits);
                    u32 peers = __match_any_sync(act, d);
                    int leader = __ffs(peers) - 1;
                    u32 old = 0;
                    if ((int)lane == leader) {
                        old = whist[wid][d];
                        whist[wid][d] = old + __popc(peers);
                    }
                    old = __shfl_sync(peers, old, leader);
                    rk[it] = old + __popc(peers & lanemask_lt());
                    dg[it] = d;
                    inmask |= 1u << it;
                }
                __syncwarp();
            }
            __syncthreads();
            u32 dtot;
            {
                u32 run = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { u32 t = whist[w][tid]; whist[w][tid] = run; run += t; }
                dtot = run;
                u32 tot;
                lbase[tid] = block_excl_scan(run, ws, &tot);
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                if ((inmask >> it) & 1u) {
                    u32 pos = lbase[dg[it]] + whist[wid][dg[it]] + rk[it];
#pragma unroll
                    for (int j = 0; j < KW; ++j) skeys[(size_t)pos * KW + j] = key[it].w[j];
                    if (PAY) spay[pos] = pay[it];
                }
            }
            __syncthreads();
            u32 rem = n - tbase;
            u32 cnt = rem < (u32)TILE ? rem : (u32)TILE;
            for (u32 i = tid; i < cnt; i += 256) {
                Kmer<KW> x = km_load<KW>(skeys, i);
                u32 d = key_digit<KW>(x, bit, nbits);
                u32 dst = gbase[d] + (i - lbase[d]);
                km_store<KW>(kout, dst, x);
                if (PAY) pout[dst] = spay[i];
            }
            __syncthreads();
            gbase[tid] += dtot;  // running base of digit `tid` for the next tile
            __syncthreads();
        }
        __threadfence();
        __syncthreads();
        u64 *t = kin; kin = kout; kout = t;
        t = pin; pin = pout; pout = t;
    }
}

// ---- rank sort: n <= RANK_SORT_N records, one launch over the whole GPU ---------------------------
// rank(i) = #{ j : key_j < key_i, or key_j == key_i and j < i } over the masked key bits: the same
// permutation as a stable LSD radix sort over those bits, in O(n^2 / threads) compares instead of
// ceil(bits / 8) dependent single-CTA passes.
#define RANK_SORT_N 32768
template <int KW>
struct KeyMask {
    u64 m[KW];
};
// 8 records per CTA, one warp per record: lane `sub` compares against the tile entries q = sub (mod 32)
template <int KW, bool PAY>
__global__ void __launch_bounds__(256) rank_sort_kernel(const u64 *__restrict__ kin, u64 *__restrict__ kout, const u64 *__restrict__ pin,
                                                        u64 *__restrict__ pout, u32 n, KeyMask<KW> km) {
    __shared__ u64 tile[256 * KW];
    const u32 i = blockIdx.x * 8 + (threadIdx.x >> 5), sub = threadIdx.x & 31u;
    const bool live = i < n;
    Kmer<KW> x, xm;
#pragma unroll
    for (int j = 0; j < KW; ++j) { x.w[j] = 0; xm.w[j] = 0; }
    if (live) {
        x = km_load<KW>(kin, i);
#pragma unroll
        for (int j = 0; j < KW; ++j) xm.w[j] = x.w[j] & km.m[j];
    }
    u32 rank = 0;
    for (u32 t0 = 0; t0 < n; t0 += 256) {
        __syncthreads();
        if (t0 + threadIdx.x < n) {
#pragma unroll
            for (int j = 0; j < KW; ++j) tile[threadIdx.x * KW + j] = kin[(u64)(t0 + threadIdx.x) * KW + j] & km.m[j];
        }
        __syncthreads();
        const u32 cnt = n - t0 < 256u ? n - t0 : 256u;
        if (live) {
#pragma unroll 4
            for (u32 q = sub; q < cnt; q += 32) {
                Kmer<KW> y;
#pragma unroll
                for (int j = 0; j < KW; ++j) y.w[j] = tile[q * KW + j];
                const int c = km_cmp<KW>(y, xm);
                rank += (c < 0 || (c == 0 && t0 + q < i)) ? 1u : 0u;
            }
        }
    }
    rank = warp_sum(rank);
    if (live && sub == 0) {
        km_store<KW>(kout, rank, x);
        if (PAY) pout[rank] = pin[i];
    }
}

// Stable LSD radix sort over the given bit ranges (least significant range first).
// On return the sorted data is in *keys (and *pay); *keys_alt / *pay_alt hold scratch.
template <int KW>
static void radix_sort(gg_ctx *ctx, u64 **keys, u64 **keys_alt, u64 **pay, u64 **pay_alt, u64 n,
                       const std::vector<BitRange> &ranges) {
    if (n < 2) return;
    if (n >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "radix_sort: %llu keys exceed the 2^32-1 limit", (unsigned long long)n);
    constexpr int TILE = RadixCfg<KW>::TILE;
    const bool has_pay = pay != nullptr && *pay != nullptr;
    u64 ntiles = ceil_div(n, TILE);
    DBuf<u32> hist(ctx, 256 * ntiles);
    size_t smem = (size_t)TILE * KW * 8 + (has_pay ? (size_t)TILE * 8 : 0);
    if (has_pay) {
        set_max_smem(ctx, radix_scatter_kernel<KW, true>, smem);
        set_max_smem(ctx, radix_small_kernel<KW, true>, smem);
    } else {
        set_max_smem(ctx, radix_scatter_kernel<KW, false>, smem);
        set_max_smem(ctx, radix_small_kernel<KW, false>, smem);
    }
    if (n <= RANK_SORT_N) {
        KeyMask<KW> km;
        for (int j = 0; j < KW; ++j) km.m[j] = 0;
        for (const BitRange &r : ranges)
            for (int bit = r.lo; bit < r.hi; ++bit) km.m[KW - 1 - (bit >> 6)] |= 1ULL << (bit & 63);
        // ranges are given least significant first and must not interleave within a word in a way that changes
        // the lexicographic order: every caller passes disjoint ranges in ascending significance
        if (has_pay) LAUNCH(ctx, (rank_sort_kernel<KW, true>), (unsigned)ceil_div(n, 8), 256, 0, *keys, *keys_alt, *pay, *pay_alt, (u32)n, km);
        else LAUNCH(ctx, (rank_sort_kernel<KW, false>), (unsigned)ceil_div(n, 8), 256, 0, *keys, *keys_alt, (const u64 *)nullptr, (u64 *)nullptr, (u32)n, km);
        std::swap(*keys, *keys_alt);
        if (has_pay) std::swap(*pay, *pay_alt);
        return;
    }
    if (n <= RADIX_SMALL_N) {
        RadixPasses pl;
        pl.n = 0;
        for (const BitRange &r : ranges)
            for (int bit = r.lo; bit < r.hi; bit += 8) {
                if (pl.n >= RADIX_MAX_PASSES) GG_THROW(GG_ERR_LIMIT, "too many radix passes");
                pl.bit[pl.n] = (short)bit;
                pl.nbits[pl.n] = (signed char)(r.hi - bit < 8 ? r.hi - bit : 8);
                ++pl.n;
            }
        if (has_pay) LAUNCH(ctx, (radix_small_kernel<KW, true>), 1, 256, smem, *keys, *keys_alt, *pay, *pay_alt, (u32)n, pl);
        else LAUNCH(ctx, (radix_small_kernel<KW, false>), 1, 256, smem, *keys, *keys_alt, (u64 *)nullptr, (u64 *)nullptr, (u32)n, pl);
        if (pl.n & 1) {
            std::swap(*keys, *keys_alt);
            if (has_pay) std::swap(*pay, *pay_alt);
        }
        return;
    }
    for (const BitRange &r : ranges) {
        for (int bit = r.lo; bit < r.hi; bit += 8) {
            int nbits = r.hi - bit < 8 ? r.hi - bit : 8;
            LAUNCH(ctx, radix_hist_kernel<KW>, (unsigned)ntiles, 256, 0, *keys, n, bit, nbits, hist.p, ntiles);
            exclusive_scan_u32(ctx, hist.p, hist.p, 256 * ntiles, nullptr);
            if (has_pay) {
                LAUNCH(ctx, (radix_scatter_kernel<KW, true>), (unsigned)ntiles, 256, smem, *keys, *keys_alt, *pay, *pay_alt, n, bit, nbits, hist.p, ntiles);
                std::swap(*pay, *pay_alt);
            } else {
                LAUNCH(ctx, (radix_scatter_kernel<KW, false>), (unsigned)ntiles, 256, smem, *keys, *keys_alt, (const u64 *)nullptr, (u64 *)nullptr, n, bit, nbits, hist.p, ntiles);
            }
            std::swap(*keys, *keys_alt);
        }
    }
}

// ---------------------------------- run-length counting ----------------------------------
template <int W>
struct RunHeadF {
    const u64 *keys;
    u64 *ukeys;
    u32 *headpos;
    __device__ bool pred(u64 i) const {
        if (i == 0) return true;
        return !km_eq<W>(km_load<W>(keys, i), km_load<W>(keys, i - 1));
    }
    __device__ void emit(u64 i, u64 dst) const {
        km_store<W>(ukeys, dst, km_load<W>(keys, i));
        headpos[dst] = (u32)i;
    }
};
__global__ void run_counts_kernel(const u32 *__restrict__ headpos, u64 nu, u64 n, u32 *__restrict__ counts) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    u64 next = (i + 1 < nu) ? headpos[i + 1] : n;
    counts[i] = (u32)(next - headpos[i]);
}
__global__ void fill_counts_kernel(u32 *counts, u64 n, u32 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = v;
}

// sorted keys[n] -> unique keys + counts (new buffers); returns U
template <int W>
static u64 run_length_count(gg_ctx *ctx, const u64 *sorted, u64 n, DBuf<u64> &ukeys, DBuf<u32> &counts) {
    if (n == 0) { ukeys.alloc(ctx, 1); counts.alloc(ctx, 1); return 0; }
    RunHeadF<W> f{sorted, nullptr, nullptr};
    Compactor<RunHeadF<W>> cp(ctx, n, "runhead");
    u64 nu = cp.count(f);
    ukeys.alloc(ctx, nu * W);
    DBuf<u32> headpos(ctx, nu);
    counts.alloc(ctx, nu);
    f.ukeys = ukeys.p;
    f.headpos = headpos.p;
    cp.emit(f);
    LAUNCH(ctx, run_counts_kernel, (unsigned)ceil_div(nu, 256), 256, 0, headpos.p, nu, n, counts.p);
    return nu;
}

// ---------------------------------- min-count filter -------------------------------------
template <int W>
struct FilterF {
    const u64 *keys;
    const u32 *counts;
    u64 *okeys;
    u32 *ocounts;
    u64 min_freq;
    int u8sat;
    __device__ bool pred(u64 i) const {
        u64 c = counts[i];
        if (u8sat && c > 255) c = 255;
        return c >= min_freq;
    }
    __device__ void emit(u64 i, u64 dst) const {
        km_store<W>(okeys, dst, km_load<W>(keys, i));
        ocounts[dst] = counts[i];
    }
};

__global__ void spectrum_kernel(const u32 *__restrict__ counts, u64 n, unsigned long long *__restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u32 c = counts[i];
        atomicAdd(&h[c > 255 ? 255 : c], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
