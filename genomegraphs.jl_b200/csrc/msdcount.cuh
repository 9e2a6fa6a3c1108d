// msdcount.cuh -- K2+K3 fused for k <= 32 (one u64 word per k-mer): counting without a
// global sort of all instances.
//
// Replaces (reference) KmerAnalysis serial_mem: collect every canonical k-mer -> sort! ->
// run-length collapse (docs/src/man/assembly.md:126-128).  Output is identical (sorted unique
// k-mers + counts); the route is an MSD formulation sized for HBM traffic:
//
//   A  msd_hist_kernel      roll over the packed reads: histogram of the top 14 key bits and a
//                           distinct-count estimate (linear counting over a hash-sampled slice of
//                           the key space)
//      plan                 children of ~2048 DISTINCT keys: b0 (8 preferred, <= 12) + b1 (<= 10) bits
//   B  msd_scatter_kernel   one roll per 32 windows (keys stay in registers), tile staged in
//                           shared memory grouped by bucket, one reservation per (tile, bucket);
//                           multi-GPU: written straight into the owner GPU's buffer (PEER)
//   C  msd_hist1_kernel     (only when b0 + b1 > 14; otherwise pass A's histogram is collapsed)
//                           per level-0 bucket, histogram of the next b1 bits
//   D  msd_scatter1_kernel  same staging scheme, keys -> 2^(b0+b1) ordered children
//   E  msd_leaf_kernel      one CTA per child: all instances streamed through a shared-memory
//                           hash table (capacity bounds DISTINCT keys, not instances), distinct
//                           keys put in order by order-preserving bins + rank-by-counting.  A
//                           child whose table overflows is partitioned again by the CTA (8 bits
//                           per level, depth first).
//   F  msd_gather_kernel    per-child unique runs -> dense output arrays
//
// A bucket / child may be given as several segments (virtual buckets): in the multi-GPU path
// the segments are the pieces received from the ranks (dist.cuh).
//
// Algorithmic bytes per k-mer instance (8-byte keys, U = distinct kept): A 0.47 read;
// B 0.47 read + 8 written; C 8 read; D 8 read + 8 written; E 8 read + 12 U/N written; F 24 U/N.
#pragma once
#include "pack.cuh"
#include "sort.cuh"
#include <math.h>

#define MSD_HB 14      // resolution of the pass-A histogram (bits): covers b0 + b1 of mid-size inputs, so that the
                       // level-1 histogram (pass C) falls out of pass A for free
#define MSD_B0_MAX 12  // largest level-0 digit (bins of the pass-B tile)
#define MSD_MAX_B1 10
#define MSD_UNROLL 8   // independent global loads in flight per thread in the leaf kernel
#define MSD_EMPTY 0xFFFFFFFFFFFFFFFFULL  // never a valid key: all-T k-mers are not canonical, 2k < 64 otherwise (see msd_supported)
#define MSD_GOLD 0x9E3779B97F4A7C15ULL
#define EST_LOG 20     // log2 of the bytes of the distinct-estimate map

static inline bool msd_supported(int k, int canonical) { return k <= 32 && (canonical || k < 32); }

// 32 consecutive windows starting at base 32 t (k <= 32): f(key, j) for every valid window j.
// Forward and reverse-complement words are kept LEFT-aligned (k-mer in the top 2k bits) so that every
// per-window shift has a compile-time amount: the forward word is a funnel shift of the two packed
// words, the reverse complement rolls by one base (>> 2, new complemented base into the top 2 bits).
template <class F>
__device__ __forceinline__ void roll32(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 t, int k, int canonical, F &&f) {
    const u32 vm = valid[t];
    if (vm == 0) return;
    const u64 a0 = packed[t], a1 = packed[t + 1];
    const int s = 64 - 2 * k;  // 0 .. 62
    const u64 topmask = ~0ULL << s;
    const u64 nbc = ~(k == 32 ? a1 : ((a0 << (2 * k)) | (a1 >> s)));  // complemented bases k, k+1, ... of the tile
    u64 rl = rev2_u64(~a0) << s;                                       // reverse complement of window 0
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const u64 fl = (j == 0 ? a0 : ((a0 << (2 * j)) | (a1 >> (64 - 2 * j)))) & topmask;
        if ((vm >> (31 - j)) & 1u) f(((canonical && rl < fl) ? rl : fl) >> s, j);
        if (j < 31) rl = ((rl >> 2) & topmask) | ((nbc << (2 * j)) & 0xC000000000000000ULL);
    }
}

// Pass-A form of the roll: only what the histogram needs.  The top bits of the canonical k-mer are the
// smaller of the two strands' top bits (equal tops give the same digit either way), and the sampler of the
// distinct estimate is a symmetric function of the two strands' high words -- both without the 64-bit
// compare / select / shift of the full key, which is built only for the sampled keys.
template <class F>
__device__ __forceinline__ void roll32_tops(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 t, int k, F &&f) {
    const u32 vm = valid[t];
    if (vm == 0) return;
    const u64 a0 = packed[t], a1 = packed[t + 1];
    const int s = 64 - 2 * k;
    const u64 topmask = ~0ULL << s;
    const u64 nbc = ~(k == 32 ? a1 : ((a0 << (2 * k)) | (a1 >> s)));
    u64 rl = rev2_u64(~a0) << s;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const u64 fl = (j == 0 ? a0 : ((a0 << (2 * j)) | (a1 >> (64 - 2 * j)))) & topmask;
        if ((vm >> (31 - j)) & 1u) f(fl, rl);
        if (j < 31) rl = ((rl >> 2) & topmask) | ((nbc << (2 * j)) & 0xC000000000000000ULL);
    }
}

__device__ __forceinline__ u32 msd_digit0(u64 key, int sh, int b0) { return b0 ? (u32)(key >> sh) : 0u; }

// ---- pass A: 12-bit histogram + distinct estimate -------------------------------------------
// Estimate (linear counting): keys whose sampler hash falls into a 2^-est_shift slice set one byte
// of a 2^EST_LOG-byte map; with Z zero bytes left, distinct ~= 2^EST_LOG * ln(2^EST_LOG / Z) << est_shift.
// The map of several GPUs merges with an elementwise max (dist.cuh).
#define MSD_HIST_THREADS 512
// words [w0, nwords) of the packed stream (the pipelined host path calls it once per arrived chunk)
__global__ void __launch_bounds__(MSD_HIST_THREADS) msd_hist_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 w0, u64 nwords,
                                                                    int k, int canonical, int hb, u32 *__restrict__ ghist, u32 est_max,
                                                                    u8 *__restrict__ est_map, u32 *__restrict__ rows) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u32 *h = reinterpret_cast<u32 *>(smem_raw);  // 2^hb bins
    const int nb = 1 << hb;
    for (int i = threadIdx.x; i < nb; i += MSD_HIST_THREADS) h[i] = 0;
    __syncthreads();
    const int s = 64 - 2 * k;
    for (u64 t = w0 + (u64)blockIdx.x * MSD_HIST_THREADS + threadIdx.x; t < nwords; t += (u64)gridDim.x * MSD_HIST_THREADS)
        roll32_tops(packed, valid, t, k, [&](u64 fl, u64 rl) {
            const u32 fh = (u32)(fl >> 32), rh = (u32)(rl >> 32);
            const u32 top = canonical ? min(fh, rh) : fh;   // left-aligned: the digit is its top hb bits
            atomicAdd(&h[hb ? (top >> (32 - hb)) : 0u], 1u);
            const u32 p = (canonical ? fh + rh : fh) * 0x9E3779B1u;  // a sampler of the canonical k-mer, not an identity
            if (p <= est_max) {
                const u64 key = ((canonical && rl < fl) ? rl : fl) >> s;
                est_map[(u32)((key * MSD_GOLD) >> (64 - EST_LOG))] = 1;
            }
        });
    __syncthreads();
    if (rows) {  // chunked launches: every CTA owns one row (plain adds, launches are stream ordered); reduced once at the end
        u32 *row = rows + (size_t)blockIdx.x * nb;
        for (int i = threadIdx.x; i < nb; i += MSD_HIST_THREADS)
            if (h[i]) row[i] += h[i];
    } else {
        for (int i = threadIdx.x; i < nb; i += MSD_HIST_THREADS)
            if (h[i]) atomicAdd(&ghist[i], h[i]);
    }
}
__global__ void msd_hist_rows_reduce_kernel(const u32 *__restrict__ rows, u32 nrows, u32 nb, u32 *__restrict__ ghist) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    u32 s = 0;
    for (u32 r = 0; r < nrows; ++r) s += rows[(size_t)r * nb + i];
    ghist[i] += s;
}
// out[0] += number of set map bytes; out[1] += sum of the histogram (N)
__global__ void __launch_bounds__(256) msd_est_count_kernel(const u8 *__restrict__ est_map, const u32 *__restrict__ hist,
                                                            u32 nb, unsigned long long *__restrict__ out) {
    u64 occ = 0, n = 0;
    const uint4 *m4 = reinterpret_cast<const uint4 *>(est_map);
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < (1u << EST_LOG) / 16; i += gridDim.x * 256) {
        uint4 v = m4[i];
        occ += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);  // bytes are 0 or 1
    }
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < nb; i += gridDim.x * 256) n += hist[i];
    for (int d = 16; d > 0; d >>= 1) { occ += __shfl_xor_sync(FULL_MASK, occ, d); n += __shfl_xor_sync(FULL_MASK, n, d); }
    if (lane_id() == 0) { if (occ) atomicAdd(&out[0], (unsigned long long)occ); if (n) atomicAdd(&out[1], (unsigned long long)n); }
}
// 2^hb-bin histogram -> 2^b0 bins (b0 <= hb)
__global__ void msd_collapse_kernel(const u32 *__restrict__ hist, int hb, int b0, u32 *__restrict__ out) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > (1u << b0)) return;
    u32 s = 0;
    if (b < (1u << b0)) {
        const u32 g = 1u << (hb - b0);
        for (u32 i = 0; i < g; ++i) s += hist[b * g + i];
    }
    out[b] = s;
}

// shared tail of the two scatter kernels: bins counted in hist[0..nb) -> per-bin reservation in
// the global cursors; hist becomes the in-tile running cursor, delta the global-minus-local base
template <int BPT>
__device__ __forceinline__ u32 msd_reserve(u32 *hist, u32 *delta, int nb, u32 *__restrict__ cursor, u32 *ws) {
    const int bpt = (nb + 255) / 256;  // <= BPT
    const int lo = min(nb, (int)threadIdx.x * bpt), hi = min(nb, lo + bpt);
    u32 s = 0;
    u32 cnts[BPT], res[BPT];
#pragma unroll
    for (int q = 0; q < BPT; ++q) { cnts[q] = (lo + q < hi) ? hist[lo + q] : 0; s += cnts[q]; }
#pragma unroll
    for (int q = 0; q < BPT; ++q) res[q] = cnts[q] ? atomicAdd(&cursor[lo + q], cnts[q]) : 0;  // independent: all in flight together
    u32 total;
    u32 run = block_excl_scan(s, ws, &total);
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
        if (lo + q < hi) {
            delta[lo + q] = res[q] - run;
            hist[lo + q] = run;
            run += cnts[q];
        }
    }
    __syncthreads();
    return total;
}

// ---- pass B: tile = 256 words = 8192 window positions ------------------------------------------
#define MSD_TILE_KEYS 8192
// Multi-GPU: the buckets are owned by ranks (contiguous ranges first[q] .. first[q + 1]); with PEER the
// tile is written straight into the owner's receive buffer over NVLink (peer pointers from CUDA IPC),
// so the level-0 scatter IS the exchange.  The cursors then hold offsets inside the owners' buffers.
#define MSD_MAX_PEERS 8
struct MsdPeers {
    u64 *ptr[MSD_MAX_PEERS];
    u32 first[MSD_MAX_PEERS + 1];
    int n;
};
__device__ __forceinline__ int msd_owner(const MsdPeers &pt, u32 d) {
    int o = 0;
#pragma unroll
    for (int q = 1; q < MSD_MAX_PEERS; ++q) o += (q < pt.n && d >= pt.first[q]) ? 1 : 0;
    return o;
}
// SPEC (speculative bucket capacities, pipelined host path): the launch covers the packed words [w0, nwords), a
// bucket ends at bend[d], and a key that would land beyond it is dropped and raises *ovf (the caller then repeats
// pass B with exact offsets).
template <bool PEER, bool SPEC = false>
__global__ void __launch_bounds__(256, 2) msd_scatter_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 nwords,
                                                             int k, int canonical, int b0, u32 *__restrict__ cursor, u64 *__restrict__ out,
                                                             MsdPeers pt, u64 w0 = 0, const u32 *__restrict__ bend = nullptr,
                                                             u32 *__restrict__ ovf = nullptr) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);                      // MSD_TILE_KEYS
    u32 *hist = reinterpret_cast<u32 *>(skeys + MSD_TILE_KEYS);          // 2^MSD_B0_MAX
    u32 *delta = hist + (1 << MSD_B0_MAX);                               // 2^MSD_B0_MAX
    __shared__ u32 ws[33];
    const int nb = 1 << b0, sh = 2 * k - b0;
    const u64 t = (SPEC ? w0 : 0) + (u64)blockIdx.x * 256 + threadIdx.x;
    for (int i = threadIdx.x; i < nb; i += 256) hist[i] = 0;
    __syncthreads();
    // one roll: the 32 keys of the thread stay in registers between the count and the placement
    u64 keys[32];
    u32 vm = 0;
    if (t < nwords) {
        vm = valid[t];
        roll32(packed, valid, t, k, canonical, [&](u64 key, int j) { keys[j] = key; atomicAdd(&hist[msd_digit0(key, sh, b0)], 1u); });
    }
    __syncthreads();
    const u32 total = msd_reserve<(1 << MSD_B0_MAX) / 256>(hist, delta, nb, cursor, ws);
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if ((vm >> (31 - j)) & 1u) skeys[atomicAdd(&hist[msd_digit0(keys[j], sh, b0)], 1u)] = keys[j];
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        const u64 key = skeys[i];
        const u32 d = msd_digit0(key, sh, b0);
        if (PEER) pt.ptr[msd_owner(pt, d)][(u32)(delta[d] + i)] = key;
        else if (SPEC) {
            const u32 pos = delta[d] + i;
            if (pos < bend[d]) out[pos] = key;
            else *ovf = 1u;
        } else out[delta[d] + i] = key;
    }
}
// Speculative capacities from the histogram of the first chunk: bucket shares of `total_cap` key slots in
// proportion to (count + 64).  One CTA; h14 = 2^hb-bin histogram of the sample.  start has nb + 1 entries.
__global__ void __launch_bounds__(1024) msd_spec_caps_kernel(const u32 *__restrict__ h14, int hb, int b0, u64 total_cap, u32 *__restrict__ start,
                                                             u32 *__restrict__ cursor, u32 *__restrict__ bend) {
    __shared__ u32 bin[1 << MSD_B0_MAX];
    __shared__ u32 ws[33];
    __shared__ double s_scale;
    const u32 nb = 1u << b0, g = 1u << (hb - b0);
    u32 mysum = 0;
    for (u32 b = threadIdx.x; b < nb; b += 1024) {
        u32 c = 64;
        for (u32 i = 0; i < g; ++i) c += h14[b * g + i];
        bin[b] = c;
        mysum += c;
    }
    u32 tot;
    block_excl_scan(mysum, ws, &tot);
    if (threadIdx.x == 0) s_scale = (double)total_cap / (double)(tot ? tot : 1);
    __syncthreads();
    for (u32 b = threadIdx.x; b < nb; b += 1024) bin[b] = (u32)((double)bin[b] * s_scale);
    __syncthreads();
    if (threadIdx.x == 0) {  // nb <= 4096: a serial scan is fine here
        u32 run = 0;
        for (u32 b = 0; b < nb; ++b) { start[b] = run; cursor[b] = run; run += bin[b]; bend[b] = run; }
        start[nb] = run;
    }
}
__global__ void msd_spec_counts_kernel(const u32 *__restrict__ start, const u32 *__restrict__ cursor, u32 nb, u32 *__restrict__ vcnt) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nb) vcnt[b] = b < nb ? cursor[b] - start[b] : 0;
}

// ---- level 1 over virtual buckets ---------------------------------------------------------------
// Virtual bucket vb = (local bucket bl) * nsrc + src: keys IN[vbeg[vb] .. vbeg[vb] + vcnt[vb]), all
// of level-0 bucket bl.  Single GPU: nsrc = 1.  Tiles never straddle a virtual bucket.
#define MSD_T1 4096  // keys per tile (16 per thread)
__global__ void msd_tilecount_kernel(const u32 *__restrict__ vcnt, u32 nvb, u32 *__restrict__ tcount) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nvb) tcount[b] = b < nvb ? (vcnt[b] + MSD_T1 - 1) / MSD_T1 : 0;
}
__device__ __forceinline__ bool msd_tile_range(const u32 *__restrict__ tstart, const u32 *__restrict__ vbeg, const u32 *__restrict__ vcnt,
                                               u32 nvb, u32 t, u32 &vb, u32 &s, u32 &e) {
    if (t >= tstart[nvb]) return false;
    u32 lo = 0, hi = nvb;  // largest vb with tstart[vb] <= t
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (tstart[mid] <= t) lo = mid; else hi = mid;
    }
    vb = lo;
    const u32 b0 = vbeg[vb];
    s = b0 + (t - tstart[vb]) * MSD_T1;
    e = min(b0 + vcnt[vb], s + MSD_T1);
    return true;
}
__device__ __forceinline__ u32 msd_digit1(u64 key, int sh1, u32 dm) { return dm ? ((u32)(key >> sh1) & dm) : 0u; }

__global__ void __launch_bounds__(256) msd_hist1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart,
                                                        const u32 *__restrict__ vbeg, const u32 *__restrict__ vcnt, u32 nvb, u32 nsrc,
                                                        int sh1, int b1, u32 *__restrict__ hist1) {
    __shared__ u32 h[1 << MSD_MAX_B1];
    u32 vb, s, e;
    if (!msd_tile_range(tstart, vbeg, vcnt, nvb, blockIdx.x, vb, s, e)) return;
    const int nb = 1 << b1;
    const u32 dm = (u32)nb - 1u;
    for (int i = threadIdx.x; i < nb; i += 256) h[i] = 0;
    __syncthreads();
    u64 kk[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { u32 i = s + u * 256 + threadIdx.x; kk[u] = i < e ? A[i] : 0; }
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) atomicAdd(&h[msd_digit1(kk[u], sh1, dm)], 1u);
    __syncthreads();
    const u64 cb = (u64)(vb / nsrc) << b1;
    for (int i = threadIdx.x; i < nb; i += 256)
        if (h[i]) atomicAdd(&hist1[cb + i], h[i]);
}
__global__ void __launch_bounds__(256) msd_scatter1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart,
                                                           const u32 *__restrict__ vbeg, const u32 *__restrict__ vcnt, u32 nvb, u32 nsrc,
                                                           int sh1, int b1, u32 *__restrict__ cursor1, u64 *__restrict__ B) {
    __shared__ u64 skeys[MSD_T1];
    __shared__ u32 hist[1 << MSD_MAX_B1];
    __shared__ u32 delta[1 << MSD_MAX_B1];
    __shared__ u32 ws[33];
    u32 vb, s, e;
    if (!msd_tile_range(tstart, vbeg, vcnt, nvb, blockIdx.x, vb, s, e)) return;
    const int nb = 1 << b1;
    const u32 dm = (u32)nb - 1u;
    for (int i = threadIdx.x; i < nb; i += 256) hist[i] = 0;
    __syncthreads();
    u64 kk[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { u32 i = s + u * 256 + threadIdx.x; kk[u] = i < e ? A[i] : 0; }
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) atomicAdd(&hist[msd_digit1(kk[u], sh1, dm)], 1u);
    __syncthreads();
    const u32 total = msd_reserve<(1 << MSD_MAX_B1) / 256>(hist, delta, nb, cursor1 + ((u64)(vb / nsrc) << b1), ws);
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) skeys[atomicAdd(&hist[msd_digit1(kk[u], sh1, dm)], 1u)] = kk[u];
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        u64 key = skeys[i];
        B[delta[msd_digit1(key, sh1, dm)] + i] = key;
    }
}
__global__ void msd_copy_u32_kernel(const u32 *__restrict__ in, u64 n, u32 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ------------------------------- pass E: one CTA per child ---------------------------------------
#define MSD_MAXD 9
#define LEAF_BINS 1024
#define LEAF_MAX_PROBE 96   // a longer probe sequence means the table is too full: partition instead
#define FSCR_STRIDE 260     // per-depth scratch: 257 offsets (+ padding)

template <int THREADS, int LOG_HT>
struct LeafSmem {
    static constexpr int HT = 1 << LOG_HT;
    u64 tab[HT];
    u32 tcnt[HT];
    u32 bins[LEAF_BINS + 1];
    u32 bincur[LEAF_BINS];
    u32 pcur[256];
    u64 red[64];
    u32 ws[33];
    u32 s_task, s_ovf, s_nd;
    u32 sub_beg, sub_cnt;
    u64 s_min, s_max;
};

// table slot of a key: folded 32-bit multiplicative hash (one IMUL instead of a 64-bit multiply chain)
template <int LOG_HT>
__device__ __forceinline__ u32 leaf_slot(u64 key) {
    return (((u32)key ^ (u32)(key >> 32)) * 0x9E3779B1u) >> (32 - LOG_HT);
}

template <int THREADS>
__device__ __forceinline__ u32 block_scan_array(u32 *arr, int n, u32 *ws) {
    // exclusive scan of arr[0..n) in place (n <= 8 * THREADS); returns the total
    const int bpt = (n + THREADS - 1) / THREADS;
    const int lo = min(n, (int)threadIdx.x * bpt), hi = min(n, lo + bpt);
    u32 s = 0;
    for (int i = lo; i < hi; ++i) s += arr[i];
    u32 total;
    u32 run = block_excl_scan(s, ws, &total);
    for (int i = lo; i < hi; ++i) { u32 c = arr[i]; arr[i] = run; run += c; }
    __syncthreads();
    return total;
}

// Count the keys of nseg segments IN[segbeg[s] .. +segcnt[s]) in the hash table; the distinct keys
// with count >= min_freq are put in order and written to OK/OC[obase ..); SK[obase ..) is staging.
// Returns false (nothing written) when the table got too full; *nd_out = number written.
// SK may alias the input; OK must not alias SK.
template <int THREADS, int LOG_HT>
__device__ bool msd_leaf(LeafSmem<THREADS, LOG_HT> &S, const u64 *IN, const u32 *segbeg, const u32 *segcnt, u32 nseg, u64 *SK,
                         u64 *OK, u32 *OC, u32 obase, u32 min_freq, u32 *nd_out) {
    constexpr int HT = 1 << LOG_HT;
    constexpr u32 HM = HT - 1;
    {
        ulonglong2 *t2 = reinterpret_cast<ulonglong2 *>(S.tab);
        for (int i = threadIdx.x; i < HT / 2; i += THREADS) t2[i] = make_ulonglong2(MSD_EMPTY, MSD_EMPTY);
        uint4 *c4 = reinterpret_cast<uint4 *>(S.tcnt);
        for (int i = threadIdx.x; i < HT / 4; i += THREADS) c4[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) S.s_ovf = 0;
    }
    __syncthreads();
    volatile u32 *ovf = &S.s_ovf;
    for (u32 sg = 0; sg < nseg; ++sg) {
        const u64 *bp = IN + segbeg[sg];
        const u32 n = segcnt[sg];
        u64 cur[MSD_UNROLL], nxt[MSD_UNROLL];
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = threadIdx.x + u * THREADS; cur[u] = i < n ? bp[i] : MSD_EMPTY; }
        for (u32 i0 = threadIdx.x; i0 < n; i0 += MSD_UNROLL * THREADS) {
            const u32 i1 = i0 + MSD_UNROLL * THREADS;  // next batch in flight while this one is inserted
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i1 + u * THREADS; nxt[u] = i < n ? bp[i] : MSD_EMPTY; }
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u) {
                const u64 key = cur[u];
                if (key == MSD_EMPTY) continue;
                u32 slot = leaf_slot<LOG_HT>(key);
                int probes = 0;
                for (;;) {
                    u64 c = S.tab[slot];
                    if (c == MSD_EMPTY) c = atomicCAS((unsigned long long *)&S.tab[slot], (unsigned long long)MSD_EMPTY, (unsigned long long)key);
                    if (c == MSD_EMPTY || c == key) { atomicAdd(&S.tcnt[slot], 1u); break; }
                    slot = (slot + 1) & HM;
                    if (++probes > LEAF_MAX_PROBE) { *ovf = 1; break; }
                }
            }
            if (*ovf) break;
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u) cur[u] = nxt[u];
        }
        if (*ovf) break;
    }
    __syncthreads();
    if (S.s_ovf) return false;
    // ---- range of the kept keys (their common prefix is skipped by the ordering bins) ----
    u64 mn = ~0ULL, mx = 0;
    for (int slot = threadIdx.x; slot < HT; slot += THREADS) {
        const u64 key = S.tab[slot];
        if (key != MSD_EMPTY && S.tcnt[slot] >= min_freq) { mn = key < mn ? key : mn; mx = key > mx ? key : mx; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 a = __shfl_xor_sync(FULL_MASK, mn, d), b = __shfl_xor_sync(FULL_MASK, mx, d);
        mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    }
    if (lane_id() == 0) { S.red[threadIdx.x >> 5] = mn; S.red[32 + (threadIdx.x >> 5)] = mx; }
    for (int i = threadIdx.x; i <= LEAF_BINS; i += THREADS) S.bins[i] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        u64 a = threadIdx.x < THREADS / 32 ? S.red[threadIdx.x] : ~0ULL, b = threadIdx.x < THREADS / 32 ? S.red[32 + threadIdx.x] : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            u64 a2 = __shfl_xor_sync(FULL_MASK, a, d), b2 = __shfl_xor_sync(FULL_MASK, b, d);
            a = a2 < a ? a2 : a; b = b2 > b ? b2 : b;
        }
        if (threadIdx.x == 0) { S.s_min = a; S.s_max = b; }
    }
    __syncthreads();
    mn = S.s_min; mx = S.s_max;
    if (mn > mx) { *nd_out = 0; return true; }  // nothing kept
    const int hi = 64 - __clzll((long long)(mn ^ mx));  // bits [0, hi) vary (0 when a single key)
    const int bb = hi < 10 ? hi : 10, bsh = hi - bb;
    const u32 bm = (1u << bb) - 1u;
    const int nbins = 1 << bb;
    for (int slot = threadIdx.x; slot < HT; slot += THREADS) {
        const u64 key = S.tab[slot];
        if (key != MSD_EMPTY && S.tcnt[slot] >= min_freq) atomicAdd(&S.bins[(u32)(key >> bsh) & bm], 1u);
    }
    __syncthreads();
    const u32 nd = block_scan_array<THREADS>(S.bins, nbins, S.ws);
    for (int i = threadIdx.x; i < nbins; i += THREADS) S.bincur[i] = S.bins[i];
    if (threadIdx.x == 0) S.bins[nbins] = nd;
    __syncthreads();
    u64 *sk = SK + obase;
    for (int slot = threadIdx.x; slot < HT; slot += THREADS) {
        const u64 key = S.tab[slot];
        if (key != MSD_EMPTY && S.tcnt[slot] >= min_freq) sk[atomicAdd(&S.bincur[(u32)(key >> bsh) & bm], 1u)] = key;
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < nd; j += THREADS) {
        const u64 key = sk[j];
        const u32 b = (u32)(key >> bsh) & bm;
        const u32 lo = S.bins[b], hi2 = S.bins[b + 1];
        u32 rank = 0;
        for (u32 q = lo; q < hi2; ++q) rank += sk[q] < key;
        u32 slot = leaf_slot<LOG_HT>(key);
        while (S.tab[slot] != key) slot = (slot + 1) & HM;
        OK[obase + lo + rank] = key;
        OC[obase + lo + rank] = S.tcnt[slot];
    }
    __syncthreads();
    *nd_out = nd;
    return true;
}

// radix partition of nseg segments by key bits [sh, sh + bits) into DST[dbase ..); the 257 child
// offsets (relative) go to offs (global scratch)
template <int THREADS, int LOG_HT>
__device__ void msd_partition(LeafSmem<THREADS, LOG_HT> &S, const u64 *SRC, const u32 *segbeg, const u32 *segcnt, u32 nseg, u64 *DST,
                              u32 dbase, int sh, int bits, u32 *offs) {
    const u32 dm = (1u << bits) - 1u;
    for (int i = threadIdx.x; i <= 256; i += THREADS) S.bins[i] = 0;
    __syncthreads();
    for (u32 sg = 0; sg < nseg; ++sg) {
        const u64 *sp = SRC + segbeg[sg];
        const u32 n = segcnt[sg];
        for (u32 i0 = threadIdx.x; i0 < n; i0 += MSD_UNROLL * THREADS) {
            u64 kk[MSD_UNROLL];
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i0 + u * THREADS; kk[u] = i < n ? sp[i] : 0; }
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u)
                if (i0 + u * THREADS < n) atomicAdd(&S.bins[(u32)(kk[u] >> sh) & dm], 1u);
        }
    }
    __syncthreads();
    const u32 total = block_scan_array<THREADS>(S.bins, 256, S.ws);
    if (threadIdx.x < 256) { S.pcur[threadIdx.x] = S.bins[threadIdx.x]; offs[threadIdx.x] = S.bins[threadIdx.x]; }
    if (threadIdx.x == 0) offs[256] = total;
    __syncthreads();
    u64 *dp = DST + dbase;
    for (u32 sg = 0; sg < nseg; ++sg) {
        const u64 *sp = SRC + segbeg[sg];
        const u32 n = segcnt[sg];
        for (u32 i0 = threadIdx.x; i0 < n; i0 += MSD_UNROLL * THREADS) {
            u64 kk[MSD_UNROLL];
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i0 + u * THREADS; kk[u] = i < n ? sp[i] : 0; }
#pragma unroll
            for (int u = 0; u < MSD_UNROLL; ++u)
                if (i0 + u * THREADS < n) dp[atomicAdd(&S.pcur[(u32)(kk[u] >> sh) & dm], 1u)] = kk[u];
        }
    }
    __syncthreads();
}

// IN: buffer holding the children (child c = nseg segments segbeg/segcnt[c * nseg ..]);
// O / OC: output keys / counts, child c writes [off[c], off[c] + nu[c]); SB: scratch keys, same size.
// O may alias IN when nseg == 1 and segbeg == off (counted in place).
template <int THREADS, int LOG_HT>
__global__ void __launch_bounds__(THREADS) msd_leaf_kernel(const u64 *IN, u64 *SB, u64 *O, u32 *OC, const u32 *__restrict__ segbeg,
                                                           const u32 *__restrict__ segcnt, u32 nseg, const u32 *__restrict__ off,
                                                           u32 nchild, int shiftc, u32 *__restrict__ nu_out, u32 *ticket, u32 min_freq,
                                                           u32 *fscr, u32 test_ovf_mod) {
    extern __shared__ __align__(16) u8 smem_raw[];
    LeafSmem<THREADS, LOG_HT> &S = *reinterpret_cast<LeafSmem<THREADS, LOG_HT> *>(smem_raw);
    u32 *myscr = fscr + (size_t)blockIdx.x * MSD_MAXD * FSCR_STRIDE;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) S.s_task = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 c = S.s_task;
        if (c >= nchild) break;
        const u32 ob = off[c], bsize = off[c + 1] - ob;
        if (bsize == 0) { if (threadIdx.x == 0) nu_out[c] = 0; continue; }
        if (shiftc == 0) {  // no varying bits: the whole child is one k-mer
            if (threadIdx.x == 0) {
                u32 keep = bsize >= min_freq;
                if (keep) { O[ob] = IN[segbeg[(size_t)c * nseg]]; OC[ob] = bsize; }
                nu_out[c] = keep;
            }
            continue;
        }
        u32 nd;
        if (!(test_ovf_mod && c % test_ovf_mod == 0) &&  // tests: exercise the in-CTA partition fallback
            msd_leaf<THREADS, LOG_HT>(S, IN, segbeg + (size_t)c * nseg, segcnt + (size_t)c * nseg, nseg, SB, O, OC, ob, min_freq, &nd)) {
            if (threadIdx.x == 0) nu_out[c] = nd;
            continue;
        }
        // ---- too many distinct keys for the table: partition again, depth first (uniform control flow) ----
        u32 f_start[MSD_MAXD], f_cur[MSD_MAXD];
        int f_sb[MSD_MAXD], f_bits[MSD_MAXD], f_dst[MSD_MAXD];  // dst: 0 = SB, 1 = O
        u32 s_out = 0;
        int depth = 0;
        f_start[0] = ob; f_sb[0] = shiftc; f_bits[0] = shiftc < 8 ? shiftc : 8; f_dst[0] = 0; f_cur[0] = 0;
        msd_partition<THREADS, LOG_HT>(S, IN, segbeg + (size_t)c * nseg, segcnt + (size_t)c * nseg, nseg, SB, ob, f_sb[0] - f_bits[0], f_bits[0], myscr);
        while (depth >= 0) {
            if (f_cur[depth] == 256u) { --depth; continue; }
            const u32 d = f_cur[depth]++;
            const u32 *fo = myscr + depth * FSCR_STRIDE;
            const u32 s = fo[d], n = fo[d + 1] - s;
            if (n == 0) continue;
            u64 *buf = f_dst[depth] ? O : SB;
            const u32 abs0 = f_start[depth] + s;
            const int child_shift = f_sb[depth] - f_bits[depth];
            if (child_shift == 0 || depth + 1 >= MSD_MAXD) {  // all n keys are the same k-mer
                __syncthreads();
                if (threadIdx.x == 0 && n >= min_freq) { u64 key = buf[abs0]; O[ob + s_out] = key; OC[ob + s_out] = n; }
                if (n >= min_freq) s_out += 1;
                __syncthreads();
                continue;
            }
            if (threadIdx.x == 0) { S.sub_beg = abs0; S.sub_cnt = n; }
            __syncthreads();
            if (msd_leaf<THREADS, LOG_HT>(S, buf, &S.sub_beg, &S.sub_cnt, 1, SB, O, OC, ob + s_out, min_freq, &nd)) {
                s_out += nd;
                continue;
            }
            f_start[depth + 1] = abs0; f_sb[depth + 1] = child_shift; f_bits[depth + 1] = child_shift < 8 ? child_shift : 8;
            f_dst[depth + 1] = 1 - f_dst[depth]; f_cur[depth + 1] = 0;
            msd_partition<THREADS, LOG_HT>(S, buf, &S.sub_beg, &S.sub_cnt, 1, f_dst[depth + 1] ? O : SB, abs0,
                                           f_sb[depth + 1] - f_bits[depth + 1], f_bits[depth + 1], myscr + (depth + 1) * FSCR_STRIDE);
            ++depth;
        }
        __syncthreads();
        if (threadIdx.x == 0) nu_out[c] = s_out;
    }
}

__global__ void __launch_bounds__(256) msd_gather_kernel(const u64 *__restrict__ X, const u32 *__restrict__ cnt, const u32 *__restrict__ off,
                                                         const u32 *__restrict__ nu, const u32 *__restrict__ ustart, u32 nchild,
                                                         u64 *__restrict__ okeys, u32 *__restrict__ ocnt) {
    for (u32 t = blockIdx.x; t < nchild; t += gridDim.x) {
        const u32 n = nu[t], s = off[t], d = ustart[t];
        for (u32 i = threadIdx.x; i < n; i += 256) {
            okeys[d + i] = X[s + i];
            ocnt[d + i] = cnt[s + i];
        }
    }
}

// ------------------------------------- host side -------------------------------------------------
struct MsdTune {
    int b0_max = 8;          // preferred level-0 digit bits (grows up to MSD_B0_MAX when level 1 is full)
    int b1_max = MSD_MAX_B1; // level-1 digit bits
    u32 dt = 2048;           // target DISTINCT keys per child: 1/4 of the leaf table (8192 slots), which leaves room for the
                             // ~1.75x skew of canonical k-mers towards small leading bases
    u32 it_log2 = 16;        // target instances per child (bounds the work of one CTA)
    int variant = 0;         // leaf geometry, see msd_launch_leaf
    u32 test_ovf_mod = 0;    // tests only (GG_MSD_TEST_OVERFLOW=m): every child with index % m == 0 takes the overflow path
};
static MsdTune msd_tune() {
    MsdTune t;
    if (const char *e = getenv("GG_MSD_B0")) t.b0_max = atoi(e);
    if (const char *e = getenv("GG_MSD_B1")) t.b1_max = atoi(e);
    if (const char *e = getenv("GG_MSD_DT")) t.dt = (u32)atoi(e);
    if (const char *e = getenv("GG_MSD_IT")) t.it_log2 = (u32)atoi(e);
    if (const char *e = getenv("GG_MSD_VARIANT")) t.variant = atoi(e);
    if (const char *e = getenv("GG_MSD_TEST_OVERFLOW")) t.test_ovf_mod = (u32)atoi(e);
    if (t.b0_max > MSD_B0_MAX) t.b0_max = MSD_B0_MAX;
    if (t.b0_max < 0) t.b0_max = 0;
    if (t.b1_max > MSD_MAX_B1) t.b1_max = MSD_MAX_B1;
    if (t.b1_max < 0) t.b1_max = 0;
    if (t.dt < 16) t.dt = 16;
    return t;
}

struct MsdPlan {
    int hb = 0;      // histogram bits of pass A
    int b0 = 0, b1 = 0;
    int shiftc = 0;  // key bits below the child digits
};
static inline int msd_hist_bits(int k) { return 2 * k < MSD_HB ? 2 * k : MSD_HB; }
// children of ~dt distinct keys and <= 2^it_log2 instances; hb = bits of the pass-A histogram
static MsdPlan msd_plan(u64 distinct_est, u64 n_instances, int k, const MsdTune &t, int hb) {
    MsdPlan p;
    p.hb = hb;
    int bits = 0;
    while (bits < 22 && (distinct_est >> bits) > t.dt) ++bits;
    while (bits < 22 && (n_instances >> bits) > (1ULL << t.it_log2)) ++bits;
    // level 0 prefers few buckets (long runs per tile, cheap reservations); level 1 takes the rest,
    // and level 0 widens only when level 1 is at its maximum
    p.b0 = bits < t.b0_max ? bits : t.b0_max;
    p.b1 = bits - p.b0;
    if (p.b1 > t.b1_max) { p.b0 += p.b1 - t.b1_max; p.b1 = t.b1_max; }
    if (p.b0 > MSD_B0_MAX) p.b0 = MSD_B0_MAX;
    if (p.b0 > p.hb) p.b0 = p.hb;
    if (p.b1 > 2 * k - p.b0) p.b1 = 2 * k - p.b0;
    p.shiftc = 2 * k - p.b0 - p.b1;
    return p;
}

static inline int msd_est_shift(u64 nbases) {
    int s = 0;
    while (s < 30 && (nbases >> s) > (1ULL << (EST_LOG - 1))) ++s;
    return s;
}
static inline u64 msd_est_value(u64 set_bytes, int est_shift) {
    const double m = (double)(1ULL << EST_LOG);
    double z = m - (double)set_bytes;
    if (z < 1.0) z = 1.0;
    return (u64)(m * log(m / z)) << est_shift;
}
// pass A, launch: hist (2^hb + 1 entries, last 0) and the estimate map (2^EST_LOG bytes); est_shift
// must be the same on every GPU whose maps are merged
static void msd_pass_a_alloc(gg_ctx *ctx, int k, DBuf<u32> &hist, DBuf<u8> &est) {
    const u32 nb = 1u << msd_hist_bits(k);
    hist.alloc(ctx, nb + 1);
    est.alloc(ctx, 1ULL << EST_LOG);
    dfill(ctx, hist.p, (nb + 1) * sizeof(u32), 0u);
    dfill(ctx, est.p, 1ULL << EST_LOG, 0u);
}
// histogram + estimate over the packed words [w0, w1), accumulated into hist / est
// rows (optional): 3 * num_sms per-CTA histogram rows for chunked launches (see msd_hist_kernel)
static void msd_pass_a_range(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, int est_shift, u32 *hist, u8 *est, u64 w0,
                             u64 w1, u32 *rows = nullptr) {
    if (w1 <= w0) return;
    u64 hgrid = ceil_div(w1 - w0, MSD_HIST_THREADS);
    const u64 gmax = (u64)ctx->num_sms * 3;  // rows, when given, hold one row per CTA of this grid
    if (hgrid > gmax) hgrid = gmax;
    const size_t smemA = sizeof(u32) << MSD_HB;
    set_max_smem(ctx, msd_hist_kernel, smemA);
    LAUNCH(ctx, msd_hist_kernel, (unsigned)hgrid, MSD_HIST_THREADS, smemA, pr.packed.p, valid, w0, w1, k, canonical, msd_hist_bits(k), hist,
           est_shift ? (0xFFFFFFFFu >> est_shift) : 0xFFFFFFFFu, est, rows);
}
static void msd_pass_a_launch(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, int est_shift, DBuf<u32> &hist,
                              DBuf<u8> &est) {
    msd_pass_a_alloc(ctx, k, hist, est);
    msd_pass_a_range(ctx, pr, valid, k, canonical, est_shift, hist.p, est.p, 0, pr.nwords);
}
// pass A, read back: N = sum of hist, distinct estimate from the map
static void msd_pass_a_read(gg_ctx *ctx, const u32 *hist, u32 nb, const u8 *est, int est_shift, u64 *N, u64 *distinct_est) {
    DBuf<u64> res(ctx, 2);
    CK(cudaMemsetAsync(res.p, 0, 2 * sizeof(u64), ctx->stream));
    LAUNCH(ctx, msd_est_count_kernel, (unsigned)(ctx->num_sms * 2), 256, 0, est, hist, nb, (unsigned long long *)res.p);
    CK(cudaMemcpyAsync(ctx->h_scalar, res.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *distinct_est = msd_est_value(ctx->h_scalar[0], est_shift);
    *N = ctx->h_scalar[1];
}

// pass B: keys scattered into 2^b0 buckets; hist0 (nb0 + 1, last 0) and bstart (nb0 + 1) describe them
// With `peers` (multi-GPU, dist.cuh) the keys go to the owners' receive buffers at the offsets in
// peer_cursor (one per bucket, device) and A stays empty.
static void msd_pass_b(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, const MsdPlan &pl, const u32 *hist,
                       u64 N, DBuf<u32> &hist0, DBuf<u32> &bstart, DBuf<u64> &A, const MsdPeers *peers = nullptr, u32 *peer_cursor = nullptr) {
    const u32 nb0 = 1u << pl.b0;
    hist0.alloc(ctx, nb0 + 1);
    bstart.alloc(ctx, nb0 + 1);
    DBuf<u32> cursor(ctx, nb0);
    LAUNCH(ctx, msd_collapse_kernel, (unsigned)ceil_div(nb0 + 1, 256), 256, 0, hist, pl.hb, pl.b0, hist0.p);
    exclusive_scan_u32(ctx, hist0.p, bstart.p, nb0 + 1, nullptr);
    LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nb0, 256), 256, 0, bstart.p, (u64)nb0, cursor.p);
    A.alloc(ctx, (N && !peers) ? N : 1);
    if (N == 0) return;
    size_t smemB = (size_t)MSD_TILE_KEYS * 8 + 2 * (size_t)(1 << MSD_B0_MAX) * 4;
    set_max_smem(ctx, msd_scatter_kernel<false>, smemB);
    set_max_smem(ctx, msd_scatter_kernel<true>, smemB);
    MsdPeers none;
    memset(&none, 0, sizeof(none));
    if (peers)
        LAUNCH(ctx, msd_scatter_kernel<true>, (unsigned)ceil_div(pr.nwords, 256), 256, smemB, pr.packed.p, valid, pr.nwords, k, canonical, pl.b0,
               peer_cursor, (u64 *)nullptr, *peers);
    else
        LAUNCH(ctx, msd_scatter_kernel<false>, (unsigned)ceil_div(pr.nwords, 256), 256, smemB, pr.packed.p, valid, pr.nwords, k, canonical, pl.b0,
               cursor.p, A.p, none);
}

// speculative pass B over the packed words [w0, w1) (see msd_scatter_kernel<.., SPEC>)
static void msd_pass_b_spec(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, int b0, u32 *cursor, const u32 *bend,
                            u32 *ovf, u64 *A, u64 w0, u64 w1) {
    if (w1 <= w0) return;
    const size_t smemB = (size_t)MSD_TILE_KEYS * 8 + 2 * (size_t)(1 << MSD_B0_MAX) * 4;
    set_max_smem(ctx, msd_scatter_kernel<false, true>, smemB);
    MsdPeers none;
    memset(&none, 0, sizeof(none));
    LAUNCH(ctx, (msd_scatter_kernel<false, true>), (unsigned)ceil_div(w1 - w0, 256), 256, smemB, pr.packed.p, valid, w1, k, canonical, b0, cursor,
           A, none, w0, bend, ovf);
}

__global__ void msd_child_total_kernel(const u32 *__restrict__ vcnt, u32 nsrc, u32 nchild, u32 *__restrict__ ctot) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nchild) return;
    u32 s = 0;
    if (c < nchild) for (u32 i = 0; i < nsrc; ++i) s += vcnt[(size_t)c * nsrc + i];
    ctot[c] = s;
}

template <int THREADS, int LOG_HT>
static void msd_launch_leaf(gg_ctx *ctx, int ctas_per_sm, const u64 *IN, u64 *SB, u64 *O, u32 *OC, const u32 *segbeg, const u32 *segcnt,
                            u32 nseg, const u32 *off, u32 nchild, int shiftc, u32 *nu, u32 *ticket, u32 min_freq, u32 test_ovf_mod) {
    size_t smem = sizeof(LeafSmem<THREADS, LOG_HT>);
    set_max_smem(ctx, msd_leaf_kernel<THREADS, LOG_HT>, smem);
    u32 grid = (u32)ctx->num_sms * ctas_per_sm;
    if (grid > nchild) grid = nchild;
    DBuf<u32> fscr(ctx, (u64)grid * MSD_MAXD * FSCR_STRIDE);
    LAUNCH(ctx, (msd_leaf_kernel<THREADS, LOG_HT>), grid, THREADS, smem, IN, SB, O, OC, segbeg, segcnt, nseg, off, nchild, shiftc, nu, ticket,
           min_freq, fscr.p, test_ovf_mod);
}

// Level-0 buckets -> sorted unique keys + counts (count >= min_freq).
// IN holds n_in keys; local bucket bl (0 .. nb_local) consists of the nsrc virtual buckets
// vb = bl * nsrc + s: IN[vbeg[vb] .. + vcnt[vb]).  IN is consumed (used as scratch).
// child_hist (optional, nchild + 1 entries, last 0): instances per child when pass A already knows them.
static u64 msd_finish(gg_ctx *ctx, u64 *IN, u64 n_in, const u32 *vbeg, const u32 *vcnt, u32 nb_local, u32 nsrc, const MsdPlan &pl,
                      const MsdTune &tune, u32 min_freq, DBuf<u64> &ukeys, DBuf<u32> &ucounts, const u32 *child_hist = nullptr,
                      bool contiguous = true) {
    if (n_in == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    const u32 nvb = nb_local * nsrc;
    const u64 nchild = (u64)nb_local << pl.b1;
    DBuf<u64> S1(ctx, n_in), S2;
    DBuf<u32> OC(ctx, n_in), off(ctx, nchild + 1), ticket(ctx, 1), nu(ctx, nchild + 1), ustart(ctx, nchild + 1);
    DBuf<u64> tot(ctx, 1);
    const u64 *leaf_in;
    u64 *leaf_sb, *leaf_o;
    const u32 *segbeg, *segcnt;
    u32 nseg;
    DBuf<u32> hist1, ctot;
    stage_begin(ctx, ST_SORT);
    if (pl.b1 > 0) {
        // ---- C, D: level-1 partition into S1 ----
        DBuf<u32> tcount(ctx, nvb + 1), tstart(ctx, nvb + 1), cursor1(ctx, nchild);
        LAUNCH(ctx, msd_tilecount_kernel, (unsigned)ceil_div(nvb + 1, 256), 256, 0, vcnt, nvb, tcount.p);
        exclusive_scan_u32(ctx, tcount.p, tstart.p, nvb + 1, nullptr);
        const unsigned tgrid = (unsigned)(ceil_div(n_in, MSD_T1) + nvb);
        const u32 *h1 = child_hist;
        if (!h1) {  // pass C: the pass-A histogram is too coarse for b0 + b1 bits
            hist1.alloc(ctx, nchild + 1);
            CK(cudaMemsetAsync(hist1.p, 0, (nchild + 1) * sizeof(u32), ctx->stream));
            LAUNCH(ctx, msd_hist1_kernel, tgrid, 256, 0, IN, tstart.p, vbeg, vcnt, nvb, nsrc, pl.shiftc, pl.b1, hist1.p);
            h1 = hist1.p;
        }
        exclusive_scan_u32(ctx, h1, off.p, nchild + 1, nullptr);
        LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nchild, 256), 256, 0, off.p, nchild, cursor1.p);
        LAUNCH(ctx, msd_scatter1_kernel, tgrid, 256, 0, IN, tstart.p, vbeg, vcnt, nvb, nsrc, pl.shiftc, pl.b1, cursor1.p, S1.p);
        leaf_in = S1.p; leaf_o = S1.p; leaf_sb = IN;
        segbeg = off.p; segcnt = h1; nseg = 1;
    } else {
        ctot.alloc(ctx, nchild + 1);
        LAUNCH(ctx, msd_child_total_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, vcnt, nsrc, (u32)nchild, ctot.p);
        exclusive_scan_u32(ctx, ctot.p, off.p, nchild + 1, nullptr);
        leaf_in = IN; leaf_sb = S1.p;
        segbeg = vbeg; segcnt = vcnt; nseg = nsrc;
        if (nsrc == 1 && contiguous) leaf_o = IN;  // contiguous buckets: off == vbeg, counted in place
        else { S2.alloc(ctx, n_in); leaf_o = S2.p; }
    }
    // ---- E: leaves ----
    CK(cudaMemsetAsync(ticket.p, 0, sizeof(u32), ctx->stream));
    if (tune.variant == 1)
        msd_launch_leaf<1024, 14>(ctx, 1, leaf_in, leaf_sb, leaf_o, OC.p, segbeg, segcnt, nseg, off.p, (u32)nchild, pl.shiftc, nu.p, ticket.p, min_freq, tune.test_ovf_mod);
    else if (tune.variant == 2)
        msd_launch_leaf<256, 12>(ctx, 4, leaf_in, leaf_sb, leaf_o, OC.p, segbeg, segcnt, nseg, off.p, (u32)nchild, pl.shiftc, nu.p, ticket.p, min_freq, tune.test_ovf_mod);
    else
        msd_launch_leaf<512, 13>(ctx, 2, leaf_in, leaf_sb, leaf_o, OC.p, segbeg, segcnt, nseg, off.p, (u32)nchild, pl.shiftc, nu.p, ticket.p, min_freq, tune.test_ovf_mod);
    stage_end(ctx, ST_SORT);
    // ---- F: gather ----
    stage_begin(ctx, ST_COUNT);
    exclusive_scan_u32(ctx, nu.p, ustart.p, nchild, tot.p);
    const u64 U = read_scalar<u64>(ctx, tot.p);
    ukeys.alloc(ctx, U ? U : 1);
    ucounts.alloc(ctx, U ? U : 1);
    if (U) {
        u32 ggrid = nchild < (u64)ctx->num_sms * 16 ? (u32)nchild : (u32)ctx->num_sms * 16;
        LAUNCH(ctx, msd_gather_kernel, ggrid, 256, 0, leaf_o, OC.p, off.p, nu.p, ustart.p, (u32)nchild, ukeys.p, ucounts.p);
    }
    stage_end(ctx, ST_COUNT);
    return U;
}

// state of a speculative pass B that ran while the reads were still arriving (pipelined host path)
struct MsdSpec {
    int b0 = 0;
    DBuf<u64> A;                          // total_cap key slots
    DBuf<u32> start, cursor, bend, ovf;   // nb + 1, nb, nb, 1
    bool used = false;
};
// everything after pass A: hist / est hold the accumulated 14-bit histogram and estimate map.  With `spec`,
// pass B has already happened into speculative bucket regions; it is kept if nothing overflowed and the plan
// agrees with its digit width, and repeated exactly otherwise.
static u64 msd_count_after_a(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u32> &hist,
                             DBuf<u8> &est, int est_shift, DBuf<u64> &ukeys, DBuf<u32> &ucounts, u64 *n_instances, MsdSpec *spec = nullptr) {
    const MsdTune tune = msd_tune();
    DBuf<u32> hist0, bstart;
    DBuf<u64> A;
    u64 N = 0, dest = 0;
    msd_pass_a_read(ctx, hist.p, 1u << msd_hist_bits(k), est.p, est_shift, &N, &dest);
    est.release();
    *n_instances = N;
    if (N == 0) { stage_end(ctx, ST_EXTRACT); ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    if (N >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu k-mer instances exceed the 2^32-1 per-call limit", (unsigned long long)N);
    const MsdPlan pl = msd_plan(dest, N, k, tune, msd_hist_bits(k));
    bool keep_spec = spec && spec->used && spec->b0 == pl.b0 && !getenv("GG_PIPE_TEST_OVERFLOW");
    if (keep_spec) keep_spec = read_scalar<u32>(ctx, spec->ovf.p) == 0;
    DBuf<u32> child_hist;
    if (pl.b1 > 0 && pl.b0 + pl.b1 <= pl.hb) {  // pass C for free: collapse the pass-A histogram to b0 + b1 bits
        const u32 nchild = 1u << (pl.b0 + pl.b1);
        child_hist.alloc(ctx, nchild + 1);
        LAUNCH(ctx, msd_collapse_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, hist.p, pl.hb, pl.b0 + pl.b1, child_hist.p);
    }
    if (keep_spec) {
        const u32 nb0 = 1u << pl.b0;
        DBuf<u32> vcnt(ctx, nb0 + 1);
        LAUNCH(ctx, msd_spec_counts_kernel, (unsigned)ceil_div(nb0 + 1, 256), 256, 0, spec->start.p, spec->cursor.p, nb0, vcnt.p);
        stage_end(ctx, ST_EXTRACT);
        return msd_finish(ctx, spec->A.p, spec->A.n, spec->start.p, vcnt.p, nb0, 1, pl, tune, min_freq, ukeys, ucounts, child_hist.p, false);
    }
    if (spec) { spec->A.release(); }
    msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N, hist0, bstart, A);
    stage_end(ctx, ST_EXTRACT);
    return msd_finish(ctx, A.p, N, bstart.p, hist0.p, 1u << pl.b0, 1, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
}

// packed reads -> sorted unique k-mers + counts (count >= min_freq only).  msd_supported(k, canonical).
// Returns U; *n_instances = number of k-mer windows.
static u64 msd_count(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq,
                     DBuf<u64> &ukeys, DBuf<u32> &ucounts, u64 *n_instances) {
    *n_instances = 0;
    if (pr.nwords == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    stage_begin(ctx, ST_EXTRACT);
    DBuf<u32> hist;
    DBuf<u8> est;
    const int est_shift = msd_est_shift(pr.nbases);
    msd_pass_a_launch(ctx, pr, valid, k, canonical, est_shift, hist, est);
    return msd_count_after_a(ctx, pr, valid, k, canonical, min_freq, hist, est, est_shift, ukeys, ucounts, n_instances);
}
