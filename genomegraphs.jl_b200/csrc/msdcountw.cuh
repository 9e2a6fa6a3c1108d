// msdcountw.cuh -- the MSD counting route of msdcount.cuh for 2- and 4-word keys (32 < k <= 128).
//
// Same passes (A histogram + distinct estimate, B level-0 scatter, C/D level 1, E one CTA per child,
// F gather) and the same plan; what differs from the one-word kernels:
//   * windows are extracted per position (kmer_at: funnel shift of W + 1 packed words, word-wise
//     reverse complement) instead of rolled, and pass B keeps the extracted keys of its tile in a
//     second shared-memory buffer between the count and the placement (two-word keys) or extracts
//     them a second time (four-word keys: the buffer would halve the tile, which costs more there);
//   * the leaf table (double hashing) claims a slot with a 32-bit tag (atomicCAS exists for one word only), the claimer
//     stores the full key, and a second pass over the child's instances verifies key equality while
//     it counts.  Two different keys with the same tag on the same probe path (astronomically rare)
//     surface in that pass as "key not found" and the child is handed to the overflow path;
//   * a child that overflows (table too full / tag clash) is finished from the host with the generic
//     LSD radix sort + run-length count restricted to that child's segment.
#pragma once
#include "msdcount.cuh"

template <int W>
struct MsdW {
    static constexpr int TILE_WORDS = 64;               // packed words per CTA tile in passes A / B
    static constexpr bool STASH = W == 2;               // pass B keeps the extracted keys in a second buffer (else re-extracts)
    static constexpr int TPW = 256 / TILE_WORDS;        // threads per packed word when the valid bits are compacted
    static constexpr int PPT = 32 / TPW;                // window positions per thread in that step
    static constexpr int TILE_KEYS = TILE_WORDS * 32;   // 2048 window starts: 32 KB (W = 2) / 64 KB (W = 4) per key buffer
    static constexpr int KPT = 16 / W;                  // keys per thread in the level-1 tiles
    static constexpr int T1 = 256 * KPT;
    static constexpr int LOG_HT = W == 2 ? 12 : 11;     // leaf table slots: 64 KB of keys
};
#define MSDW_OVERFLOW 0xFFFFFFFFu
#define MSDW_HB 14      // pass-A histogram bits of the multi-word path (doubles as the level-1 histogram when b0 + b1 <= 14)
#define MSDW_B0_MAX 12  // largest level-0 digit (bins of the pass-B tile)
static inline int msdw_hist_bits(int k) { return 2 * k < MSDW_HB ? 2 * k : MSDW_HB; }

// key bits [lowbit, lowbit + nbits), nbits <= 24
template <int W>
__device__ __forceinline__ u32 kw_bits(const Kmer<W> &x, int lowbit, int nbits) {
    if (nbits == 0) return 0u;
    const int wi = W - 1 - (lowbit >> 6), sh = lowbit & 63;
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == wi) lo = x.w[i];
        if (i == wi - 1) hi = x.w[i];
    }
    u64 v = lo >> sh;
    if (sh) v |= hi << (64 - sh);
    return (u32)v & ((1u << nbits) - 1u);
}
// Hash of a multi-word key: folded 64x64 -> 128-bit multiplies (high ^ low half) chained over the
// words, then a murmur3-style finaliser.  The tag is its low half, so every input difference must
// reach the low bits: a plain multiply only diffuses upwards, and k-mers that differ by substitutions
// in the top bases of two words (common among error variants of one genomic k-mer) then cancel each
// other -- measured on real keys before this form was adopted.
__device__ __forceinline__ u64 mum64(u64 a, u64 b) { return __umul64hi(a, b) ^ (a * b); }
template <int W>
__device__ __forceinline__ u64 kw_hash(const Kmer<W> &x) {
    const u64 c[4] = {0x9E3779B97F4A7C15ULL, 0xBF58476D1CE4E5B9ULL, 0x94D049BB133111EBULL, 0xD6E8FEB86659FD93ULL};
    u64 h = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) h = mum64(h ^ x.w[i], c[i]);
    h ^= h >> 33;
    h *= 0xFF51AFD7ED558CCDULL;
    h ^= h >> 33;
    return h;
}
// valid-window bits of the PPT positions owned by sub-thread `sub` of packed word t (MSB = first position)
template <int W>
__device__ __forceinline__ u32 msdw_vm(const u32 *__restrict__ valid, u64 t, u32 sub) {
    constexpr int PPT = MsdW<W>::PPT;
    return (valid[t] >> (32 - (sub + 1) * PPT)) & ((1u << PPT) - 1u);
}

// Valid windows are sparse for long k (24 of 150 positions at k = 127) and come in runs, so a thread
// that walked its own PPT positions would leave most lanes idle.  Each CTA tile therefore first
// compacts the valid positions of its TILE_KEYS window starts into a shared list (block scan of
// popcounts); every lane then extracts one listed window at a time.
template <int W>
__device__ __forceinline__ u32 msdw_compact_tile(const u32 *__restrict__ valid, u64 nwords, u64 tile, u16 *plist, u32 *ws) {
    constexpr int PPT = MsdW<W>::PPT;
    const u64 t = tile * MsdW<W>::TILE_WORDS + threadIdx.x / MsdW<W>::TPW;
    const u32 sub = threadIdx.x % MsdW<W>::TPW;
    u32 vm = t < nwords ? msdw_vm<W>(valid, t, sub) : 0u;
    u32 total;
    u32 base = block_excl_scan((u32)__popc(vm), ws, &total);
    while (vm) {
        const int b = 31 - __clz(vm);
        vm &= ~(1u << b);
        plist[base++] = (u16)(threadIdx.x * PPT + (PPT - 1 - b));
    }
    __syncthreads();
    return total;
}

// ---- pass A ---------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) msdw_hist_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 nwords, int k,
                                                        int canonical, int hb, u32 *__restrict__ ghist, u32 est_max,
                                                        u8 *__restrict__ est_map) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u32 *h = reinterpret_cast<u32 *>(smem_raw);  // 2^hb bins
    __shared__ u16 plist[MsdW<W>::TILE_KEYS];
    __shared__ u32 ws[33];
    const int nb = 1 << hb;
    for (int i = threadIdx.x; i < nb; i += 256) h[i] = 0;
    __syncthreads();
    const u64 ntiles = (nwords + MsdW<W>::TILE_WORDS - 1) / MsdW<W>::TILE_WORDS;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u32 np = msdw_compact_tile<W>(valid, nwords, tile, plist, ws);
        const u64 pbase = tile * MsdW<W>::TILE_KEYS;
        for (u32 q = threadIdx.x; q < np; q += 256) {
            const Kmer<W> key = kmer_at<W>(packed, pbase + plist[q], k, canonical, nullptr);
            atomicAdd(&h[kw_bits<W>(key, 2 * k - hb, hb)], 1u);
            if ((u32)key.w[W - 1] * 0x9E3779B1u <= est_max) est_map[(u32)(kw_hash<W>(key) >> (64 - EST_LOG))] = 1;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < nb; i += 256)
        if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// ---- pass B ---------------------------------------------------------------------------------------
// PEER (multi-GPU): the buckets are owned by ranks and the tile's runs are written straight into the owners' receive
// buffers over NVLink (the cursors then hold key offsets inside those buffers): the level-0 scatter IS the exchange.
template <int W, bool PEER>
__global__ void __launch_bounds__(256, 2) msdw_scatter_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 nwords,
                                                              int k, int canonical, int b0, u32 *__restrict__ cursor,
                                                              u64 *__restrict__ out, MsdPeers pt) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);                                  // TILE_KEYS * W: keys grouped by bucket
    u64 *akeys = skeys + (size_t)MsdW<W>::TILE_KEYS * W;                             // STASH: TILE_KEYS * W keys as extracted
    u32 *hist = reinterpret_cast<u32 *>(akeys + (MsdW<W>::STASH ? (size_t)MsdW<W>::TILE_KEYS * W : 0));  // 2^MSDW_B0_MAX
    u32 *delta = hist + (1 << MSDW_B0_MAX);                                          // 2^MSDW_B0_MAX
    u16 *plist = reinterpret_cast<u16 *>(delta + (1 << MSDW_B0_MAX));                // TILE_KEYS
    __shared__ u32 ws[33];
    const int nb = 1 << b0, lowbit = 2 * k - b0;
    for (int i = threadIdx.x; i < nb; i += 256) hist[i] = 0;
    const u32 np = msdw_compact_tile<W>(valid, nwords, blockIdx.x, plist, ws);  // ends with a barrier
    const u64 pbase = (u64)blockIdx.x * MsdW<W>::TILE_KEYS;
    for (u32 q = threadIdx.x; q < np; q += 256) {
        const Kmer<W> key = kmer_at<W>(packed, pbase + plist[q], k, canonical, nullptr);
        if (MsdW<W>::STASH) km_store<W>(akeys, q, key);  // extracted once, kept in shared memory until it is placed
        atomicAdd(&hist[kw_bits<W>(key, lowbit, b0)], 1u);
    }
    __syncthreads();
    const u32 total = msd_reserve<(1 << MSDW_B0_MAX) / 256>(hist, delta, nb, cursor, ws);
    for (u32 q = threadIdx.x; q < np; q += 256) {
        const Kmer<W> key = MsdW<W>::STASH ? km_load<W>(akeys, q) : kmer_at<W>(packed, pbase + plist[q], k, canonical, nullptr);
        km_store<W>(skeys, atomicAdd(&hist[kw_bits<W>(key, lowbit, b0)], 1u), key);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        const Kmer<W> key = km_load<W>(skeys, i);
        const u32 d = kw_bits<W>(key, lowbit, b0);
        km_store<W>(PEER ? pt.ptr[msd_owner(pt, d)] : out, (u32)(delta[d] + i), key);
    }
}

// ---- level 1 --------------------------------------------------------------------------------------
template <int W>
__global__ void msdw_tilecount_kernel(const u32 *__restrict__ vcnt, u32 nvb, u32 *__restrict__ tcount) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nvb) tcount[b] = b < nvb ? (vcnt[b] + MsdW<W>::T1 - 1) / MsdW<W>::T1 : 0;
}
template <int W>
__device__ __forceinline__ bool msdw_tile_range(const u32 *__restrict__ tstart, const u32 *__restrict__ vbeg, const u32 *__restrict__ vcnt,
                                                u32 nvb, u32 t, u32 &vb, u32 &s, u32 &e) {
    if (t >= tstart[nvb]) return false;
    u32 lo = 0, hi = nvb;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (tstart[mid] <= t) lo = mid; else hi = mid;
    }
    vb = lo;
    const u32 b0 = vbeg[vb];
    s = b0 + (t - tstart[vb]) * MsdW<W>::T1;
    e = min(b0 + vcnt[vb], s + MsdW<W>::T1);
    return true;
}
template <int W>
__global__ void __launch_bounds__(256) msdw_hist1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart, const u32 *__restrict__ vbeg,
                                                         const u32 *__restrict__ vcnt, u32 nvb, u32 nsrc, int sh1, int b1,
                                                         u32 *__restrict__ hist1) {
    constexpr int KPT = MsdW<W>::KPT;
    __shared__ u32 h[1 << MSD_MAX_B1];
    u32 vb, s, e;
    if (!msdw_tile_range<W>(tstart, vbeg, vcnt, nvb, blockIdx.x, vb, s, e)) return;
    const int nb = 1 << b1;
    for (int i = threadIdx.x; i < nb; i += 256) h[i] = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
        const u32 i = s + u * 256 + threadIdx.x;
        if (i < e) atomicAdd(&h[kw_bits<W>(km_load<W>(A, i), sh1, b1)], 1u);
    }
    __syncthreads();
    const u64 cb = (u64)(vb / nsrc) << b1;
    for (int i = threadIdx.x; i < nb; i += 256)
        if (h[i]) atomicAdd(&hist1[cb + i], h[i]);
}
template <int W>
__global__ void __launch_bounds__(256) msdw_scatter1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart,
                                                            const u32 *__restrict__ vbeg, const u32 *__restrict__ vcnt, u32 nvb, u32 nsrc,
                                                            int sh1, int b1, u32 *__restrict__ cursor1, u64 *__restrict__ B) {
    constexpr int KPT = MsdW<W>::KPT;
    __shared__ __align__(16) u64 skeys[MsdW<W>::T1 * W];
    __shared__ u32 hist[1 << MSD_MAX_B1];
    __shared__ u32 delta[1 << MSD_MAX_B1];
    __shared__ u32 ws[33];
    u32 vb, s, e;
    if (!msdw_tile_range<W>(tstart, vbeg, vcnt, nvb, blockIdx.x, vb, s, e)) return;
    const int nb = 1 << b1;
    for (int i = threadIdx.x; i < nb; i += 256) hist[i] = 0;
    __syncthreads();
    Kmer<W> kk[KPT];
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
        const u32 i = s + u * 256 + threadIdx.x;
        if (i < e) kk[u] = km_load<W>(A, i);
    }
#pragma unroll
    for (int u = 0; u < KPT; ++u)
        if (s + u * 256 + threadIdx.x < e) atomicAdd(&hist[kw_bits<W>(kk[u], sh1, b1)], 1u);
    __syncthreads();
    const u32 total = msd_reserve<(1 << MSD_MAX_B1) / 256>(hist, delta, nb, cursor1 + ((u64)(vb / nsrc) << b1), ws);
#pragma unroll
    for (int u = 0; u < KPT; ++u)
        if (s + u * 256 + threadIdx.x < e) km_store<W>(skeys, atomicAdd(&hist[kw_bits<W>(kk[u], sh1, b1)], 1u), kk[u]);
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        const Kmer<W> key = km_load<W>(skeys, i);
        km_store<W>(B, (u32)(delta[kw_bits<W>(key, sh1, b1)] + i), key);
    }
}

// ---- pass E: one CTA per child ----------------------------------------------------------------------
template <int W>
struct LeafSmemW {
    static constexpr int HT = 1 << MsdW<W>::LOG_HT;
    u64 tab[HT * W];
    u32 tcnt[HT];
    u32 ttag[HT];  // 0 = empty, else the claimer's tag (odd)
    u32 bins[LEAF_BINS + 1];
    u32 bincur[LEAF_BINS];
    u32 ws[33];
    u32 s_task, s_ovf;
};

// child = X[start .. start + n): counted in place (unique keys + counts to X / OC[start ..)); SB is scratch.
// Returns the number of kept distinct keys, or MSDW_OVERFLOW (nothing written).
template <int W, int THREADS>
__device__ u32 msdw_leaf(LeafSmemW<W> &S, u64 *X, u64 *SB, u32 *OC, u32 start, u32 n, int shift, u32 min_freq) {
    constexpr int LOG_HT = MsdW<W>::LOG_HT;
    constexpr int HT = 1 << LOG_HT;
    constexpr u32 HM = HT - 1;
    for (int i = threadIdx.x; i < HT; i += THREADS) { S.ttag[i] = 0; S.tcnt[i] = 0; }
    if (threadIdx.x == 0) S.s_ovf = 0;
    __syncthreads();
    volatile u32 *ovf = &S.s_ovf;
    // pass 1: claim one slot per distinct tag; the claimer stores the key
    for (u32 i = threadIdx.x; i < n; i += THREADS) {
        const Kmer<W> key = km_load<W>(X, (u64)start + i);
        const u64 h = kw_hash<W>(key);
        const u32 tag = (u32)h | 1u;
        u32 slot = (u32)(h >> (64 - LOG_HT));
        const u32 step = ((u32)(h >> 24) | 1u) & HM;  // double hashing: an odd stride visits every slot, no primary clustering
        for (int probes = 0;; ++probes) {
            u32 t = S.ttag[slot];
            if (t == 0) {
                t = atomicCAS(&S.ttag[slot], 0u, tag);
                if (t == 0) { km_store<W>(S.tab, slot, key); break; }
            }
            if (t == tag) break;
            slot = (slot + step) & HM;
            if (probes > LEAF_MAX_PROBE) { *ovf = 1; break; }
        }
        if (*ovf) break;
    }
    __syncthreads();
    if (S.s_ovf) return MSDW_OVERFLOW;
    // pass 2: count, comparing the full key
    for (u32 i = threadIdx.x; i < n; i += THREADS) {
        const Kmer<W> key = km_load<W>(X, (u64)start + i);
        const u64 h = kw_hash<W>(key);
        const u32 tag = (u32)h | 1u;
        u32 slot = (u32)(h >> (64 - LOG_HT));
        const u32 step = ((u32)(h >> 24) | 1u) & HM;  // double hashing: an odd stride visits every slot, no primary clustering
        for (int probes = 0;; ++probes) {
            const u32 t = S.ttag[slot];
            if (t == tag && km_eq<W>(km_load<W>(S.tab, slot), key)) { atomicAdd(&S.tcnt[slot], 1u); break; }
            if (t == 0 || probes > LEAF_MAX_PROBE + 1) { *ovf = 1; break; }  // empty slot: this key never got one (tag clash)  // tag clash: this key never got a slot
            slot = (slot + step) & HM;
        }
        if (*ovf) break;
    }
    for (int i = threadIdx.x; i <= LEAF_BINS; i += THREADS) S.bins[i] = 0;
    __syncthreads();
    if (S.s_ovf) return MSDW_OVERFLOW;
    // order the kept keys: order-preserving bins over the top varying bits, rank by counting inside a bin
    const int bb = shift < 10 ? shift : 10, blow = shift - bb;
    const int nbins = 1 << bb;
    for (int slot = threadIdx.x; slot < HT; slot += THREADS)
        if (S.ttag[slot] && S.tcnt[slot] >= min_freq) atomicAdd(&S.bins[kw_bits<W>(km_load<W>(S.tab, slot), blow, bb)], 1u);
    __syncthreads();
    const u32 nd = block_scan_array<THREADS>(S.bins, nbins, S.ws);
    for (int i = threadIdx.x; i < nbins; i += THREADS) S.bincur[i] = S.bins[i];
    if (threadIdx.x == 0) S.bins[nbins] = nd;
    __syncthreads();
    for (int slot = threadIdx.x; slot < HT; slot += THREADS)
        if (S.ttag[slot] && S.tcnt[slot] >= min_freq) {
            const Kmer<W> key = km_load<W>(S.tab, slot);
            km_store<W>(SB, (u64)start + atomicAdd(&S.bincur[kw_bits<W>(key, blow, bb)], 1u), key);
        }
    __syncthreads();
    for (u32 j = threadIdx.x; j < nd; j += THREADS) {
        const Kmer<W> key = km_load<W>(SB, (u64)start + j);
        const u32 b = kw_bits<W>(key, blow, bb);
        const u32 lo = S.bins[b], hi = S.bins[b + 1];
        u32 rank = 0;
        for (u32 q = lo; q < hi; ++q) rank += km_lt<W>(km_load<W>(SB, (u64)start + q), key) ? 1u : 0u;
        const u64 h = kw_hash<W>(key);
        const u32 tag = (u32)h | 1u;
        u32 slot = (u32)(h >> (64 - LOG_HT));
        const u32 step = ((u32)(h >> 24) | 1u) & HM;  // double hashing: an odd stride visits every slot, no primary clustering
        while (!(S.ttag[slot] == tag && km_eq<W>(km_load<W>(S.tab, slot), key))) slot = (slot + step) & HM;
        km_store<W>(X, (u64)start + lo + rank, key);
        OC[start + lo + rank] = S.tcnt[slot];
    }
    __syncthreads();
    return nd;
}

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) msdw_leaf_kernel(u64 *X, u64 *SB, u32 *OC, const u32 *__restrict__ off, u32 nchild, int shiftc,
                                                            u32 *__restrict__ nu_out, u32 *ticket, u32 min_freq, u32 test_ovf_mod) {
    extern __shared__ __align__(16) u8 smem_raw[];
    LeafSmemW<W> &S = *reinterpret_cast<LeafSmemW<W> *>(smem_raw);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) S.s_task = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 c = S.s_task;
        if (c >= nchild) break;
        const u32 ob = off[c], n = off[c + 1] - ob;
        if (n == 0) { if (threadIdx.x == 0) nu_out[c] = 0; continue; }
        if (shiftc == 0) {  // no varying bits: the whole child is one k-mer (already at X[ob])
            if (threadIdx.x == 0) {
                const u32 keep = n >= min_freq;
                if (keep) OC[ob] = n;
                nu_out[c] = keep;
            }
            continue;
        }
        const u32 nd = (test_ovf_mod && c % test_ovf_mod == 0) ? MSDW_OVERFLOW  // tests: exercise the overflow path
                                                               : msdw_leaf<W, THREADS>(S, X, SB, OC, ob, n, shiftc, min_freq);
        if (threadIdx.x == 0) nu_out[c] = nd;
    }
}
template <int W>
__global__ void __launch_bounds__(256) msdw_gather_kernel(const u64 *__restrict__ X, const u32 *__restrict__ cnt, const u32 *__restrict__ off,
                                                          const u32 *__restrict__ nu, const u32 *__restrict__ ustart, u32 nchild,
                                                          u64 *__restrict__ okeys, u32 *__restrict__ ocnt) {
    for (u32 t = blockIdx.x; t < nchild; t += gridDim.x) {
        const u32 n = nu[t], s = off[t], d = ustart[t];
        for (u32 i = threadIdx.x; i < n; i += 256) {
            km_store<W>(okeys, (u64)d + i, km_load<W>(X, (u64)s + i));
            ocnt[d + i] = cnt[s + i];
        }
    }
}
struct MsdwOverflowF {
    const u32 *nu;
    u32 *list;
    __device__ bool pred(u64 c) const { return nu[c] == MSDW_OVERFLOW; }
    __device__ void emit(u64 c, u64 dst) const { list[dst] = (u32)c; }
};

// ------------------------------------- host side -------------------------------------------------
template <int W>
static MsdTune msdw_tune() {
    MsdTune tune = msd_tune();
    tune.dt = tune.dt / W < 16 ? 16 : tune.dt / W;  // the leaf table holds 8192 / W slots
    if (tune.b0_max > MSDW_B0_MAX) tune.b0_max = MSDW_B0_MAX;
    return tune;
}
// pass A, launch (see msd_pass_a_launch)
template <int W>
static void msdw_pass_a_launch(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, int est_shift, DBuf<u32> &hist,
                               DBuf<u8> &est) {
    const int hb = msdw_hist_bits(k);
    const u32 nbh = 1u << hb;
    hist.alloc(ctx, nbh + 1);
    est.alloc(ctx, 1ULL << EST_LOG);
    CK(cudaMemsetAsync(hist.p, 0, (nbh + 1) * sizeof(u32), ctx->stream));
    CK(cudaMemsetAsync(est.p, 0, 1ULL << EST_LOG, ctx->stream));
    if (pr.nwords == 0) return;
    u64 hgrid = ceil_div(pr.nwords, MsdW<W>::TILE_WORDS);
    if (hgrid > (u64)ctx->num_sms * 3) hgrid = (u64)ctx->num_sms * 3;  // 64 KB of histogram per CTA
    const size_t smemA = sizeof(u32) << MSDW_HB;
    set_max_smem(ctx, msdw_hist_kernel<W>, smemA);
    prof_bytes(ctx, 12.0 * (double)pr.nwords);   // packed words + window masks read once
    LAUNCH(ctx, msdw_hist_kernel<W>, (unsigned)hgrid, 256, smemA, pr.packed.p, valid, pr.nwords, k, canonical, hb, hist.p,
           est_shift ? (0xFFFFFFFFu >> est_shift) : 0xFFFFFFFFu, est.p);
}
// pass B: keys scattered into 2^b0 buckets of A (N * W words)
template <int W>
static void msdw_pass_b(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, const MsdPlan &pl, const u32 *hist, u64 N,
                        DBuf<u32> &hist0, DBuf<u32> &bstart, DBuf<u64> &A, const MsdPeers *peers = nullptr, u32 *peer_cursor = nullptr) {
    const u32 nb0 = 1u << pl.b0;
    hist0.alloc(ctx, nb0 + 1);
    bstart.alloc(ctx, nb0 + 1);
    DBuf<u32> cursor(ctx, nb0);
    LAUNCH(ctx, msd_collapse_kernel, (unsigned)ceil_div(nb0 + 1, 256), 256, 0, hist, pl.hb, pl.b0, hist0.p);
    exclusive_scan_u32(ctx, hist0.p, bstart.p, nb0 + 1, nullptr);
    LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nb0, 256), 256, 0, bstart.p, (u64)nb0, cursor.p);
    A.alloc(ctx, ((N && !peers) ? N : 1) * W);
    if (N == 0) return;
    const size_t smemB = (MsdW<W>::STASH ? 2 : 1) * (size_t)MsdW<W>::TILE_KEYS * W * 8 + 2 * (size_t)(1 << MSDW_B0_MAX) * 4 + (size_t)MsdW<W>::TILE_KEYS * 2;
    set_max_smem(ctx, (msdw_scatter_kernel<W, false>), smemB);
    set_max_smem(ctx, (msdw_scatter_kernel<W, true>), smemB);
    prof_bytes(ctx, 12.0 * (double)pr.nwords + 8.0 * W * (double)N);   // masks + packed words in, every key out once
    MsdPeers none;
    memset(&none, 0, sizeof(none));
    if (peers)
        LAUNCH(ctx, (msdw_scatter_kernel<W, true>), (unsigned)ceil_div(pr.nwords, MsdW<W>::TILE_WORDS), 256, smemB, pr.packed.p, valid, pr.nwords, k,
               canonical, pl.b0, peer_cursor, (u64 *)nullptr, *peers);
    else
        LAUNCH(ctx, (msdw_scatter_kernel<W, false>), (unsigned)ceil_div(pr.nwords, MsdW<W>::TILE_WORDS), 256, smemB, pr.packed.p, valid, pr.nwords, k,
               canonical, pl.b0, cursor.p, A.p, none);
}
// Level-0 buckets -> sorted unique keys + counts (see msd_finish).  With nsrc > 1 (multi-GPU) level 1 always
// runs (with b1 = 0 bits if need be): it merges the pieces of a bucket into one contiguous child.
template <int W>
static u64 msdw_finish(gg_ctx *ctx, u64 *IN, u64 n_in, const u32 *vbeg, const u32 *vcnt, u32 nb_local, u32 nsrc, const MsdPlan &pl,
                       const MsdTune &tune, u32 min_freq, DBuf<u64> &ukeys, DBuf<u32> &ucounts, const u32 *child_hist = nullptr) {
    if (n_in == 0) { ukeys.alloc(ctx, W); ucounts.alloc(ctx, 1); return 0; }
    const u32 nvb = nb_local * nsrc;
    const u64 nchild = (u64)nb_local << pl.b1;
    stage_begin(ctx, ST_SORT);
    DBuf<u32> off(ctx, nchild + 1), OC(ctx, n_in), ticket(ctx, 1), nu(ctx, nchild + 1), ustart(ctx, nchild + 1);
    DBuf<u64> tot(ctx, 1), B(ctx, n_in * W);
    u64 *X = IN, *SB = B.p;
    if (pl.b1 > 0 || nsrc > 1) {
        DBuf<u32> tcount(ctx, nvb + 1), tstart(ctx, nvb + 1), cursor1(ctx, nchild), hist1;
        LAUNCH(ctx, msdw_tilecount_kernel<W>, (unsigned)ceil_div(nvb + 1, 256), 256, 0, vcnt, nvb, tcount.p);
        exclusive_scan_u32(ctx, tcount.p, tstart.p, nvb + 1, nullptr);
        const unsigned tgrid = (unsigned)(ceil_div(n_in, MsdW<W>::T1) + nvb);
        const u32 *h1 = child_hist;
        if (!h1) {  // pass C: pass A's histogram is too coarse for b0 + b1 bits
            hist1.alloc(ctx, nchild + 1);
            CK(cudaMemsetAsync(hist1.p, 0, (nchild + 1) * sizeof(u32), ctx->stream));
            prof_bytes(ctx, 8.0 * W * (double)n_in);
            LAUNCH(ctx, msdw_hist1_kernel<W>, tgrid, 256, 0, IN, tstart.p, vbeg, vcnt, nvb, nsrc, pl.shiftc, pl.b1, hist1.p);
            h1 = hist1.p;
        }
        exclusive_scan_u32(ctx, h1, off.p, nchild + 1, nullptr);
        LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nchild, 256), 256, 0, off.p, nchild, cursor1.p);
        prof_bytes(ctx, 16.0 * W * (double)n_in);   // every key in and out once
        LAUNCH(ctx, msdw_scatter1_kernel<W>, tgrid, 256, 0, IN, tstart.p, vbeg, vcnt, nvb, nsrc, pl.shiftc, pl.b1, cursor1.p, B.p);
        X = B.p; SB = IN;
    } else {
        LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, vbeg, nchild + 1, off.p);  // contiguous buckets: off == vbeg
    }
    // ---- E ----
    CK(cudaMemsetAsync(ticket.p, 0, sizeof(u32), ctx->stream));
    {
        const size_t smem = sizeof(LeafSmemW<W>);
        set_max_smem(ctx, msdw_leaf_kernel<W, 512>, smem);
        u32 grid = (u32)ctx->num_sms * 2;
        if (grid > nchild) grid = (u32)nchild;
        prof_bytes(ctx, 16.0 * W * (double)n_in);   // two sweeps over the child's keys (claim, then count); + (8 W + 4) per kept key, added below
        LAUNCH(ctx, (msdw_leaf_kernel<W, 512>), grid, 512, smem, X, SB, OC.p, off.p, (u32)nchild, pl.shiftc, nu.p, ticket.p, min_freq, tune.test_ovf_mod);
    }
    // children the table could not hold: generic sort + run-length count of that segment
    {
        DBuf<u32> olist(ctx, nchild);
        MsdwOverflowF of{nu.p, olist.p};
        Compactor<MsdwOverflowF> oc(ctx, nchild, "msdw_overflow");
        const u64 novf = oc.count(of);
        if (getenv("GG_MSD_DEBUG"))
            fprintf(stderr, "[msdw W=%d] n_in=%llu b0=%d b1=%d shiftc=%d children=%llu overflowed=%llu\n", W, (unsigned long long)n_in, pl.b0, pl.b1,
                    pl.shiftc, (unsigned long long)nchild, (unsigned long long)novf);
        if (novf) {
            oc.emit(of);
            std::vector<u32> h_list(novf), h_off(nchild + 1);
            CK(cudaMemcpyAsync(h_list.data(), olist.p, novf * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(h_off.data(), off.p, (nchild + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (u32 c : h_list) {
                const u64 s = h_off[c], n = h_off[c + 1] - h_off[c];
                u64 *a = X + s * W, *b = SB + s * W;
                std::vector<BitRange> ranges = {{0, pl.shiftc}};
                radix_sort<W>(ctx, &a, &b, nullptr, nullptr, n, ranges);
                DBuf<u64> uk;
                DBuf<u32> uc;
                const u64 nuq = run_length_count<W>(ctx, a, n, uk, uc);
                FilterF<W> ff{uk.p, uc.p, X + s * W, OC.p + s, min_freq, 0};
                Compactor<FilterF<W>> fc(ctx, nuq, "filter");
                const u32 kept = (u32)fc.count(ff);
                fc.emit(ff);
                CK(cudaMemcpyAsync(nu.p + c, &kept, sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
        }
    }
    stage_end(ctx, ST_SORT);
    // ---- F ----
    stage_begin(ctx, ST_COUNT);
    exclusive_scan_u32(ctx, nu.p, ustart.p, nchild, tot.p);
    const u64 U = read_scalar<u64>(ctx, tot.p);
    ukeys.alloc(ctx, (U ? U : 1) * W);
    ucounts.alloc(ctx, U ? U : 1);
    if (U) {
        u32 ggrid = nchild < (u64)ctx->num_sms * 16 ? (u32)nchild : (u32)ctx->num_sms * 16;
        prof_bytes(ctx, 2.0 * (8.0 * W + 4.0) * (double)U);
        LAUNCH(ctx, msdw_gather_kernel<W>, ggrid, 256, 0, X, OC.p, off.p, nu.p, ustart.p, (u32)nchild, ukeys.p, ucounts.p);
    }
    stage_end(ctx, ST_COUNT);
    return U;
}

// packed reads -> sorted unique k-mers + counts (count >= min_freq only), 32 < k <= 128.
template <int W>
static u64 msdw_count(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u64> &ukeys,
                      DBuf<u32> &ucounts, u64 *n_instances) {
    const MsdTune tune = msdw_tune<W>();
    *n_instances = 0;
    if (pr.nwords == 0) { ukeys.alloc(ctx, W); ucounts.alloc(ctx, 1); return 0; }
    stage_begin(ctx, ST_EXTRACT);
    const int hb = msdw_hist_bits(k);
    DBuf<u32> hist, hist0, bstart;
    DBuf<u8> est;
    DBuf<u64> A;
    const int est_shift = msd_est_shift(pr.nbases);
    msdw_pass_a_launch<W>(ctx, pr, valid, k, canonical, est_shift, hist, est);
    u64 N = 0, dest = 0;
    msd_pass_a_read(ctx, hist.p, 1u << hb, est.p, est_shift, &N, &dest);
    est.release();
    *n_instances = N;
    if (N == 0) { stage_end(ctx, ST_EXTRACT); ukeys.alloc(ctx, W); ucounts.alloc(ctx, 1); return 0; }
    if (N >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu k-mer instances exceed the 2^32-1 per-call limit", (unsigned long long)N);
    const MsdPlan pl = msd_plan(dest, N, k, tune, hb);
    msdw_pass_b<W>(ctx, pr, valid, k, canonical, pl, hist.p, N, hist0, bstart, A);
    DBuf<u32> child_hist;
    if (pl.b1 > 0 && pl.b0 + pl.b1 <= pl.hb) {  // pass C for free: collapse the pass-A histogram to b0 + b1 bits
        const u32 nchild = 1u << (pl.b0 + pl.b1);
        child_hist.alloc(ctx, nchild + 1);
        LAUNCH(ctx, msd_collapse_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, hist.p, pl.hb, pl.b0 + pl.b1, child_hist.p);
    }
    stage_end(ctx, ST_EXTRACT);
    return msdw_finish<W>(ctx, A.p, N, bstart.p, hist0.p, 1u << pl.b0, 1, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
}
