// msdcount.cuh -- K2+K3 fused for k <= 32 (one u64 word per k-mer): counting without a
// global sort of all instances.
//
// Replaces (reference) KmerAnalysis serial_mem: collect every canonical k-mer -> sort! ->
// run-length collapse (docs/src/man/assembly.md:126-128).  Output is identical (sorted unique
// k-mers + counts); the route is an MSD formulation sized for HBM traffic:
//
//   A  msd_hist_kernel      roll over the packed reads, histogram of the top b0 key bits
//   B  msd_scatter_kernel   one roll per 32 windows (keys stay in registers), tile staged in
//                           shared memory grouped by bucket, one reservation per (tile, bucket)
//   C  msd_hist1_kernel     per level-0 bucket, histogram of the next b1 bits
//   D  msd_scatter1_kernel  same staging scheme, keys -> 2^(b0+b1) ordered children
//   E  task list            runs of consecutive small children (<= CAP keys together)
//   F  msd_task_kernel      one CTA per task: keys counted in a shared-memory hash table, the
//                           distinct keys put in order by order-preserving bins + rank-by-counting
//                           and written in place.  A child larger than CAP is partitioned again
//                           by the CTA (8 bits per level, depth first); a range with no key bits
//                           left is one run.
//   G  msd_gather_kernel    per-task unique runs -> dense output arrays
//
// Algorithmic bytes per k-mer instance (8-byte keys): A 0.31 read; B 0.31 read + 8 written;
// C 8 read; D 8 read + 8 written; F 8 read + 12 U/N written; G 24 U/N.
#pragma once
#include "pack.cuh"
#include "sort.cuh"

#define MSD_MAX_B0 11
#define MSD_MAX_B1 10
#define MSD_UNROLL 8  // independent global loads in flight per thread in the task kernel
#define MSD_EMPTY 0xFFFFFFFFFFFFFFFFULL  // never a valid key: all-T k-mers are not canonical, and 2k < 64 otherwise

// 32 consecutive windows starting at base 32 t (k <= 32): f(key, j) for every valid window j
template <class F>
__device__ __forceinline__ void roll32(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 t, int k, int canonical, F &&f) {
    const u32 vm = valid[t];
    if (vm == 0) return;
    const u64 a0 = packed[t], a1 = packed[t + 1];
    const u64 mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    const int top = 2 * (k - 1);
    u64 fw = a0 >> (64 - 2 * k);
    u64 rc = rev2_u64(~fw) >> (64 - 2 * k);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if ((vm >> (31 - j)) & 1u) f((canonical && rc < fw) ? rc : fw, j);
        if (j < 31) {
            int o = k + j;  // base index relative to 32 t, 1 .. 62
            u64 w = o < 32 ? a0 : a1;
            u64 nb = (w >> (62 - 2 * (o & 31))) & 3ULL;
            fw = ((fw << 2) | nb) & mask;
            rc = (rc >> 2) | ((3ULL - nb) << top);
        }
    }
}

__device__ __forceinline__ u32 msd_digit0(u64 key, int sh, int b0) { return b0 ? (u32)(key >> sh) : 0u; }

__global__ void __launch_bounds__(256) msd_hist_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 nwords,
                                                       int k, int canonical, int b0, u32 *__restrict__ ghist) {
    __shared__ u32 h[1 << MSD_MAX_B0];
    const int nb = 1 << b0, sh = 2 * k - b0;
    for (int i = threadIdx.x; i < nb; i += 256) h[i] = 0;
    __syncthreads();
    for (u64 t = (u64)blockIdx.x * 256 + threadIdx.x; t < nwords; t += (u64)gridDim.x * 256)
        roll32(packed, valid, t, k, canonical, [&](u64 key, int) { atomicAdd(&h[msd_digit0(key, sh, b0)], 1u); });
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += 256)
        if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// shared tail of the two scatter kernels: bins counted in hist[0..nb) -> per-bin reservation in
// the global cursors, hist becomes the in-tile running cursor, delta the global-minus-local base
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

// tile = 256 words = 8192 window positions
#define MSD_TILE_KEYS 8192
__global__ void __launch_bounds__(256) msd_scatter_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 nwords,
                                                          int k, int canonical, int b0, u32 *__restrict__ cursor, u64 *__restrict__ out) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *skeys = reinterpret_cast<u64 *>(smem_raw);                      // MSD_TILE_KEYS
    u32 *hist = reinterpret_cast<u32 *>(skeys + MSD_TILE_KEYS);          // 2^b0
    u32 *delta = hist + (1 << MSD_MAX_B0);                               // 2^b0
    __shared__ u32 ws[33];
    const int nb = 1 << b0, sh = 2 * k - b0;
    const u64 t = (u64)blockIdx.x * 256 + threadIdx.x;
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
    const u32 total = msd_reserve<8>(hist, delta, nb, cursor, ws);
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if ((vm >> (31 - j)) & 1u) skeys[atomicAdd(&hist[msd_digit0(keys[j], sh, b0)], 1u)] = keys[j];
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        u64 key = skeys[i];
        out[delta[msd_digit0(key, sh, b0)] + i] = key;
    }
}

// ------------------------------ level 1: bucket-aligned tiles --------------------------------
#define MSD_T1 4096  // keys per tile (16 per thread)
__global__ void msd_tilecount_kernel(const u32 *__restrict__ bstart, u32 nb, u32 *__restrict__ tcount) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nb) tcount[b] = b < nb ? (bstart[b + 1] - bstart[b] + MSD_T1 - 1) / MSD_T1 : 0;
}
// tile t -> (bucket, key range)
__device__ __forceinline__ bool msd_tile_range(const u32 *__restrict__ tstart, const u32 *__restrict__ bstart, u32 nb, u32 t, u32 &b,
                                               u32 &s, u32 &e) {
    if (t >= tstart[nb]) return false;
    u32 lo = 0, hi = nb;  // largest b with tstart[b] <= t
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (tstart[mid] <= t) lo = mid; else hi = mid;
    }
    b = lo;
    s = bstart[b] + (t - tstart[b]) * MSD_T1;
    e = min(bstart[b + 1], s + MSD_T1);
    return true;
}
__global__ void __launch_bounds__(256) msd_hist1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart,
                                                        const u32 *__restrict__ bstart, u32 nb0, int sh1, int b1,
                                                        u32 *__restrict__ hist1) {
    __shared__ u32 h[1 << MSD_MAX_B1];
    u32 b, s, e;
    if (!msd_tile_range(tstart, bstart, nb0, blockIdx.x, b, s, e)) return;
    const int nb = 1 << b1;
    const u32 dm = (u32)nb - 1u;
    for (int i = threadIdx.x; i < nb; i += 256) h[i] = 0;
    __syncthreads();
    u64 kk[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { u32 i = s + u * 256 + threadIdx.x; kk[u] = i < e ? A[i] : 0; }
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) atomicAdd(&h[(u32)(kk[u] >> sh1) & dm], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += 256)
        if (h[i]) atomicAdd(&hist1[((u64)b << b1) + i], h[i]);
}
__global__ void __launch_bounds__(256) msd_scatter1_kernel(const u64 *__restrict__ A, const u32 *__restrict__ tstart,
                                                           const u32 *__restrict__ bstart, u32 nb0, int sh1, int b1,
                                                           u32 *__restrict__ cursor1, u64 *__restrict__ B) {
    __shared__ u64 skeys[MSD_T1];
    __shared__ u32 hist[1 << MSD_MAX_B1];
    __shared__ u32 delta[1 << MSD_MAX_B1];
    __shared__ u32 ws[33];
    u32 b, s, e;
    if (!msd_tile_range(tstart, bstart, nb0, blockIdx.x, b, s, e)) return;
    const int nb = 1 << b1;
    const u32 dm = (u32)nb - 1u;
    for (int i = threadIdx.x; i < nb; i += 256) hist[i] = 0;
    __syncthreads();
    u64 kk[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { u32 i = s + u * 256 + threadIdx.x; kk[u] = i < e ? A[i] : 0; }
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) atomicAdd(&hist[(u32)(kk[u] >> sh1) & dm], 1u);
    __syncthreads();
    const u32 total = msd_reserve<4>(hist, delta, nb, cursor1 + ((u64)b << b1), ws);
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (s + u * 256 + threadIdx.x < e) skeys[atomicAdd(&hist[(u32)(kk[u] >> sh1) & dm], 1u)] = kk[u];
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += 256) {
        u64 key = skeys[i];
        B[delta[(u32)(key >> sh1) & dm] + i] = key;
    }
}
__global__ void msd_copy_u32_kernel(const u32 *__restrict__ in, u64 n, u32 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ---- task list: consecutive children are merged while they stay small -----------------------
// A child larger than SMALL keys is a task of its own; otherwise children whose first key falls
// in the same window of GW keys form one task, so a merged task holds <= GW + SMALL <= CAP keys.
struct MsdTaskF {
    const u32 *off;  // nchild + 1
    u32 *tasks;
    u32 gw, small;
    __device__ bool pred(u64 c) const {
        if (c == 0) return true;
        u32 o0 = off[c - 1], o1 = off[c], o2 = off[c + 1];
        return (o2 - o1) > small || (o1 - o0) > small || (o1 / gw) != (o0 / gw);
    }
    __device__ void emit(u64 c, u64 dst) const { tasks[dst] = (u32)c; }
};

// ------------------------------- per-task CTA ------------------------------------------------
#define MSD_MAXD 9
struct MsdFrame {
    u32 start, size;      // range (absolute key index) being partitioned
    int shift_before;     // key bits [0, shift_before) still vary inside the range
    int bits;             // digit = bits [shift_before - bits, shift_before)
    int dst;              // buffer (0 = the task's own buffer, 1 = the other one) holding the children
    u32 ntasks;
    u32 off[257];         // child offsets relative to start
    u16 task[258];        // first child of every task; task[ntasks] = nchildren
};

template <int THREADS, int CAP, int HT>
struct MsdSmem {
    static constexpr int BINS = 1024;
    u64 tab[HT];
    u64 lkey[CAP];
    u32 tcnt[HT];
    u32 lcnt[CAP];
    u32 bins[BINS + 1];
    u32 bincur[BINS];
    u32 pcur[256];
    u32 ws[33];
    u32 s_task, s_nd, s_out;
    MsdFrame frames[MSD_MAXD];
};

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

template <int THREADS, int CAP, int HT>
__device__ void msd_partition(MsdSmem<THREADS, CAP, HT> &S, MsdFrame &f, const u64 *src, u64 *dst) {
    const int nch = 1 << f.bits, sh = f.shift_before - f.bits;
    const u32 dm = (u32)nch - 1u;
    for (int i = threadIdx.x; i <= 256; i += THREADS) f.off[i] = 0;
    __syncthreads();
    const u64 *sp = src + f.start;
    for (u32 i0 = threadIdx.x; i0 < f.size; i0 += MSD_UNROLL * THREADS) {  // batch the loads: memory-level parallelism
        u64 kk[MSD_UNROLL];
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i0 + u * THREADS; kk[u] = i < f.size ? sp[i] : 0; }
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u)
            if (i0 + u * THREADS < f.size) atomicAdd(&f.off[(u32)(kk[u] >> sh) & dm], 1u);
    }
    __syncthreads();
    block_scan_array<THREADS>(f.off, 256, S.ws);
    if (threadIdx.x < 256) S.pcur[threadIdx.x] = f.off[threadIdx.x];
    if (threadIdx.x == 0) f.off[256] = f.size;
    __syncthreads();
    u64 *dp = dst + f.start;
    for (u32 i0 = threadIdx.x; i0 < f.size; i0 += MSD_UNROLL * THREADS) {
        u64 kk[MSD_UNROLL];
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i0 + u * THREADS; kk[u] = i < f.size ? sp[i] : 0; }
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u)
            if (i0 + u * THREADS < f.size) dp[atomicAdd(&S.pcur[(u32)(kk[u] >> sh) & dm], 1u)] = kk[u];
    }
    __syncthreads();
    // sub-tasks: same merging rule as MsdTaskF with GW = SMALL = CAP / 2
    constexpr u32 G = CAP / 2;
    u32 flag = 0;
    const int c = threadIdx.x;
    if (c < 256) {
        u32 sz = f.off[c + 1] - f.off[c];
        bool big = sz > G;
        bool prevbig = c > 0 && (f.off[c] - f.off[c - 1]) > G;
        flag = (c == 0 || big || prevbig || (f.off[c] / G != f.off[c - 1] / G)) ? 1u : 0u;
    }
    u32 total;
    u32 idx = block_excl_scan(flag, S.ws, &total);
    if (flag) f.task[idx] = (u16)c;
    if (threadIdx.x == 0) { f.task[total] = 256; f.ntasks = total; }
    __syncthreads();
}

// count the n keys of buf[start ..) in the hash table, order the distinct ones, append them
template <int THREADS, int CAP, int HT>
__device__ void msd_leaf(MsdSmem<THREADS, CAP, HT> &S, const u64 *buf, u32 start, u32 n, int shift, u64 *okeys, u32 *ocnt, u32 obase,
                         u32 min_freq) {
    constexpr int LOG_HT = (HT == 2048) ? 11 : (HT == 4096) ? 12 : (HT == 8192) ? 13 : (HT == 16384) ? 14 : 0;
    static_assert(LOG_HT != 0 && HT % THREADS == 0 && CAP <= HT * 3 / 4, "unsupported leaf geometry");
    for (int i = threadIdx.x; i < HT; i += THREADS) { S.tab[i] = MSD_EMPTY; S.tcnt[i] = 0; }
    if (threadIdx.x == 0) S.s_nd = 0;
    __syncthreads();
    const u64 *bp = buf + start;
    for (u32 i0 = threadIdx.x; i0 < n; i0 += MSD_UNROLL * THREADS) {
        u64 kk[MSD_UNROLL];
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u) { u32 i = i0 + u * THREADS; kk[u] = i < n ? bp[i] : MSD_EMPTY; }
#pragma unroll
        for (int u = 0; u < MSD_UNROLL; ++u) {
            const u64 key = kk[u];
            if (key == MSD_EMPTY) continue;
            u32 slot = (u32)((key * 0x9E3779B97F4A7C15ULL) >> (64 - LOG_HT));
            for (;;) {
                u64 cur = S.tab[slot];
                if (cur == MSD_EMPTY) cur = atomicCAS((unsigned long long *)&S.tab[slot], (unsigned long long)MSD_EMPTY, (unsigned long long)key);
                if (cur == MSD_EMPTY || cur == key) { atomicAdd(&S.tcnt[slot], 1u); break; }
                slot = (slot + 1) & (HT - 1);
            }
        }
    }
    __syncthreads();
    // compact the occupied slots that pass the min-count filter (warp-aggregated append)
    for (int base = 0; base < HT; base += THREADS) {
        int slot = base + threadIdx.x;
        u64 key = S.tab[slot];
        u32 c = S.tcnt[slot];
        bool keep = key != MSD_EMPTY && c >= min_freq;
        u32 b = __ballot_sync(FULL_MASK, keep);
        u32 wbase = 0;
        if (lane_id() == 0 && b) wbase = atomicAdd(&S.s_nd, (u32)__popc(b));
        wbase = __shfl_sync(FULL_MASK, wbase, 0);
        if (keep) {
            u32 p = wbase + __popc(b & lanemask_lt());
            S.lkey[p] = key;
            S.lcnt[p] = c;
        }
    }
    __syncthreads();
    const u32 nd = S.s_nd;
    if (nd == 0) return;
    // order-preserving bins over the top bb varying bits, then rank by counting inside the bin
    int bb = 0;
    while ((1u << (bb + 3)) <= nd && bb < 10) ++bb;  // ~4..8 keys per bin on uniform keys
    if (bb > shift) bb = shift;
    const int nbins = 1 << bb, bsh = shift - bb;
    const u32 bm = (u32)nbins - 1u;
    for (int i = threadIdx.x; i <= nbins; i += THREADS) S.bins[i] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < nd; i += THREADS) atomicAdd(&S.bins[bb ? ((u32)(S.lkey[i] >> bsh) & bm) : 0u], 1u);
    __syncthreads();
    block_scan_array<THREADS>(S.bins, nbins, S.ws);
    for (int i = threadIdx.x; i < nbins; i += THREADS) S.bincur[i] = S.bins[i];
    if (threadIdx.x == 0) S.bins[nbins] = nd;
    __syncthreads();
    u64 *lkey2 = S.tab;   // the table is dead now: reuse it for the binned order
    u32 *lcnt2 = S.tcnt;
    for (u32 i = threadIdx.x; i < nd; i += THREADS) {
        u64 key = S.lkey[i];
        u32 r = atomicAdd(&S.bincur[bb ? ((u32)(key >> bsh) & bm) : 0u], 1u);
        lkey2[r] = key;
        lcnt2[r] = S.lcnt[i];
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < nd; j += THREADS) {
        u64 key = lkey2[j];
        u32 b = bb ? ((u32)(key >> bsh) & bm) : 0u;
        u32 lo = S.bins[b], hi = S.bins[b + 1], rank = 0;
        for (u32 q = lo; q < hi; ++q) rank += lkey2[q] < key;
        okeys[obase + lo + rank] = key;
        ocnt[obase + lo + rank] = lcnt2[j];
    }
    __syncthreads();
}

// X: buffer holding the children (read, then overwritten in place with the unique keys);
// Y: the other buffer (scratch of the depth-first fallback).
template <int THREADS, int CAP, int HT>
__global__ void __launch_bounds__(THREADS) msd_task_kernel(u64 *X, u64 *Y, u32 *cnt_out, const u32 *__restrict__ off1,
                                                           const u32 *__restrict__ tasks, u32 ntasks, u32 nchild, int shiftc,
                                                           u32 *__restrict__ nu_out, u32 *ticket, u32 min_freq) {
    extern __shared__ __align__(16) u8 smem_raw[];
    MsdSmem<THREADS, CAP, HT> &S = *reinterpret_cast<MsdSmem<THREADS, CAP, HT> *>(smem_raw);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { S.s_task = atomicAdd(ticket, 1u); S.s_out = 0; }
        __syncthreads();
        const u32 t = S.s_task;
        if (t >= ntasks) break;
        const u32 c0 = tasks[t], c1 = (t + 1 < ntasks) ? tasks[t + 1] : nchild;
        const u32 bs = off1[c0], bsize = off1[c1] - bs;
        if (bsize == 0) { if (threadIdx.x == 0) nu_out[t] = 0; continue; }
        // key bits that vary inside the task: the child's own bits plus the differing child-index bits
        const u32 x = c0 ^ (c1 - 1);
        const int shift0 = shiftc + (x ? 32 - __clz(x) : 0);
        if (bsize <= (u32)CAP) {
            msd_leaf<THREADS, CAP, HT>(S, X, bs, bsize, shift0, X, cnt_out, bs, min_freq);
            if (threadIdx.x == 0) nu_out[t] = S.s_nd;
            continue;
        }
        // a single child larger than CAP: partition it again, depth first
        if (shift0 == 0) {  // no varying bits: the whole child is one k-mer
            if (threadIdx.x == 0) {
                u32 keep = bsize >= min_freq;
                if (keep) cnt_out[bs] = bsize;  // X[bs] already holds the key
                nu_out[t] = keep;
            }
            continue;
        }
        u32 cur[MSD_MAXD];
        int depth = 0;
        if (threadIdx.x == 0) {
            MsdFrame &f = S.frames[0];
            f.start = bs; f.size = bsize; f.shift_before = shift0; f.bits = shift0 < 8 ? shift0 : 8; f.dst = 1;
        }
        __syncthreads();
        msd_partition<THREADS, CAP, HT>(S, S.frames[0], X, Y);
        cur[0] = 0;
        while (depth >= 0) {
            MsdFrame &f = S.frames[depth];
            if (cur[depth] == f.ntasks) { --depth; continue; }
            const u32 d0 = f.task[cur[depth]], d1 = f.task[cur[depth] + 1];
            ++cur[depth];
            const u32 s = f.off[d0], n = f.off[d1] - s;
            if (n == 0) continue;
            if (n <= (u32)CAP) {
                const u32 ob = bs + S.s_out;
                msd_leaf<THREADS, CAP, HT>(S, f.dst ? Y : X, f.start + s, n, f.shift_before, X, cnt_out, ob, min_freq);
                if (threadIdx.x == 0) S.s_out += S.s_nd;
                __syncthreads();
                continue;
            }
            const int child_shift = f.shift_before - f.bits;
            if (child_shift == 0 || depth + 1 >= MSD_MAXD) {
                // all n keys of this child are the same k-mer (8 levels of 8 bits cover the 64-bit key)
                if (threadIdx.x == 0 && n >= min_freq) {
                    u64 key = (f.dst ? Y : X)[f.start + s];
                    X[bs + S.s_out] = key;
                    cnt_out[bs + S.s_out] = n;
                    S.s_out += 1;
                }
                __syncthreads();
                continue;
            }
            if (threadIdx.x == 0) {
                MsdFrame &g = S.frames[depth + 1];
                g.start = f.start + s; g.size = n; g.shift_before = child_shift; g.bits = child_shift < 8 ? child_shift : 8; g.dst = 1 - f.dst;
            }
            __syncthreads();
            msd_partition<THREADS, CAP, HT>(S, S.frames[depth + 1], f.dst ? Y : X, f.dst ? X : Y);
            ++depth;
            cur[depth] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) nu_out[t] = S.s_out;
    }
}

__global__ void __launch_bounds__(256) msd_gather_kernel(const u64 *__restrict__ X, const u32 *__restrict__ cnt, const u32 *__restrict__ off1,
                                                         const u32 *__restrict__ tasks, const u32 *__restrict__ nu,
                                                         const u32 *__restrict__ ustart, u32 ntasks, u64 *__restrict__ okeys,
                                                         u32 *__restrict__ ocnt) {
    for (u32 t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const u32 n = nu[t], s = off1[tasks[t]], d = ustart[t];
        for (u32 i = threadIdx.x; i < n; i += 256) {
            okeys[d + i] = X[s + i];
            ocnt[d + i] = cnt[s + i];
        }
    }
}

struct MsdTune {
    int b0_max = 10;       // level-0 digit bits (<= MSD_MAX_B0)
    int b1_max = 9;        // level-1 digit bits (<= MSD_MAX_B1)
    int child_log2 = 9;    // target keys per child = 2^child_log2
    int variant = 0;       // task kernel geometry, see msd_launch_tasks
    int gw = 0, small = 0; // task merging window / small-child threshold (0 = from the variant)
};
static MsdTune msd_tune() {
    MsdTune t;
    if (const char *e = getenv("GG_MSD_B0")) t.b0_max = atoi(e);
    if (const char *e = getenv("GG_MSD_B1")) t.b1_max = atoi(e);
    if (const char *e = getenv("GG_MSD_CHILD")) t.child_log2 = atoi(e);
    if (const char *e = getenv("GG_MSD_VARIANT")) t.variant = atoi(e);
    if (const char *e = getenv("GG_MSD_GW")) t.gw = atoi(e);
    if (const char *e = getenv("GG_MSD_SMALL")) t.small = atoi(e);
    if (t.b0_max > MSD_MAX_B0) t.b0_max = MSD_MAX_B0;
    if (t.b0_max < 0) t.b0_max = 0;
    if (t.b1_max > MSD_MAX_B1) t.b1_max = MSD_MAX_B1;
    if (t.b1_max < 0) t.b1_max = 0;
    return t;
}

template <int THREADS, int CAP, int HT>
static void msd_launch_tasks(gg_ctx *ctx, int ctas_per_sm, u64 *X, u64 *Y, u32 *cnt, const u32 *off1, const u32 *tasks, u32 ntasks,
                             u32 nchild, int shiftc, u32 *nu, u32 *ticket, u32 min_freq) {
    size_t smem = sizeof(MsdSmem<THREADS, CAP, HT>);
    static bool set = false;
    if (!set) {
        CK(cudaFuncSetAttribute(msd_task_kernel<THREADS, CAP, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = true;
    }
    u32 grid = (u32)ctx->num_sms * ctas_per_sm;
    if (grid > ntasks) grid = ntasks;
    LAUNCH(ctx, (msd_task_kernel<THREADS, CAP, HT>), grid, THREADS, smem, X, Y, cnt, off1, tasks, ntasks, nchild, shiftc, nu, ticket, min_freq);
}

// packed reads -> sorted unique k-mers + counts (count >= min_freq only).  k <= 32.
// Returns U; *n_instances = number of k-mer windows.
static u64 msd_count(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq,
                     DBuf<u64> &ukeys, DBuf<u32> &ucounts, u64 *n_instances) {
    const MsdTune tune = msd_tune();
    const u64 nwords = pr.nwords;
    *n_instances = 0;
    if (nwords == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    // digit widths from the number of windows (upper bound of N): children of ~2^child_log2 keys
    int bt = 0;
    while (bt < 21 && (pr.nbases >> (bt + tune.child_log2)) > 0) ++bt;
    int b0 = (bt + 1) / 2;
    if (b0 > tune.b0_max) b0 = tune.b0_max;
    if (b0 > 2 * k) b0 = 2 * k;
    int b1 = bt - b0;
    if (b1 > tune.b1_max) b1 = tune.b1_max;
    if (b1 > 2 * k - b0) b1 = 2 * k - b0;
    const u32 nb0 = 1u << b0;
    const u64 nchild = (u64)nb0 << b1;
    const int shiftc = 2 * k - b0 - b1;
    // ---- A: histogram ----
    stage_begin(ctx, ST_EXTRACT);
    DBuf<u32> hist(ctx, nb0 + 1), bstart(ctx, nb0 + 1), cursor(ctx, nb0), tcount(ctx, nb0 + 1), tstart(ctx, nb0 + 1), ticket(ctx, 1);
    DBuf<u64> tot(ctx, 1);
    CK(cudaMemsetAsync(hist.p, 0, (nb0 + 1) * sizeof(u32), ctx->stream));
    u64 hgrid = ceil_div(nwords, 256);
    if (hgrid > (u64)ctx->num_sms * 8) hgrid = (u64)ctx->num_sms * 8;
    LAUNCH(ctx, msd_hist_kernel, (unsigned)hgrid, 256, 0, pr.packed.p, valid, nwords, k, canonical, b0, hist.p);
    exclusive_scan_u32(ctx, hist.p, bstart.p, nb0 + 1, tot.p);
    const u64 N = read_scalar<u64>(ctx, tot.p);
    *n_instances = N;
    if (N == 0) { stage_end(ctx, ST_EXTRACT); ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    if (N >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu k-mer instances exceed the 2^32-1 per-call limit", (unsigned long long)N);
    // ---- B: extract + scatter ----
    DBuf<u64> A(ctx, N), B(ctx, N);
    DBuf<u32> cnt(ctx, N);
    LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nb0, 256), 256, 0, bstart.p, (u64)nb0, cursor.p);
    size_t smemB = (size_t)MSD_TILE_KEYS * 8 + 2 * (size_t)(1 << MSD_MAX_B0) * 4;
    static bool setB = false;
    if (!setB) { CK(cudaFuncSetAttribute(msd_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB)); setB = true; }
    LAUNCH(ctx, msd_scatter_kernel, (unsigned)ceil_div(nwords, 256), 256, smemB, pr.packed.p, valid, nwords, k, canonical, b0, cursor.p, A.p);
    stage_end(ctx, ST_EXTRACT);
    // ---- C, D: level-1 partition (children in X) ----
    stage_begin(ctx, ST_SORT);
    DBuf<u32> off1(ctx, nchild + 1), tasks(ctx, nchild + 1);
    u64 *X = A.p, *Y = B.p;
    if (b1 > 0) {
        DBuf<u32> hist1(ctx, nchild + 1), cursor1(ctx, nchild);
        CK(cudaMemsetAsync(hist1.p, 0, (nchild + 1) * sizeof(u32), ctx->stream));
        LAUNCH(ctx, msd_tilecount_kernel, (unsigned)ceil_div(nb0 + 1, 256), 256, 0, bstart.p, nb0, tcount.p);
        exclusive_scan_u32(ctx, tcount.p, tstart.p, nb0 + 1, nullptr);
        const unsigned tgrid = (unsigned)(ceil_div(N, MSD_T1) + nb0);
        LAUNCH(ctx, msd_hist1_kernel, tgrid, 256, 0, A.p, tstart.p, bstart.p, nb0, shiftc, b1, hist1.p);
        exclusive_scan_u32(ctx, hist1.p, off1.p, nchild + 1, nullptr);
        LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nchild, 256), 256, 0, off1.p, nchild, cursor1.p);
        LAUNCH(ctx, msd_scatter1_kernel, tgrid, 256, 0, A.p, tstart.p, bstart.p, nb0, shiftc, b1, cursor1.p, B.p);
        X = B.p;
        Y = A.p;
    } else {
        LAUNCH(ctx, msd_copy_u32_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, bstart.p, nchild + 1, off1.p);
    }
    // ---- E: tasks ----
    int variant = tune.variant;
    u32 cap = variant == 1 ? 6144u : variant == 2 ? 1536u : 3072u;
    MsdTaskF tf{off1.p, tasks.p, tune.gw ? (u32)tune.gw : cap * 2 / 3, tune.small ? (u32)tune.small : cap / 3};
    if (tf.gw + tf.small > cap) GG_THROW(GG_ERR_ARG, "GG_MSD_GW + GG_MSD_SMALL exceed the leaf capacity %u", cap);
    Compactor<MsdTaskF> tc(ctx, nchild, "msd_tasks");
    const u32 ntasks = (u32)tc.count(tf);
    tc.emit(tf);
    // ---- F: per-task counting ----
    DBuf<u32> nu(ctx, ntasks + 1), ustart(ctx, ntasks + 1);
    CK(cudaMemsetAsync(ticket.p, 0, sizeof(u32), ctx->stream));
    if (variant == 1) msd_launch_tasks<1024, 6144, 8192>(ctx, 1, X, Y, cnt.p, off1.p, tasks.p, ntasks, (u32)nchild, shiftc, nu.p, ticket.p, min_freq);
    else if (variant == 2) msd_launch_tasks<256, 1536, 2048>(ctx, 3, X, Y, cnt.p, off1.p, tasks.p, ntasks, (u32)nchild, shiftc, nu.p, ticket.p, min_freq);
    else if (variant == 3) msd_launch_tasks<256, 3072, 4096>(ctx, 2, X, Y, cnt.p, off1.p, tasks.p, ntasks, (u32)nchild, shiftc, nu.p, ticket.p, min_freq);
    else msd_launch_tasks<512, 3072, 4096>(ctx, 2, X, Y, cnt.p, off1.p, tasks.p, ntasks, (u32)nchild, shiftc, nu.p, ticket.p, min_freq);
    stage_end(ctx, ST_SORT);
    // ---- G: gather ----
    stage_begin(ctx, ST_COUNT);
    exclusive_scan_u32(ctx, nu.p, ustart.p, ntasks, tot.p);
    const u64 U = read_scalar<u64>(ctx, tot.p);
    ukeys.alloc(ctx, U ? U : 1);
    ucounts.alloc(ctx, U ? U : 1);
    if (U) {
        u32 ggrid = ntasks < (u32)ctx->num_sms * 16 ? ntasks : (u32)ctx->num_sms * 16;
        LAUNCH(ctx, msd_gather_kernel, ggrid, 256, 0, X, cnt.p, off1.p, tasks.p, nu.p, ustart.p, ntasks, ukeys.p, ucounts.p);
    }
    stage_end(ctx, ST_COUNT);
    return U;
}
