// skm.cuh -- K2+K3+K4 for 17 <= k <= 32: counting by MINIMIZER-partitioned super-k-mers.
//
// Replaces (reference) KmerAnalysis serial_mem: collect every canonical k-mer -> sort! ->
// run-length collapse (docs/src/man/assembly.md:126-128), followed by the min-count filter of
// dbg!(sg, counted_kmers, min_freq) (src/dbg.jl:275-276).  Output is identical (sorted unique
// k-mers + counts); the route never materialises the 8-byte k-mer instances in HBM:
//
//   S  skm_scan_kernel<COUNT>   (sample) records per level-0 bin -> speculative bin capacities
//   0  skm_scan_kernel<SCATTER> roll over the packed reads; every window gets the minimum, over its
//                               k-m+1 m-mers, of a strand-symmetric m-mer hash (its minimizer);
//                               consecutive windows with one minimizer form a SUPER-K-MER, stored as
//                               one 16-byte record (<= 48 bases + length + level-1 digit) in the
//                               level-0 bin its minimizer hashes to: ~1.9 B per k-mer instead of 8.
//                               A k-mer and its reverse complement have the same m-mer set, so all
//                               instances of a canonical k-mer meet in one bin.  Multi-GPU: the bins
//                               are owned by ranks and the records are stored straight into the
//                               owner's buffer over NVLink (dist.cuh).
//   1  skm_l1_kernel            records -> nb0 x nb1 leaf bins (exact sizes from pass 0's histogram)
//   E  skm_leaf_kernel          one CTA per leaf bin: records expanded to canonical k-mers in
//                               registers, counted in a shared-memory hash table; keys with
//                               count >= min_freq appended (unordered) to the kept list
//   F  skm_sort_*               kept (k-mer, count) pairs -> ascending k-mer order: one partition by
//                               the top key bits + one shared-memory ordering pass per bucket
//
// Algorithmic bytes per k-mer instance (R = records per instance ~ 2 / (WIN + 1), U' = kept keys):
// S+0: 0.375 B/base read + 16 R written; 1: 32 R; E: 16 R read + 12 U'/N written; F: 48 U'/N.
#pragma once
#include "msdcount.cuh"

#define SKM_THREADS 256
#define SKM_STAGE_CAP 8704    // runs a CTA stages in shared memory before it reserves space in the bins (>= 8192 + 256: one word per thread always fits)
#define SKM_CUR_STRIDE 64     // the level-0 cursors sit 256 bytes apart: neighbours would serialise in one L2 slice
#define SKM_MAX_NB0 512       // level-0 bins
#define SKM_MAX_NB1 8192      // leaf bins per level-0 bin (13 bits of the record's meta word)
#define SKM_ALIGN 8192        // level-0 regions are multiples of this many records (tiles of pass 1 never straddle a region)
#define SKM_MAXPROBE 64

static inline bool skm_supported(int k, int canonical) { return k >= 17 && k <= 32 && (canonical || k < 32); }
static inline int skm_logw(int k) { return k >= 25 ? 4 : 3; }   // WIN = 2^logw + 1 = k - m + 1 m-mer positions per window
// consecutive packed words (32 windows each) per thread: a CTA tile yields ~3 K records in both geometries
static inline int skm_twords(int k) { return skm_logw(k) == 4 ? 4 : 2; }
static inline u64 skm_tile_words(int k) { return (u64)SKM_THREADS * skm_twords(k); }

// bin digits of a minimizer hash: the minimum of WIN hashes is far from uniform, its finalised image is
__host__ __device__ __forceinline__ u32 skm_digits(u32 h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

struct SkmGeom {
    int k, canonical;
    u32 nb0, nb1;   // level-0 bins x leaf bins per level-0 bin (any counts: the digits are range-reduced by multiply-high)
    // rounds > 1 (inputs whose leaf bins would outnumber 512 x 8192): a call handles the super-k-mers whose minimizer
    // digit falls into slice `round` of `rounds` equal slices and skips the others; the rounds' tables are disjoint
    u32 rounds = 1, round = 0;
};
// bins of a minimizer: d uniform in [0, 2^32) -> bin0 = floor(d * nb0 / 2^32); the low word of that product is uniform
// again -> bin1 likewise; `rest` = what is left for further splits
__device__ __forceinline__ void skm_bins(u32 d, u32 nb0, u32 nb1, u32 &bin0, u32 &bin1, u32 &rest) {
    const u64 p0 = (u64)d * nb0;
    bin0 = (u32)(p0 >> 32);
    const u64 p1 = (u64)(u32)p0 * nb1;
    bin1 = (u32)(p1 >> 32);
    rest = (u32)p1;
}

// multi-GPU: rank q owns the level-0 bins [first[q], first[q + 1]); base0 / cap0 then describe this source's
// private region of every bin inside the owner's buffer
struct SkmPeers {
    ulonglong2 *ptr[MSD_MAX_PEERS];
    u32 first[MSD_MAX_PEERS + 1];
    int n;
    int bulk;   // runs go to the owner with bulk asynchronous copies (TMA) out of a shared-memory staging area
};
// shared memory -> global (possibly a peer's memory over NVLink) bulk copy: ONE transaction-sized request per run instead
// of one 16-byte store per record; bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_store_s2g(void *dst, const void *src_smem, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((u32)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // copies complete
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ int skm_owner(const SkmPeers &pt, u32 b) {
    int o = 0;
#pragma unroll
    for (int q = 1; q < MSD_MAX_PEERS; ++q) o += (q < pt.n && b >= pt.first[q]) ? 1 : 0;
    return o;
}

// One record: `len` consecutive windows starting at base `start` (len + k - 1 <= 48 bases, left-aligned in 128
// bits: .x = bases 0..31, top half of .y = bases 32..47); low 32 bits of .y = meta: len - 1 (5 bits) |
// leaf bin inside the level-0 bin << 5 (13 bits) | further digit bits << 18.
__device__ __forceinline__ ulonglong2 skm_make_record(const u64 *__restrict__ packed, u64 start, u32 len, int k, u32 meta) {
    const u32 nb = len + (u32)k - 1u;  // 17 .. 48
    const u64 j = start >> 5;
    const int r2 = 2 * (int)(start & 31);
    const u64 w0 = packed[j], w1 = packed[j + 1], w2 = packed[j + 2];
    u64 hi = r2 ? ((w0 << r2) | (w1 >> (64 - r2))) : w0;
    u64 lo = r2 ? ((w1 << r2) | (w2 >> (64 - r2))) : w1;
    if (nb <= 32) { hi &= ~0ULL << (64 - 2 * nb); lo = 0; }
    else lo &= ~0ULL << (64 - 2 * (nb - 32));
    return make_ulonglong2(hi, lo | (u64)meta);
}

struct SkmScanSmem {
    u32 dh[SKM_STAGE_CAP];              // staged run: minimizer hash
    u32 ds[SKM_STAGE_CAP];              // staged run: (first base - first base of the tile) << 5 | (windows - 1)
    u32 sq[32 * SKM_THREADS];           // per-thread queue: minimizer hash of every run that starts in the current word
    u32 hist[SKM_MAX_NB0];              // runs per level-0 bin, then the running cursor inside the reserved space
    u32 gbase[SKM_MAX_NB0];             // first record slot of this tile's runs in every bin (absolute index into the buffer)
    u32 glim[SKM_MAX_NB0];              // end of the bin's region
    u32 cnt[SKM_MAX_NB0];               // runs of this tile per bin (bulk route)
    u32 hstart[256];                    // bulk route: start of the bin's run inside the staging area (bins of the current half)
    u32 s_n;
};
#define SKM_BULK_CAP 2048   // records the staging area (the per-thread queues of phase 1, free by then) holds: one half of the bins at a time
// MODE 0: count records per level-0 bin (hist0); MODE 1: scatter the records.
// The launch covers the packed words [w0, w1); CTA c handles tile number c * stride of 256 * TWORDS words
// (stride > 1: sampling for the capacity estimate).
// Phase 1 (every lane busy, no calls, no divergent loops): roll the m-mer hashes of a word, sliding minimum, and note
// every run (start, length, minimizer hash) in the CTA's staging list.  Phases B-D run over the staged runs with all
// lanes: bin digits + per-bin count, ONE global reservation per (tile, bin) -- cursors 256 bytes apart, so that they
// sit in different L2 slices -- then the records are built from the packed words and stored next to each other.
// A tile with more runs than the list holds (only adversarial input: > 0.26 runs per window) is redone word by word.
template <int LOGW, int MODE, bool PEER>
__global__ void __launch_bounds__(SKM_THREADS, 2) skm_scan_kernel(const u64 *__restrict__ packed, const u32 *__restrict__ valid, u64 w0, u64 w1,
                                                                  u32 stride, SkmGeom g, u32 *__restrict__ hist0, u32 *__restrict__ cur0,
                                                                  const u32 *__restrict__ base0, const u32 *__restrict__ cap0,
                                                                  ulonglong2 *__restrict__ out, u32 *__restrict__ hist1,
                                                                  unsigned long long *__restrict__ totals /* [0] instances, [1] records */,
                                                                  u32 *__restrict__ ovf, SkmPeers pt) {
    constexpr int WIN = (1 << LOGW) + 1;
    constexpr int NH = 32 + WIN - 1;      // m-mer positions the 32 windows of a word look at
    constexpr int TWORDS = LOGW == 4 ? 4 : 2;
    constexpr u32 NOSLOT = 0xFFFFFFFFu;
    extern __shared__ __align__(16) u8 smem_raw[];
    SkmScanSmem &S = *reinterpret_cast<SkmScanSmem *>(smem_raw);
    const int k = g.k, m = k - WIN + 1;   // 9 <= m <= 16
    const u32 nb0 = g.nb0;
    const u32 mmask = m == 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    const u32 rmul = 1u << (2 * m - 2);
    const u64 tile_w0 = w0 + ((u64)blockIdx.x * stride) * (SKM_THREADS * TWORDS);
    const u64 tbase = tile_w0 + (u64)threadIdx.x * TWORDS;
    u32 tot_inst = 0, tot_rec = 0;
    int it0 = 0, it1 = TWORDS;
    bool fallback = false;
    for (;;) {
        for (u32 i = threadIdx.x; i < nb0; i += SKM_THREADS) S.hist[i] = 0;
        if (threadIdx.x == 0) S.s_n = 0;
        __syncthreads();
        // ---------------- phase 1 ----------------
        u32 rl = 0, rh = 0, rslot = NOSLOT;   // open run: windows so far, minimizer hash, staging slot
#pragma unroll 1
        for (int it = it0; it < it1; ++it) {
            const u64 t = tbase + it;
            const u32 vm = t < w1 ? valid[t] : 0u;
            __syncwarp();
            if (vm != 0) {
                const u64 a0 = packed[t], a1 = packed[t + 1];
                // hash of the canonical m-mer at every position 0 .. NH-1 (forward and reverse complement rolled)
                u32 h[NH];
                {
                    u32 f = (u32)(a0 >> (64 - 2 * m));
                    u32 r = (u32)rev2_u64(~a0) & mmask;
                    const u64 xh = (a0 << (2 * m)) | (a1 >> (64 - 2 * m)), xl = a1 << (2 * m);  // bases m, m+1, ... left-aligned
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        h[q] = min(f, r) * 0x9E3779B1u;
                        if (q + 1 < NH) {
                            const u32 nbq = q < 32 ? ((u32)(xh >> (62 - 2 * q)) & 3u) : ((u32)(xl >> (62 - 2 * (q - 32))) & 3u);
                            f = ((f << 2) | nbq) & mmask;
                            r = (r >> 2) + (3u - nbq) * rmul;
                        }
                    }
                }
                // sliding minimum over WIN positions: doubling to 2^LOGW, then one more
#pragma unroll
                for (int s = 0; s < LOGW; ++s) {
                    const int d = 1 << s;
#pragma unroll
                    for (int q = 0; q < NH; ++q)
                        if (q + 2 * d - 1 < NH) h[q] = min(h[q], h[q + d]);
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) h[q] = min(h[q], h[q + 1]);
                // runs: consecutive valid windows with one minimizer hash, at most WIN long.  Sweep over the windows
                // (straight-line code): mask of the run starts, their hashes queued in order.
                u32 nm = 0, nq = 0;
                {
                    u32 l = rl, hh = rh;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const bool v = (vm >> (31 - j)) & 1u;
                        const bool cont = v && l != 0 && h[j] == hh && l < (u32)WIN;
                        if (v && !cont) {
                            S.sq[nq * SKM_THREADS + threadIdx.x] = h[j];
                            ++nq;
                            nm |= 1u << j;
                            hh = h[j];
                            l = 0;
                        }
                        l = v ? l + 1 : 0;
                    }
                }
                const u32 cm = __brev(vm) & ~nm;   // bit j <=> window j continues the run of window j - 1
                if (rl != 0) {                     // the carried run takes the leading continuation windows and ends in this word (WIN < 32)
                    rl += (u32)__ffs((int)~cm) - 1u;
                    if (rslot < SKM_STAGE_CAP) S.ds[rslot] += rl - 1u;
                    rl = 0;
                }
                const u32 sbase = nq ? atomicAdd(&S.s_n, nq) : 0u;   // one reservation per thread and word
                const u32 rel = (u32)(t - tile_w0) << 5;
                u32 mask = nm, qi = 0;
                while (mask) {
                    const int j = __ffs((int)mask) - 1;
                    mask &= mask - 1;
                    const u32 rest = j == 31 ? 0u : (cm >> (j + 1));
                    const u32 len = (u32)__ffs((int)~rest);  // 1 + continuation windows (zeros are shifted in at the top of rest)
                    const u32 slot = sbase + qi;
                    const u32 hq = S.sq[qi * SKM_THREADS + threadIdx.x];
                    ++qi;
                    if ((u32)j + len == 32u) {  // still open at the end of the word: the length is added when it ends
                        if (slot < SKM_STAGE_CAP) { S.dh[slot] = hq; S.ds[slot] = (rel + (u32)j) << 5; }
                        rl = len; rh = hq; rslot = slot;
                    } else {
                        if (slot < SKM_STAGE_CAP) { S.dh[slot] = hq; S.ds[slot] = ((rel + (u32)j) << 5) | (len - 1u); }
                    }
                }
            } else if (rl != 0) {  // no k-mer starts in this word: the open run ends here
                if (rslot < SKM_STAGE_CAP) S.ds[rslot] += rl - 1u;
                rl = 0;
            }
        }
        __syncwarp();
        if (rl != 0) {
            if (rslot < SKM_STAGE_CAP) S.ds[rslot] += rl - 1u;
        }
        __syncthreads();
        const u32 ntot = S.s_n;
        if (ntot > SKM_STAGE_CAP) {  // cannot happen with one word per thread (<= 8192 + 256 runs)
            __syncthreads();
            fallback = true;
            it0 = 0; it1 = 1;
            continue;
        }
        // ---------------- phase B: bin of every staged run (the hash is replaced by its bin digits; runs of other rounds
        // are marked and skipped from here on) ----------------
        for (u32 i = threadIdx.x; i < ntot; i += SKM_THREADS) {
            u32 d = skm_digits(S.dh[i]);
            if (g.rounds > 1) {
                const u64 pr = (u64)d * g.rounds;
                if ((u32)(pr >> 32) != g.round) { S.ds[i] = 0xFFFFFFFFu; continue; }
                d = (u32)pr;   // the low word is uniform again
            }
            S.dh[i] = d;
            tot_inst += (S.ds[i] & 31u) + 1u;
            ++tot_rec;
            atomicAdd(&S.hist[__umulhi(d, nb0)], 1u);
        }
        __syncthreads();
        if (MODE == 0) {
            for (u32 b = threadIdx.x; b < nb0; b += SKM_THREADS)
                if (S.hist[b]) atomicAdd(&hist0[b], S.hist[b]);
        } else {
            // ---------------- phase C: one reservation per (tile, bin) ----------------
            for (u32 b = threadIdx.x; b < nb0; b += SKM_THREADS) {
                const u32 c = S.hist[b];
                const u32 cap = cap0[b], base = base0[b];
                const u32 gpos = c ? atomicAdd(&cur0[(u64)b * SKM_CUR_STRIDE], c) : 0u;
                if (c && (u64)gpos + c > (u64)cap) *ovf = 1u;
                S.gbase[b] = base + min(gpos, cap);
                S.glim[b] = base + cap;
                S.cnt[b] = (c && (u64)gpos + c <= (u64)cap) ? c : 0u;   // (an overflowing run is dropped: the scatter is repeated anyway)
                S.hist[b] = 0;
            }
            __syncthreads();
            if (PEER && pt.bulk) {
                // ---------------- phase D, bulk route: the records of 256 bins at a time are staged bin by bin, then every bin's
                // run leaves with one bulk copy into the owner's buffer ----------------
                __shared__ u32 ws[33];
                ulonglong2 *srec = reinterpret_cast<ulonglong2 *>(S.sq);
                static_assert(sizeof(S.sq) >= SKM_BULK_CAP * sizeof(ulonglong2), "staging area");
                const u64 tile_b0 = tile_w0 << 5;
                for (u32 h0 = 0; h0 < nb0; h0 += 256) {
                    const u32 bme = h0 + threadIdx.x;
                    const u32 c = bme < nb0 ? S.cnt[bme] : 0u;
                    u32 tot;
                    const u32 ex = block_excl_scan(c, ws, &tot);
                    S.hstart[threadIdx.x] = ex;
                    __syncthreads();
                    const bool staged = tot <= SKM_BULK_CAP;   // (uniform) otherwise this half goes out record by record
                    for (u32 i = threadIdx.x; i < ntot; i += SKM_THREADS) {
                        const u32 dsv = S.ds[i];
                        if (dsv == 0xFFFFFFFFu) continue;   // another round's run
                        u32 b, bin1, rest;
                        skm_bins(S.dh[i], nb0, g.nb1, b, bin1, rest);
                        if (b < h0 || b >= h0 + 256) continue;
                        const u32 r = atomicAdd(&S.hist[b], 1u);
                        const u32 len = (dsv & 31u) + 1u;
                        const u32 meta = (len - 1u) | (bin1 << 5) | ((rest >> 18) << 18);
                        atomicAdd(&hist1[(u64)b * g.nb1 + bin1], 1u);
                        const u32 p = S.gbase[b] + r;
                        if (p < S.glim[b]) {
                            const ulonglong2 rec = skm_make_record(packed, tile_b0 + (dsv >> 5), len, k, meta);
                            if (staged && S.cnt[b]) srec[S.hstart[b - h0] + r] = rec;
                            else pt.ptr[skm_owner(pt, b)][p] = rec;
                        }
                    }
                    fence_async_smem();
                    __syncthreads();
                    if (staged && c) {
                        bulk_store_s2g(pt.ptr[skm_owner(pt, bme)] + S.gbase[bme], srec + ex, c * (u32)sizeof(ulonglong2));
                        bulk_commit();
                        bulk_wait_read();
                    }
                    __syncthreads();
                }
                bulk_wait_all();
            } else {
            // ---------------- phase D: records built and stored, grouped by bin ----------------
            const u64 tile_b0 = tile_w0 << 5;
            for (u32 i = threadIdx.x; i < ntot; i += SKM_THREADS) {
                const u32 dsv = S.ds[i];
                if (dsv == 0xFFFFFFFFu) continue;   // another round's run
                u32 b, bin1, rest;
                skm_bins(S.dh[i], nb0, g.nb1, b, bin1, rest);
                const u32 p = S.gbase[b] + atomicAdd(&S.hist[b], 1u);
                const u32 len = (dsv & 31u) + 1u;
                const u32 meta = (len - 1u) | (bin1 << 5) | ((rest >> 18) << 18);
                atomicAdd(&hist1[(u64)b * g.nb1 + bin1], 1u);
                if (p < S.glim[b]) {
                    const ulonglong2 rec = skm_make_record(packed, tile_b0 + (dsv >> 5), len, k, meta);
                    if (PEER) pt.ptr[skm_owner(pt, b)][p] = rec;
                    else out[p] = rec;
                }
            }
            }
        }
        if (!fallback || it1 == TWORDS) break;
        __syncthreads();
        it0 = it1; it1 = it0 + 1;
    }
    // totals: one atomic per warp
    for (int d = 16; d > 0; d >>= 1) { tot_inst += __shfl_xor_sync(FULL_MASK, tot_inst, d); tot_rec += __shfl_xor_sync(FULL_MASK, tot_rec, d); }
    if (lane_id() == 0 && tot_rec) { atomicAdd(&totals[0], (unsigned long long)tot_inst); atomicAdd(&totals[1], (unsigned long long)tot_rec); }
}

// Capacities of the level-0 regions from a (sampled) histogram: region r gets
// round_up(scale * c * 1.25 + 4 * scale * sqrt(c + 1) + 1024, SKM_ALIGN) records (exact counts: round_up(c, SKM_ALIGN)).
// One CTA.  base has nreg + 1 entries; *total_cap gets the sum.
__global__ void __launch_bounds__(1024) skm_caps_kernel(const u32 *__restrict__ hist, u32 nreg, float scale, int exact, u32 *__restrict__ base,
                                                        u32 *__restrict__ cap, u64 *__restrict__ total_cap, u32 force_cap) {
    __shared__ u32 ws[33];
    __shared__ u64 s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 r0 = 0; r0 < nreg; r0 += 1024) {
        const u32 r = r0 + threadIdx.x;
        u32 c = 0;
        if (r < nreg) {
            const u64 cnt = hist[r];
            const float hv = (float)cnt;
            double want = exact ? (double)cnt : (double)(scale * hv * 1.25f + 4.0f * scale * sqrtf(hv + 1.0f) + 1024.0f);
            if (force_cap) want = force_cap;  // tests: provoke the overflow path
            u64 w = (u64)want;
            w = ((w + SKM_ALIGN - 1) / SKM_ALIGN) * SKM_ALIGN;
            if (w == 0) w = SKM_ALIGN;
            c = (u32)(w / SKM_ALIGN);  // in units of SKM_ALIGN: the block scan stays within 32 bits
        }
        u32 tot;
        const u32 ex = block_excl_scan(c, ws, &tot);
        if (r < nreg) {
            const u64 bb = (s_run + ex) * SKM_ALIGN;
            base[r] = bb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)bb;
            cap[r] = c * SKM_ALIGN;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const u64 bb = s_run * SKM_ALIGN;
        base[nreg] = bb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)bb;
        *total_cap = bb;
    }
}

// ---- pass 1: level-0 regions -> leaf bins ---------------------------------------------------------
// Region r (a level-0 bin, or a (source, bin) pair on the multi-GPU path: source-major, so that one source's regions are
// one contiguous segment) holds cnt[r * cnt_stride] records at base[r]; its records belong to the local level-0 bin
// r % nb_local.  tile_reg[t] = region of the t-th tile of SKM_ALIGN slots.
__global__ void skm_tilemap_kernel(const u32 *__restrict__ base, u32 nreg, u32 *__restrict__ tile_reg) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreg) return;
    for (u32 t = base[r] / SKM_ALIGN; t < base[r + 1] / SKM_ALIGN; ++t) tile_reg[t] = r;
}
// One CTA per tile of SKM_ALIGN record slots.  Two sweeps over the tile's records (the second one hits L2): count per
// leaf bin in shared memory, ONE global reservation per (tile, leaf bin), then the records are stored, those of one bin next
// to each other.  (Per-record global atomics serialise on the few cursors one level-0 bin's tiles share.)
#define SKM_L1_THREADS 512
__global__ void __launch_bounds__(SKM_L1_THREADS) skm_l1_kernel(const ulonglong2 *__restrict__ in, const u32 *__restrict__ base,
                                                                const u32 *__restrict__ cnt, u32 cnt_stride, const u32 *__restrict__ tile_reg,
                                                                u32 nb_local, u32 nb1, const u32 *__restrict__ lbeg, u32 *__restrict__ cur1,
                                                                ulonglong2 *__restrict__ out, unsigned long long *__restrict__ inst_total) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u32 *hist = reinterpret_cast<u32 *>(smem_raw);   // nb1
    u32 *gpos = hist + nb1;                          // nb1
    const u64 tile0 = (u64)blockIdx.x * SKM_ALIGN;
    const u32 r = tile_reg[blockIdx.x];
    const u32 n = cnt[(u64)r * cnt_stride];
    const u32 rel0 = (u32)(tile0 - base[r]);
    if (rel0 >= n) return;
    const u32 m = min(n - rel0, (u32)SKM_ALIGN);     // records of this tile
    const u64 leaf0 = (u64)(r % nb_local) * nb1;
    const ulonglong2 *tp = in + tile0;
    for (u32 i = threadIdx.x; i < nb1; i += SKM_L1_THREADS) hist[i] = 0;
    __syncthreads();
    // the tile's records stay in registers between the count and the placement (16 independent 128-bit loads per thread):
    // a second sweep over the 128 KB tile missed L2 and doubled the kernel's DRAM reads (ncu, profiles/README.md)
    constexpr int PER = SKM_ALIGN / SKM_L1_THREADS;
    ulonglong2 rec[PER];
    u32 inst = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const u32 i = threadIdx.x + (u32)q * SKM_L1_THREADS;
        rec[q] = i < m ? tp[i] : make_ulonglong2(0, 0);
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const u32 i = threadIdx.x + (u32)q * SKM_L1_THREADS;
        if (i < m) {
            const u32 meta = (u32)rec[q].y;
            atomicAdd(&hist[(meta >> 5) & (SKM_MAX_NB1 - 1u)], 1u);
            inst += (meta & 31u) + 1u;
        }
    }
    if (inst_total) {   // multi-GPU: the owner learns how many k-mer instances it received
        inst = warp_sum(inst);
        if (lane_id() == 0 && inst) atomicAdd(inst_total, (unsigned long long)inst);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < nb1; i += SKM_L1_THREADS) {
        const u32 c = hist[i];
        gpos[i] = c ? lbeg[leaf0 + i] + atomicAdd(&cur1[leaf0 + i], c) : 0u;
        hist[i] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const u32 i = threadIdx.x + (u32)q * SKM_L1_THREADS;
        if (i < m) {
            const u32 b = ((u32)rec[q].y >> 5) & (SKM_MAX_NB1 - 1u);
            out[(u64)gpos[b] + atomicAdd(&hist[b], 1u)] = rec[q];
        }
    }
}
__global__ void skm_gather_strided_kernel(const u32 *__restrict__ in, u32 stride, u32 n, u32 *__restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[(u64)i * stride];
}

// ---- pass E: one CTA per leaf bin -------------------------------------------------------------------
template <int THREADS, int LOG_HT>
struct SkmLeafSmem {
    static constexpr int HT = 1 << LOG_HT;
    static constexpr int LIST = HT / 4 * 3;
    u64 tab[HT];
    u32 tcnt[HT];      // occurrences beyond the first one
    u16 list[LIST];    // claimed slots of the current bin: the sweep and the clean-up touch only these
    u32 s_ovf[2], s_nd[2], s_kpos, s_newblk;   // s_ovf / s_nd alternate between consecutive bins (set / reset without an extra barrier:
                                               // a slow warp may still read bin i's values when thread 0 prepares bin i + 1)
    unsigned long long s_kbase;       // the CTA's current block of the kept list
};
#define SKM_KBLOCK 32768   // kept-list slots a CTA reserves at a time (unused slots are marked empty and skipped by the sort)
__device__ __forceinline__ u32 skm_slot(u64 key, int lh) { return (((u32)key ^ (u32)(key >> 32)) * 0x9E3779B1u) >> (32 - lh); }

// Leaf bin l = records recs[lbeg[l] .. + lcnt[l]); CTA c takes the bins c, c + grid, ... (the bins are near-uniform).
// (Measured alternatives on C3, 19.0 ms as is: no list and a 128-bit sweep of the whole table after every bin 21.8 ms --
// the sweep sits on the critical path between two barriers, the list upkeep overlaps with the insertions; geometries
// 256 x 4096 slots x 4 CTAs 21.4, 256 x 2048 x 8 19.3, 512 x 8192 x 2 22.8, 256 x 8192 x 2 31.5 ms: what counts is the number
// of resident warps, the kernel is bound by instruction issue, ~180 warp instructions per 32 insertions.)
// The table is cleared once per CTA; every bin leaves it clean again by resetting the slots it claimed, so the cost of
// a bin is proportional to its instances + distinct keys, not to the table size.  Kept keys (count >= min_freq) are
// appended to okeys / ocnts at *kept_cursor.  A bin whose table fills up is put on the overflow list instead.
template <int THREADS, int LOG_HT, int LOGW>
__global__ void __launch_bounds__(THREADS) skm_leaf_kernel(const ulonglong2 *__restrict__ recs, const u32 *__restrict__ lbeg,
                                                           const u32 *__restrict__ lcnt, u32 nleaf, int k, int canonical, u32 min_freq,
                                                           u64 *__restrict__ okeys, u32 *__restrict__ ocnts, unsigned long long *kept_cursor,
                                                           u32 *__restrict__ ovf_list, u32 *ovf_n, u32 test_ovf_mod) {
    constexpr u32 HT = 1u << LOG_HT, HM = HT - 1u, LIST = SkmLeafSmem<THREADS, LOG_HT>::LIST;
    extern __shared__ __align__(16) u8 smem_raw[];
    SkmLeafSmem<THREADS, LOG_HT> &S = *reinterpret_cast<SkmLeafSmem<THREADS, LOG_HT> *>(smem_raw);
    const int s = 64 - 2 * k;
    const u64 topmask = ~0ULL << s, kmask = ~0ULL >> s;
    {
        ulonglong2 *t2 = reinterpret_cast<ulonglong2 *>(S.tab);
        for (u32 i = threadIdx.x; i < HT / 2; i += THREADS) t2[i] = make_ulonglong2(MSD_EMPTY, MSD_EMPTY);
        uint4 *c4 = reinterpret_cast<uint4 *>(S.tcnt);
        for (u32 i = threadIdx.x; i < HT / 4; i += THREADS) c4[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { S.s_nd[0] = 0; S.s_nd[1] = 0; S.s_kpos = SKM_KBLOCK; S.s_kbase = 0; }  // no block yet
    }
    u32 par = 0;
    unsigned long long ctot = 0;   // (thread 0) kept pairs of this CTA's finished blocks
    for (u32 leaf = blockIdx.x; leaf < nleaf; leaf += gridDim.x) {
        const u32 n = lcnt[leaf];
        if (n == 0) continue;
        const ulonglong2 *rp = recs + lbeg[leaf];
        if (threadIdx.x == 0) S.s_ovf[par] = (test_ovf_mod && leaf % test_ovf_mod == 0) ? 1u : 0u;
        __syncthreads();  // table clean, flags set, the previous bin's sweep is over
        if (threadIdx.x == 0) S.s_newblk = (S.s_kpos + LIST > SKM_KBLOCK) ? 1u : 0u;  // this bin may keep up to LIST keys (read after the next barrier)
        volatile u32 *ovf = &S.s_ovf[par];
        u32 *ndp = &S.s_nd[par];
        // (the flag is polled with warp votes: lanes of one warp must agree on leaving, they meet again in the shuffles)
        if (!__any_sync(FULL_MASK, *ovf != 0)) {
            // A warp takes 32 records at a time and spreads their windows over its lanes: lane i handles instance p + i of
            // the chunk (record and window found with one OR-reduction + ballot; the record travels by shuffle), so every
            // lane inserts, whatever the record lengths are.
            // (a bin holds a few hundred records: the warps share them evenly, `per` <= 32 records each at a time, so that
            // none of them sits out the bin)
            const u32 per = min(32u, (n + (THREADS / 32) - 1u) / (THREADS / 32));
            for (u32 c0 = (threadIdx.x >> 5) * per; c0 < n; c0 += (THREADS / 32) * per) {
                const u32 ri = c0 + lane_id();
                const bool mine = lane_id() < per && ri < n;
                ulonglong2 rec = make_ulonglong2(0, 0);
                if (mine) rec = rp[ri];
                const u32 len = mine ? ((u32)rec.y & 31u) + 1u : 0u;
                const u32 incl = warp_incl_scan(len), start = incl - len;
                const u32 T = __shfl_sync(FULL_MASK, incl, 31);
                const u64 a0 = rec.x;
                const u32 a1h = (u32)(rec.y >> 32);
                for (u32 p = 0; p < T; p += 32) {
                    const u32 rel = start - p;  // wraps to a huge value for records that start before p
                    const u32 smask = __reduce_or_sync(FULL_MASK, (len != 0 && rel < 32u) ? (1u << rel) : 0u);
                    const u32 nbefore = __popc(__ballot_sync(FULL_MASK, len != 0 && start <= p));
                    const u32 le = lane_id() == 31 ? 0xFFFFFFFFu : ((2u << lane_id()) - 1u);
                    const u32 r = (nbefore - 1u + __popc(smask & 0xFFFFFFFEu & le)) & 31u;
                    const u32 W0 = __shfl_sync(FULL_MASK, (u32)(a0 >> 32), r);   // bases 0..15 of the record
                    const u32 W1 = __shfl_sync(FULL_MASK, (u32)a0, r);           // 16..31
                    const u32 W2 = __shfl_sync(FULL_MASK, a1h, r);               // 32..47
                    const u32 st = __shfl_sync(FULL_MASK, start, r);
                    const u32 q = p + lane_id();
                    u32 claimed = 0xFFFFFFFFu;   // slot this lane claimed (first occurrence of its key in the bin)
                    if (q < T) {
                        const u32 j2 = 2u * (q - st);  // window q - st of record r: 64 bits from bit offset j2 <= 32 (two funnel shifts)
                        const u64 fl = (((u64)__funnelshift_lc(W1, W0, j2) << 32) | __funnelshift_lc(W2, W1, j2)) & topmask;
                        const u64 fw = fl >> s;
                        const u64 rc = rev2_u64(~fl) & kmask;
                        const u64 key = (canonical && rc < fw) ? rc : fw;
                        u32 slot = skm_slot(key, LOG_HT);
                        int probes = 0;
                        for (;;) {
                            u64 c = S.tab[slot];
                            if (c == MSD_EMPTY) {
                                c = atomicCAS((unsigned long long *)&S.tab[slot], (unsigned long long)MSD_EMPTY, (unsigned long long)key);
                                if (c == MSD_EMPTY) { claimed = slot; break; }
                            }
                            if (c == key) { atomicAdd(&S.tcnt[slot], 1u); break; }
                            slot = (slot + 1) & HM;
                            if (++probes > SKM_MAXPROBE) { *ovf = 1; break; }
                        }
                    }
                    // the claimed slots go on the bin's list: one reservation per warp and round
                    const u32 cm = __ballot_sync(FULL_MASK, claimed != 0xFFFFFFFFu);
                    if (cm) {
                        u32 lb = 0;
                        if (lane_id() == 0) lb = atomicAdd(ndp, (u32)__popc(cm));
                        lb = __shfl_sync(FULL_MASK, lb, 0) + __popc(cm & lanemask_lt());
                        if (claimed != 0xFFFFFFFFu) { if (lb < LIST) S.list[lb] = (u16)claimed; else *ovf = 1; }
                    }
                }
                if (__any_sync(FULL_MASK, *ovf != 0)) break;
            }
        }
        __syncthreads();
        const u32 nd_all = *ndp, nd = min(nd_all, LIST);
        const bool over = *ovf != 0;
        if (S.s_newblk) {  // (uniform) the bin's keys may not fit the rest of the CTA's block: next block
            const u32 kp = S.s_kpos;
            const unsigned long long kb = S.s_kbase;
            __syncthreads();
            for (u32 i = kp + threadIdx.x; i < SKM_KBLOCK; i += THREADS) okeys[kb + i] = MSD_EMPTY;   // unused tail: skipped later
            if (threadIdx.x == 0) {
                if (kp < SKM_KBLOCK) ctot += kp;
                S.s_kbase = atomicAdd(kept_cursor, (unsigned long long)SKM_KBLOCK);
                S.s_kpos = 0;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            S.s_nd[par ^ 1u] = 0;
            if (over) ovf_list[atomicAdd(ovf_n, 1u)] = leaf;
        }
        // ---- one sweep over the claimed slots: kept keys -> the CTA's block of the kept list; the table is left clean ----
        if (nd_all > LIST) {  // slots were claimed that the list does not hold (the bin is on the overflow list)
            ulonglong2 *t2 = reinterpret_cast<ulonglong2 *>(S.tab);
            for (u32 i = threadIdx.x; i < HT / 2; i += THREADS) t2[i] = make_ulonglong2(MSD_EMPTY, MSD_EMPTY);
            uint4 *c4 = reinterpret_cast<uint4 *>(S.tcnt);
            for (u32 i = threadIdx.x; i < HT / 4; i += THREADS) c4[i] = make_uint4(0, 0, 0, 0);
        } else {
            const unsigned long long kb = S.s_kbase;
            for (u32 i = threadIdx.x; i < nd; i += THREADS) {
                const u32 slot = S.list[i];
                const u32 cnt = S.tcnt[slot] + 1u;
                if (!over && cnt >= min_freq) {
                    const u32 o = atomicAdd(&S.s_kpos, 1u);
                    okeys[kb + o] = S.tab[slot];
                    ocnts[kb + o] = cnt;
                }
                S.tab[slot] = MSD_EMPTY;
                S.tcnt[slot] = 0;
            }
        }
        par ^= 1u;
    }
    // the rest of the CTA's last block
    __syncthreads();
    if (S.s_kpos < SKM_KBLOCK) {
        const unsigned long long kb = S.s_kbase;
        for (u32 i = S.s_kpos + threadIdx.x; i < SKM_KBLOCK; i += THREADS) okeys[kb + i] = MSD_EMPTY;
        if (threadIdx.x == 0) ctot += S.s_kpos;
    }
    if (threadIdx.x == 0 && ctot) atomicAdd(kept_cursor + 1, ctot);   // [1]: kept pairs, [0]: slots handed out
}

// ---- overflowed leaf bins: all their k-mers as plain keys (finished with the generic sort + run-length path) ----
__global__ void skm_ovf_size_kernel(const ulonglong2 *__restrict__ recs, const u32 *__restrict__ lbeg, const u32 *__restrict__ lcnt,
                                    const u32 *__restrict__ ovf_list, u32 n_ovf, unsigned long long *__restrict__ total) {
    for (u32 q = blockIdx.x; q < n_ovf; q += gridDim.x) {
        const u32 leaf = ovf_list[q];
        const ulonglong2 *rp = recs + lbeg[leaf];
        u32 mine = 0;
        for (u32 i = threadIdx.x; i < lcnt[leaf]; i += blockDim.x) mine += ((u32)rp[i].y & 31u) + 1u;
        mine = warp_sum(mine);
        if (lane_id() == 0 && mine) atomicAdd(total, (unsigned long long)mine);
    }
}
__global__ void skm_ovf_expand_kernel(const ulonglong2 *__restrict__ recs, const u32 *__restrict__ lbeg, const u32 *__restrict__ lcnt,
                                      const u32 *__restrict__ ovf_list, u32 n_ovf, int k, int canonical, u64 *__restrict__ out,
                                      unsigned long long *__restrict__ cursor) {
    const int s = 64 - 2 * k;
    const u64 topmask = ~0ULL << s;
    for (u32 q = blockIdx.x; q < n_ovf; q += gridDim.x) {
        const u32 leaf = ovf_list[q];
        const ulonglong2 *rp = recs + lbeg[leaf];
        for (u32 i = threadIdx.x; i < lcnt[leaf]; i += blockDim.x) {
            const ulonglong2 rec = rp[i];
            const u32 len = ((u32)rec.y & 31u) + 1u;
            const u64 a0 = rec.x, a1 = rec.y & 0xFFFFFFFF00000000ULL;
            u64 o = atomicAdd(cursor, (unsigned long long)len);
            for (u32 j = 0; j < len; ++j) {
                const u64 fl = (j == 0 ? a0 : ((a0 << (2 * j)) | (a1 >> (64 - 2 * j)))) & topmask;
                const u64 fw = fl >> s;
                Kmer<1> x; x.w[0] = fw;
                const u64 rc = km_rc<1>(x, k).w[0];
                out[o + j] = (canonical && rc < fw) ? rc : fw;
            }
        }
    }
}

// ---- pass F: kept pairs -> ascending key order ------------------------------------------------------
#define SKM_SORT_CAP 8192   // pairs one CTA orders in shared memory
// ---- staged partition of (key, count) pairs by a key digit -------------------------------------------------------
// The slots are cut into tiles of PS_TILE; tile t belongs to region tile_reg[t] (a bucket of the previous pass; null: one
// region) that holds rcnt[r] pairs from rbase[r] on (a multiple of PS_TILE).  MODE 0 adds the tile's digit histogram to
// hist[r * nb + digit].  MODE 1 stages the tile's pairs in shared memory grouped by digit (rank inside the group from the
// shared-memory histogram atomics), reserves ONE range per (tile, digit) in the output bucket and writes the groups as
// runs -- instead of one global atomic and one 12-byte scattered store per pair.  Empty slots (MSD_EMPTY) are skipped.
#define PS_TILE 8192
#define PS_THREADS 512
#define PS_MAX_NB 512
template <int MODE>
__global__ void __launch_bounds__(PS_THREADS) pair_part_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ cnts, const u32 *__restrict__ tile_reg,
                                                               const u32 *__restrict__ rbase, const u32 *__restrict__ rcnt, u64 n_all, u64 kbase, int sh, u32 nb,
                                                               u32 *__restrict__ hist, const u32 *__restrict__ ostart, u64 *__restrict__ okeys,
                                                               u32 *__restrict__ ocnts) {
    constexpr int PER = PS_TILE / PS_THREADS;
    extern __shared__ __align__(16) u8 smem_raw[];
    u32 *h = reinterpret_cast<u32 *>(smem_raw);           // PS_MAX_NB: pairs per digit, then the running start
    u32 *gp = h + PS_MAX_NB;                              // PS_MAX_NB: output position of the group minus its start in the tile
    u64 *sk = reinterpret_cast<u64 *>(gp + PS_MAX_NB);    // PS_TILE
    u32 *sc = reinterpret_cast<u32 *>(sk + PS_TILE);      // PS_TILE
    __shared__ u32 ws[33];
    const u64 tile0 = (u64)blockIdx.x * PS_TILE;
    u32 r = 0;
    u64 m = n_all - tile0 < PS_TILE ? n_all - tile0 : PS_TILE;
    if (tile_reg) {
        r = tile_reg[blockIdx.x];
        const u64 rel0 = tile0 - rbase[r];
        if (rel0 >= rcnt[r]) return;
        m = rcnt[r] - rel0 < PS_TILE ? rcnt[r] - rel0 : PS_TILE;
    }
    const u32 dm = nb - 1u;
    for (u32 i = threadIdx.x; i < nb; i += PS_THREADS) h[i] = 0;
    __syncthreads();
    u64 kr[PER];
    u32 pos[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const u32 j = threadIdx.x + (u32)i * PS_THREADS;
        kr[i] = j < m ? keys[tile0 + j] : MSD_EMPTY;
        pos[i] = kr[i] != MSD_EMPTY ? atomicAdd(&h[(u32)((kr[i] - kbase) >> sh) & dm], 1u) : 0u;
    }
    __syncthreads();
    if (MODE == 0) {
        for (u32 b = threadIdx.x; b < nb; b += PS_THREADS)
            if (h[b]) atomicAdd(&hist[(u64)r * nb + b], h[b]);
        return;
    }
    {   // starts of the groups inside the tile; one reservation per (tile, digit)
        const u32 c = threadIdx.x < nb ? h[threadIdx.x] : 0u;
        u32 tot;
        const u32 ex = block_excl_scan(c, ws, &tot);
        if (threadIdx.x < nb) {
            h[threadIdx.x] = ex;
            const u64 ob = (u64)r * nb + threadIdx.x;
            gp[threadIdx.x] = c ? ostart[ob] + atomicAdd(&hist[ob], c) - ex : 0u;   // hist doubles as the cursor array (zeroed)
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const u32 j = threadIdx.x + (u32)i * PS_THREADS;
            if (kr[i] != MSD_EMPTY) {
                const u32 q = h[(u32)((kr[i] - kbase) >> sh) & dm] + pos[i];
                sk[q] = kr[i];
                sc[q] = cnts[tile0 + j];
            }
        }
        __syncthreads();
        for (u32 q = threadIdx.x; q < tot; q += PS_THREADS) {
            const u64 key = sk[q];
            const u32 o = gp[(u32)((key - kbase) >> sh) & dm] + q;
            okeys[o] = key;
            ocnts[o] = sc[q];
        }
    }
}
// bucket starts of a pass: exact (compact) or every bucket rounded up to whole tiles (its tiles then belong to it alone)
__global__ void __launch_bounds__(1024) pair_starts_kernel(const u32 *__restrict__ hist, u32 n, int padded, u32 *__restrict__ start, u64 *__restrict__ total) {
    __shared__ u32 ws[33];
    __shared__ u64 s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 b0 = 0; b0 < n; b0 += 1024) {
        const u32 b = b0 + threadIdx.x;
        u32 c = b < n ? hist[b] : 0u;
        if (padded) c = (c + PS_TILE - 1) / PS_TILE;   // in tiles: stays within 32 bits
        u32 tot;
        const u32 ex = block_excl_scan(c, ws, &tot);
        if (b < n) {
            const u64 v = (s_run + ex) * (padded ? (u64)PS_TILE : 1ULL);
            start[b] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)v;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const u64 v = s_run * (padded ? (u64)PS_TILE : 1ULL);
        start[n] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)v;
        *total = v;
    }
}
__global__ void pair_tilemap_kernel(const u32 *__restrict__ start, u32 nreg, u32 *__restrict__ tile_reg) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreg) return;
    for (u32 t = start[r] / PS_TILE; t < start[r + 1] / PS_TILE; ++t) tile_reg[t] = r;
}
// max bucket size -> *mx; buckets larger than SKM_SORT_CAP -> big list (first 1024 of them)
__global__ void skm_sort_big_kernel(const u32 *__restrict__ hist, u32 nb, u32 *__restrict__ big, u32 *__restrict__ nbig) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb && hist[b] > SKM_SORT_CAP) {
        const u32 p = atomicAdd(nbig, 1u);
        if (p < 1024) big[p] = b;
    }
}
// One CTA per bucket (grid-stride): pairs IN[bstart[b] ..) of one top-bits bucket are ordered by key into OUT.
// Keys are distinct.  Order-preserving bins on the bits below the bucket digit + rank-by-counting inside a bin.
__global__ void __launch_bounds__(256) skm_sort_bucket_kernel(const u64 *__restrict__ ik, const u32 *__restrict__ ic, const u32 *__restrict__ bstart,
                                                              u32 nb, u64 kbase, int sh, u64 *__restrict__ ok, u32 *__restrict__ oc) {
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *sk = reinterpret_cast<u64 *>(smem_raw);              // SKM_SORT_CAP
    u32 *sc = reinterpret_cast<u32 *>(sk + SKM_SORT_CAP);     // SKM_SORT_CAP
    __shared__ u32 bins[1025], bincur[1024], ws[33];
    const int bb = sh < 10 ? sh : 10, bsh = sh - bb;
    const u32 bm = (1u << bb) - 1u;
    const int nbins = 1 << bb;
    for (u32 b = blockIdx.x; b < nb; b += gridDim.x) {
        const u32 s0 = bstart[b], n = bstart[b + 1] - s0;
        if (n == 0) continue;
        if (n > SKM_SORT_CAP) continue;  // handled by the host (generic radix sort of that range)
        if (n == 1) { if (threadIdx.x == 0) { ok[s0] = ik[s0]; oc[s0] = ic[s0]; } continue; }
        __syncthreads();
        for (int i = threadIdx.x; i <= nbins; i += 256) bins[i] = 0;
        __syncthreads();
        for (u32 i = threadIdx.x; i < n; i += 256) atomicAdd(&bins[(u32)((ik[s0 + i] - kbase) >> bsh) & bm], 1u);
        __syncthreads();
        {   // exclusive scan of bins[0 .. nbins) in place
            const int bpt = (nbins + 255) / 256;
            const int lo = min(nbins, (int)threadIdx.x * bpt), hi = min(nbins, lo + bpt);
            u32 s = 0;
            for (int i = lo; i < hi; ++i) s += bins[i];
            u32 tot;
            u32 run = block_excl_scan(s, ws, &tot);
            for (int i = lo; i < hi; ++i) { const u32 c = bins[i]; bins[i] = run; bincur[i] = run; run += c; }
            if (threadIdx.x == 0) bins[nbins] = n;
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < n; i += 256) {
            const u64 key = ik[s0 + i];
            const u32 p = atomicAdd(&bincur[(u32)((key - kbase) >> bsh) & bm], 1u);
            sk[p] = key;
            sc[p] = ic[s0 + i];
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < n; i += 256) {
            const u64 key = sk[i];
            const u32 bi = (u32)((key - kbase) >> bsh) & bm;
            const u32 lo = bins[bi], hi = bins[bi + 1];
            u32 rank = 0;
            for (u32 q = lo; q < hi; ++q) rank += sk[q] < key;
            ok[s0 + lo + rank] = key;
            oc[s0 + lo + rank] = sc[i];
        }
    }
}
__global__ void skm_widen_kernel(const u32 *__restrict__ in, u64 n, u64 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void skm_narrow_kernel(const u64 *__restrict__ in, u64 n, u32 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)in[i];
}

// number of k-mer windows of the packed words [w0, w1): popcount of the window masks
__global__ void __launch_bounds__(256) skm_popc_kernel(const u32 *__restrict__ valid, u64 w0, u64 w1, unsigned long long *__restrict__ total) {
    u64 c = 0;
    for (u64 t = w0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; t < w1; t += (u64)gridDim.x * blockDim.x) c += __popc(valid[t]);
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(FULL_MASK, c, d);
    if (lane_id() == 0 && c) atomicAdd(total, (unsigned long long)c);
}

// ------------------------------------- host side -------------------------------------------------
struct SkmTune {
    u32 bin_inst = 2048;   // target k-mer instances per leaf bin (distinct <= instances <= half the leaf table for most bins)
    int variant = 0;       // leaf geometry
    u32 sample = 16;       // capacity estimate from every `sample`-th tile
    u32 test_ovf_mod = 0;  // tests (GG_SKM_TEST_OVERFLOW=m): every leaf bin with index % m == 0 takes the overflow path
    u32 force_cap = 0;     // tests (GG_SKM_TEST_CAP=c): level-0 capacities of c records -> provokes the exact repeat
};
static SkmTune skm_tune() {
    SkmTune t;
    if (const char *e = getenv("GG_SKM_VARIANT")) t.variant = atoi(e);
    if (t.variant == 2 || t.variant == 5) t.bin_inst = 4096;
    if (t.variant == 3) t.bin_inst = 1024;
    if (const char *e = getenv("GG_SKM_BIN")) t.bin_inst = (u32)atoi(e);
    if (const char *e = getenv("GG_SKM_SAMPLE")) t.sample = (u32)atoi(e);
    if (const char *e = getenv("GG_SKM_TEST_OVERFLOW")) t.test_ovf_mod = (u32)atoi(e);
    if (const char *e = getenv("GG_SKM_TEST_CAP")) t.force_cap = (u32)atoi(e);
    if (t.bin_inst < 64) t.bin_inst = 64;
    if (t.sample < 1) t.sample = 1;
    return t;
}
// bin counts from the number of instances: leaf bins of ~bin_inst instances
static SkmGeom skm_geom(u64 n_inst, int k, int canonical, const SkmTune &t) {
    SkmGeom g;
    g.k = k; g.canonical = canonical;
    u64 nleaf = ceil_div(n_inst ? n_inst : 1, t.bin_inst);
    g.nb0 = (u32)(nleaf < SKM_MAX_NB0 ? nleaf : SKM_MAX_NB0);
    const u64 nb1 = ceil_div(nleaf, g.nb0);
    g.nb1 = (u32)(nb1 < SKM_MAX_NB1 ? nb1 : SKM_MAX_NB1);
    return g;
}
// exact number of k-mer windows of the packed words [w0, w1) (one pass over the 4-byte window masks, one synchronisation)
static u64 skm_count_windows(gg_ctx *ctx, const u32 *valid, u64 w0, u64 w1) {
    if (w1 <= w0) return 0;
    DBuf<u64> tot(ctx, 1);
    dfill(ctx, tot.p, sizeof(u64), 0u);
    const u64 grid = std::min<u64>(ceil_div(w1 - w0, 256 * 8), (u64)ctx->num_sms * 16);
    prof_bytes(ctx, 4.0 * (double)(w1 - w0));
    LAUNCH(ctx, skm_popc_kernel, (unsigned)grid, 256, 0, valid, w0, w1, (unsigned long long *)tot.p);
    return read_scalar<u64>(ctx, tot.p);
}

template <int MODE, bool PEER>
static void skm_scan_launch(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, u64 w0, u64 w1, u32 stride, const SkmGeom &g, u32 *hist0, u32 *cur0,
                            const u32 *base0, const u32 *cap0, ulonglong2 *out, u32 *hist1, u64 *totals, u32 *ovf, const SkmPeers *pt) {
    if (w1 <= w0) return;
    const u64 ntiles = ceil_div(w1 - w0, skm_tile_words(g.k));
    const u64 grid = ceil_div(ntiles, stride);
    SkmPeers none;
    memset(&none, 0, sizeof(none));
    const size_t smem = sizeof(SkmScanSmem);
    if (skm_logw(g.k) == 4) {
        set_max_smem(ctx, skm_scan_kernel<4, MODE, PEER>, smem);
        LAUNCH(ctx, (skm_scan_kernel<4, MODE, PEER>), (unsigned)grid, SKM_THREADS, smem, pr.packed.p, valid, w0, w1, stride, g, hist0, cur0, base0, cap0, out,
               hist1, (unsigned long long *)totals, ovf, pt ? *pt : none);
    } else {
        set_max_smem(ctx, skm_scan_kernel<3, MODE, PEER>, smem);
        LAUNCH(ctx, (skm_scan_kernel<3, MODE, PEER>), (unsigned)grid, SKM_THREADS, smem, pr.packed.p, valid, w0, w1, stride, g, hist0, cur0, base0, cap0, out,
               hist1, (unsigned long long *)totals, ovf, pt ? *pt : none);
    }
}

template <int THREADS, int LOG_HT>
static void skm_launch_leaf(gg_ctx *ctx, int ctas_per_sm, const ulonglong2 *recs, const u32 *lbeg, const u32 *lcnt, u32 nleaf, const SkmGeom &g,
                            u32 min_freq, u64 *okeys, u32 *ocnts, u64 *kept_cursor, u32 *ovf_list, u32 *ovf_n, u32 test_ovf_mod) {
    const size_t smem = sizeof(SkmLeafSmem<THREADS, LOG_HT>);
    u32 grid = (u32)ctx->num_sms * ctas_per_sm;
    if (grid > nleaf) grid = nleaf;
    if (skm_logw(g.k) == 4) {
        set_max_smem(ctx, skm_leaf_kernel<THREADS, LOG_HT, 4>, smem);
        LAUNCH(ctx, (skm_leaf_kernel<THREADS, LOG_HT, 4>), grid, THREADS, smem, recs, lbeg, lcnt, nleaf, g.k, g.canonical, min_freq, okeys, ocnts,
               (unsigned long long *)kept_cursor, ovf_list, ovf_n, test_ovf_mod);
    } else {
        set_max_smem(ctx, skm_leaf_kernel<THREADS, LOG_HT, 3>, smem);
        LAUNCH(ctx, (skm_leaf_kernel<THREADS, LOG_HT, 3>), grid, THREADS, smem, recs, lbeg, lcnt, nleaf, g.k, g.canonical, min_freq, okeys, ocnts,
               (unsigned long long *)kept_cursor, ovf_list, ovf_n, test_ovf_mod);
    }
}

struct SkmLiveF {   // compaction of the kept list: drop the unused slots
    const u64 *k;
    const u32 *c;
    u64 *ok;
    u32 *oc;
    __device__ bool pred(u64 i) const { return k[i] != MSD_EMPTY; }
    __device__ void emit(u64 i, u64 dst) const { ok[dst] = k[i]; oc[dst] = c[i]; }
};
// kept list (unordered; nslots slots in kk / kc, unused ones hold MSD_EMPTY) -> sorted ukeys / ucounts (new buffers);
// returns the number of pairs
// (kbase, kbits): every key lies in [kbase, kbase + 2^kbits) -- the whole key space by default; a rank of the multi-GPU
// path hands over its key range, so that the partition digits are taken below the range's common prefix
static u64 skm_sort_pairs(gg_ctx *ctx, u64 *kk, u32 *kc, u64 nslots, u64 n_pairs, int k, DBuf<u64> &ukeys, DBuf<u32> &ucounts, u64 kbase = 0,
                          int kbits = -1) {
    if (kbits < 0 || kbits > 2 * k) { kbits = 2 * k; kbase = 0; }
    if (nslots == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    if (nslots >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu kept-list slots exceed the 2^32-1 per-GPU limit", (unsigned long long)nslots);
    auto generic = [&](u64 *keys_io, u32 *cnts_io, u64 cnt, u64 *ok, u32 *oc) {  // LSD radix sort with the count as payload
        if (cnt == 0) return;
        DBuf<u64> alt(ctx, cnt), pay(ctx, cnt), pay_alt(ctx, cnt);
        LAUNCH(ctx, skm_widen_kernel, (unsigned)ceil_div(cnt, 256), 256, 0, cnts_io, cnt, pay.p);
        u64 *a = keys_io, *b = alt.p, *p = pay.p, *q = pay_alt.p;
        std::vector<BitRange> ranges = {{0, 2 * k}};
        radix_sort<1>(ctx, &a, &b, &p, &q, cnt, ranges);
        CK(cudaMemcpyAsync(ok, a, cnt * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        LAUNCH(ctx, skm_narrow_kernel, (unsigned)ceil_div(cnt, 256), 256, 0, p, cnt, oc);
    };
    auto compact_generic = [&]() -> u64 {  // drop the unused slots, then the generic sort
        SkmLiveF f{kk, kc, nullptr, nullptr};
        Compactor<SkmLiveF> cp(ctx, nslots, "live");
        const u64 n = cp.count(f);
        DBuf<u64> ck(ctx, n ? n : 1);
        DBuf<u32> cc(ctx, n ? n : 1);
        f.ok = ck.p; f.oc = cc.p;
        cp.emit(f);
        ukeys.alloc(ctx, n ? n : 1);
        ucounts.alloc(ctx, n ? n : 1);
        generic(ck.p, cc.p, n, ukeys.p, ucounts.p);
        CK(cudaStreamSynchronize(ctx->stream));  // ck / cc go back to the allocator
        return n;
    };
    int tb = 0;
    while (tb < 24 && tb < kbits && (n_pairs >> tb) > 1024) ++tb;
    if (tb < 4) return compact_generic();
    // ---- MSD partition on the top tb key bits in passes of <= 9 bits (staged tiles), then one ordering pass per bucket ----
    const int npass = tb <= 9 ? 1 : tb <= 18 ? 2 : 3;
    const int sh = kbits - tb;
    const u32 nb = 1u << tb;
    DBuf<u32> hist, bstart, big(ctx, 1024), nbig(ctx, 1);
    DBuf<u64> tk, tot(ctx, 1);
    DBuf<u32> tc;
    dfill(ctx, nbig.p, sizeof(u32), 0u);
    const size_t ps_smem = (size_t)PS_MAX_NB * 8 + (size_t)PS_TILE * 12;
    set_max_smem(ctx, pair_part_kernel<0>, ps_smem);
    set_max_smem(ctx, pair_part_kernel<1>, ps_smem);
    u64 n = 0;
    {
        const u64 *ik = kk;
        const u32 *ic = kc;
        u64 in_slots = nslots;
        DBuf<u32> rbase, rcnt, tile_reg;   // regions of the current input (none for the first pass)
        u32 nreg = 1;
        int bits_done = 0;
        for (int pass = 0; pass < npass; ++pass) {
            const int bits = (tb - bits_done + (npass - pass) - 1) / (npass - pass);
            const int psh = kbits - bits_done - bits;
            const u32 pnb = 1u << bits;
            const bool last = pass == npass - 1;
            const u64 nbuck = (u64)nreg * pnb;
            DBuf<u32> h(ctx, nbuck + 1), st(ctx, nbuck + 1);
            dfill(ctx, h.p, (nbuck + 1) * sizeof(u32), 0u);
            const unsigned ntile = (unsigned)ceil_div(in_slots, PS_TILE);
            prof_bytes(ctx, 8.0 * (double)in_slots);
            LAUNCH(ctx, pair_part_kernel<0>, ntile, PS_THREADS, ps_smem, ik, ic, tile_reg.p, rbase.p, rcnt.p, in_slots, kbase, psh, pnb, h.p, nullptr, nullptr, nullptr);
            LAUNCH(ctx, pair_starts_kernel, 1, 1024, 0, h.p, (u32)nbuck, last ? 0 : 1, st.p, tot.p);
            if (last) LAUNCH(ctx, skm_sort_big_kernel, (unsigned)ceil_div(nbuck, 256), 256, 0, h.p, (u32)nbuck, big.p, nbig.p);
            const u64 out_slots = read_scalar<u64>(ctx, tot.p);
            if (out_slots >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu padded pair slots exceed the 2^32-1 per-GPU limit", (unsigned long long)out_slots);
            DBuf<u64> ok(ctx, out_slots + 1);
            DBuf<u32> oc(ctx, out_slots + 1), cnt_keep(ctx, nbuck + 1);
            CK(cudaMemcpyAsync(cnt_keep.p, h.p, (nbuck + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
            dfill(ctx, h.p, (nbuck + 1) * sizeof(u32), 0u);   // now the cursors
            prof_bytes(ctx, 12.0 * (double)in_slots + 12.0 * (double)n_pairs);
            LAUNCH(ctx, pair_part_kernel<1>, ntile, PS_THREADS, ps_smem, ik, ic, tile_reg.p, rbase.p, rcnt.p, in_slots, kbase, psh, pnb, h.p, st.p, ok.p, oc.p);
            if (last) { n = out_slots; bstart = std::move(st); }
            else {   // the buckets become the regions of the next pass
                tile_reg.alloc(ctx, out_slots / PS_TILE + 1);
                LAUNCH(ctx, pair_tilemap_kernel, (unsigned)ceil_div(nbuck, 128), 128, 0, st.p, (u32)nbuck, tile_reg.p);
                rbase = std::move(st);
                rcnt = std::move(cnt_keep);
                nreg = (u32)nbuck;
            }
            CK(cudaStreamSynchronize(ctx->stream));   // the previous pass's buffers go back to the allocator
            tk = std::move(ok); tc = std::move(oc);
            ik = tk.p; ic = tc.p;
            in_slots = out_slots;
            bits_done += bits;
        }
    }
    const u32 h_nbig = read_scalar<u32>(ctx, nbig.p);
    if (h_nbig > 1024) return compact_generic();  // hopeless skew: sort everything the slow way
    ukeys.alloc(ctx, n ? n : 1);
    ucounts.alloc(ctx, n ? n : 1);
    if (n == 0) return 0;
    const size_t smem = (size_t)SKM_SORT_CAP * 12;
    set_max_smem(ctx, skm_sort_bucket_kernel, smem);
    prof_bytes(ctx, 24.0 * (double)n);
    LAUNCH(ctx, skm_sort_bucket_kernel, (unsigned)std::min<u64>(nb, (u64)ctx->num_sms * 8), 256, smem, tk.p, tc.p, bstart.p, nb, kbase, sh, ukeys.p, ucounts.p);
    if (h_nbig) {
        std::vector<u32> hb(h_nbig), hs(2);
        CK(cudaMemcpyAsync(hb.data(), big.p, h_nbig * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (u32 b : hb) {
            CK(cudaMemcpyAsync(hs.data(), bstart.p + b, 2 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            generic(tk.p + hs[0], tc.p + hs[0], (u64)hs[1] - hs[0], ukeys.p + hs[0], ucounts.p + hs[0]);
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));  // tk / tc go back to the allocator
    return n;
}

// Level-0 state of a scatter (also filled chunk by chunk by the pipelined host path).
struct SkmL0 {
    SkmGeom g;
    DBuf<u32> hist0, cur0, base0, cap0, hist1, ovf;   // cur0: strided (SKM_CUR_STRIDE)
    DBuf<u64> totals, total_cap;   // totals: [0] instances, [1] records
    DBuf<ulonglong2> rec0;
    u64 cap_total = 0;
    bool exact = false;
    void alloc_sample(gg_ctx *ctx) {   // g is set
        hist0.alloc(ctx, g.nb0 + 1);
        totals.alloc(ctx, 2);
        dfill(ctx, hist0.p, (u64)(g.nb0 + 1) * sizeof(u32), 0u);
        dfill(ctx, totals.p, 2 * sizeof(u64), 0u);
    }
};
// after the sample pass (hist0 filled by skm_scan_kernel<.., 0, ..> over 1 / scale of the input): capacities, record
// buffer, zeroed scatter state.  One stream synchronisation.
static void skm_l0_prepare(gg_ctx *ctx, SkmL0 &L, double scale, bool exact, const SkmTune &tune) {
    L.total_cap.alloc(ctx, 1);
    const u32 nb0 = L.g.nb0;
    const u64 nleaf = (u64)L.g.nb0 * L.g.nb1;
    L.base0.alloc(ctx, nb0 + 1); L.cap0.alloc(ctx, nb0); L.cur0.alloc(ctx, (u64)nb0 * SKM_CUR_STRIDE); L.hist1.alloc(ctx, nleaf + 1);
    L.ovf.alloc(ctx, 1);
    LAUNCH(ctx, skm_caps_kernel, 1, 1024, 0, L.hist0.p, nb0, (float)scale, exact ? 1 : 0, L.base0.p, L.cap0.p, L.total_cap.p,
           exact ? 0u : tune.force_cap);
    L.cap_total = read_scalar<u64>(ctx, L.total_cap.p);
    if (L.cap_total >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu super-k-mer records exceed the 2^32-1 per-GPU limit", (unsigned long long)L.cap_total);
    L.rec0.alloc(ctx, L.cap_total ? L.cap_total : 1);
    dfill(ctx, L.cur0.p, (u64)nb0 * SKM_CUR_STRIDE * sizeof(u32), 0u);
    dfill(ctx, L.hist1.p, (nleaf + 1) * sizeof(u32), 0u);
    dfill(ctx, L.ovf.p, sizeof(u32), 0u);
    dfill(ctx, L.totals.p, 2 * sizeof(u64), 0u);
    L.exact = exact;
}

// Everything after a successful level-0 scatter: regions (base / cnt, nreg = nsrc * nb_local regions of `in`, source-major) ->
// sorted unique keys + counts.  hist1: records per leaf bin of the local bins (nleaf entries).  n_inst: k-mer instances in
// the regions, or ~0 when unknown (multi-GPU owner: counted by level 1).
static u64 skm_finish(gg_ctx *ctx, const ulonglong2 *in, const u32 *base, const u32 *cnt, u32 cnt_stride, u32 nb_local, u32 nsrc, u64 in_cap,
                      const SkmGeom &g, const u32 *hist1, u64 n_inst, u64 n_rec, u32 min_freq, const SkmTune &tune, DBuf<u64> &ukeys,
                      DBuf<u32> &ucounts, u64 *n_inst_out = nullptr) {
    if (n_rec == 0 || nb_local == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); if (n_inst_out) *n_inst_out = 0; return 0; }
    const u64 nleaf = (u64)nb_local * g.nb1;
    stage_begin(ctx, ST_SORT);
    DBuf<u32> lbeg(ctx, nleaf + 1), cur1, dense;
    DBuf<ulonglong2> rec1;
    const ulonglong2 *leaf_recs = in;
    const u32 *leaf_beg = base, *leaf_cnt = cnt;
    if (g.nb1 > 1 || nsrc > 1) {
        exclusive_scan_u32(ctx, hist1, lbeg.p, nleaf + 1, nullptr);
        cur1.alloc(ctx, nleaf);
        dfill(ctx, cur1.p, nleaf * sizeof(u32), 0u);
        rec1.alloc(ctx, n_rec);
        const u32 ntile = (u32)ceil_div(in_cap, SKM_ALIGN);
        DBuf<u32> tile_reg(ctx, ntile + 1);
        LAUNCH(ctx, skm_tilemap_kernel, (unsigned)ceil_div((u64)nb_local * nsrc, 128), 128, 0, base, nb_local * nsrc, tile_reg.p);
        prof_bytes(ctx, 32.0 * (double)n_rec);
        const size_t smem1 = (size_t)g.nb1 * 8;
        set_max_smem(ctx, skm_l1_kernel, (size_t)SKM_MAX_NB1 * 8);
        DBuf<u64> itot;
        if (n_inst == ~0ULL) { itot.alloc(ctx, 1); dfill(ctx, itot.p, sizeof(u64), 0u); }
        LAUNCH(ctx, skm_l1_kernel, ntile, SKM_L1_THREADS, smem1, in, base, cnt, cnt_stride, tile_reg.p, nb_local, g.nb1, lbeg.p, cur1.p, rec1.p,
               (unsigned long long *)itot.p);
        if (n_inst == ~0ULL) n_inst = read_scalar<u64>(ctx, itot.p);
        leaf_recs = rec1.p; leaf_beg = lbeg.p; leaf_cnt = hist1;
    } else if (cnt_stride != 1) {
        dense.alloc(ctx, nleaf);
        LAUNCH(ctx, skm_gather_strided_kernel, (unsigned)ceil_div(nleaf, 256), 256, 0, cnt, cnt_stride, (u32)nleaf, dense.p);
        leaf_cnt = dense.p;
    }
    if (n_inst == ~0ULL) GG_THROW(GG_ERR_ARG, "skm_finish: instance count needed when level 1 is skipped");
    if (n_inst_out) *n_inst_out = n_inst;
    gg_trace(ctx, "skm level 1");
    // ---- E: leaves ----
    // every kept key has >= min_freq instances; the CTAs of the leaf kernel take the kept list in blocks of SKM_KBLOCK slots and
    // leave up to 3/4 of a leaf table unused at the end of each
    const u64 kept_bound = n_inst / (min_freq ? min_freq : 1) + 1;
    // (a block is given up with < LIST <= 6144 free slots: at most 6144 / (32768 - 6144) = 23 % of the slots stay unused)
    const u64 kept_cap = kept_bound + kept_bound / 4 + (u64)(ctx->num_sms * 8 + 2) * SKM_KBLOCK;
    DBuf<u64> kk(ctx, kept_cap), cursor(ctx, 2);
    DBuf<u32> kc(ctx, kept_cap), ovf_n(ctx, 2), ovf_list(ctx, nleaf);
    dfill(ctx, cursor.p, 2 * sizeof(u64), 0u);
    dfill(ctx, ovf_n.p, 2 * sizeof(u32), 0u);
    prof_bytes(ctx, 16.0 * (double)n_rec);
    if (tune.variant == 1)   // 4 CTAs / SM, 4096-slot tables, bins of <= 2048 instances
        skm_launch_leaf<256, 12>(ctx, 4, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    else if (tune.variant == 3)   // 8 CTAs / SM, 2048-slot tables, bins of <= 1024 instances
        skm_launch_leaf<256, 11>(ctx, 8, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    else if (tune.variant == 4)   // 4 CTAs / SM of 512 threads, 4096-slot tables
        skm_launch_leaf<512, 12>(ctx, 4, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    else if (tune.variant == 2)
        skm_launch_leaf<256, 13>(ctx, 2, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    else if (tune.variant == 5)   // 2 CTAs / SM, 8192-slot tables, bins of <= 4096 instances
        skm_launch_leaf<512, 13>(ctx, 2, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    else                     // 4 CTAs / SM of 512 threads (64 warps per SM), 4096-slot tables, bins of <= 2048 instances
        skm_launch_leaf<512, 12>(ctx, 4, leaf_recs, leaf_beg, leaf_cnt, (u32)nleaf, g, min_freq, kk.p, kc.p, cursor.p, ovf_list.p, ovf_n.p, tune.test_ovf_mod);
    CK(cudaMemcpyAsync(ctx->h_scalar, cursor.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_scalar + 2, ovf_n.p, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u64 U = ctx->h_scalar[0];   // slots of the kept list handed out so far (unused ones are marked empty)
    u64 n_pairs = ctx->h_scalar[1];
    const u32 n_ovf = (u32)ctx->h_scalar[2];
    gg_trace(ctx, "skm leaves");
    if (n_ovf) {  // bins whose table filled up: their k-mers as plain keys -> radix sort + run-length count + filter
        DBuf<u64> tot(ctx, 2);
        dfill(ctx, tot.p, 2 * sizeof(u64), 0u);
        LAUNCH(ctx, skm_ovf_size_kernel, (unsigned)std::min<u32>(n_ovf, (u32)ctx->num_sms * 8), 256, 0, leaf_recs, leaf_beg, leaf_cnt, ovf_list.p, n_ovf,
               (unsigned long long *)tot.p);
        const u64 n_keys = read_scalar<u64>(ctx, tot.p);
        DBuf<u64> keys(ctx, n_keys + 1), alt(ctx, n_keys + 1);
        LAUNCH(ctx, skm_ovf_expand_kernel, (unsigned)std::min<u32>(n_ovf, (u32)ctx->num_sms * 8), 256, 0, leaf_recs, leaf_beg, leaf_cnt, ovf_list.p, n_ovf, g.k,
               g.canonical, keys.p, (unsigned long long *)(tot.p + 1));
        u64 *a = keys.p, *b = alt.p;
        std::vector<BitRange> ranges = {{0, 2 * g.k}};
        radix_sort<1>(ctx, &a, &b, nullptr, nullptr, n_keys, ranges);
        DBuf<u64> uk;
        DBuf<u32> uc;
        const u64 nu = run_length_count<1>(ctx, a, n_keys, uk, uc);
        FilterF<1> f{uk.p, uc.p, kk.p + U, kc.p + U, (u64)min_freq, 0};
        Compactor<FilterF<1>> cp(ctx, nu, "filter");
        const u64 nk = cp.count(f);
        cp.emit(f);
        U += nk;
        n_pairs += nk;
    }
    gg_trace(ctx, "skm overflow bins");
    stage_end(ctx, ST_SORT);
    // ---- F: order the kept pairs ----
    stage_begin(ctx, ST_COUNT);
    U = skm_sort_pairs(ctx, kk.p, kc.p, U, n_pairs, g.k, ukeys, ucounts);
    prof_add_bytes(ctx, "skm_leaf_kernel", 12.0 * (double)U);
    gg_trace(ctx, "skm sort");
    stage_end(ctx, ST_COUNT);
    return U;
}

// packed reads -> sorted unique k-mers + counts (count >= min_freq only).  skm_supported(k, canonical).
static u64 skm_count(gg_ctx *ctx, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u64> &ukeys, DBuf<u32> &ucounts,
                     u64 *n_instances) {
    *n_instances = 0;
    if (pr.nwords == 0) { ukeys.alloc(ctx, 1); ucounts.alloc(ctx, 1); return 0; }
    const SkmTune tune = skm_tune();
    stage_begin(ctx, ST_EXTRACT);
    const u64 ntiles = ceil_div(pr.nwords, skm_tile_words(k));
    u32 stride = ntiles >= 256 ? tune.sample : 1;
    SkmL0 L;
    L.g = skm_geom(skm_count_windows(ctx, valid, 0, pr.nwords), k, canonical, tune);
    for (int attempt = 0;; ++attempt) {
        L.alloc_sample(ctx);
        prof_bytes(ctx, 12.0 * (double)pr.nwords / stride);
        skm_scan_launch<0, false>(ctx, pr, valid, 0, pr.nwords, stride, L.g, L.hist0.p, nullptr, nullptr, nullptr, nullptr, nullptr, L.totals.p, nullptr, nullptr);
        gg_trace(ctx, "skm sample");
        skm_l0_prepare(ctx, L, (double)stride, stride == 1 && !(attempt == 0 && tune.force_cap), tune);
        prof_bytes(ctx, 12.0 * (double)pr.nwords);  // + 16 B per record, added once the count is known
        skm_scan_launch<1, false>(ctx, pr, valid, 0, pr.nwords, 1, L.g, nullptr, L.cur0.p, L.base0.p, L.cap0.p, L.rec0.p, L.hist1.p, L.totals.p, L.ovf.p,
                                  nullptr);
        CK(cudaMemcpyAsync(ctx->h_scalar, L.totals.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_scalar + 2, L.ovf.p, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        gg_trace(ctx, "skm scatter");
        if ((u32)ctx->h_scalar[2] == 0) break;
        if (L.exact || attempt > 0) GG_THROW(GG_ERR_CUDA, "super-k-mer scatter overflowed exact capacities");
        stride = 1;  // a bin outgrew its sampled share: repeat with exact counts
        L.rec0.release();
    }
    const u64 n_inst = ctx->h_scalar[0], n_rec = ctx->h_scalar[1];
    prof_add_bytes(ctx, "skm_scan_kernel", 16.0 * (double)n_rec);
    *n_instances = n_inst;
    stage_end(ctx, ST_EXTRACT);
    return skm_finish(ctx, L.rec0.p, L.base0.p, L.cur0.p, SKM_CUR_STRIDE, L.g.nb0, 1, L.cap_total, L.g, L.hist1.p, n_inst, n_rec, min_freq, tune, ukeys,
                      ucounts);
}
