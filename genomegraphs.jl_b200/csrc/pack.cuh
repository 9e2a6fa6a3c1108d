// pack.cuh -- K1: ASCII reads -> 2-bit packed words + validity masks; K2: canonical k-mer
// extraction from the packed stream.
//
// Replaces (reference): FASTX parse + ReadDatastores encode behind read_datastore
// (src/GenomeGraphs.jl:79-83) and BioSequences each(M, seq) + canonical
// (in-repo statement src/GraphIndexes.jl:78-82).
//
// Device layout (all "first base most significant" so a window is one funnel shift):
//   packed[j] (u64): bases 32j .. 32j+31, base 32j in bits 63:62
//   bad[j]    (u32): bit (31-i) set  <=> base 32j+i is not A/C/G/T (padding is all bad)
//   brk[j]    (u32): bit (31-i) set  <=> base 32j+i is the first base of a read
//   valid[j]  (u32): bit (31-i) set  <=> the k-window starting at base 32j+i is a k-mer
#pragma once
#include "prims.cuh"

#define PACK_PAD_WORDS 8  // zero/bad padding words after the last real word (k <= 128 windows)

struct PackedReads {
    u64 nbases = 0;    // = nbytes
    u64 nwords = 0;    // words holding real bases: ceil(nbases / 32)
    DBuf<u64> packed;  // nwords + PACK_PAD_WORDS
    DBuf<u32> bad;     // nwords + PACK_PAD_WORDS
    DBuf<u32> brk;     // nwords + PACK_PAD_WORDS
};

__device__ __forceinline__ void pack_byte(u32 c, u32 &code, u32 &isbad) {
    u32 up = c & 0xDFu;
    u32 t = (up >> 1) & 3u;
    code = t ^ (t >> 1);  // A0 C1 G2 T3
    isbad = !(up == 'A' || up == 'C' || up == 'G' || up == 'T');
}

// four ASCII bases (byte 0 = first base) at once: 8 code bits (first base in the top 2 bits) and 4 bad bits
// (first base in bit 3).  Codes come from bits 1-2 of the upper-cased letter; validity by comparing with the
// letter that code stands for (PRMT table lookup), all four bytes in one 32-bit word.
__device__ __forceinline__ void pack4(u32 w, u32 &code8, u32 &bad4) {
    const u32 up = w & 0xDFDFDFDFu;
    const u32 t = (up >> 1) & 0x03030303u;                   // A0 C1 T2 G3
    const u32 c = t ^ ((t >> 1) & 0x01010101u);              // A0 C1 G2 T3
    code8 = (c * 0x40100401u) >> 24;
    const u32 expect = __byte_perm(0x47544341u, 0u, t | (t >> 12));  // "ACTG"[t] for bases 0, 2, 1, 3
    const u32 x = __byte_perm(up, 0u, 0x3120u) ^ expect;
    const u32 nz = ((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u) >> 7;  // 1 per mismatching byte (bases 0, 2, 1, 3)
    bad4 = ((nz * 0x08020401u) >> 24) & 0xFu;
}

// one thread = 16 bases (one 128-bit load when aligned and fully inside the buffer)
// groups [g0, ngroups) are packed by this launch (g0 even): the pipelined host path packs chunk by chunk
__global__ void __launch_bounds__(256) pack_kernel(const u8 *__restrict__ seq, u64 nbytes, int aligned16,
                                                   u32 *__restrict__ packed32, u32 *__restrict__ bad, u64 g0, u64 ngroups) {
    u64 g = g0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;  // 16-base group index
    u32 pk = 0, bd = 0xFFFFu;
    if (g < ngroups) {
        u64 b0 = g * 16;
        u32 w[4];
        if (aligned16 && b0 + 16 <= nbytes) {
            uint4 v = *reinterpret_cast<const uint4 *>(seq + b0);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            w[0] = w[1] = w[2] = w[3] = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                u32 c = (b0 + i < nbytes) ? seq[b0 + i] : 0u;
                w[i >> 2] |= c << (8 * (i & 3));
            }
        }
        bd = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            u32 code8, bad4;
            pack4(w[q], code8, bad4);
            pk |= code8 << (24 - 8 * q);
            bd |= bad4 << (12 - 4 * q);
        }
    }
    // packed u64 word j = [group 2j | group 2j+1]; little-endian halves: u32 index 2j+1 is the high half
    if (g < ngroups) packed32[(g & ~1ULL) + (1 - (g & 1))] = pk;
    u32 other = __shfl_down_sync(FULL_MASK, bd, 1);
    if (g < ngroups + (ngroups & 1) && !(g & 1)) bad[g >> 1] = (bd << 16) | (other & 0xFFFFu);
}
// 4-bit input (the reference's read store keeps LongSequence{DNAAlphabet{4}}, src/GenomeGraphs.jl:79-83: BioSequences
// packs 16 bases per UInt64, base i in bits 4i .. 4i+3, A = 0001, C = 0010, G = 0100, T = 1000, anything else
// ambiguous): one thread = one input word = 16 bases; same outputs as pack_kernel.
__global__ void __launch_bounds__(256) pack4bit_kernel(const u64 *__restrict__ data4, u64 nbases, u32 *__restrict__ packed32, u32 *__restrict__ bad,
                                                       u64 g0, u64 ngroups) {
    u64 g = g0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 pk = 0, bd = 0xFFFFu;
    if (g < ngroups) {
        const u64 w = data4[g];
        bd = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const u32 x = (u32)(w >> (4 * i)) & 15u;
            const u32 code = ((x >> 1) - (x >> 3)) & 3u;            // 1 -> 0, 2 -> 1, 4 -> 2, 8 -> 3
            const u32 isbad = (x == 0u || (x & (x - 1u)) != 0u || g * 16 + i >= nbases) ? 1u : 0u;
            pk |= (isbad ? 0u : code) << (30 - 2 * i);
            bd |= isbad << (15 - i);
        }
    }
    if (g < ngroups) packed32[(g & ~1ULL) + (1 - (g & 1))] = pk;
    u32 other = __shfl_down_sync(FULL_MASK, bd, 1);
    if (g < ngroups + (ngroups & 1) && !(g & 1)) bad[g >> 1] = (bd << 16) | (other & 0xFFFFu);
}
__global__ void brk_kernel(const u64 *__restrict__ offs, u64 nreads, u64 nbytes, u32 *__restrict__ brk) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    u64 p = offs[r];
    if (p < nbytes) atomicOr(&brk[p >> 5], 1u << (31 - (p & 31)));
}
__global__ void fill_u32_kernel(u32 *p, u64 n, u32 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct Bits6 {
    u32 w[6];
};
// result[pos] = x[pos + s] (positions are MSB-first), zeros shifted in; 0 <= s < 192
__device__ __forceinline__ Bits6 shl6(const Bits6 &x, int s) {
    Bits6 r;
    int ws = s >> 5, bs = s & 31;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        u32 a = 0, b = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j == i + ws) a = x.w[j];
            if (j == i + ws + 1) b = x.w[j];
        }
        r.w[i] = __funnelshift_l(b, a, bs);  // upper 32 bits of (a:b) << bs, bs in 0..31
    }
    return r;
}
// OR over windows of m consecutive positions: r[p] = OR_{i<m} x[p+i]  (m >= 0)
__device__ __forceinline__ Bits6 window_or6(Bits6 x, int m) {
    Bits6 acc;
    if (m <= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) acc.w[i] = 0;
        return acc;
    }
    acc = x;
    int sz = 1;
    while (2 * sz <= m) {
        Bits6 t = shl6(acc, sz);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc.w[i] |= t.w[i];
        sz *= 2;
    }
    if (sz < m) {
        Bits6 t = shl6(acc, m - sz);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc.w[i] |= t.w[i];
    }
    return acc;
}
// one thread = 32 window start positions.  A window [p, p+k) is a k-mer iff it holds no bad
// base and no read starts strictly inside it.
__global__ void __launch_bounds__(256) valid_kernel(const u32 *__restrict__ bad, const u32 *__restrict__ brk, u64 nwords,
                                                    int k, u32 *__restrict__ valid) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    Bits6 B, R;
#pragma unroll
    for (int i = 0; i < 6; ++i) { B.w[i] = bad[j + i]; R.w[i] = brk[j + i]; }  // padding keeps this in bounds
    Bits6 R1 = shl6(R, 1), M;
#pragma unroll
    for (int i = 0; i < 6; ++i) M.w[i] = B.w[i] | R1.w[i];
    u32 inv = window_or6(M, k - 1).w[0] | shl6(B, k - 1).w[0];
    valid[j] = ~inv;
}

// k <= 32: every window lies inside two mask words, so the sliding OR runs on one u64
__global__ void __launch_bounds__(256) valid32_kernel(const u32 *__restrict__ bad, const u32 *__restrict__ brk, u64 w0, u64 nwords, int k,
                                                      u32 *__restrict__ valid) {
    u64 j = w0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    const u64 B = ((u64)bad[j] << 32) | bad[j + 1], R = ((u64)brk[j] << 32) | brk[j + 1];  // MSB-first positions 0 .. 63
    const int m = k - 1;
    u64 acc = 0;
    if (m > 0) {
        acc = B | (R << 1);
        int sz = 1;
        while (2 * sz <= m) { acc |= acc << sz; sz *= 2; }
        if (sz < m) acc |= acc << (m - sz);
    }
    const u64 inv = acc | (B << m);
    valid[j] = ~(u32)(inv >> 32);
}

// buffers of the packed form, cleared / padded; read starts marked (needs only the offsets)
// 32 < k <= 64: three mask words (96 positions) in one 128-bit value
__global__ void __launch_bounds__(256) valid64_kernel(const u32 *__restrict__ bad, const u32 *__restrict__ brk, u64 nwords, int k,
                                                      u32 *__restrict__ valid) {
    typedef unsigned __int128 u128;
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    const u128 B = ((u128)bad[j] << 96) | ((u128)bad[j + 1] << 64) | ((u128)bad[j + 2] << 32);  // MSB-first positions 0 .. 95
    const u128 R = ((u128)brk[j] << 96) | ((u128)brk[j + 1] << 64) | ((u128)brk[j + 2] << 32);
    const int m = k - 1;
    u128 acc = B | (R << 1);
    int sz = 1;
    while (2 * sz <= m) { acc |= acc << sz; sz *= 2; }
    if (sz < m) acc |= acc << (m - sz);
    const u128 inv = acc | (B << m);
    valid[j] = ~(u32)(inv >> 96);
}

static void pack_prepare(gg_ctx *ctx, const u64 *d_offs, u64 nreads, u64 nbytes, PackedReads &pr) {
    pr.nbases = nbytes;
    pr.nwords = ceil_div(nbytes, 32);
    u64 tot = pr.nwords + PACK_PAD_WORDS;
    pr.packed.alloc(ctx, tot);
    pr.bad.alloc(ctx, tot);
    pr.brk.alloc(ctx, tot);
    dfill(ctx, pr.packed.p, tot * sizeof(u64), 0u);
    dfill(ctx, pr.bad.p, tot * sizeof(u32), 0xFFFFFFFFu);
    dfill(ctx, pr.brk.p, tot * sizeof(u32), 0u);
    if (nbytes && nreads) LAUNCH(ctx, brk_kernel, (unsigned)ceil_div(nreads, 256), 256, 0, d_offs, nreads, nbytes, pr.brk.p);
}
// pack the 16-base groups [g0, g1) (g0 even; g1 = number of groups when it is the last piece)
static void pack_range(gg_ctx *ctx, const u8 *d_seq, u64 nbytes, PackedReads &pr, u64 g0, u64 g1) {
    if (g1 <= g0) return;
    const int aligned = (((uintptr_t)d_seq) & 15) == 0;
    // an even number of threads so that the shuffle partner of every even group exists
    const u64 nthreads = (g1 - g0) + ((g1 - g0) & 1);
    LAUNCH(ctx, pack_kernel, (unsigned)ceil_div(nthreads, 256), 256, 0, d_seq, nbytes, aligned, (u32 *)pr.packed.p, pr.bad.p, g0, g1);
}
static void pack_reads(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, PackedReads &pr) {
    pack_prepare(ctx, d_offs, nreads, nbytes, pr);
    pack_range(ctx, d_seq, nbytes, pr, 0, ceil_div(nbytes, 16));
}

// enc: 0 = ASCII bytes, 1 = 4-bit (d_seq is then an array of u64, 16 bases each; nbytes = number of bases)
static void pack_input(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, int enc, PackedReads &pr) {
    if (enc == 0) { pack_reads(ctx, d_seq, d_offs, nreads, nbytes, pr); return; }
    pack_prepare(ctx, d_offs, nreads, nbytes, pr);
    const u64 ng = ceil_div(nbytes, 16);
    if (ng) LAUNCH(ctx, pack4bit_kernel, (unsigned)ceil_div(ng + (ng & 1), 256), 256, 0, (const u64 *)d_seq, nbytes, (u32 *)pr.packed.p, pr.bad.p, 0ULL, ng);
}

// valid-window mask for a given k (nwords u32)
static void window_valid_mask(gg_ctx *ctx, const PackedReads &pr, int k, DBuf<u32> &valid) {
    valid.alloc(ctx, pr.nwords ? pr.nwords : 1);
    if (pr.nwords && k <= 32) LAUNCH(ctx, valid32_kernel, (unsigned)ceil_div(pr.nwords, 256), 256, 0, pr.bad.p, pr.brk.p, 0ULL, pr.nwords, k, valid.p);
    else if (pr.nwords && k <= 64) LAUNCH(ctx, valid64_kernel, (unsigned)ceil_div(pr.nwords, 256), 256, 0, pr.bad.p, pr.brk.p, pr.nwords, k, valid.p);
    else if (pr.nwords) LAUNCH(ctx, valid_kernel, (unsigned)ceil_div(pr.nwords, 256), 256, 0, pr.bad.p, pr.brk.p, pr.nwords, k, valid.p);
}

// ---------------------------------- K2: k-mer at position p -------------------------------
// forward k-mer of the window starting at base p, right-aligned
template <int W>
__device__ __forceinline__ Kmer<W> kmer_fw_at(const u64 *__restrict__ packed, u64 p, int k) {
    u64 j = p >> 5;
    int r2 = 2 * (int)(p & 31);
    u64 a[W + 1];
#pragma unroll
    for (int i = 0; i <= W; ++i) a[i] = packed[j + i];
    Kmer<W> t;
#pragma unroll
    for (int i = 0; i < W; ++i) t.w[i] = r2 ? ((a[i] << r2) | (a[i + 1] >> (64 - r2))) : a[i];
    return km_shr<W>(t, 64 * W - 2 * k);
}
template <int W>
__device__ __forceinline__ Kmer<W> kmer_at(const u64 *__restrict__ packed, u64 p, int k, int canonical, bool *fw_is_canonical) {
    Kmer<W> fw = kmer_fw_at<W>(packed, p, k);
    Kmer<W> rc = km_rc<W>(fw, k);
    bool c = km_cmp<W>(fw, rc) <= 0;
    if (fw_is_canonical) *fw_is_canonical = c;
    return (canonical && !c) ? rc : fw;
}

// compaction functor: all k-mers in read order
template <int W>
struct ExtractF {
    const u32 *valid;
    const u64 *packed;
    u64 *out;
    int k, canonical;
    __device__ bool pred(u64 p) const { return (valid[p >> 5] >> (31 - (p & 31))) & 1u; }
    __device__ void emit(u64 p, u64 dst) const { km_store<W>(out, dst, kmer_at<W>(packed, p, k, canonical, nullptr)); }
};
