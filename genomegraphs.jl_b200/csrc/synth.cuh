// synth.cuh -- deterministic synthetic genome / read generator.
// Counter-based: every byte is a pure function of (seed, index), so the host and device
// variants produce identical buffers and nothing has to be shipped to the GPU box.
#pragma once
#include "common.cuh"

__host__ __device__ __forceinline__ u64 sy_mix(u64 z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 sy_key(u64 seed, u64 stream) { return sy_mix(seed + stream * 0x9E3779B97F4A7C15ULL); }
__host__ __device__ __forceinline__ u64 sy_rnd(u64 key, u64 i) { return sy_mix(key + i * 0xD6E8FEB86659FD93ULL); }
__host__ __device__ __forceinline__ u64 sy_mulhi(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
__host__ __device__ __forceinline__ u32 sy_genome_base(u64 gkey, u64 i) { return (u32)(sy_rnd(gkey, i >> 5) >> (2 * (i & 31))) & 3u; }

struct SynthParams {
    u64 gkey, skey, dkey, ekey, nkey;
    u64 genome_len, nreads;
    u32 read_len, frag_len, err_ppm, n_ppm;
    int paired;
};
static SynthParams synth_params(u64 seed, u64 genome_len, u64 nreads, u32 read_len, int paired, u32 frag_len, u32 err_ppm, u32 n_ppm) {
    SynthParams P;
    P.gkey = sy_key(seed, 1); P.skey = sy_key(seed, 2); P.dkey = sy_key(seed, 3); P.ekey = sy_key(seed, 4); P.nkey = sy_key(seed, 5);
    P.genome_len = genome_len; P.nreads = nreads; P.read_len = read_len;
    P.frag_len = paired ? frag_len : read_len; P.err_ppm = err_ppm; P.n_ppm = n_ppm; P.paired = paired;
    return P;
}
// base p of read r
__host__ __device__ __forceinline__ u8 synth_base(const SynthParams &P, u64 r, u32 p) {
    u64 j = P.paired ? (r >> 1) : r;
    u32 mate = P.paired ? (u32)(r & 1) : 0u;
    u64 F = P.frag_len;
    u64 start = sy_mulhi(sy_rnd(P.skey, j), P.genome_len - F + 1);
    u32 strand = (u32)(sy_rnd(P.dkey, j) & 1u);
    u64 t = mate ? (F - 1 - p) : p;           // offset inside the oriented fragment
    u32 comp = mate;                          // mate 2 is the reverse complement of the fragment end
    u64 g = strand ? (start + F - 1 - t) : (start + t);
    comp ^= strand;
    u32 b = sy_genome_base(P.gkey, g);
    if (comp) b = 3u - b;
    u64 idx = r * (u64)P.read_len + p;
    if (P.err_ppm) {
        u64 e = sy_rnd(P.ekey, idx);
        if ((u32)(((u64)(u32)e * 1000000ULL) >> 32) < P.err_ppm) b = (b + 1u + (u32)((e >> 32) % 3u)) & 3u;
    }
    if (P.n_ppm) {
        u64 e = sy_rnd(P.nkey, idx);
        if ((u32)(((u64)(u32)e * 1000000ULL) >> 32) < P.n_ppm) return (u8)'N';
    }
    return (u8)(0x54474341u >> (8 * b));
}
__global__ void synth_reads_kernel(SynthParams P, u64 first_read, u8 *__restrict__ out) {
    u64 total = P.nreads * (u64)P.read_len;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x)
        out[i] = synth_base(P, first_read + i / P.read_len, (u32)(i % P.read_len));
}
