// common.cuh -- context, error handling, stream-ordered allocation, k-mer word type.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/ggb200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define GG_NSTAGE 9
enum { ST_PACK = 0, ST_EXTRACT, ST_SORT, ST_COUNT, ST_FILTER, ST_LOOKUP, ST_UNITIG, ST_LINKS, ST_TOTAL };

struct gg_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;       // second copy stream: read offsets travel beside the read bytes
    cudaEvent_t aux_ev = nullptr;
    std::vector<cudaEvent_t> chunk_ev;  // one per in-flight chunk of the pipelined host front end
    std::string err;
    // timing of the last pipeline call
    cudaEvent_t ev[2 * GG_NSTAGE];
    bool ev_used[GG_NSTAGE];
    float ms[GG_NSTAGE];
    u64 launches = 0;
    // small pinned scratch for device->host scalars
    u64 *h_scalar = nullptr;  // 64 x u64
    // optional per-kernel profiling (gg_ctx_set_profile): events around every launch
    bool profile = false;
    struct ProfRec { std::string name; cudaEvent_t a, b; double bytes; };
    double next_bytes = 0.0;  // algorithmic bytes of the next launch (prof_bytes); consumed by LAUNCH
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    std::string prof_text;
    // user events (gg_ctx_event_record / gg_ctx_event_elapsed_ms)
    cudaEvent_t user_ev[8];
    struct GgArena *arena = nullptr;  // caching device allocator (below)
    int live_handles = 0;             // gg_counts / gg_graph / gg_umi / gg_comm objects that still own memory of this context
    bool dying = false;               // gg_ctx_destroy was called while handles were alive: the last handle tears the context down
    int coop_blocks = -1;             // co-resident CTAs per SM of the cooperative list-ranking kernel on this device (-1: not asked yet)
    std::set<const void *> smem_attr_done;  // kernels whose dynamic shared-memory limit was raised on THIS device
};

// Handles keep their context alive: host languages with garbage collection (the Julia layer's finalizers) free
// objects in no particular order, so gg_ctx_destroy only marks a context that still has live handles and the last
// handle's release performs the teardown.
static void gg_ctx_destroy_now(gg_ctx *ctx);
struct CtxRef {
    gg_ctx *p = nullptr;
    CtxRef() {}
    CtxRef(const CtxRef &) = delete;
    CtxRef &operator=(const CtxRef &) = delete;
    CtxRef &operator=(gg_ctx *c) { release(); p = c; if (p) ++p->live_handles; return *this; }
    operator gg_ctx *() const { return p; }
    gg_ctx *operator->() const { return p; }
    void release() {
        gg_ctx *c = p;
        p = nullptr;
        if (c && --c->live_handles == 0 && c->dying) gg_ctx_destroy_now(c);
    }
    ~CtxRef() { release(); }
};

struct GgError {
    int code;
    std::string msg;
};

#define GG_THROW(code, ...)                                   \
    do {                                                      \
        char _b[512];                                         \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                \
        throw GgError{code, std::string(_b)};                 \
    } while (0)

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            GG_THROW(_e == cudaErrorMemoryAllocation ? GG_ERR_OOM : GG_ERR_CUDA, "%s:%d %s: %s", \
                     __FILE__, __LINE__, #call, cudaGetErrorString(_e));                      \
    } while (0)

// every kernel launch goes through this so that gpu_launches can be reported
static inline cudaEvent_t prof_event(gg_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->ev_pool.empty()) { e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
#define LAUNCH(ctx, kernel, grid, block, smem, ...) LAUNCH_NAMED(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)
#define LAUNCH_NAMED(ctx, kname, kernel, grid, block, smem, ...)                     \
    do {                                                                             \
        cudaEvent_t _pa = nullptr, _pb = nullptr;                                    \
        if ((ctx)->profile) { _pa = prof_event(ctx); _pb = prof_event(ctx); cudaEventRecord(_pa, (ctx)->stream); } \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);             \
        (ctx)->launches++;                                                           \
        if ((ctx)->profile) { cudaEventRecord(_pb, (ctx)->stream); (ctx)->prof.push_back({std::string(kname), _pa, _pb, (ctx)->next_bytes}); } \
        (ctx)->next_bytes = 0.0;                                                     \
        CK(cudaGetLastError());                                                      \
    } while (0)

// Algorithmic (compulsory) bytes of the NEXT launch, for the per-kernel roofline of the profile: every input element
// read once + every output element written once, evaluated with the launch's own element counts.
static inline void prof_bytes(gg_ctx *ctx, double bytes) { ctx->next_bytes = bytes; }
// bytes that are only known after the launch (output sizes): added to the latest profiled launch whose name contains `name`
static inline void prof_add_bytes(gg_ctx *ctx, const char *name, double bytes) {
    if (!ctx->profile) return;
    for (size_t i = ctx->prof.size(); i-- > 0;)
        if (ctx->prof[i].name.find(name) != std::string::npos) { ctx->prof[i].bytes += bytes; return; }
}

static inline u64 ceil_div(u64 a, u64 b) { return (a + b - 1) / b; }
// GG_TRACE=1: synchronise and print a marker (debugging aid for hangs on the GPU box; never set while timing)
static inline void gg_trace(gg_ctx *ctx, const char *what) {
    static int on = -1;
    if (on < 0) on = getenv("GG_TRACE") ? 1 : 0;
    if (!on) return;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[gg_trace] %s: %s\n", what, cudaGetErrorString(e));
    fflush(stderr);
}

// Raise a kernel's dynamic shared-memory limit once per context: the attribute belongs to the device, so a process
// that drives several GPUs must set it on each of them.
template <class K>
static inline void set_max_smem(gg_ctx *ctx, K kernel, size_t bytes) {
    const void *key = (const void *)kernel;
    if (ctx->smem_attr_done.count(key)) return;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        char b[256];
        snprintf(b, sizeof(b), "cudaFuncSetAttribute(%zu bytes of shared memory): %s", bytes, cudaGetErrorString(e));
        throw GgError{GG_ERR_CUDA, std::string(b)};
    }
    ctx->smem_attr_done.insert(key);
}

// Fill with a kernel instead of cudaMemsetAsync, so that the main stream carries no copy-engine style work in the
// part of the host path that overlaps with the host-to-device transfer (see count_kmers_host_pipelined).
__global__ void gg_fill_kernel(uint4 *p, u64 n16, u32 v) {
    const uint4 x = make_uint4(v, v, v, v);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64)gridDim.x * blockDim.x) p[i] = x;
}

// Device memory: a per-context caching allocator.  Every pipeline call allocates the same sequence of
// temporaries, so after the first call each request is served from the free list without touching
// the driver (cudaMallocAsync pools re-map physical memory between differently sized requests, and
// those driver calls serialise across the processes of a multi-GPU job: multi-millisecond stalls).
// All work of a context runs on ONE stream, so a block freed by the host can be handed out again at
// once: later kernels are ordered after the earlier users of the block.
struct GgArena {
    std::multimap<size_t, void *> free_blocks;   // size -> block
    std::map<void *, size_t> live;               // block -> size (allocated and cached blocks alike)
    size_t cached_bytes = 0, total_bytes = 0;
    u64 driver_allocs = 0;  // cudaMalloc calls so far (0 per call in steady state)
};
static inline GgArena &gg_arena(gg_ctx *ctx) {
    if (!ctx->arena) ctx->arena = new GgArena();
    return *ctx->arena;
}
static void arena_release_cached(gg_ctx *ctx) {
    GgArena &a = gg_arena(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : a.free_blocks) { cudaFree(kv.second); a.total_bytes -= kv.first; a.live.erase(kv.second); }
    a.free_blocks.clear();
    a.cached_bytes = 0;
}
static void *arena_alloc(gg_ctx *ctx, size_t bytes) {
    GgArena &a = gg_arena(ctx);
    bytes = (bytes + 511) & ~(size_t)511;
    if (bytes == 0) bytes = 512;
    auto it = a.free_blocks.lower_bound(bytes);
    // reuse a cached block unless it is much larger than the request (keeps big blocks for big requests)
    if (it != a.free_blocks.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
        void *p = it->second;
        a.cached_bytes -= it->first;
        a.free_blocks.erase(it);
        return p;
    }
    void *p = nullptr;
    a.driver_allocs++;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and retry once
        cudaGetLastError();
        arena_release_cached(ctx);
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        GG_THROW(e == cudaErrorMemoryAllocation ? GG_ERR_OOM : GG_ERR_CUDA, "cudaMalloc of %zu bytes: %s", bytes, cudaGetErrorString(e));
    }
    a.live[p] = bytes;
    a.total_bytes += bytes;
    return p;
}
static void arena_free(gg_ctx *ctx, void *p) {
    GgArena &a = gg_arena(ctx);
    auto it = a.live.find(p);
    if (it == a.live.end()) { cudaFree(p); return; }  // not ours (never happens)
    a.free_blocks.emplace(it->second, p);
    a.cached_bytes += it->second;
}
template <class T>
static T *dalloc(gg_ctx *ctx, u64 count) {
    return (T *)arena_alloc(ctx, (size_t)(count * sizeof(T)));
}
static inline void dfree(gg_ctx *ctx, void *p) {
    if (p) arena_free(ctx, p);
}
// RAII temp buffer
template <class T>
struct DBuf {
    gg_ctx *ctx = nullptr;
    T *p = nullptr;
    u64 n = 0;
    DBuf() {}
    DBuf(gg_ctx *c, u64 count) : ctx(c), n(count) { p = dalloc<T>(c, count); }
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
    DBuf &operator=(DBuf &&o) noexcept {
        if (this != &o) { release(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; }
        return *this;
    }
    void alloc(gg_ctx *c, u64 count) { release(); ctx = c; n = count; p = dalloc<T>(c, count); }
    void release() { if (p) { dfree(ctx, p); p = nullptr; } }
    T *take() { T *r = p; p = nullptr; return r; }
    ~DBuf() { release(); }
};

// device fill of `bytes` (rounded up to 16; every arena block is a multiple of 512 bytes) with a 32-bit pattern
static inline void dfill(gg_ctx *ctx, void *p, u64 bytes, u32 pattern) {
    const u64 n16 = ceil_div(bytes, 16);
    if (n16 == 0) return;
    u64 grid = ceil_div(n16, 256);
    if (grid > (u64)ctx->num_sms * 16) grid = (u64)ctx->num_sms * 16;
    LAUNCH(ctx, gg_fill_kernel, (unsigned)grid, 256, 0, (uint4 *)p, n16, pattern);
}
static inline void stage_begin(gg_ctx *ctx, int s) {
    if (!ctx->ev_used[s]) { cudaEventRecord(ctx->ev[2 * s], ctx->stream); ctx->ev_used[s] = true; }
}
static inline void stage_end(gg_ctx *ctx, int s) { cudaEventRecord(ctx->ev[2 * s + 1], ctx->stream); }

// read a device scalar (blocking on the context's stream)
template <class T>
static T read_scalar(gg_ctx *ctx, const T *d) {
    CK(cudaMemcpyAsync(ctx->h_scalar, d, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    T v;
    memcpy(&v, ctx->h_scalar, sizeof(T));
    return v;
}

// ------------------------------------------------------------------------------------------
// k-mer value: W little-endian u64 words, w[0] MOST significant; 2k bits right-aligned.
// ------------------------------------------------------------------------------------------
template <int W>
struct Kmer {
    u64 w[W];
};

__host__ __device__ __forceinline__ u64 rev2_u64(u64 x) {  // reverse the 32 2-bit groups
#ifdef __CUDA_ARCH__
    u64 y = __brevll(x);
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
#else
    x = __builtin_bswap64(x);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    return x;
#endif
}

template <int W>
__host__ __device__ __forceinline__ int km_cmp(const Kmer<W> &a, const Kmer<W> &b) {
#pragma unroll
    for (int i = 0; i < W; ++i)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}
template <int W>
__host__ __device__ __forceinline__ bool km_eq(const Kmer<W> &a, const Kmer<W> &b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < W; ++i) e = e && (a.w[i] == b.w[i]);
    return e;
}
template <int W>
__host__ __device__ __forceinline__ bool km_lt(const Kmer<W> &a, const Kmer<W> &b) { return km_cmp(a, b) < 0; }

// logical right shift of the W*64-bit value by s bits (0 <= s < 64*W)
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_shr(const Kmer<W> &x, int s) {
    Kmer<W> r;
    int ws = s >> 6, bs = s & 63;
#pragma unroll
    for (int i = W - 1; i >= 0; --i) {
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {  // select without dynamic register indexing
            if (j == i - ws) lo = x.w[j];
            if (j == i - ws - 1) hi = x.w[j];
        }
        r.w[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
    return r;
}
// two-word values: one 128-bit shift instead of the generic word-select form
template <>
__host__ __device__ __forceinline__ Kmer<2> km_shr<2>(const Kmer<2> &x, int s) {
    const unsigned __int128 v = (((unsigned __int128)x.w[0] << 64) | x.w[1]) >> s;
    Kmer<2> r;
    r.w[0] = (u64)(v >> 64);
    r.w[1] = (u64)v;
    return r;
}
// four-word values: branch on the word part of the shift (uniform within a kernel: it only depends on k), so that
// every word index below is a compile-time constant
template <>
__host__ __device__ __forceinline__ Kmer<4> km_shr<4>(const Kmer<4> &x, int s) {
    const int ws = s >> 6, bs = s & 63;
    Kmer<4> r;
#define GG_SHR4(i, lo, hi) r.w[i] = bs ? (((lo) >> bs) | ((hi) << (64 - bs))) : (lo)
    switch (ws) {
        case 0: GG_SHR4(3, x.w[3], x.w[2]); GG_SHR4(2, x.w[2], x.w[1]); GG_SHR4(1, x.w[1], x.w[0]); GG_SHR4(0, x.w[0], 0ULL); break;
        case 1: GG_SHR4(3, x.w[2], x.w[1]); GG_SHR4(2, x.w[1], x.w[0]); GG_SHR4(1, x.w[0], 0ULL); r.w[0] = 0; break;
        case 2: GG_SHR4(3, x.w[1], x.w[0]); GG_SHR4(2, x.w[0], 0ULL); r.w[1] = 0; r.w[0] = 0; break;
        default: GG_SHR4(3, x.w[0], 0ULL); r.w[2] = 0; r.w[1] = 0; r.w[0] = 0; break;
    }
#undef GG_SHR4
    return r;
}
// low 2k bits mask applied in place
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_mask(const Kmer<W> &x, int k) {
    Kmer<W> r;
    int bits = 2 * k;
#pragma unroll
    for (int i = W - 1; i >= 0; --i) {
        int b = bits - 64 * (W - 1 - i);  // bits available to word i
        u64 m = b >= 64 ? ~0ULL : (b > 0 ? ((1ULL << b) - 1) : 0ULL);
        r.w[i] = x.w[i] & m;
    }
    return r;
}
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_rc(const Kmer<W> &x, int k) {
    Kmer<W> t;
#pragma unroll
    for (int i = 0; i < W; ++i) t.w[W - 1 - i] = rev2_u64(~x.w[i]);
    return km_shr<W>(t, 64 * W - 2 * k);
}
// ((x << 2) | b) masked to 2k bits : forward neighbour (src/dbg.jl:30-34)
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_fw_nb(const Kmer<W> &x, u32 b, int k) {
    Kmer<W> r;
#pragma unroll
    for (int i = 0; i < W; ++i) r.w[i] = (x.w[i] << 2) | ((i + 1 < W) ? (x.w[i + 1] >> 62) : (u64)b);
    return km_mask<W>(r, k);
}
// (x >> 2) | (b << 2(k-1)) : backward neighbour (src/dbg.jl:36-42)
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_bw_nb(const Kmer<W> &x, u32 b, int k) {
    Kmer<W> r;
#pragma unroll
    for (int i = W - 1; i >= 0; --i) r.w[i] = (x.w[i] >> 2) | ((i > 0) ? (x.w[i - 1] << 62) : 0ULL);
    int pos = 2 * (k - 1);
#pragma unroll
    for (int i = 0; i < W; ++i)
        if (i == W - 1 - (pos >> 6)) r.w[i] |= ((u64)b) << (pos & 63);
    return r;
}
// 2-bit code of base j (0 = first base) of a k-mer
template <int W>
__host__ __device__ __forceinline__ u32 km_base(const Kmer<W> &x, int j, int k) {
    int pos = 2 * (k - 1 - j);
    u64 v = 0;
#pragma unroll
    for (int i = 0; i < W; ++i)
        if (i == W - 1 - (pos >> 6)) v = x.w[i];
    return (u32)((v >> (pos & 63)) & 3ULL);
}
template <int W>
__host__ __device__ __forceinline__ Kmer<W> km_load(const u64 *p, u64 i) {
    Kmer<W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.w[j] = p[i * W + j];
    return r;
}
template <int W>
__host__ __device__ __forceinline__ void km_store(u64 *p, u64 i, const Kmer<W> &x) {
#pragma unroll
    for (int j = 0; j < W; ++j) p[i * W + j] = x.w[j];
}
#ifdef __CUDACC__
// vectorised variants (keys are 8*W-byte aligned)
template <>
__device__ __forceinline__ Kmer<2> km_load<2>(const u64 *p, u64 i) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p + i * 2);
    Kmer<2> r; r.w[0] = v.x; r.w[1] = v.y; return r;
}
template <>
__device__ __forceinline__ void km_store<2>(u64 *p, u64 i, const Kmer<2> &x) {
    *reinterpret_cast<ulonglong2 *>(p + i * 2) = make_ulonglong2(x.w[0], x.w[1]);
}
template <>
__device__ __forceinline__ Kmer<4> km_load<4>(const u64 *p, u64 i) {
    const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(p + i * 4);
    ulonglong2 a = q[0], b = q[1];
    Kmer<4> r; r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y; return r;
}
template <>
__device__ __forceinline__ void km_store<4>(u64 *p, u64 i, const Kmer<4> &x) {
    ulonglong2 *q = reinterpret_cast<ulonglong2 *>(p + i * 4);
    q[0] = make_ulonglong2(x.w[0], x.w[1]);
    q[1] = make_ulonglong2(x.w[2], x.w[3]);
}
#endif

static inline int kmer_words(int k) { return k < 1 ? 0 : k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 0; }

#define DISPATCH_W(W_, ...)                              \
    switch (W_) {                                        \
        case 1: { constexpr int W = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int W = 2; __VA_ARGS__; } break; \
        case 4: { constexpr int W = 4; __VA_ARGS__; } break; \
        default: GG_THROW(GG_ERR_ARG, "unsupported word count %d", (int)(W_)); \
    }
