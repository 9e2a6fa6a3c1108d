// ggb200.cu -- C ABI of libggb200.so (see include/ggb200.h).  Single translation unit:
// the kernels live in the .cuh files included below.
#include "common.cuh"
#include "prims.cuh"
#include "pack.cuh"
#include "sort.cuh"
#include "msdcount.cuh"
#include "msdcountw.cuh"
#include "skm.cuh"
#include "dist.cuh"
#include "graph.cuh"
#include "graphdist.cuh"
#include "graphedit.cuh"
#include "umi.cuh"
#include "synth.cuh"
#include "gfa.cuh"

struct gg_counts {
    CtxRef ctx;
    int k = 0, W = 0;
    u64 n_unique = 0, n_instances = 0;
    u64 *d_keys = nullptr;   // n_unique * W, sorted ascending
    u32 *d_counts = nullptr; // n_unique
};

#define API_BEGIN(ctx_) \
    gg_ctx *_c = (ctx_); \
    if (!_c) return GG_ERR_ARG; \
    try {
#define API_END \
    } catch (const GgError &e) { \
        _c->err = e.msg; \
        cudaGetLastError(); \
        return e.code; \
    } catch (const std::bad_alloc &) { \
        _c->err = "host allocation failed"; \
        return GG_ERR_OOM; \
    } catch (...) { \
        _c->err = "unknown failure"; \
        return GG_ERR_CUDA; \
    } \
    return GG_OK;

static void timing_reset(gg_ctx *ctx) {
    for (int i = 0; i < GG_NSTAGE; ++i) { ctx->ev_used[i] = false; ctx->ms[i] = 0.f; }
    ctx->launches = 0;
    for (auto &r : ctx->prof) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->prof.clear();
}
static void profile_collect(gg_ctx *ctx) {
    if (ctx->prof.empty()) return;
    struct Acc { u64 n; double ms, bytes; };
    std::map<std::string, Acc> acc;
    std::vector<std::string> order;
    for (auto &r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto it = acc.find(r.name);
        if (it == acc.end()) { acc[r.name] = {1, ms, r.bytes}; order.push_back(r.name); }
        else { it->second.n++; it->second.ms += ms; it->second.bytes += r.bytes; }
        ctx->ev_pool.push_back(r.a);
        ctx->ev_pool.push_back(r.b);
    }
    ctx->prof.clear();
    ctx->prof_text.clear();
    char line[512];
    for (auto &n : order) {
        snprintf(line, sizeof(line), "%s\t%llu\t%.6f\t%.0f\n", n.c_str(), (unsigned long long)acc[n].n, acc[n].ms, acc[n].bytes);
        ctx->prof_text += line;
    }
}
static void timing_collect(gg_ctx *ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    profile_collect(ctx);
    for (int i = 0; i < GG_NSTAGE; ++i)
        if (ctx->ev_used[i]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]) == cudaSuccess) ctx->ms[i] = ms;
        }
}

extern "C" {

int gg_version(void) { return 100; }
int gg_kmer_words(int k) { return kmer_words(k); }

int gg_ctx_create(int device, gg_ctx **out) {
    if (!out) return GG_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return GG_ERR_CUDA;  // no CPU fallback
    }
    gg_ctx *ctx = new (std::nothrow) gg_ctx();
    if (!ctx) return GG_ERR_OOM;
    ctx->device = device;
    try {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        ctx->num_sms = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->aux_ev, cudaEventDisableTiming));
        for (int i = 0; i < 2 * GG_NSTAGE; ++i) CK(cudaEventCreate(&ctx->ev[i]));
        for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&ctx->user_ev[i]));
        CK(cudaHostAlloc((void **)&ctx->h_scalar, 64 * sizeof(u64), cudaHostAllocDefault));
        timing_reset(ctx);
    } catch (const GgError &) {
        delete ctx;
        cudaGetLastError();
        return GG_ERR_CUDA;
    }
    *out = ctx;
    return GG_OK;
}
void gg_ctx_destroy(gg_ctx *ctx) {
    if (!ctx) return;
    if (ctx->live_handles > 0) { ctx->dying = true; return; }   // the last handle to go finishes the job (CtxRef::release)
    gg_ctx_destroy_now(ctx);
}
}  // extern "C"
static void gg_ctx_destroy_now(gg_ctx *ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2 * GG_NSTAGE; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->user_ev[i]);
    for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    cudaFreeHost(ctx->h_scalar);
    if (ctx->arena) {  // handles still alive keep their blocks listed in `live`; everything is returned to the driver here
        for (auto &kv : ctx->arena->live) cudaFree(kv.first);
        delete ctx->arena;
    }
    cudaStreamDestroy(ctx->stream);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->aux_ev) cudaEventDestroy(ctx->aux_ev);
    for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
    delete ctx;
}
extern "C" {
const char *gg_last_error(gg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gg_host_alloc(gg_ctx *ctx, uint64_t bytes, void **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    CK(cudaSetDevice(_c->device));
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    API_END
}
void gg_host_free(gg_ctx *ctx, void *p) { (void)ctx; if (p) cudaFreeHost(p); }
int gg_dev_alloc(gg_ctx *ctx, uint64_t bytes, void **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    CK(cudaSetDevice(_c->device));
    CK(cudaMalloc(out, bytes ? bytes : 16));
    API_END
}
void gg_dev_free(gg_ctx *ctx, void *p) { (void)ctx; if (p) cudaFree(p); }
int gg_memcpy_h2d(gg_ctx *ctx, void *dst, const void *src, uint64_t bytes) {
    API_BEGIN(ctx)
    CK(cudaSetDevice(_c->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
int gg_memcpy_d2h(gg_ctx *ctx, void *dst, const void *src, uint64_t bytes) {
    API_BEGIN(ctx)
    CK(cudaSetDevice(_c->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
int gg_ctx_set_profile(gg_ctx *ctx, int enabled) {
    if (!ctx) return GG_ERR_ARG;
    ctx->profile = enabled != 0;
    return GG_OK;
}
const char *gg_ctx_kernel_profile(gg_ctx *ctx) { return ctx ? ctx->prof_text.c_str() : ""; }
int gg_ctx_event_record(gg_ctx *ctx, int slot) {
    API_BEGIN(ctx)
    if (slot < 0 || slot >= 8) GG_THROW(GG_ERR_ARG, "event slot out of range");
    CK(cudaSetDevice(_c->device));
    CK(cudaEventRecord(_c->user_ev[slot], _c->stream));
    API_END
}
int gg_ctx_event_elapsed_ms(gg_ctx *ctx, int slot_a, int slot_b, float *ms) {
    API_BEGIN(ctx)
    if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !ms) GG_THROW(GG_ERR_ARG, "bad event slot");
    CK(cudaSetDevice(_c->device));
    CK(cudaEventSynchronize(_c->user_ev[slot_b]));
    CK(cudaEventElapsedTime(ms, _c->user_ev[slot_a], _c->user_ev[slot_b]));
    API_END
}
int gg_ctx_sync(gg_ctx *ctx) {
    API_BEGIN(ctx)
    CK(cudaSetDevice(_c->device));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
int gg_ctx_last_timings(gg_ctx *ctx, float ms[9], uint64_t *kernel_launches) {
    if (!ctx) return GG_ERR_ARG;
    if (ms) for (int i = 0; i < GG_NSTAGE; ++i) ms[i] = ctx->ms[i];
    if (kernel_launches) *kernel_launches = ctx->launches;
    return GG_OK;
}

}  // extern "C"

// ------------------------------------ pipelines -------------------------------------------
static void check_k(int k) {
    if (kmer_words(k) == 0) GG_THROW(GG_ERR_ARG, "k = %d is outside 1..%d", k, GG_MAX_K);
}

// device reads -> all k-mers in read order (device), returns N
template <int W>
static u64 extract_all(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, int k, int canonical, DBuf<u64> &keys, u64 extra_words) {
    stage_begin(ctx, ST_PACK);
    PackedReads pr;
    pack_reads(ctx, d_seq, d_offs, nreads, nbytes, pr);
    DBuf<u32> valid;
    window_valid_mask(ctx, pr, k, valid);
    stage_end(ctx, ST_PACK);
    stage_begin(ctx, ST_EXTRACT);
    ExtractF<W> ef{valid.p, pr.packed.p, nullptr, k, canonical};
    Compactor<ExtractF<W>> cp(ctx, nbytes, "extract");
    u64 n = cp.count(ef);
    keys.alloc(ctx, n * W + extra_words);
    ef.out = keys.p;
    cp.emit(ef);
    stage_end(ctx, ST_EXTRACT);
    return n;
}

static inline u32 eff_min_freq(u64 min_freq) {
    u32 mf = min_freq > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)min_freq;
    return mf < 1 ? 1 : mf;
}
// dbg!(sg, counted_kmers, min_freq) compares freq(x), the UInt8-saturated count of a MerCount (KmerAnalysis, recalled):
// min(count, 255) >= min_freq.  For min_freq <= 255 that equals count >= min_freq; above it nothing is kept.
static inline u64 dbg_min_freq(u64 min_freq) { return min_freq > 255 ? 0xFFFFFFFFull : min_freq; }
// min_freq > 1 drops rarer k-mers inside the counting kernels (fused filter of the reads -> graph
// pipeline); *filtered tells the caller whether that happened.
template <int W>
static gg_counts *count_kmers_dev_impl(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, int k, int canonical,
                                       u64 min_freq = 1, bool *filtered = nullptr, int enc = 0) {
    gg_counts *c = new gg_counts();
    c->ctx = ctx; c->k = k; c->W = W;
    if (filtered) *filtered = false;
    try {
        if (W == 1 && skm_supported(k, canonical) && !getenv("GG_FORCE_LSD") && !getenv("GG_FORCE_MSD")) {
            stage_begin(ctx, ST_PACK);
            PackedReads pr;
            pack_input(ctx, d_seq, d_offs, nreads, nbytes, enc, pr);
            DBuf<u32> valid;
            window_valid_mask(ctx, pr, k, valid);
            stage_end(ctx, ST_PACK);
            DBuf<u64> ukeys;
            DBuf<u32> counts;
            c->n_unique = skm_count(ctx, pr, valid.p, k, canonical, eff_min_freq(min_freq), ukeys, counts, &c->n_instances);
            c->d_keys = ukeys.take();
            c->d_counts = counts.take();
            if (filtered) *filtered = true;
            return c;
        }
        if (W == 1 && msd_supported(k, canonical) && !getenv("GG_FORCE_LSD")) {
            stage_begin(ctx, ST_PACK);
            PackedReads pr;
            pack_input(ctx, d_seq, d_offs, nreads, nbytes, enc, pr);
            DBuf<u32> valid;
            window_valid_mask(ctx, pr, k, valid);
            stage_end(ctx, ST_PACK);
            DBuf<u64> ukeys;
            DBuf<u32> counts;
            u32 mf = min_freq > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)min_freq;
            c->n_unique = msd_count(ctx, pr, valid.p, k, canonical, mf < 1 ? 1 : mf, ukeys, counts, &c->n_instances);
            c->d_keys = ukeys.take();
            c->d_counts = counts.take();
            if (filtered) *filtered = true;
            return c;
        }
        if constexpr (W > 1) {
            if (!getenv("GG_FORCE_LSD")) {
                stage_begin(ctx, ST_PACK);
                PackedReads pr;
                pack_input(ctx, d_seq, d_offs, nreads, nbytes, enc, pr);
                DBuf<u32> valid;
                window_valid_mask(ctx, pr, k, valid);
                stage_end(ctx, ST_PACK);
                DBuf<u64> ukeys;
                DBuf<u32> counts;
                u32 mf = min_freq > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)min_freq;
                c->n_unique = msdw_count<W>(ctx, pr, valid.p, k, canonical, mf < 1 ? 1 : mf, ukeys, counts, &c->n_instances);
                c->d_keys = ukeys.take();
                c->d_counts = counts.take();
                if (filtered) *filtered = true;
                return c;
            }
        }
        if (enc != 0) GG_THROW(GG_ERR_ARG, "4-bit input is not available together with GG_FORCE_LSD");
        DBuf<u64> keys;
        u64 n = extract_all<W>(ctx, d_seq, d_offs, nreads, nbytes, k, canonical, keys, 0);
        c->n_instances = n;
        stage_begin(ctx, ST_SORT);
        DBuf<u64> alt(ctx, n * W);
        u64 *a = keys.p, *b = alt.p;
        std::vector<BitRange> ranges = {{0, 2 * k}};
        radix_sort<W>(ctx, &a, &b, nullptr, nullptr, n, ranges);
        stage_end(ctx, ST_SORT);
        stage_begin(ctx, ST_COUNT);
        DBuf<u64> ukeys;
        DBuf<u32> counts;
        c->n_unique = run_length_count<W>(ctx, a, n, ukeys, counts);
        c->d_keys = ukeys.take();
        c->d_counts = counts.take();
        stage_end(ctx, ST_COUNT);
        return c;
    } catch (...) {
        dfree(ctx, c->d_keys); dfree(ctx, c->d_counts);
        delete c;
        throw;
    }
}

template <int W>
static void filter_impl(gg_counts *c, u64 min_freq, int u8sat) {
    gg_ctx *ctx = c->ctx;
    stage_begin(ctx, ST_FILTER);
    FilterF<W> f{c->d_keys, c->d_counts, nullptr, nullptr, min_freq, u8sat};
    Compactor<FilterF<W>> cp(ctx, c->n_unique, "filter");
    u64 nk = cp.count(f);
    DBuf<u64> okeys(ctx, nk * W);
    DBuf<u32> ocounts(ctx, nk);
    f.okeys = okeys.p; f.ocounts = ocounts.p;
    cp.emit(f);
    dfree(ctx, c->d_keys); dfree(ctx, c->d_counts);
    c->d_keys = okeys.take();
    c->d_counts = ocounts.take();
    c->n_unique = nk;
    stage_end(ctx, ST_FILTER);
}

template <int W>
__global__ void is_sorted_kernel(const u64 *__restrict__ keys, u64 n, u32 *__restrict__ flag) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    if (km_cmp<W>(km_load<W>(keys, i), km_load<W>(keys, i + 1)) > 0) *flag = 1;
}

// read_offsets (host): offs[0] == 0 and non-decreasing (the reference's per-read sequences cannot overlap or leave gaps)
static void check_offsets_host(gg_ctx *ctx, const u64 *offs, u64 nreads) {
    bool ok = offs[0] == 0;
    u64 bad = 0;
    for (u64 i = 0; ok && i < nreads; ++i)
        if (offs[i] > offs[i + 1]) { ok = false; bad = i; }
    if (!ok) {
        cudaStreamSynchronize(ctx->aux);  // the buffers go back to the allocator when the caller's frame unwinds
        cudaStreamSynchronize(ctx->stream);
        if (offs[0] != 0) GG_THROW(GG_ERR_ARG, "read_offsets[0] must be 0 (it is %llu)", (unsigned long long)offs[0]);
        GG_THROW(GG_ERR_ARG, "read_offsets must be non-decreasing (read %llu)", (unsigned long long)bad);
    }
}
// read_offsets (device, the *_dev entry points): offs[0] == 0, non-decreasing, offs[nreads] == nbytes
__global__ void check_offsets_kernel(const u64 *__restrict__ offs, u64 nreads, u64 nbytes, u32 *__restrict__ bad) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nreads) return;
    if (i == 0 && offs[0] != 0) *bad = 1u;
    if (i == nreads) { if (offs[nreads] != nbytes) *bad = 1u; }
    else if (offs[i] > offs[i + 1]) *bad = 1u;
}
static void check_offsets_dev(gg_ctx *ctx, const u64 *d_offs, u64 nreads, u64 nbytes) {
    if (!d_offs) GG_THROW(GG_ERR_ARG, "read_offsets is null");
    DBuf<u32> bad(ctx, 1);
    dfill(ctx, bad.p, sizeof(u32), 0u);
    LAUNCH(ctx, check_offsets_kernel, (unsigned)ceil_div(nreads + 1, 256), 256, 0, d_offs, nreads, nbytes, bad.p);
    if (read_scalar<u32>(ctx, bad.p)) GG_THROW(GG_ERR_ARG, "device read_offsets must start at 0, be non-decreasing and end at nbytes");
}
// stage device input from host buffers (padded so that 128-bit loads stay in bounds)
struct StagedReads {
    DBuf<u8> seq;
    DBuf<u64> offs;
    u64 nbytes = 0;
};
static void stage_reads(gg_ctx *ctx, const u8 *seq, const u64 *offs, u64 nreads, StagedReads &s) {
    if (!offs) GG_THROW(GG_ERR_ARG, "read_offsets is null");
    s.nbytes = offs[nreads];
    if (s.nbytes && !seq) GG_THROW(GG_ERR_ARG, "seq_bytes is null");
    s.seq.alloc(ctx, s.nbytes + 32);
    s.offs.alloc(ctx, nreads + 1);
    // the bytes go first (asynchronous from pinned memory); the offsets travel on the second copy stream and the
    // host checks them meanwhile, so neither the check nor the small copy adds to the transfer time
    CK(cudaEventRecord(ctx->aux_ev, ctx->stream));        // the blocks may still be in use by earlier work of this context
    CK(cudaStreamWaitEvent(ctx->aux, ctx->aux_ev, 0));
    if (s.nbytes) CK(cudaMemcpyAsync(s.seq.p, seq, s.nbytes, cudaMemcpyHostToDevice, ctx->aux));
    CK(cudaMemcpyAsync(s.offs.p, offs, (nreads + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->aux));
    CK(cudaEventRecord(ctx->aux_ev, ctx->aux));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev, 0));
    check_offsets_host(ctx, offs, nreads);
}

// Host buffers -> counts with the transfer pipelined against the front of the path (k <= 32 MSD route):
// the read bytes travel in chunks on the copy stream; as soon as chunk c has landed it is packed, and the
// window masks + pass A (histogram, distinct estimate) of chunk c - 1 run (they look one word ahead, hence the
// lag).  When the last byte arrives only pass B onwards remains.
static u64 h2d_chunk_bytes() {  // a multiple of 8192 (whole packed words and pack-thread pairs); GG_H2D_CHUNK_KB for tests
    if (const char *e = getenv("GG_H2D_CHUNK_KB")) {
        u64 kb = (u64)atoll(e);
        if (kb >= 8) return (kb / 8) * 8192;
    }
    return 16ULL << 20;
}
// Front of the pipelined host path, shared by the two counting routes: copies on the copy stream, chunk c packed as
// soon as it has landed, window masks of chunk c - 1 right after (they look one word ahead), then
// on_range(ci, w0, w1) for the packed words of chunk ci = c - 1.
template <class OnRange>
static void host_pipeline_front(gg_ctx *ctx, const u8 *seq, const u64 *offs, u64 nreads, u64 nbytes, int k, DBuf<u8> &dseq, DBuf<u64> &doffs,
                                PackedReads &pr, DBuf<u32> &valid, u64 chunk_bytes, OnRange on_range) {
    const u64 nch = ceil_div(nbytes, chunk_bytes);
    while (ctx->chunk_ev.size() < nch) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
    }
    // EVERY copy goes to the copy stream (offsets first, then the chunks) and the main stream only waits on events.
    // Measured: one host-to-device copy issued on the main stream itself is enough to hold all its later kernels
    // back until the copy stream has drained -- no overlap at all.
    CK(cudaEventRecord(ctx->aux_ev, ctx->stream));      // the copy stream starts after the earlier work of this context
    CK(cudaStreamWaitEvent(ctx->aux, ctx->aux_ev, 0));
    CK(cudaMemcpyAsync(doffs.p, offs, (nreads + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->aux));
    CK(cudaEventRecord(ctx->aux_ev, ctx->aux));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev, 0));
    for (u64 ch = 0; ch < nch; ++ch) {
        const u64 b0 = ch * chunk_bytes, b1 = b0 + chunk_bytes < nbytes ? b0 + chunk_bytes : nbytes;
        CK(cudaMemcpyAsync(dseq.p + b0, seq + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->aux));
        CK(cudaEventRecord(ctx->chunk_ev[ch], ctx->aux));
    }
    check_offsets_host(ctx, offs, nreads);  // while the bytes are in flight
    stage_begin(ctx, ST_PACK);
    pack_prepare(ctx, doffs.p, nreads, nbytes, pr);
    valid.alloc(ctx, pr.nwords ? pr.nwords : 1);
    stage_begin(ctx, ST_EXTRACT);
    for (u64 ch = 0; ch <= nch; ++ch) {
        if (ch < nch) {
            const u64 b0 = ch * chunk_bytes, b1 = b0 + chunk_bytes < nbytes ? b0 + chunk_bytes : nbytes;
            CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[ch], 0));
            pack_range(ctx, dseq.p, nbytes, pr, b0 / 16, ceil_div(b1, 16));
        }
        if (ch > 0) {  // chunk ch - 1: its look-ahead word is packed now (or it is the last chunk)
            const u64 w0 = (ch - 1) * (chunk_bytes / 32), w1 = ch < nch ? ch * (chunk_bytes / 32) : pr.nwords;
            if (w1 > w0) {
                LAUNCH(ctx, valid32_kernel, (unsigned)ceil_div(w1 - w0, 256), 256, 0, pr.bad.p, pr.brk.p, w0, w1, k, valid.p);
                on_range(ch - 1, w0, w1);
            }
        }
    }
}

// Host buffers -> counts with the transfer pipelined against the front of the path (one-word keys):
// the read bytes travel in chunks on the copy stream; as soon as chunk c has landed it is packed, the window masks of
// chunk c - 1 follow, and the chunk is scattered at once (super-k-mer route: level-0 capacities from the first chunk's
// exact record histogram; MSD route: pass A + speculative pass B).  When the last byte arrives only the tail remains.
static gg_counts *count_kmers_host_pipelined(gg_ctx *ctx, const u8 *seq, const u64 *offs, u64 nreads, int k, int canonical, u64 min_freq,
                                             bool *filtered) {
    if (!offs) GG_THROW(GG_ERR_ARG, "read_offsets is null");
    const u64 nbytes = offs[nreads];
    if (nbytes && !seq) GG_THROW(GG_ERR_ARG, "seq_bytes is null");
    gg_counts *c = new gg_counts();
    c->ctx = ctx; c->k = k; c->W = 1;
    if (filtered) *filtered = true;
    try {
        DBuf<u8> dseq(ctx, nbytes + 32);
        DBuf<u64> doffs(ctx, nreads + 1);
        const u64 GG_H2D_CHUNK = h2d_chunk_bytes();
        const u64 nch = ceil_div(nbytes, GG_H2D_CHUNK);
        PackedReads pr;
        DBuf<u32> valid;
        DBuf<u64> ukeys;
        DBuf<u32> counts;
        const u32 mf = eff_min_freq(min_freq);
        if (skm_supported(k, canonical) && !getenv("GG_FORCE_MSD")) {
            const SkmTune tune = skm_tune();
            SkmL0 L;
            const bool speculate = nch >= 3 && !getenv("GG_NO_SPEC");
            bool scattered = false;
            host_pipeline_front(ctx, seq, offs, nreads, nbytes, k, dseq, doffs, pr, valid, GG_H2D_CHUNK, [&](u64 ci, u64 w0, u64 w1) {
                if (!speculate) return;
                if (ci == 0) {  // geometry and capacities from the first chunk's windows and records, scaled to the whole input
                    const double scale = (double)pr.nwords / (double)(w1 - w0);
                    L.g = skm_geom((u64)((double)skm_count_windows(ctx, valid.p, w0, w1) * scale), k, canonical, tune);
                    L.alloc_sample(ctx);
                    skm_scan_launch<0, false>(ctx, pr, valid.p, w0, w1, 1, L.g, L.hist0.p, nullptr, nullptr, nullptr, nullptr, nullptr, L.totals.p, nullptr,
                                              nullptr);
                    skm_l0_prepare(ctx, L, scale, false, tune);
                    scattered = true;
                }
                skm_scan_launch<1, false>(ctx, pr, valid.p, w0, w1, 1, L.g, nullptr, L.cur0.p, L.base0.p, L.cap0.p, L.rec0.p, L.hist1.p, L.totals.p,
                                          L.ovf.p, nullptr);
            });
            stage_end(ctx, ST_PACK);
            bool done = false;
            if (scattered) {
                CK(cudaMemcpyAsync(ctx->h_scalar, L.totals.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(ctx->h_scalar + 2, L.ovf.p, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                if ((u32)ctx->h_scalar[2] == 0 && !getenv("GG_PIPE_TEST_OVERFLOW")) {
                    const u64 n_inst = ctx->h_scalar[0], n_rec = ctx->h_scalar[1];
                    c->n_instances = n_inst;
                    stage_end(ctx, ST_EXTRACT);
                    c->n_unique = skm_finish(ctx, L.rec0.p, L.base0.p, L.cur0.p, SKM_CUR_STRIDE, L.g.nb0, 1, L.cap_total, L.g, L.hist1.p, n_inst, n_rec,
                                             mf, tune, ukeys, counts);
                    done = true;
                } else L.rec0.release();  // a bin outgrew the share the first chunk predicted (reads sorted by content, say): exact repeat
            }
            if (!done) c->n_unique = skm_count(ctx, pr, valid.p, k, canonical, mf, ukeys, counts, &c->n_instances);
            c->d_keys = ukeys.take();
            c->d_counts = counts.take();
            return c;
        }
        DBuf<u32> hist;
        DBuf<u8> est;
        msd_pass_a_alloc(ctx, k, hist, est);
        const int est_shift = msd_est_shift(nbytes);
        const u32 nbh = 1u << msd_hist_bits(k);
        const u32 nrows = 3u * (u32)ctx->num_sms;
        DBuf<u32> rows(ctx, (u64)nrows * nbh);  // per-CTA histogram rows of the chunked pass A
        dfill(ctx, rows.p, (u64)nrows * nbh * sizeof(u32), 0u);
        // Speculative pass B: with at least three chunks, the first chunk's histogram predicts the bucket shares (reads
        // arrive in no particular order), every bucket gets its share of nbytes * 1.1 key slots (the number of windows
        // is below nbytes), and chunk c - 1 is scattered as soon as its masks exist.  An overflowing bucket (input
        // sorted by position, say) only costs a repeat of pass B with exact offsets.
        MsdSpec spec;
        const int spec_b0 = msd_tune().b0_max;
        const bool speculate = nch >= 3 && spec_b0 <= msd_hist_bits(k) && spec_b0 <= MSD_B0_MAX && !getenv("GG_NO_SPEC") &&
                               nbytes + nbytes / 10 + 4096 < 0xFFFFFFFFull;  // the speculative bucket offsets are 32-bit
        if (speculate) {
            const u32 nb0 = 1u << spec_b0;
            spec.b0 = spec_b0;
            spec.A.alloc(ctx, nbytes + nbytes / 10 + 4096);
            spec.start.alloc(ctx, nb0 + 1); spec.cursor.alloc(ctx, nb0); spec.bend.alloc(ctx, nb0); spec.ovf.alloc(ctx, 1);
            dfill(ctx, spec.ovf.p, sizeof(u32), 0u);
        }
        host_pipeline_front(ctx, seq, offs, nreads, nbytes, k, dseq, doffs, pr, valid, GG_H2D_CHUNK, [&](u64 ci, u64 w0, u64 w1) {
            msd_pass_a_range(ctx, pr, valid.p, k, canonical, est_shift, hist.p, est.p, w0, w1, rows.p);
            if (speculate) {
                if (ci == 0) {  // capacities from the first chunk's histogram
                    DBuf<u32> h0(ctx, nbh);
                    dfill(ctx, h0.p, (u64)nbh * sizeof(u32), 0u);
                    LAUNCH(ctx, msd_hist_rows_reduce_kernel, (unsigned)ceil_div(nbh, 256), 256, 0, rows.p, nrows, nbh, h0.p);
                    LAUNCH(ctx, msd_spec_caps_kernel, 1, 1024, 0, h0.p, msd_hist_bits(k), spec_b0, (u64)spec.A.n, spec.start.p, spec.cursor.p, spec.bend.p);
                    spec.used = true;
                }
                msd_pass_b_spec(ctx, pr, valid.p, k, canonical, spec_b0, spec.cursor.p, spec.bend.p, spec.ovf.p, spec.A.p, w0, w1);
            }
        });
        LAUNCH(ctx, msd_hist_rows_reduce_kernel, (unsigned)ceil_div(nbh, 256), 256, 0, rows.p, nrows, nbh, hist.p);
        rows.release();
        stage_end(ctx, ST_PACK);
        if (pr.nwords == 0) { ukeys.alloc(ctx, 1); counts.alloc(ctx, 1); stage_end(ctx, ST_EXTRACT); }
        else c->n_unique = msd_count_after_a(ctx, pr, valid.p, k, canonical, mf, hist, est, est_shift, ukeys, counts, &c->n_instances,
                                             speculate ? &spec : nullptr);
        c->d_keys = ukeys.take();
        c->d_counts = counts.take();
        return c;
    } catch (...) {
        cudaStreamSynchronize(ctx->aux);  // no copy may still be writing into a block that goes back to the allocator
        dfree(ctx, c->d_keys); dfree(ctx, c->d_counts);
        delete c;
        throw;
    }
}
static inline bool host_pipeline_ok(int k, int canonical) {
    return kmer_words(k) == 1 && msd_supported(k, canonical) && !getenv("GG_FORCE_LSD") && !getenv("GG_NO_PIPELINE");
}

extern "C" {

int gg_count_kmers_dev(gg_ctx *ctx, const uint8_t *d_seq, const uint64_t *d_offs, uint64_t nreads, uint64_t nbytes, int k,
                       int canonical, gg_counts **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    check_offsets_dev(_c, d_offs, nreads, nbytes);
    stage_begin(_c, ST_TOTAL);
    DISPATCH_W(kmer_words(k), *out = count_kmers_dev_impl<W>(_c, d_seq, d_offs, nreads, nbytes, k, canonical));
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}
int gg_count_kmers(gg_ctx *ctx, const uint8_t *seq, const uint64_t *offs, uint64_t nreads, int k, int canonical, gg_counts **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    if (host_pipeline_ok(k, canonical)) {
        stage_begin(_c, ST_TOTAL);
        *out = count_kmers_host_pipelined(_c, seq, offs, nreads, k, canonical, 1, nullptr);
        stage_end(_c, ST_TOTAL);
        timing_collect(_c);
        return GG_OK;
    }
    StagedReads s;
    stage_reads(_c, seq, offs, nreads, s);
    stage_begin(_c, ST_TOTAL);
    DISPATCH_W(kmer_words(k), *out = count_kmers_dev_impl<W>(_c, s.seq.p, s.offs.p, nreads, s.nbytes, k, canonical));
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}
int gg_counts_from_kmers(gg_ctx *ctx, const uint64_t *words, const uint32_t *counts, uint64_t n, int k, gg_counts **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    if (n && !words) GG_THROW(GG_ERR_ARG, "kmer_words is null");
    CK(cudaSetDevice(_c->device));
    int W = kmer_words(k);
    gg_counts *c = new gg_counts();
    c->ctx = _c; c->k = k; c->W = W; c->n_unique = n; c->n_instances = n;
    try {
        DBuf<u64> keys(_c, n * W), alt(_c, n * W), pay, pay_alt;
        DBuf<u32> cnt(_c, n);
        if (n) CK(cudaMemcpyAsync(keys.p, words, n * W * sizeof(u64), cudaMemcpyHostToDevice, _c->stream));
        if (n && counts) CK(cudaMemcpyAsync(cnt.p, counts, n * sizeof(u32), cudaMemcpyHostToDevice, _c->stream));
        else if (n) LAUNCH(_c, fill_counts_kernel, (unsigned)ceil_div(n, 256), 256, 0, cnt.p, n, 1u);
        // issorted(kmerlist) || sort!(kmerlist)  (src/dbg.jl:98-100)
        DBuf<u32> flag(_c, 1);
        CK(cudaMemsetAsync(flag.p, 0, sizeof(u32), _c->stream));
        if (n > 1) {
            DISPATCH_W(W, LAUNCH(_c, is_sorted_kernel<W>, (unsigned)ceil_div(n, 256), 256, 0, keys.p, n, flag.p));
        }
        u64 *a = keys.p, *b = alt.p;
        if (n > 1 && read_scalar<u32>(_c, flag.p)) {
            if (counts) GG_THROW(GG_ERR_ARG, "k-mers with counts must be sorted ascending");
            std::vector<BitRange> ranges = {{0, 2 * k}};
            DISPATCH_W(W, radix_sort<W>(_c, &a, &b, nullptr, nullptr, n, ranges));
        }
        if (a == keys.p) { c->d_keys = keys.take(); } else { c->d_keys = alt.take(); }
        c->d_counts = cnt.take();
        CK(cudaStreamSynchronize(_c->stream));
    } catch (...) {
        delete c;
        throw;
    }
    *out = c;
    API_END
}
int gg_counts_sizes(gg_counts *c, uint64_t *n_unique, uint64_t *n_instances, int *k, int *words) {
    if (!c) return GG_ERR_ARG;
    if (n_unique) *n_unique = c->n_unique;
    if (n_instances) *n_instances = c->n_instances;
    if (k) *k = c->k;
    if (words) *words = c->W;
    return GG_OK;
}
int gg_counts_filter(gg_counts *c, uint64_t min_freq, int u8_saturated) {
    if (!c) return GG_ERR_ARG;
    API_BEGIN(c->ctx)
    CK(cudaSetDevice(_c->device));
    DISPATCH_W(c->W, filter_impl<W>(c, min_freq, u8_saturated));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
__global__ void sat_u8_kernel(const u32 *__restrict__ in, u64 n, u8 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u8)(in[i] > 255u ? 255u : in[i]);
}
int gg_counts_export(gg_counts *c, uint64_t *kmer_words_out, uint8_t *counts_u8, uint32_t *counts_u32) {
    if (!c) return GG_ERR_ARG;
    API_BEGIN(c->ctx)
    CK(cudaSetDevice(_c->device));
    u64 n = c->n_unique;
    if (n && kmer_words_out) CK(cudaMemcpyAsync(kmer_words_out, c->d_keys, n * c->W * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    if (n && counts_u32) CK(cudaMemcpyAsync(counts_u32, c->d_counts, n * sizeof(u32), cudaMemcpyDeviceToHost, _c->stream));
    if (n && counts_u8) {
        DBuf<u8> t(_c, n);
        LAUNCH(_c, sat_u8_kernel, (unsigned)ceil_div(n, 256), 256, 0, c->d_counts, n, t.p);
        CK(cudaMemcpyAsync(counts_u8, t.p, n, cudaMemcpyDeviceToHost, _c->stream));
        CK(cudaStreamSynchronize(_c->stream));
    }
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
int gg_counts_spectrum(gg_counts *c, uint64_t hist[256]) {
    if (!c || !hist) return GG_ERR_ARG;
    API_BEGIN(c->ctx)
    CK(cudaSetDevice(_c->device));
    DBuf<u64> h(_c, 256);
    CK(cudaMemsetAsync(h.p, 0, 256 * sizeof(u64), _c->stream));
    if (c->n_unique) {
        u64 blocks = ceil_div(c->n_unique, 256 * 16);
        if (blocks > (u64)_c->num_sms * 8) blocks = (u64)_c->num_sms * 8;
        LAUNCH(_c, spectrum_kernel, (unsigned)blocks, 256, 0, c->d_counts, c->n_unique, (unsigned long long *)h.p);
    }
    CK(cudaMemcpyAsync(hist, h.p, 256 * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
}  // extern "C"
// spectra-cn (docs/src/man/assembly.md:221-228: spectra(reads_kcount, graph_kcount)): joint histogram of the UInt8-saturated
// counts of every k-mer in two tables; a k-mer missing from one table counts 0 there.  hist[a * 256 + b].
// A_ROWS: every k-mer of table A, looked up in B.  Otherwise: only the k-mers of A that B lacks (their row is 0).
template <int W, bool A_ROWS>
__global__ void __launch_bounds__(256) spectrum_cn_kernel(const u64 *__restrict__ ka, const u32 *__restrict__ ca, u64 na, const u64 *__restrict__ kb,
                                                          const u32 *__restrict__ cb, const u32 *__restrict__ startb, int k, int pbits,
                                                          unsigned long long *__restrict__ hist) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 bin = 0xFFFFFFFFu;
    if (i < na) {
        const u32 idx = find_key<W>(kb, startb, k, pbits, km_load<W>(ka, i));
        const u32 mine = ca[i] > 255u ? 255u : ca[i];
        if (A_ROWS) bin = mine * 256u + (idx == V_NONE ? 0u : (cb[idx] > 255u ? 255u : cb[idx]));
        else if (idx == V_NONE) bin = mine;
    }
    // the mass sits in a few bins: one atomic per warp and bin
    const u32 peers = __match_any_sync(FULL_MASK, bin);
    if (bin != 0xFFFFFFFFu && lane_id() == (u32)__ffs((int)peers) - 1u) atomicAdd(&hist[bin], (unsigned long long)__popc(peers));
}
template <int W>
static void spectrum_cn_impl(gg_ctx *ctx, gg_counts *a, gg_counts *b, u64 *d_hist) {
    auto index = [&](gg_counts *t, DBuf<u32> &start, int &pbits) {
        pbits = t->n_unique ? ilog2_ceil(t->n_unique) - 1 : 0;
        if (pbits > 24) pbits = 24;
        if (pbits > 2 * t->k) pbits = 2 * t->k;
        if (pbits < 0) pbits = 0;
        start.alloc(ctx, (1ULL << pbits) + 1);
        LAUNCH(ctx, prefix_index_kernel<W>, (unsigned)ceil_div(t->n_unique + 1, 256), 256, 0, t->d_keys, t->n_unique, t->k, pbits, start.p);
    };
    DBuf<u32> sa, sb;
    int pa = 0, pb = 0;
    index(a, sa, pa);
    index(b, sb, pb);
    if (a->n_unique)
        LAUNCH(ctx, (spectrum_cn_kernel<W, true>), (unsigned)ceil_div(a->n_unique, 256), 256, 0, a->d_keys, a->d_counts, a->n_unique, b->d_keys, b->d_counts,
               sb.p, a->k, pb, (unsigned long long *)d_hist);
    if (b->n_unique)
        LAUNCH(ctx, (spectrum_cn_kernel<W, false>), (unsigned)ceil_div(b->n_unique, 256), 256, 0, b->d_keys, b->d_counts, b->n_unique, a->d_keys, a->d_counts,
               sa.p, a->k, pa, (unsigned long long *)d_hist);
    CK(cudaStreamSynchronize(ctx->stream));
}
extern "C" {
int gg_counts_spectrum_cn(gg_counts *a, gg_counts *b, uint64_t hist[65536]) {
    if (!a || !b || !hist) return GG_ERR_ARG;
    API_BEGIN(a->ctx)
    if ((gg_ctx *)b->ctx != _c || a->k != b->k) GG_THROW(GG_ERR_ARG, "the two count tables must share the context and k");
    CK(cudaSetDevice(_c->device));
    DBuf<u64> h(_c, 65536);
    CK(cudaMemsetAsync(h.p, 0, 65536 * sizeof(u64), _c->stream));
    DISPATCH_W(a->W, spectrum_cn_impl<W>(_c, a, b, h.p));
    CK(cudaMemcpyAsync(hist, h.p, 65536 * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
void gg_counts_free(gg_counts *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    dfree(c->ctx, c->d_keys);
    dfree(c->ctx, c->d_counts);
    delete c;
}

int gg_extract_kmers(gg_ctx *ctx, const uint8_t *seq, const uint64_t *offs, uint64_t nreads, int k, int canonical,
                     uint64_t *out_words, uint64_t *n_kmers) {
    API_BEGIN(ctx)
    if (!n_kmers) GG_THROW(GG_ERR_ARG, "n_kmers is null");
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    StagedReads s;
    stage_reads(_c, seq, offs, nreads, s);
    int W = kmer_words(k);
    DBuf<u64> keys;
    u64 n = 0;
    DISPATCH_W(W, n = extract_all<W>(_c, s.seq.p, s.offs.p, nreads, s.nbytes, k, canonical, keys, 0));
    if (out_words && n) {
        if (*n_kmers < n) GG_THROW(GG_ERR_ARG, "output holds %llu k-mers, %llu needed", (unsigned long long)*n_kmers, (unsigned long long)n);
        CK(cudaMemcpyAsync(out_words, keys.p, n * W * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    }
    CK(cudaStreamSynchronize(_c->stream));
    *n_kmers = n;
    API_END
}

// ------------------------------------- graph ----------------------------------------------
}  // extern "C"
// sorted unique canonical k-mers (the whole table, on every rank of cm) -> graph.  The sharded formulation
// (graphdist.cuh) is the default, on one GPU too; inputs with perfect circles take graph.cuh's.
template <int W>
static gg_graph *build_graph_auto(gg_ctx *ctx, const u64 *keys, u64 U, int k, gg_comm *cm = nullptr) {
    if (!getenv("GG_GRAPH_REPLICATED")) {
        gg_graph *g = build_graph_sharded<W>(ctx, keys, U, k, cm);
        if (g) return g;
    }
    return build_graph<W>(ctx, keys, U, k, cm);
}
extern "C" {
int gg_dbg_build(gg_ctx *ctx, gg_counts *counts, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out || !counts) GG_THROW(GG_ERR_ARG, "null argument");
    *out = nullptr;
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    stage_begin(_c, ST_TOTAL);
    DISPATCH_W(counts->W, *out = build_graph_auto<W>(_c, counts->d_keys, counts->n_unique, counts->k));
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}
static void dbg_from_reads_dev(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, int k, u64 min_freq, gg_graph **out, int enc = 0) {
    gg_counts *c = nullptr;
    min_freq = dbg_min_freq(min_freq);
    stage_begin(ctx, ST_TOTAL);
    bool filtered = false;
    DISPATCH_W(kmer_words(k), c = count_kmers_dev_impl<W>(ctx, d_seq, d_offs, nreads, nbytes, k, 1, min_freq, &filtered, enc));
    try {
        if (!filtered) { DISPATCH_W(c->W, filter_impl<W>(c, min_freq, 1)); }
        DISPATCH_W(c->W, *out = build_graph_auto<W>(ctx, c->d_keys, c->n_unique, c->k));
    } catch (...) {
        gg_counts_free(c);
        throw;
    }
    gg_counts_free(c);
    stage_end(ctx, ST_TOTAL);
    timing_collect(ctx);
}
int gg_dbg_from_reads_dev(gg_ctx *ctx, const uint8_t *d_seq, const uint64_t *d_offs, uint64_t nreads, uint64_t nbytes, int k,
                          uint64_t min_freq, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    check_offsets_dev(_c, d_offs, nreads, nbytes);
    dbg_from_reads_dev(_c, d_seq, d_offs, nreads, nbytes, k, min_freq, out);
    API_END
}
int gg_dbg_from_reads(gg_ctx *ctx, const uint8_t *seq, const uint64_t *offs, uint64_t nreads, int k, uint64_t min_freq, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    if (host_pipeline_ok(k, 1)) {
        stage_begin(_c, ST_TOTAL);
        gg_counts *c = count_kmers_host_pipelined(_c, seq, offs, nreads, k, 1, dbg_min_freq(min_freq), nullptr);
        try {
            DISPATCH_W(c->W, *out = build_graph_auto<W>(_c, c->d_keys, c->n_unique, c->k));
        } catch (...) {
            gg_counts_free(c);
            throw;
        }
        gg_counts_free(c);
        stage_end(_c, ST_TOTAL);
        timing_collect(_c);
        return GG_OK;
    }
    StagedReads s;
    stage_reads(_c, seq, offs, nreads, s);
    dbg_from_reads_dev(_c, s.seq.p, s.offs.p, nreads, s.nbytes, k, min_freq, out);
    API_END
}
// 4-bit packed host input: half the bytes of ASCII over PCIe; base offsets instead of byte offsets
static void stage_reads_4bit(gg_ctx *ctx, const u64 *data4, const u64 *offs, u64 nreads, StagedReads &s) {
    if (!offs) GG_THROW(GG_ERR_ARG, "read_offsets is null");
    s.nbytes = offs[nreads];   // bases
    if (s.nbytes && !data4) GG_THROW(GG_ERR_ARG, "packed sequence data is null");
    const u64 nw = ceil_div(s.nbytes, 16);
    s.seq.alloc(ctx, (nw + 2) * sizeof(u64));
    s.offs.alloc(ctx, nreads + 1);
    CK(cudaEventRecord(ctx->aux_ev, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->aux, ctx->aux_ev, 0));
    if (nw) CK(cudaMemcpyAsync(s.seq.p, data4, nw * sizeof(u64), cudaMemcpyHostToDevice, ctx->aux));
    CK(cudaMemcpyAsync(s.offs.p, offs, (nreads + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->aux));
    CK(cudaEventRecord(ctx->aux_ev, ctx->aux));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev, 0));
    check_offsets_host(ctx, offs, nreads);
}
int gg_count_kmers_4bit(gg_ctx *ctx, const uint64_t *data4, const uint64_t *offs, uint64_t nreads, int k, int canonical, gg_counts **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    StagedReads s;
    stage_reads_4bit(_c, data4, offs, nreads, s);
    stage_begin(_c, ST_TOTAL);
    DISPATCH_W(kmer_words(k), *out = count_kmers_dev_impl<W>(_c, s.seq.p, s.offs.p, nreads, s.nbytes, k, canonical, 1, nullptr, 1));
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}
int gg_dbg_from_reads_4bit(gg_ctx *ctx, const uint64_t *data4, const uint64_t *offs, uint64_t nreads, int k, uint64_t min_freq, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    StagedReads s;
    stage_reads_4bit(_c, data4, offs, nreads, s);
    dbg_from_reads_dev(_c, s.seq.p, s.offs.p, nreads, s.nbytes, k, min_freq, out, 1);
    API_END
}
// find_tip_nodes (src/Graphs.jl:778-816): node n (not deleted, length <= min_size) is a tip when it has one forward link
// and no backward link and the node it leads to has more than one link on that side, the mirror case, or no link at all
__global__ void tip_kernel(const u64 *__restrict__ node_off, const u64 *__restrict__ row, const i64 *__restrict__ tri, u64 nn, u64 min_size,
                           u32 *__restrict__ flag) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const u64 len = node_off[i + 1] - node_off[i];
    u32 tip = 0;
    if (len != 0 && len <= min_size) {
        const i64 n = (i64)(i + 1);
        u32 fw = 0, bw = 0;
        i64 dfw = 0, dbw = 0;
        for (u64 q = row[i]; q < row[i + 1]; ++q) {
            const i64 s = tri[3 * q], d = tri[3 * q + 1];
            if (s == -n) { if (!fw) dfw = d; ++fw; }   // is_forwards_from (src/Graphs.jl:188)
            if (s == n) { if (!bw) dbw = d; ++bw; }    // is_backwards_from (:191)
        }
        auto side = [&](i64 m) {   // links of node |m| whose source is m: backward_links(sg, m) = forward_links(sg, -m)
            const u64 j = (u64)(m < 0 ? -m : m) - 1;
            u32 c = 0;
            for (u64 q = row[j]; q < row[j + 1]; ++q) c += tri[3 * q] == m ? 1u : 0u;
            return c;
        };
        if (fw == 1 && bw == 0 && side(dfw) > 1) tip = 1;
        if (fw == 0 && bw == 1 && side(dbw) > 1) tip = 1;
        if (fw == 0 && bw == 0) tip = 1;
    }
    flag[i] = tip;
}
struct TipF {
    const u32 *flag;
    i64 *out;
    __device__ bool pred(u64 i) const { return flag[i] != 0; }
    __device__ void emit(u64 i, u64 dst) const { out[dst] = (i64)(i + 1); }
};
int gg_graph_find_tip_nodes(gg_graph *g, uint64_t min_size, int64_t *tips, uint64_t *n_tips) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    if (!n_tips) GG_THROW(GG_ERR_ARG, "n_tips is null");
    CK(cudaSetDevice(_c->device));
    const u64 nn = g->n_nodes;
    u64 nt = 0;
    if (nn) {
        DBuf<u32> flag(_c, nn);
        LAUNCH(_c, tip_kernel, (unsigned)ceil_div(nn, 256), 256, 0, g->d_node_off, g->d_link_row, g->d_link_tri, nn, (u64)min_size, flag.p);
        TipF f{flag.p, nullptr};
        Compactor<TipF> cp(_c, nn, "tips");
        nt = cp.count(f);
        if (tips && nt) {
            if (*n_tips < nt) GG_THROW(GG_ERR_ARG, "output holds %llu node ids, %llu needed", (unsigned long long)*n_tips, (unsigned long long)nt);
            DBuf<i64> out(_c, nt);
            f.out = out.p;
            cp.emit(f);
            CK(cudaMemcpyAsync(tips, out.p, nt * sizeof(i64), cudaMemcpyDeviceToHost, _c->stream));
        }
        CK(cudaStreamSynchronize(_c->stream));
    }
    *n_tips = nt;
    API_END
}
// remove_node! (src/Graphs.jl:459-466) for every id (+n and -n both name node n): a new handle, `g` stays as it is
int gg_graph_remove_nodes(gg_graph *g, const int64_t *ids, uint64_t n_ids, gg_graph **out) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    if (!out || (n_ids && !ids)) GG_THROW(GG_ERR_ARG, "null argument");
    *out = nullptr;
    CK(cudaSetDevice(_c->device));
    const u64 nn = g->n_nodes;
    for (u64 t = 0; t < n_ids; ++t) {
        const i64 x = ids[t];
        if (x == 0 || (u64)(x < 0 ? -x : x) > nn) GG_THROW(GG_ERR_ARG, "node id %lld outside 1..%llu", (long long)x, (unsigned long long)nn);
    }
    DBuf<u32> dead(_c, nn + 1);
    CK(cudaMemsetAsync(dead.p, 0, (nn + 1) * sizeof(u32), _c->stream));
    if (n_ids) {
        DBuf<i64> d_ids(_c, n_ids);
        CK(cudaMemcpyAsync(d_ids.p, ids, n_ids * sizeof(i64), cudaMemcpyHostToDevice, _c->stream));
        LAUNCH(_c, ge_mark_ids_kernel, (unsigned)ceil_div(n_ids, 256), 256, 0, d_ids.p, (u64)n_ids, dead.p);
        CK(cudaStreamSynchronize(_c->stream));   // ids is the caller's pageable memory
    }
    *out = ge_remove_dead(_c, g, dead.p);
    API_END
}
// collapse_all_unitigs!(sg, min_nodes, consume = true) (src/Graphs.jl:528-539)
int gg_graph_collapse_all_unitigs(gg_graph *g, uint64_t min_nodes, gg_graph **out, uint64_t *n_new_nodes) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    CK(cudaSetDevice(_c->device));
    u64 m = 0;
    *out = ge_collapse(_c, g, min_nodes, &m);
    if (n_new_nodes) *n_new_nodes = m;
    API_END
}
// remove_tips!(sdg, min_size) (src/remove_tips.jl:2-30): find the tips, remove them, collapse the paths that appear,
// until no tip is left.  *n_passes counts like the reference's log (1 = nothing to do).
int gg_graph_remove_tips(gg_graph *g, uint64_t min_size, gg_graph **out, uint64_t *n_passes, uint64_t *n_tips_removed) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    CK(cudaSetDevice(_c->device));
    gg_graph *cur = nullptr, *mid = nullptr;
    u64 passes = 1, removed = 0;
    try {
        for (;;) {
            const gg_graph *src = cur ? cur : g;
            const u64 nn = src->n_nodes;
            if (nn == 0) break;
            if (nn >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "too many nodes");
            DBuf<u32> flag(_c, nn + 1), scratch(_c, nn + 1);
            DBuf<u64> tot(_c, 1);
            LAUNCH(_c, tip_kernel, (unsigned)ceil_div(nn, 256), 256, 0, src->d_node_off, src->d_link_row, src->d_link_tri, nn, (u64)min_size, flag.p);
            exclusive_scan_u32(_c, flag.p, scratch.p, nn, tot.p);
            const u64 nt = read_scalar<u64>(_c, tot.p);
            if (nt == 0) break;
            mid = ge_remove_dead(_c, src, flag.p);
            if (cur) { ge_free_graph(_c, cur); cur = nullptr; }
            cur = ge_collapse(_c, mid, 2, nullptr);
            ge_free_graph(_c, mid);
            mid = nullptr;
            removed += nt;
            ++passes;
        }
        if (!cur) cur = ge_clone_graph(_c, g);
    } catch (...) {
        if (cur) ge_free_graph(_c, cur);
        if (mid) ge_free_graph(_c, mid);
        throw;
    }
    *out = cur;
    if (n_passes) *n_passes = passes;
    if (n_tips_removed) *n_tips_removed = removed;
    API_END
}
int gg_graph_sizes(gg_graph *g, uint64_t *n_nodes, uint64_t *total_bases, uint64_t *n_link_entries, int *k) {
    if (!g) return GG_ERR_ARG;
    if (n_nodes) *n_nodes = g->n_nodes;
    if (total_bases) *total_bases = g->total_bases;
    if (n_link_entries) *n_link_entries = g->n_link_entries;
    if (k) *k = g->k;
    return GG_OK;
}
int gg_graph_export(gg_graph *g, uint64_t *node_offsets, uint8_t *bases_ascii, uint64_t *link_row_offsets, int64_t *link_triples) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    CK(cudaSetDevice(_c->device));
    if (node_offsets) CK(cudaMemcpyAsync(node_offsets, g->d_node_off, (g->n_nodes + 1) * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    if (bases_ascii && g->total_bases) CK(cudaMemcpyAsync(bases_ascii, g->d_bases, g->total_bases, cudaMemcpyDeviceToHost, _c->stream));
    if (link_row_offsets) CK(cudaMemcpyAsync(link_row_offsets, g->d_link_row, (g->n_nodes + 1) * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    if (link_triples && g->n_link_entries)
        CK(cudaMemcpyAsync(link_triples, g->d_link_tri, 3 * g->n_link_entries * sizeof(i64), cudaMemcpyDeviceToHost, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
__global__ void checksum_kernel(const u8 *__restrict__ bases, u64 nb, const i64 *__restrict__ tri, u64 nt, unsigned long long *out) {
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (u64)gridDim.x * blockDim.x)
        acc += sy_mix(i * 4 + (u64)bases[i]);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (u64)gridDim.x * blockDim.x)
        acc += sy_mix(0x5bd1e995ULL + i) ^ sy_mix((u64)tri[i]);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)acc);
}
int gg_graph_checksum(gg_graph *g, uint64_t *checksum) {
    if (!g || !checksum) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    CK(cudaSetDevice(_c->device));
    DBuf<u64> acc(_c, 1);
    CK(cudaMemsetAsync(acc.p, 0, sizeof(u64), _c->stream));
    LAUNCH(_c, checksum_kernel, (unsigned)(_c->num_sms * 4), 256, 0, g->d_bases, g->total_bases, g->d_link_tri, 3 * g->n_link_entries, (unsigned long long *)acc.p);
    *checksum = read_scalar<u64>(_c, acc.p) + g->n_nodes * 0x9E3779B97F4A7C15ULL;
    API_END
}
void gg_graph_free(gg_graph *g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    dfree(g->ctx, g->d_node_off); dfree(g->ctx, g->d_bases); dfree(g->ctx, g->d_link_row); dfree(g->ctx, g->d_link_tri);
    delete g;
}

// ------------------------------------- persistence ------------------------------------------
int gg_gfa1_write(const char *filename, uint64_t n_nodes, const uint64_t *node_offsets, const uint8_t *bases_ascii,
                  const uint64_t *link_row_offsets, const int64_t *link_triples) {
    return gfa1_write(filename, n_nodes, node_offsets, bases_ascii, link_row_offsets, link_triples);
}
int gg_graph_write_gfa1(gg_graph *g, const char *filename) {
    if (!g) return GG_ERR_ARG;
    API_BEGIN(g->ctx)
    if (!filename) GG_THROW(GG_ERR_ARG, "null file name");
    std::vector<u64> off(g->n_nodes + 1), row(g->n_nodes + 1);
    std::vector<u8> bases(g->total_bases ? g->total_bases : 1);
    std::vector<i64> tri(g->n_link_entries ? 3 * g->n_link_entries : 1);
    const int rc = gg_graph_export(g, off.data(), bases.data(), row.data(), tri.data());
    if (rc != GG_OK) return rc;
    if (gfa1_write(filename, g->n_nodes, off.data(), bases.data(), row.data(), tri.data()) != GG_OK)
        GG_THROW(GG_ERR_ARG, "cannot write %s.gfa / %s.fasta", filename, filename);
    API_END
}
int gg_graph_from_arrays(gg_ctx *ctx, int k, uint64_t n_nodes, const uint64_t *node_offsets, const uint8_t *bases_ascii,
                         const uint64_t *link_row_offsets, const int64_t *link_triples, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out || !node_offsets || !link_row_offsets) GG_THROW(GG_ERR_ARG, "null argument");
    *out = nullptr;
    if (node_offsets[0] != 0 || link_row_offsets[0] != 0) GG_THROW(GG_ERR_ARG, "offset arrays must start at 0");
    for (u64 i = 0; i < n_nodes; ++i)
        if (node_offsets[i] > node_offsets[i + 1] || link_row_offsets[i] > link_row_offsets[i + 1]) GG_THROW(GG_ERR_ARG, "offset arrays must be non-decreasing");
    const u64 nb = node_offsets[n_nodes], nl = link_row_offsets[n_nodes];
    if ((nb && !bases_ascii) || (nl && !link_triples)) GG_THROW(GG_ERR_ARG, "null argument");
    for (u64 q = 0; q < nl; ++q) {
        const i64 s = link_triples[3 * q], d = link_triples[3 * q + 1];
        if (s == 0 || d == 0 || (u64)(s < 0 ? -s : s) > n_nodes || (u64)(d < 0 ? -d : d) > n_nodes) GG_THROW(GG_ERR_ARG, "link %llu names a node outside 1..n", (unsigned long long)q);
    }
    CK(cudaSetDevice(_c->device));
    gg_graph *g = new gg_graph();
    g->ctx = _c; g->k = k; g->n_nodes = n_nodes; g->total_bases = nb; g->n_link_entries = nl;
    try {
        g->d_node_off = dalloc<u64>(_c, n_nodes + 1);
        g->d_link_row = dalloc<u64>(_c, n_nodes + 1);
        g->d_bases = dalloc<u8>(_c, nb + 16);
        g->d_link_tri = dalloc<i64>(_c, 3 * (nl ? nl : 1));
        CK(cudaMemcpyAsync(g->d_node_off, node_offsets, (n_nodes + 1) * sizeof(u64), cudaMemcpyHostToDevice, _c->stream));
        CK(cudaMemcpyAsync(g->d_link_row, link_row_offsets, (n_nodes + 1) * sizeof(u64), cudaMemcpyHostToDevice, _c->stream));
        if (nb) CK(cudaMemcpyAsync(g->d_bases, bases_ascii, nb, cudaMemcpyHostToDevice, _c->stream));
        if (nl) CK(cudaMemcpyAsync(g->d_link_tri, link_triples, 3 * nl * sizeof(i64), cudaMemcpyHostToDevice, _c->stream));
        CK(cudaStreamSynchronize(_c->stream));
    } catch (...) {
        dfree(_c, g->d_node_off); dfree(_c, g->d_bases); dfree(_c, g->d_link_row); dfree(_c, g->d_link_tri);
        delete g;
        throw;
    }
    *out = g;
    API_END
}
int gg_graph_load_gfa1(gg_ctx *ctx, const char *gfa_file, const char *fasta_file, gg_graph **out) {
    API_BEGIN(ctx)
    if (!out || !gfa_file || !fasta_file) GG_THROW(GG_ERR_ARG, "null argument");
    *out = nullptr;
    GfaGraphHost h;
    std::string err;
    if (gfa1_read(gfa_file, fasta_file, h, err) != GG_OK) GG_THROW(GG_ERR_ARG, "%s", err.c_str());
    i64 ov = 0;   // k - 1 = the overlap of the DBG links (0 when the graph has none)
    if (!h.tri.empty()) ov = -h.tri[2];
    const u64 nn = h.node_off.size() - 1;
    const int rc = gg_graph_from_arrays(ctx, (int)(ov > 0 ? ov + 1 : 0), nn, h.node_off.data(), h.bases.data(), h.link_row.data(), h.tri.data(), out);
    if (rc != GG_OK) return rc;
    API_END
}

// ------------------------------------- UMI -------------------------------------------------
int gg_unique_mer_index(gg_ctx *ctx, const uint8_t *bases, const uint64_t *node_offsets, uint64_t n_nodes, int k, gg_umi **out) {
    API_BEGIN(ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    StagedReads s;
    stage_reads(_c, bases, node_offsets, n_nodes, s);
    DISPATCH_W(kmer_words(k), *out = build_umi<W>(_c, s.seq.p, s.offs.p, n_nodes, s.nbytes, k));
    API_END
}
int gg_umi_sizes(gg_umi *u, uint64_t *n_unique, uint64_t *n_nodes, int *k, int *words) {
    if (!u) return GG_ERR_ARG;
    if (n_unique) *n_unique = u->n_unique;
    if (n_nodes) *n_nodes = u->n_nodes;
    if (k) *k = u->k;
    if (words) *words = u->W;
    return GG_OK;
}
int gg_umi_export(gg_umi *u, uint64_t *kmer_words_out, int64_t *node, uint32_t *pos, uint64_t *unique_per_node, uint64_t *total_per_node) {
    if (!u) return GG_ERR_ARG;
    API_BEGIN(u->ctx)
    CK(cudaSetDevice(_c->device));
    u64 n = u->n_unique;
    if (n && kmer_words_out) CK(cudaMemcpyAsync(kmer_words_out, u->d_keys, n * u->W * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    if (n && (node || pos)) {
        DBuf<i64> dn(_c, n);
        DBuf<u32> dp(_c, n);
        LAUNCH(_c, umi_unpack_kernel, (unsigned)ceil_div(n, 256), 256, 0, u->d_pay, n, dn.p, dp.p);
        if (node) CK(cudaMemcpyAsync(node, dn.p, n * sizeof(i64), cudaMemcpyDeviceToHost, _c->stream));
        if (pos) CK(cudaMemcpyAsync(pos, dp.p, n * sizeof(u32), cudaMemcpyDeviceToHost, _c->stream));
        CK(cudaStreamSynchronize(_c->stream));
    }
    if (u->n_nodes && unique_per_node) CK(cudaMemcpyAsync(unique_per_node, u->d_unique_per_node, u->n_nodes * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    if (u->n_nodes && total_per_node) CK(cudaMemcpyAsync(total_per_node, u->d_total_per_node, u->n_nodes * sizeof(u64), cudaMemcpyDeviceToHost, _c->stream));
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
void gg_umi_free(gg_umi *u) {
    if (!u) return;
    cudaSetDevice(u->ctx->device);
    dfree(u->ctx, u->d_keys); dfree(u->ctx, u->d_pay); dfree(u->ctx, u->d_unique_per_node); dfree(u->ctx, u->d_total_per_node);
    delete u;
}

// ------------------------------------- multi-GPU ---------------------------------------------
int gg_comm_unique_id(uint8_t id[GG_COMM_ID_BYTES]) {
    NcclApi *nc = nccl_api();
    if (!nc || !id) return GG_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) <= GG_COMM_ID_BYTES, "unique id size");
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) return GG_ERR_CUDA;
    memset(id, 0, GG_COMM_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return GG_OK;
}
int gg_comm_create(gg_ctx *ctx, const uint8_t id[GG_COMM_ID_BYTES], int rank, int world, gg_comm **out) {
    API_BEGIN(ctx)
    if (!out || !id || world < 1 || rank < 0 || rank >= world) GG_THROW(GG_ERR_ARG, "bad communicator arguments");
    if (world > MSD_MAX_PEERS) GG_THROW(GG_ERR_LIMIT, "a communicator spans at most %d ranks (the GPUs of one box)", MSD_MAX_PEERS);
    *out = nullptr;
    NcclApi *nc = nccl_api();
    if (!nc) GG_THROW(GG_ERR_CUDA, "libnccl.so.2 could not be loaded");
    CK(cudaSetDevice(_c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    gg_comm *cm = new gg_comm();
    cm->ctx = _c; cm->rank = rank; cm->world = world;
    ncclResult_t r = nc->CommInitRank(&cm->comm, world, u, rank);
    if (r != ncclSuccess) { delete cm; GG_THROW(GG_ERR_CUDA, "ncclCommInitRank: %s", nc->GetErrorString(r)); }
    *out = cm;
    API_END
}
void gg_comm_destroy(gg_comm *cm) {
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    dist_close_peers(cm);
    if (cm->R) cudaFree(cm->R);
    if (cm->agree) cudaFree(cm->agree);
    if (cm->comm) nccl_api()->CommDestroy(cm->comm);
    delete cm;
}
int gg_dist_splitters(const uint64_t *hist, uint32_t nb, int world, uint32_t *first) {
    if (!hist || !first || world < 1) return GG_ERR_ARG;
    dist_splitters(hist, nb, world, first);
    return GG_OK;
}
}  // extern "C"

template <int W>
static gg_counts *dist_count_impl(gg_comm *cm, const u8 *d_seq, const u64 *d_offs, u64 nreads, u64 nbytes, int k, int canonical, u64 min_freq,
                                  u64 *n_glob) {
    gg_ctx *ctx = cm->ctx;
    if (W == 1 && !msd_supported(k, canonical)) GG_THROW(GG_ERR_LIMIT, "non-canonical counting at k = 32 is not covered by the multi-GPU path");
    gg_counts *c = new gg_counts();
    c->ctx = ctx; c->k = k; c->W = W;
    try {
        stage_begin(ctx, ST_PACK);
        PackedReads pr;
        pack_reads(ctx, d_seq, d_offs, nreads, nbytes, pr);
        DBuf<u32> valid;
        window_valid_mask(ctx, pr, k, valid);
        stage_end(ctx, ST_PACK);
        DBuf<u64> ukeys;
        DBuf<u32> counts;
        u32 mf = min_freq > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)min_freq;
        u64 ng = 0;
        if (W == 1 && skm_supported(k, canonical) && !getenv("GG_FORCE_MSD"))
            c->n_unique = dist_skm_count(cm, pr, valid.p, k, canonical, mf < 1 ? 1 : mf, ukeys, counts, &c->n_instances, &ng);
        else
            c->n_unique = dist_msd_count<W>(cm, pr, valid.p, k, canonical, mf < 1 ? 1 : mf, ukeys, counts, &c->n_instances, &ng);
        if (n_glob) *n_glob = ng;
        c->d_keys = ukeys.take();
        c->d_counts = counts.take();
        return c;
    } catch (...) {
        dfree(ctx, c->d_keys); dfree(ctx, c->d_counts);
        delete c;
        throw;
    }
}
// every rank's slice -> the whole table on every rank (slices are in key order by rank).
// The slices are balanced, so they are padded to the largest one and moved with ONE ncclAllGather per
// array (ring / NVLS inside NCCL) and then compacted; a mesh of sends costs several times more at 8 ranks.
__global__ void allgather_compact_kernel(const u64 *__restrict__ pk, const u32 *__restrict__ pc, u64 maxn, int W, int P,
                                         const u64 *__restrict__ n, const u64 *__restrict__ off, u64 *__restrict__ keys,
                                         u32 *__restrict__ counts) {
    for (int p = 0; p < P; ++p) {
        const u64 np = n[p], o = off[p];
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < np * W; i += (u64)gridDim.x * blockDim.x) keys[o * W + i] = pk[(u64)p * maxn * W + i];
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (u64)gridDim.x * blockDim.x) counts[o + i] = pc[(u64)p * maxn + i];
    }
}
static gg_counts *counts_allgather_impl(gg_comm *cm, gg_counts *c) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world;
    DBuf<u64> sz(ctx, 2 * P + 2);
    CK(cudaMemcpyAsync(sz.p + 2 * P, &c->n_unique, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    NCK(nc->AllGather(sz.p + 2 * P, sz.p, 1, ncclUint64, cm->comm, ctx->stream));
    std::vector<u64> n(P), off(P);
    CK(cudaMemcpyAsync(n.data(), sz.p, P * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u64 total = 0, maxn = 1;
    for (int p = 0; p < P; ++p) { off[p] = total; total += n[p]; maxn = n[p] > maxn ? n[p] : maxn; }
    CK(cudaMemcpyAsync(sz.p + P, off.data(), P * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    gg_counts *f = new gg_counts();
    f->ctx = ctx; f->k = c->k; f->W = c->W; f->n_unique = total; f->n_instances = c->n_instances;
    try {
        const int W = c->W;
        DBuf<u64> mk, pk;
        DBuf<u32> mc, pc;
        dist_guard(cm, "gathered count table", [&] {   // an allocation failure on one rank must not leave the peers in the all-gather
            f->d_keys = dalloc<u64>(ctx, total * W);
            f->d_counts = dalloc<u32>(ctx, total);
            mk.alloc(ctx, maxn * W); pk.alloc(ctx, (u64)P * maxn * W);
            mc.alloc(ctx, maxn); pc.alloc(ctx, (u64)P * maxn);
            if (c->n_unique) {
                CK(cudaMemcpyAsync(mk.p, c->d_keys, c->n_unique * W * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
                CK(cudaMemcpyAsync(mc.p, c->d_counts, c->n_unique * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
            }
        });
        NCK(nc->AllGather(mk.p, pk.p, maxn * W, ncclUint64, cm->comm, ctx->stream));
        NCK(nc->AllGather(mc.p, pc.p, maxn, ncclUint32, cm->comm, ctx->stream));
        LAUNCH(ctx, allgather_compact_kernel, (unsigned)(ctx->num_sms * 8), 256, 0, pk.p, pc.p, maxn, W, P, sz.p, sz.p + P, f->d_keys, f->d_counts);
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        dfree(ctx, f->d_keys); dfree(ctx, f->d_counts);
        delete f;
        throw;
    }
    return f;
}

extern "C" {
int gg_dist_count_kmers_dev(gg_comm *cm, const uint8_t *d_seq, const uint64_t *d_offs, uint64_t nreads, uint64_t nbytes, int k,
                            int canonical, uint64_t min_freq, gg_counts **out, uint64_t *n_instances_global) {
    if (!cm) return GG_ERR_ARG;
    API_BEGIN(cm->ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    stage_begin(_c, ST_TOTAL);
    DISPATCH_W(kmer_words(k), *out = dist_count_impl<W>(cm, d_seq, d_offs, nreads, nbytes, k, canonical, min_freq, n_instances_global));
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}
int gg_counts_allgather(gg_comm *cm, gg_counts *local, gg_counts **full) {
    if (!cm) return GG_ERR_ARG;
    API_BEGIN(cm->ctx)
    if (!local || !full) GG_THROW(GG_ERR_ARG, "null argument");
    *full = nullptr;
    CK(cudaSetDevice(_c->device));
    *full = counts_allgather_impl(cm, local);
    API_END
}
int gg_dist_dbg_from_reads_dev(gg_comm *cm, const uint8_t *d_seq, const uint64_t *d_offs, uint64_t nreads, uint64_t nbytes, int k,
                               uint64_t min_freq, gg_graph **out, uint64_t *n_instances_global) {
    if (!cm) return GG_ERR_ARG;
    API_BEGIN(cm->ctx)
    if (!out) GG_THROW(GG_ERR_ARG, "null output");
    *out = nullptr;
    check_k(k);
    CK(cudaSetDevice(_c->device));
    timing_reset(_c);
    stage_begin(_c, ST_TOTAL);
    gg_counts *c = nullptr;
    DISPATCH_W(kmer_words(k), c = dist_count_impl<W>(cm, d_seq, d_offs, nreads, nbytes, k, 1, dbg_min_freq(min_freq), n_instances_global));
    try {
        stage_begin(_c, ST_FILTER);  // the all-gather of the kept k-mers is booked under "filter"
        DBuf<u64> all;
        u64 U = 0;
        DISPATCH_W(c->W, U = keys_allgatherv<W>(cm, c->d_keys, c->n_unique, all));
        stage_end(_c, ST_FILTER);
        DISPATCH_W(c->W, *out = build_graph_auto<W>(_c, all.p, U, c->k, cm));
    } catch (...) {
        gg_counts_free(c);
        throw;
    }
    gg_counts_free(c);
    stage_end(_c, ST_TOTAL);
    timing_collect(_c);
    API_END
}

// ------------------------------------- synth ------------------------------------------------
static int synth_check(u64 genome_len, u32 read_len, int paired, u32 frag_len) {
    u64 F = paired ? frag_len : read_len;
    if (read_len == 0 || F < read_len || genome_len < F) return GG_ERR_ARG;
    return GG_OK;
}
int gg_synth_reads_range_host(uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t nreads, uint32_t read_len, int paired,
                              uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *out) {
    if (!out || synth_check(genome_len, read_len, paired, frag_len) != GG_OK || (paired && (first_read & 1))) return GG_ERR_ARG;
    SynthParams P = synth_params(seed, genome_len, nreads, read_len, paired, frag_len, err_ppm, n_ppm);
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long)nreads; ++r)
        for (u32 p = 0; p < read_len; ++p) out[(u64)r * read_len + p] = synth_base(P, first_read + (u64)r, p);
    return GG_OK;
}
int gg_synth_reads_host(uint64_t seed, uint64_t genome_len, uint64_t nreads, uint32_t read_len, int paired, uint32_t frag_len,
                        uint32_t err_ppm, uint32_t n_ppm, uint8_t *out) {
    return gg_synth_reads_range_host(seed, genome_len, 0, nreads, read_len, paired, frag_len, err_ppm, n_ppm, out);
}
int gg_synth_reads_range_dev(gg_ctx *ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t nreads, uint32_t read_len,
                             int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *d_out) {
    API_BEGIN(ctx)
    if (!d_out || synth_check(genome_len, read_len, paired, frag_len) != GG_OK || (paired && (first_read & 1)))
        GG_THROW(GG_ERR_ARG, "bad synthetic read parameters");
    CK(cudaSetDevice(_c->device));
    SynthParams P = synth_params(seed, genome_len, nreads, read_len, paired, frag_len, err_ppm, n_ppm);
    LAUNCH(_c, synth_reads_kernel, (unsigned)(_c->num_sms * 16), 256, 0, P, first_read, d_out);
    CK(cudaStreamSynchronize(_c->stream));
    API_END
}
int gg_synth_reads_dev(gg_ctx *ctx, uint64_t seed, uint64_t genome_len, uint64_t nreads, uint32_t read_len, int paired,
                       uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *d_out) {
    return gg_synth_reads_range_dev(ctx, seed, genome_len, 0, nreads, read_len, paired, frag_len, err_ppm, n_ppm, d_out);
}
int gg_synth_genome_host(uint64_t seed, uint64_t genome_len, uint8_t *out) {
    if (!out) return GG_ERR_ARG;
    u64 gkey = sy_key(seed, 1);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)genome_len; ++i) out[i] = code_to_ascii(sy_genome_base(gkey, (u64)i));
    return GG_OK;
}

}  // extern "C"
