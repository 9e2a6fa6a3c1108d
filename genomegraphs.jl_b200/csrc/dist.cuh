// dist.cuh -- multi-GPU counting: one process per GPU, k-mers range-partitioned over the ranks.
//
// The reference has no distributed code of its own; its re-exported KmerAnalysis `dist_mem`
// (src/GenomeGraphs.jl:46-47, docs/src/man/assembly.md:108-117) counts read ranges in worker
// processes and merges sorted count vectors.  Here every rank runs passes A and B of msdcount.cuh
// on its own shard of the reads, the level-0 buckets are assigned to ranks as contiguous ranges
// balanced on the all-gathered histogram, and ONE exchange (grouped ncclSend/ncclRecv over
// NVLink/NVSwitch) moves every bucket to its owner.  Because the level-0 scatter already groups
// the keys by bucket, the send buffers are plain slices of its output (no packing pass) and the
// received pieces are consumed in place as "virtual buckets".  Rank r then holds all instances of
// every k-mer of its key range: its sorted table is a contiguous slice of the global sorted table
// (concatenating the ranks' tables in rank order gives exactly the single-GPU result).
//
// NCCL is loaded with dlopen so that libggb200.so has no hard dependency on it (single-GPU use).
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <chrono>
#include "msdcount.cuh"

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi *nccl_api() {
    static NcclApi a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.h) break;
        }
        if (a.h) {
#define NCCL_SYM(field, sym) *(void **)(&a.field) = dlsym(a.h, sym)
            NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            NCCL_SYM(CommInitRank, "ncclCommInitRank");
            NCCL_SYM(CommDestroy, "ncclCommDestroy");
            NCCL_SYM(GroupStart, "ncclGroupStart");
            NCCL_SYM(GroupEnd, "ncclGroupEnd");
            NCCL_SYM(Send, "ncclSend");
            NCCL_SYM(Recv, "ncclRecv");
            NCCL_SYM(AllGather, "ncclAllGather");
            NCCL_SYM(AllReduce, "ncclAllReduce");
            NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
            a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.AllGather &&
                   a.AllReduce && a.GetErrorString;
        }
    }
    return a.ok ? &a : nullptr;
}
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t _r = (call);                                                                          \
        if (_r != ncclSuccess) GG_THROW(GG_ERR_CUDA, "%s:%d NCCL %s: %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(_r)); \
    } while (0)

struct gg_comm {
    gg_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

// Contiguous bucket ranges balanced on the global histogram: rank q owns buckets
// [first[q], first[q + 1]).  Pure host code (also exported for tests: gg_dist_splitters).
static void dist_splitters(const u64 *hist, u32 nb, int world, u32 *first) {
    u64 total = 0;
    for (u32 b = 0; b < nb; ++b) total += hist[b];
    first[0] = 0;
    u64 cum = 0;
    u32 b = 0;
    for (int q = 1; q < world; ++q) {
        // smallest b such that the keys before it reach q/world of the total (rounded to the nearer boundary)
        const u64 target = (u64)(((unsigned __int128)total * (unsigned)q) / (unsigned)world);
        while (b < nb && cum + hist[b] / 2 < target) cum += hist[b++];
        first[q] = b;
    }
    first[world] = nb;
}

// variable-size exchange: rank p's slice send[soff[p] .. + scnt[p]) -> its recv[roff[me] ..) (elements of `esz` bytes)
static void dist_alltoallv(gg_comm *cm, const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt,
                           size_t esz) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int me = cm->rank;
    if (scnt[me])
        CK(cudaMemcpyAsync((u8 *)recv + roff[me] * esz, (const u8 *)send + soff[me] * esz, scnt[me] * esz, cudaMemcpyDeviceToDevice, ctx->stream));
    NCK(nc->GroupStart());
    for (int p = 0; p < cm->world; ++p) {
        if (p == me) continue;
        if (scnt[p]) NCK(nc->Send((const u8 *)send + soff[p] * esz, scnt[p] * esz, ncclUint8, p, cm->comm, ctx->stream));
        if (rcnt[p]) NCK(nc->Recv((u8 *)recv + roff[p] * esz, rcnt[p] * esz, ncclUint8, p, cm->comm, ctx->stream));
    }
    NCK(nc->GroupEnd());
}

// instances per child of this rank's bucket range, summed over the ranks' pass-A histograms
__global__ void dist_child_hist_kernel(const u32 *__restrict__ all_hist, int P, u32 nbh, int hb, int bits, u32 child0, u32 nchild,
                                       u32 *__restrict__ out) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nchild) return;
    u32 s = 0;
    if (c < nchild) {
        const u32 g = 1u << (hb - bits), base = (child0 + c) * g;
        for (int p = 0; p < P; ++p)
            for (u32 i = 0; i < g; ++i) s += all_hist[(size_t)p * nbh + base + i];
    }
    out[c] = s;
}

// reads of this rank -> this rank's key-range slice of the global sorted (k-mer, count) table
static u64 dist_msd_count(gg_comm *cm, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u64> &ukeys,
                          DBuf<u32> &ucounts, u64 *n_instances_local, u64 *n_instances_global) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world, me = cm->rank;
    const MsdTune tune = msd_tune();
    const int hb = msd_hist_bits(k);
    const u32 nbh = 1u << hb;
    // GG_DIST_DEBUG: host wall-clock of every phase on every rank (adds stream syncs; never set while timing)
    const bool dbg = getenv("GG_DIST_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tmark[8] = {0};
    int nmark = 0;
    auto mark = [&]() { if (dbg) { cudaStreamSynchronize(ctx->stream); if (nmark < 8) tmark[nmark++] = now(); } };
    mark();
    stage_begin(ctx, ST_EXTRACT);
    // ---- global size -> sampling rate of the estimate (must agree on all ranks) ----
    DBuf<u64> sz(ctx, 1);
    CK(cudaMemcpyAsync(sz.p, &pr.nbases, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    NCK(nc->AllReduce(sz.p, sz.p, 1, ncclUint64, ncclSum, cm->comm, ctx->stream));
    const u64 nbases_glob = read_scalar<u64>(ctx, sz.p);
    const int est_shift = msd_est_shift(nbases_glob);
    mark();  // 1: size all-reduce done
    // ---- pass A on the local reads; histograms all-gathered, estimate maps merged ----
    DBuf<u32> hist, all_hist(ctx, (u64)P * nbh);
    DBuf<u8> est;
    msd_pass_a_launch(ctx, pr, valid, k, canonical, est_shift, hist, est);
    NCK(nc->AllGather(hist.p, all_hist.p, nbh, ncclUint32, cm->comm, ctx->stream));
    NCK(nc->AllReduce(est.p, est.p, 1ULL << EST_LOG, ncclUint8, ncclMax, cm->comm, ctx->stream));
    u64 N_loc = 0, dest = 0;
    msd_pass_a_read(ctx, hist.p, nbh, est.p, est_shift, &N_loc, &dest);
    est.release();
    std::vector<u32> h_all((size_t)P * nbh);
    CK(cudaMemcpyAsync(h_all.data(), all_hist.p, h_all.size() * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    mark();  // 2: pass A + histogram all-gather + estimate merge done
    u64 N_glob = 0;
    for (u32 v : h_all) N_glob += v;
    *n_instances_local = N_loc;
    *n_instances_global = N_glob;
    if (N_loc >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu k-mer instances exceed the 2^32-1 per-GPU limit", (unsigned long long)N_loc);
    const MsdPlan pl = msd_plan(dest, N_glob, k, tune, hb);
    const u32 nb0 = 1u << pl.b0, grp = 1u << (hb - pl.b0);
    // ---- per-rank histograms at level-0 resolution, owners of the buckets ----
    std::vector<u64> cnt((size_t)P * nb0, 0), gh(nb0, 0);
    for (int p = 0; p < P; ++p)
        for (u32 b = 0; b < nbh; ++b) cnt[(size_t)p * nb0 + b / grp] += h_all[(size_t)p * nbh + b];
    for (int p = 0; p < P; ++p)
        for (u32 b = 0; b < nb0; ++b) gh[b] += cnt[(size_t)p * nb0 + b];
    std::vector<u32> first(P + 1);
    dist_splitters(gh.data(), nb0, P, first.data());
    // ---- pass B: local scatter (its output doubles as the send buffer) ----
    DBuf<u32> hist0, bstart;
    DBuf<u64> A;
    msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A);
    stage_end(ctx, ST_EXTRACT);
    mark();  // 3: pass B done
    // ---- exchange ----
    stage_begin(ctx, ST_SORT);
    std::vector<u64> soff(P), scnt(P), roff(P), rcnt(P);
    {
        u64 run = 0;
        u32 b = 0;
        for (int p = 0; p < P; ++p) {
            soff[p] = run;
            for (; b < first[p + 1]; ++b) run += cnt[(size_t)me * nb0 + b];
            scnt[p] = run - soff[p];
        }
    }
    const u32 nb_local = first[me + 1] - first[me];
    u64 N_recv = 0;
    for (int p = 0; p < P; ++p) {
        roff[p] = N_recv;
        u64 c = 0;
        for (u32 b = first[me]; b < first[me + 1]; ++b) c += cnt[(size_t)p * nb0 + b];
        rcnt[p] = c;
        N_recv += c;
    }
    if (N_recv >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu received k-mer instances exceed the 2^32-1 per-GPU limit", (unsigned long long)N_recv);
    // virtual buckets of the receive buffer: vb = (b - first[me]) * P + p
    std::vector<u32> h_vbeg((size_t)nb_local * P + 1, 0), h_vcnt((size_t)nb_local * P + 1, 0);
    for (int p = 0; p < P; ++p) {
        u64 run = roff[p];
        for (u32 bl = 0; bl < nb_local; ++bl) {
            const u64 c = cnt[(size_t)p * nb0 + first[me] + bl];
            h_vbeg[(size_t)bl * P + p] = (u32)run;
            h_vcnt[(size_t)bl * P + p] = (u32)c;
            run += c;
        }
    }
    DBuf<u32> vbeg(ctx, h_vbeg.size()), vcnt(ctx, h_vcnt.size());
    CK(cudaMemcpyAsync(vbeg.p, h_vbeg.data(), h_vbeg.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(vcnt.p, h_vcnt.data(), h_vcnt.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    DBuf<u64> R(ctx, N_recv ? N_recv : 1);
    dist_alltoallv(cm, A.p, soff.data(), scnt.data(), R.p, roff.data(), rcnt.data(), sizeof(u64));
    CK(cudaStreamSynchronize(ctx->stream));  // h_vbeg / h_vcnt are read by the copies above; A is released next
    A.release();
    stage_end(ctx, ST_SORT);
    mark();  // 4: exchange done
    DBuf<u32> child_hist;
    if (pl.b1 > 0 && pl.b0 + pl.b1 <= hb && nb_local) {  // pass C for free (see msd_count)
        const u32 nchild = nb_local << pl.b1;
        child_hist.alloc(ctx, nchild + 1);
        LAUNCH(ctx, dist_child_hist_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, all_hist.p, P, nbh, hb, pl.b0 + pl.b1,
               first[me] << pl.b1, nchild, child_hist.p);
    }
    const u64 U = msd_finish(ctx, R, N_recv, vbeg.p, vcnt.p, nb_local, (u32)P, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
    mark();  // 5: level 1 + leaves + gather done
    if (dbg)
        fprintf(stderr, "[dist rank %d] t0=%.3f size-allreduce %.2f | passA+allgather %.2f | passB %.2f | exchange %.2f (send %.1f MB, recv %.1f MB) | finish %.2f ms\n",
                me, tmark[0], tmark[1] - tmark[0], tmark[2] - tmark[1], tmark[3] - tmark[2], tmark[4] - tmark[3],
                (double)(N_loc - scnt[me]) * 8e-6, (double)(N_recv - rcnt[me]) * 8e-6, tmark[5] - tmark[4]);
    return U;
}
