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
#include "msdcountw.cuh"

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
    // receive buffer of the fused scatter + exchange, mapped into every peer with CUDA IPC
    u64 *R = nullptr;
    u64 cap = 0;                       // keys
    u64 *peerR[MSD_MAX_PEERS] = {};    // peerR[rank] == R
    int ipc_state = 0;                 // 0 untried, 1 usable, -1 unavailable (NCCL send/recv is used instead)
};

static void dist_close_peers(gg_comm *cm) {
    for (int p = 0; p < cm->world && p < MSD_MAX_PEERS; ++p) {
        if (p != cm->rank && cm->peerR[p]) cudaIpcCloseMemHandle(cm->peerR[p]);
        cm->peerR[p] = nullptr;
    }
}
// Collective: every rank calls it with the same `need` (keys).  (Re)allocates the receive buffers and
// exchanges the IPC handles with an all-gather.  Returns false when peer mapping is not available.
static bool dist_ensure_recv(gg_comm *cm, u64 need) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    if (cm->ipc_state < 0 || cm->world > MSD_MAX_PEERS || getenv("GG_DIST_NO_IPC")) { cm->ipc_state = -1; return false; }
    if (cm->ipc_state == 1 && need <= cm->cap) return true;
    // every rank reaches this point with the same arguments, so the collective sequence below is consistent
    DBuf<u64> tok(ctx, 1);
    CK(cudaMemsetAsync(tok.p, 0, sizeof(u64), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dist_close_peers(cm);
    NCK(nc->AllReduce(tok.p, tok.p, 1, ncclUint64, ncclSum, cm->comm, ctx->stream));  // barrier: nobody maps the old buffers any more
    CK(cudaStreamSynchronize(ctx->stream));
    if (cm->R) { cudaFree(cm->R); cm->R = nullptr; cm->cap = 0; }
    const u64 cap = need + need / 8 + 1024;
    u64 ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc((void **)&cm->R, cap * sizeof(u64)) != cudaSuccess || cudaIpcGetMemHandle(&mine, cm->R) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DBuf<u8> hd(ctx, 64 * (u64)(cm->world + 1));
    CK(cudaMemcpyAsync(hd.p + 64 * (u64)cm->world, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    NCK(nc->AllGather(hd.p + 64 * (u64)cm->world, hd.p, 64, ncclUint8, cm->comm, ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(cm->world);
    CK(cudaMemcpyAsync(all.data(), hd.p, 64 * (u64)cm->world, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ok) {
        for (int p = 0; p < cm->world; ++p) {
            if (p == cm->rank) { cm->peerR[p] = cm->R; continue; }
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            cm->peerR[p] = (u64 *)q;
        }
    }
    // all ranks must agree: one failure anywhere switches everybody to NCCL send/recv
    CK(cudaMemcpyAsync(tok.p, &ok, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    NCK(nc->AllReduce(tok.p, tok.p, 1, ncclUint64, ncclMin, cm->comm, ctx->stream));
    const u64 all_ok = read_scalar<u64>(ctx, tok.p);
    if (!all_ok) {
        dist_close_peers(cm);
        if (cm->R) { cudaFree(cm->R); cm->R = nullptr; }
        cm->cap = 0;
        cm->ipc_state = -1;
        return false;
    }
    cm->cap = cap;
    cm->ipc_state = 1;
    return true;
}

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
template <int W>
static u64 dist_msd_count(gg_comm *cm, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u64> &ukeys,
                          DBuf<u32> &ucounts, u64 *n_instances_local, u64 *n_instances_global) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world, me = cm->rank;
    const MsdTune tune = W == 1 ? msd_tune() : msdw_tune<W>();
    const int hb = W == 1 ? msd_hist_bits(k) : msdw_hist_bits(k);
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
    if constexpr (W == 1) msd_pass_a_launch(ctx, pr, valid, k, canonical, est_shift, hist, est);
    else msdw_pass_a_launch<W>(ctx, pr, valid, k, canonical, est_shift, hist, est);
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
    // ---- who sends what to whom (every rank can compute the whole matrix from the all-gathered histograms) ----
    const u32 nb_local = first[me + 1] - first[me];
    std::vector<u64> nrecv(P, 0);
    for (int q = 0; q < P; ++q)
        for (int p = 0; p < P; ++p)
            for (u32 b = first[q]; b < first[q + 1]; ++b) nrecv[q] += cnt[(size_t)p * nb0 + b];
    u64 need = 1;
    for (int q = 0; q < P; ++q) need = nrecv[q] > need ? nrecv[q] : need;
    if (need >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu received k-mer instances exceed the 2^32-1 per-GPU limit", (unsigned long long)need);
    const u64 N_recv = nrecv[me];
    // layout of rank q's receive buffer: source-major, buckets of q's range in order inside every source segment
    auto seg_off = [&](int q, int p) { u64 o = 0; for (int pp = 0; pp < p; ++pp) for (u32 b = first[q]; b < first[q + 1]; ++b) o += cnt[(size_t)pp * nb0 + b]; return o; };
    std::vector<u64> soff(P), scnt(P), roff(P), rcnt(P);
    {
        u64 run = 0;
        u32 b = 0;
        for (int p = 0; p < P; ++p) {
            soff[p] = run;
            for (; b < first[p + 1]; ++b) run += cnt[(size_t)me * nb0 + b];
            scnt[p] = run - soff[p];
            roff[p] = seg_off(me, p);
            rcnt[p] = 0;
            for (u32 bb = first[me]; bb < first[me + 1]; ++bb) rcnt[p] += cnt[(size_t)p * nb0 + bb];
        }
    }
    // virtual buckets of MY receive buffer: vb = (b - first[me]) * P + p
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
    h_vbeg.back() = (u32)N_recv;  // with a single source the virtual buckets are contiguous and vbeg doubles as an offset array
    DBuf<u32> vbeg(ctx, h_vbeg.size()), vcnt(ctx, h_vcnt.size());
    CK(cudaMemcpyAsync(vbeg.p, h_vbeg.data(), h_vbeg.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(vcnt.p, h_vcnt.data(), h_vcnt.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    DBuf<u32> hist0, bstart;
    DBuf<u64> A, Rtmp;
    u64 *R = nullptr;
    if (W == 1 && dist_ensure_recv(cm, need)) {
        // ---- pass B fused with the exchange: tiles are written into the owners' buffers over NVLink ----
        MsdPeers pt;
        memset(&pt, 0, sizeof(pt));
        pt.n = P;
        for (int q = 0; q < P; ++q) { pt.ptr[q] = cm->peerR[q]; pt.first[q] = first[q]; }
        pt.first[P] = first[P];
        std::vector<u32> h_cur(nb0);  // where my keys of bucket b start inside the owner's buffer
        for (int q = 0; q < P; ++q) {
            u64 run = seg_off(q, me);
            for (u32 b = first[q]; b < first[q + 1]; ++b) { h_cur[b] = (u32)run; run += cnt[(size_t)me * nb0 + b]; }
        }
        DBuf<u32> pcur(ctx, nb0);
        DBuf<u64> tok(ctx, 1);
        CK(cudaMemcpyAsync(pcur.p, h_cur.data(), nb0 * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
        // the collectives of pass A ran after every rank's previous call on its stream: nobody still reads its buffer
        msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A, &pt, pcur.p);
        stage_end(ctx, ST_EXTRACT);
        mark();  // 3: pass B (scatter into the peers) done locally
        stage_begin(ctx, ST_SORT);
        CK(cudaMemsetAsync(tok.p, 0, sizeof(u64), ctx->stream));
        NCK(nc->AllReduce(tok.p, tok.p, 1, ncclUint64, ncclSum, cm->comm, ctx->stream));  // barrier: every rank's stores have landed
        CK(cudaStreamSynchronize(ctx->stream));  // h_vbeg / h_vcnt / h_cur are read by the copies above
        R = cm->R;
    } else {
        // ---- pass B locally (its output doubles as the send buffer), then one grouped ncclSend / ncclRecv exchange ----
        if constexpr (W == 1) msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A);
        else msdw_pass_b<W>(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A);
        stage_end(ctx, ST_EXTRACT);
        mark();  // 3: pass B done
        stage_begin(ctx, ST_SORT);
        Rtmp.alloc(ctx, (N_recv ? N_recv : 1) * W);
        dist_alltoallv(cm, A.p, soff.data(), scnt.data(), Rtmp.p, roff.data(), rcnt.data(), sizeof(u64) * W);
        CK(cudaStreamSynchronize(ctx->stream));  // h_vbeg / h_vcnt are read by the copies above; A is released next
        A.release();
        R = Rtmp.p;
    }
    stage_end(ctx, ST_SORT);
    mark();  // 4: exchange done
    DBuf<u32> child_hist;
    if (pl.b1 > 0 && pl.b0 + pl.b1 <= hb && nb_local) {  // pass C for free (see msd_count)
        const u32 nchild = nb_local << pl.b1;
        child_hist.alloc(ctx, nchild + 1);
        LAUNCH(ctx, dist_child_hist_kernel, (unsigned)ceil_div(nchild + 1, 256), 256, 0, all_hist.p, P, nbh, hb, pl.b0 + pl.b1,
               first[me] << pl.b1, nchild, child_hist.p);
    }
    u64 U = 0;
    if constexpr (W == 1) U = msd_finish(ctx, R, N_recv, vbeg.p, vcnt.p, nb_local, (u32)P, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
    else U = msdw_finish<W>(ctx, R, N_recv, vbeg.p, vcnt.p, nb_local, (u32)P, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
    mark();  // 5: level 1 + leaves + gather done
    if (dbg)
        fprintf(stderr, "[dist rank %d] t0=%.3f size-allreduce %.2f | passA+allgather %.2f | passB %.2f | exchange %.2f (send %.1f MB, recv %.1f MB) | finish %.2f ms\n",
                me, tmark[0], tmark[1] - tmark[0], tmark[2] - tmark[1], tmark[3] - tmark[2], tmark[4] - tmark[3],
                (double)(N_loc - scnt[me]) * 8e-6 * W, (double)(N_recv - rcnt[me]) * 8e-6 * W, tmark[5] - tmark[4]);
    return U;
}
