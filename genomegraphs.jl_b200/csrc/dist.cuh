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
#include "skm.cuh"

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
    CtxRef ctx;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // receive buffer of the fused scatter + exchange, mapped into every peer with CUDA IPC
    u64 *R = nullptr;
    u64 cap = 0;                       // keys
    u64 *peerR[MSD_MAX_PEERS] = {};    // peerR[rank] == R
    int ipc_state = 0;                 // 0 untried, 1 usable, -1 unavailable (NCCL send/recv is used instead)
    u64 *agree = nullptr;              // status word of dist_guard (plain cudaMalloc)
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

// ---- error agreement -------------------------------------------------------------------------------
// A failure that only one rank sees (out of memory, a CUDA error) must not leave the peers waiting in the next
// collective: local phases run under dist_guard, which catches the failure, lets every rank learn the worst status
// with ONE ncclAllReduce(min) and then throws on all ranks together.  `f` must not contain collectives.
template <class F>
static void dist_guard(gg_comm *cm, const char *what, F f) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    GgError mine{GG_OK, ""};
    try { f(); } catch (const GgError &e) { mine = e; cudaGetLastError(); } catch (const std::bad_alloc &) { mine = GgError{GG_ERR_OOM, "host allocation failed"}; }
    if (!cm->agree) {   // one word of plain device memory outside the arena (the arena itself may be what failed)
        if (cudaMalloc((void **)&cm->agree, sizeof(u64)) != cudaSuccess) GG_THROW(GG_ERR_OOM, "no memory for the status word");
    }
    const u64 ok = mine.code == GG_OK ? 1 : 0;
    CK(cudaMemcpyAsync(cm->agree, &ok, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    NCK(nc->AllReduce(cm->agree, cm->agree, 1, ncclUint64, ncclMin, cm->comm, ctx->stream));
    u64 all = 0;
    CK(cudaMemcpyAsync(&all, cm->agree, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (mine.code != GG_OK) throw mine;
    if (!all) GG_THROW(GG_ERR_CUDA, "a peer rank failed in: %s", what);
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
    DBuf<u32> hist, all_hist;
    DBuf<u8> est;
    dist_guard(cm, "pass A", [&] {
        all_hist.alloc(ctx, (u64)P * nbh);
        if constexpr (W == 1) msd_pass_a_launch(ctx, pr, valid, k, canonical, est_shift, hist, est);
        else msdw_pass_a_launch<W>(ctx, pr, valid, k, canonical, est_shift, hist, est);
    });
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
    // (a limit only this rank hits must reach the peers before the next collective)
    dist_guard(cm, "instance count", [&] {
        if (N_loc >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu k-mer instances exceed the 2^32-1 per-GPU limit", (unsigned long long)N_loc);
    });
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
    DBuf<u32> vbeg, vcnt;
    dist_guard(cm, "bucket tables", [&] {
        vbeg.alloc(ctx, h_vbeg.size()); vcnt.alloc(ctx, h_vcnt.size());
        CK(cudaMemcpyAsync(vbeg.p, h_vbeg.data(), h_vbeg.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(vcnt.p, h_vcnt.data(), h_vcnt.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    });
    DBuf<u32> hist0, bstart;
    DBuf<u64> A, Rtmp;
    u64 *R = nullptr;
    if (dist_ensure_recv(cm, need * W)) {
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
        dist_guard(cm, "pass B (peer scatter)", [&] {
            if constexpr (W == 1) msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A, &pt, pcur.p);
            else msdw_pass_b<W>(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A, &pt, pcur.p);
        });
        stage_end(ctx, ST_EXTRACT);
        mark();  // 3: pass B (scatter into the peers) done locally
        stage_begin(ctx, ST_SORT);
        CK(cudaMemsetAsync(tok.p, 0, sizeof(u64), ctx->stream));
        NCK(nc->AllReduce(tok.p, tok.p, 1, ncclUint64, ncclSum, cm->comm, ctx->stream));  // barrier: every rank's stores have landed
        CK(cudaStreamSynchronize(ctx->stream));  // h_vbeg / h_vcnt / h_cur are read by the copies above
        R = cm->R;
    } else {
        // ---- pass B locally (its output doubles as the send buffer), then one grouped ncclSend / ncclRecv exchange ----
        dist_guard(cm, "pass B", [&] {
            if constexpr (W == 1) msd_pass_b(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A);
            else msdw_pass_b<W>(ctx, pr, valid, k, canonical, pl, hist.p, N_loc, hist0, bstart, A);
            Rtmp.alloc(ctx, (N_recv ? N_recv : 1) * W);
        });
        stage_end(ctx, ST_EXTRACT);
        mark();  // 3: pass B done
        stage_begin(ctx, ST_SORT);
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
    dist_guard(cm, "level 1 + leaves", [&] {
        if constexpr (W == 1) U = msd_finish(ctx, R, N_recv, vbeg.p, vcnt.p, nb_local, (u32)P, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
        else U = msdw_finish<W>(ctx, R, N_recv, vbeg.p, vcnt.p, nb_local, (u32)P, pl, tune, min_freq, ukeys, ucounts, child_hist.p);
    });
    mark();  // 5: level 1 + leaves + gather done
    if (dbg)
        fprintf(stderr, "[dist rank %d] t0=%.3f size-allreduce %.2f | passA+allgather %.2f | passB %.2f | exchange %.2f (send %.1f MB, recv %.1f MB) | finish %.2f ms\n",
                me, tmark[0], tmark[1] - tmark[0], tmark[2] - tmark[1], tmark[3] - tmark[2], tmark[4] - tmark[3],
                (double)(N_loc - scnt[me]) * 8e-6 * W, (double)(N_recv - rcnt[me]) * 8e-6 * W, tmark[5] - tmark[4]);
    return U;
}

// ---- super-k-mer route over several GPUs (skm.cuh) -------------------------------------------------
// The level-0 bins of the minimizer hash are owned by ranks (rank q: bins [first[q], first[q + 1]), equal shares: the
// bin digit is a hash).  Every source gets a PRIVATE region per bin inside the owner's receive buffer -- regions are
// source-major, r = p * nbl + bl, so a source's regions are one contiguous segment -- sized from that source's own
// sampled record histogram; the capacities are all-gathered, after which every rank computes the same layout.
// skm_scan_kernel<.., PEER> then stores its records straight into the owners' buffers over NVLink (CUDA IPC mappings);
// without peer mapping the same segments move with one grouped ncclSend / ncclRecv.  The all-gather of the per-region
// record counts that follows doubles as the barrier after which the buffer is consumed (level 1 + leaves).
struct SkmFirst { u32 first[MSD_MAX_PEERS + 1]; };
// one CTA per owner q: bases of its regions; this rank keeps base0_src[b] (where ITS records of bin b start inside the
// owner's buffer) and, for q == me, mybase[r] (+ the end sentinel); rtotal[q] = slots of q's buffer
__global__ void __launch_bounds__(1024) skm_dist_layout_kernel(const u32 *__restrict__ capall, u32 nb0, int P, int me, SkmFirst fs,
                                                               u32 *__restrict__ base0_src, u32 *__restrict__ mybase, u64 *__restrict__ rtotal) {
    __shared__ u32 ws[33];
    __shared__ u64 s_run;
    const int q = blockIdx.x;
    const u32 f0 = fs.first[q], nbl = fs.first[q + 1] - f0, nreg = nbl * (u32)P;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 r0 = 0; r0 < nreg; r0 += 1024) {
        const u32 r = r0 + threadIdx.x;
        u32 c = 0, p = 0, bl = 0;
        if (r < nreg) { p = r / nbl; bl = r % nbl; c = capall[(u64)p * nb0 + f0 + bl] / SKM_ALIGN; }
        u32 tot;
        const u32 ex = block_excl_scan(c, ws, &tot);
        if (r < nreg) {
            const u64 bb = (s_run + ex) * SKM_ALIGN;
            const u32 b32 = bb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)bb;
            if ((int)p == me) base0_src[f0 + bl] = b32;
            if (q == me) mybase[r] = b32;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const u64 bb = s_run * SKM_ALIGN;
        rtotal[q] = bb;
        if (q == me) mybase[nreg] = bb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)bb;
    }
}
// what a source tells the owners after its scatter: records per bin (dense), overflow flag, instance / record totals
#define SKM_DIST_TAIL 8
__global__ void skm_dist_pack_kernel(const u32 *__restrict__ cur0, const u32 *__restrict__ cap0, u32 nb0, const u32 *__restrict__ ovf,
                                     const u64 *__restrict__ totals, u32 *__restrict__ out) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb0) out[b] = min(cur0[(u64)b * SKM_CUR_STRIDE], cap0[b]);
    if (b == 0) {
        out[nb0] = *ovf;
        out[nb0 + 1] = (u32)totals[0]; out[nb0 + 2] = (u32)(totals[0] >> 32);
        out[nb0 + 3] = (u32)totals[1]; out[nb0 + 4] = (u32)(totals[1] >> 32);
        out[nb0 + 5] = out[nb0 + 6] = out[nb0 + 7] = 0;
    }
}
__global__ void skm_dist_mycnt_kernel(const u32 *__restrict__ allc, u32 row, u32 f0, u32 nbl, u32 nreg, u32 *__restrict__ mycnt) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nreg) mycnt[r] = allc[(u64)(r / nbl) * row + f0 + r % nbl];
}

// lb[b] = first index whose key's top digit is >= b (b = 0 .. nbh): the digit histogram of a SORTED key list
__global__ void range_lb_kernel(const u64 *__restrict__ keys, u64 n, int sh, u32 nbh, u32 *__restrict__ lb) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nbh) return;
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = lo + ((hi - lo) >> 1);
        if ((keys[mid] >> sh) < (u64)b) lo = mid + 1; else hi = mid;
    }
    lb[b] = (u32)lo;
}
// Second exchange: every rank holds the sorted kept (k-mer, count) pairs of ITS minimizer bins; afterwards rank r
// holds the r-th contiguous KEY RANGE of the global sorted table (ranges balanced on the all-gathered digit
// histogram), which is the reference's output order (Vector{MerCount} sorted by k-mer) cut into P pieces.
static u64 dist_range_repartition(gg_comm *cm, DBuf<u64> &uk, DBuf<u32> &uc, u64 U, int k, DBuf<u64> &ukeys, DBuf<u32> &ucounts) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world, me = cm->rank;
    if (P == 1) { ukeys = std::move(uk); ucounts = std::move(uc); return U; }
    const int tb = 2 * k < 14 ? 2 * k : 14, sh = 2 * k - tb;
    const u32 nbh = 1u << tb, row = nbh + 1;
    DBuf<u32> lb, all_lb;
    dist_guard(cm, "range histogram", [&] {
        lb.alloc(ctx, row); all_lb.alloc(ctx, (u64)P * row);
        LAUNCH(ctx, range_lb_kernel, (unsigned)ceil_div(row, 256), 256, 0, uk.p, U, sh, nbh, lb.p);
    });
    NCK(nc->AllGather(lb.p, all_lb.p, row, ncclUint32, cm->comm, ctx->stream));
    std::vector<u32> h((size_t)P * row);
    CK(cudaMemcpyAsync(h.data(), all_lb.p, h.size() * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<u64> gh(nbh, 0);
    for (int p = 0; p < P; ++p)
        for (u32 b = 0; b < nbh; ++b) gh[b] += h[(size_t)p * row + b + 1] - h[(size_t)p * row + b];
    std::vector<u32> first(P + 1);
    dist_splitters(gh.data(), nbh, P, first.data());
    std::vector<u64> soff(P), scnt(P), roff(P), rcnt(P);
    u64 n_recv = 0;
    for (int p = 0; p < P; ++p) {
        soff[p] = h[(size_t)me * row + first[p]];
        scnt[p] = h[(size_t)me * row + first[p + 1]] - soff[p];
        roff[p] = n_recv;
        rcnt[p] = h[(size_t)p * row + first[me + 1]] - h[(size_t)p * row + first[me]];
        n_recv += rcnt[p];
    }
    DBuf<u64> rk;
    DBuf<u32> rc;
    dist_guard(cm, "range exchange buffers", [&] {
        if (n_recv >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu kept k-mers exceed the 2^32-1 per-GPU limit", (unsigned long long)n_recv);
        rk.alloc(ctx, n_recv + 1); rc.alloc(ctx, n_recv + 1);
    });
    dist_alltoallv(cm, uk.p, soff.data(), scnt.data(), rk.p, roff.data(), rcnt.data(), sizeof(u64));
    dist_alltoallv(cm, uc.p, soff.data(), scnt.data(), rc.p, roff.data(), rcnt.data(), sizeof(u32));
    u64 Uout = 0;
    dist_guard(cm, "merge of the received runs", [&] {
        CK(cudaStreamSynchronize(ctx->stream));
        uk.release(); uc.release();
        // P sorted runs -> one (partition + bucket ordering on the bits below this rank's range prefix)
        const u64 kbase = (u64)first[me] << sh;
        const u64 span = (u64)(first[me + 1] - first[me]) << sh;
        int kbits = 0;
        while (kbits < 2 * k && (1ULL << kbits) < span) ++kbits;
        Uout = skm_sort_pairs(ctx, rk.p, rc.p, n_recv, n_recv, k, ukeys, ucounts, kbase, kbits);
    });
    return Uout;
}

// reads of this rank -> this rank's key-range slice of the global sorted (k-mer, count) table; skm_supported(k, canonical)
static u64 dist_skm_count(gg_comm *cm, const PackedReads &pr, const u32 *valid, int k, int canonical, u32 min_freq, DBuf<u64> &ukeys,
                          DBuf<u32> &ucounts, u64 *n_instances_local, u64 *n_instances_global) {
    NcclApi *nc = nccl_api();
    gg_ctx *ctx = cm->ctx;
    const int P = cm->world, me = cm->rank;
    if (P > MSD_MAX_PEERS) GG_THROW(GG_ERR_LIMIT, "at most %d ranks", MSD_MAX_PEERS);
    const SkmTune tune = skm_tune();
    stage_begin(ctx, ST_EXTRACT);
    // ---- global number of windows -> one geometry on every rank ----
    DBuf<u64> sz;
    u64 win_local = 0;
    dist_guard(cm, "window count", [&] {
        sz.alloc(ctx, 2);
        win_local = skm_count_windows(ctx, valid, 0, pr.nwords);
        CK(cudaMemcpyAsync(sz.p, &win_local, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    });
    NCK(nc->AllReduce(sz.p, sz.p, 1, ncclUint64, ncclSum, cm->comm, ctx->stream));
    const u64 win_glob = read_scalar<u64>(ctx, sz.p);
    *n_instances_local = win_local;
    *n_instances_global = win_glob;
    // rounds: leaf bins of ~bin_inst instances, at most 512 x 8192 of them per round -- larger jobs are counted in rounds over
    // slices of the minimizer digit (every round rescans the reads and exchanges only its own slice's records)
    u32 rounds = (u32)ceil_div(ceil_div(win_glob ? win_glob : 1, tune.bin_inst), (u64)SKM_MAX_NB0 * SKM_MAX_NB1 * 9 / 10);
    if (const char *e = getenv("GG_SKM_ROUNDS")) rounds = (u32)atoi(e);   // tests
    if (rounds < 1) rounds = 1;
    std::vector<DBuf<u64>> round_keys(rounds);
    std::vector<DBuf<u32>> round_counts(rounds);
    std::vector<u64> round_n(rounds, 0);
    for (u32 round = 0; round < rounds; ++round) {
    SkmGeom g = skm_geom(win_glob / rounds + 1, k, canonical, tune);
    g.rounds = rounds; g.round = round;
    const u32 nb0 = g.nb0, row = nb0 + SKM_DIST_TAIL;
    const u64 nleaf = (u64)nb0 * g.nb1;
    SkmFirst fs;
    for (int q = 0; q <= P; ++q) fs.first[q] = (u32)(((u64)q * nb0) / (u64)P);
    for (int q = P + 1; q <= MSD_MAX_PEERS; ++q) fs.first[q] = nb0;
    const u32 f0 = fs.first[me], nbl = fs.first[me + 1] - f0, nreg = nbl * (u32)P;
    const u64 ntiles = ceil_div(pr.nwords, skm_tile_words(k));
    SkmL0 L;
    L.g = g;
    DBuf<u32> capall, base_src, mybase, mine, allc;
    DBuf<u64> rtotal;
    DBuf<ulonglong2> sbuf, rtmp;
    std::vector<u64> h_rtotal(P);
    std::vector<u32> h_allc;
    const ulonglong2 *R = nullptr;
    for (int attempt = 0;; ++attempt) {
        const u32 stride = (ntiles >= 256 && attempt == 0) ? tune.sample : 1;
        const bool exact = stride == 1 && !(attempt == 0 && tune.force_cap);
        // ---- sample: my records per bin -> capacities of my regions; all-gathered -> the layout of every owner's buffer ----
        dist_guard(cm, "super-k-mer sample", [&] {
            L.alloc_sample(ctx);
            prof_bytes(ctx, 12.0 * (double)pr.nwords / stride);
            skm_scan_launch<0, false>(ctx, pr, valid, 0, pr.nwords, stride, g, L.hist0.p, nullptr, nullptr, nullptr, nullptr, nullptr, L.totals.p, nullptr, nullptr);
            L.base0.alloc(ctx, nb0 + 1); L.cap0.alloc(ctx, nb0); L.total_cap.alloc(ctx, 1);
            LAUNCH(ctx, skm_caps_kernel, 1, 1024, 0, L.hist0.p, nb0, (float)stride, exact ? 1 : 0, L.base0.p, L.cap0.p, L.total_cap.p, exact ? 0u : tune.force_cap);
            capall.alloc(ctx, (u64)P * nb0); base_src.alloc(ctx, nb0); mybase.alloc(ctx, (u64)nreg + 1); rtotal.alloc(ctx, P);
            mine.alloc(ctx, row); allc.alloc(ctx, (u64)P * row);
        });
        NCK(nc->AllGather(L.cap0.p, capall.p, nb0, ncclUint32, cm->comm, ctx->stream));
        LAUNCH(ctx, skm_dist_layout_kernel, (unsigned)P, 1024, 0, capall.p, nb0, P, me, fs, base_src.p, mybase.p, rtotal.p);
        CK(cudaMemcpyAsync(h_rtotal.data(), rtotal.p, P * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        u64 need = 1;
        for (int q = 0; q < P; ++q) need = h_rtotal[q] > need ? h_rtotal[q] : need;
        if (need >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu super-k-mer records exceed the 2^32-1 per-GPU limit", (unsigned long long)need);  // same on all ranks
        const bool ipc = dist_ensure_recv(cm, 2 * need);   // collective; the buffer holds u64 words, a record is two
        // ---- scatter: records go to the owner of their minimizer's bin ----
        u64 local_cap = 0;
        dist_guard(cm, "super-k-mer scatter", [&] {
            L.cur0.alloc(ctx, (u64)nb0 * SKM_CUR_STRIDE); L.hist1.alloc(ctx, nleaf + 1); L.ovf.alloc(ctx, 1);
            dfill(ctx, L.cur0.p, (u64)nb0 * SKM_CUR_STRIDE * sizeof(u32), 0u);
            dfill(ctx, L.hist1.p, (nleaf + 1) * sizeof(u32), 0u);
            dfill(ctx, L.ovf.p, sizeof(u32), 0u);
            dfill(ctx, L.totals.p, 2 * sizeof(u64), 0u);
            prof_bytes(ctx, 12.0 * (double)pr.nwords);
            if (ipc) {
                SkmPeers pt;
                memset(&pt, 0, sizeof(pt));
                pt.n = P;
                pt.bulk = getenv("GG_SKM_NO_BULK") ? 0 : 1;
                for (int q = 0; q < P; ++q) pt.ptr[q] = reinterpret_cast<ulonglong2 *>(cm->peerR[q]);
                for (int q = 0; q <= MSD_MAX_PEERS; ++q) pt.first[q] = fs.first[q];
                skm_scan_launch<1, true>(ctx, pr, valid, 0, pr.nwords, 1, g, nullptr, L.cur0.p, base_src.p, L.cap0.p, nullptr, L.hist1.p, L.totals.p, L.ovf.p, &pt);
            } else {   // my segments next to each other in a send buffer (the local layout of skm_caps_kernel: bins in order = owners in order)
                local_cap = read_scalar<u64>(ctx, L.total_cap.p);
                if (local_cap >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu super-k-mer records exceed the 2^32-1 per-GPU limit", (unsigned long long)local_cap);
                sbuf.alloc(ctx, local_cap ? local_cap : 1);
                skm_scan_launch<1, false>(ctx, pr, valid, 0, pr.nwords, 1, g, nullptr, L.cur0.p, L.base0.p, L.cap0.p, sbuf.p, L.hist1.p, L.totals.p, L.ovf.p, nullptr);
            }
            LAUNCH(ctx, skm_dist_pack_kernel, (unsigned)ceil_div(nb0, 256), 256, 0, L.cur0.p, L.cap0.p, nb0, L.ovf.p, L.totals.p, mine.p);
        });
        // ---- what every source stored where (IPC route: also the barrier after which all stores have landed) ----
        NCK(nc->AllGather(mine.p, allc.p, row, ncclUint32, cm->comm, ctx->stream));
        h_allc.resize((size_t)P * row);
        CK(cudaMemcpyAsync(h_allc.data(), allc.p, h_allc.size() * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        bool any_ovf = false;
        for (int p = 0; p < P; ++p) any_ovf = any_ovf || h_allc[(size_t)p * row + nb0] != 0;
        if (any_ovf) {   // a region outgrew its sampled share somewhere: everybody repeats with exact counts
            if (exact || attempt > 0) GG_THROW(GG_ERR_CUDA, "super-k-mer scatter overflowed exact capacities");
            sbuf.release();
            continue;
        }
        if (ipc) R = reinterpret_cast<const ulonglong2 *>(cm->R);
        else {
            std::vector<u32> h_lbase(nb0 + 1), h_my((size_t)nreg + 1);
            dist_guard(cm, "receive buffer", [&] {
                CK(cudaMemcpyAsync(h_lbase.data(), L.base0.p, (nb0 + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(h_my.data(), mybase.p, ((size_t)nreg + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                rtmp.alloc(ctx, h_rtotal[me] ? h_rtotal[me] : 1);
            });
            std::vector<u64> soff(P), scnt(P), roff(P), rcnt(P);
            for (int q = 0; q < P; ++q) {
                soff[q] = h_lbase[fs.first[q]];
                scnt[q] = (u64)h_lbase[fs.first[q + 1]] - soff[q];
                roff[q] = h_my[(size_t)q * nbl];
                rcnt[q] = (u64)h_my[(size_t)(q + 1) * nbl] - roff[q];
            }
            dist_alltoallv(cm, sbuf.p, soff.data(), scnt.data(), rtmp.p, roff.data(), rcnt.data(), sizeof(ulonglong2));
            CK(cudaStreamSynchronize(ctx->stream));
            sbuf.release();
            R = rtmp.p;
        }
        break;
    }
    stage_end(ctx, ST_EXTRACT);
    gg_trace(ctx, "dist skm scatter + exchange");
    // ---- leaf-bin sizes: the sources' histograms summed ----
    NCK(nc->AllReduce(L.hist1.p, L.hist1.p, nleaf, ncclUint32, ncclSum, cm->comm, ctx->stream));
    u64 n_rec_recv = 0;
    for (int p = 0; p < P; ++p)
        for (u32 bl = 0; bl < nbl; ++bl) n_rec_recv += h_allc[(size_t)p * row + f0 + bl];
    DBuf<u64> uk;
    DBuf<u32> uc;
    u64 U = 0;
    dist_guard(cm, "super-k-mer counting", [&] {
        DBuf<u32> mycnt(ctx, (u64)nreg + 1);
        if (nreg) LAUNCH(ctx, skm_dist_mycnt_kernel, (unsigned)ceil_div(nreg, 256), 256, 0, allc.p, row, f0, nbl, nreg, mycnt.p);
        // instances behind the received records: counted by level 1 (a single rank knows its own total, and may skip level 1)
        const u64 inst_me = (u64)h_allc[(size_t)me * row + nb0 + 1] | ((u64)h_allc[(size_t)me * row + nb0 + 2] << 32);
        U = skm_finish(ctx, R, mybase.p, mycnt.p, 1, nbl, (u32)P, h_rtotal[me], g, L.hist1.p + (u64)f0 * g.nb1, P == 1 ? inst_me : ~0ULL, n_rec_recv, min_freq,
                       tune, uk, uc);
        CK(cudaStreamSynchronize(ctx->stream));
    });
    rtmp.release();
    round_keys[round] = std::move(uk); round_counts[round] = std::move(uc); round_n[round] = U;
    }   // rounds
    DBuf<u64> uk;
    DBuf<u32> uc;
    u64 U = round_n[0];
    if (rounds == 1) { uk = std::move(round_keys[0]); uc = std::move(round_counts[0]); }
    else {   // the rounds' tables are disjoint and each is sorted: one ordering pass over their concatenation
        dist_guard(cm, "merge of the rounds", [&] {
            U = 0;
            for (u32 r = 0; r < rounds; ++r) U += round_n[r];
            DBuf<u64> ck(ctx, U + 1);
            DBuf<u32> cc(ctx, U + 1);
            u64 at = 0;
            for (u32 r = 0; r < rounds; ++r) {
                if (round_n[r]) {
                    CK(cudaMemcpyAsync(ck.p + at, round_keys[r].p, round_n[r] * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
                    CK(cudaMemcpyAsync(cc.p + at, round_counts[r].p, round_n[r] * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
                }
                at += round_n[r];
            }
            CK(cudaStreamSynchronize(ctx->stream));
            for (u32 r = 0; r < rounds; ++r) { round_keys[r].release(); round_counts[r].release(); }
            U = skm_sort_pairs(ctx, ck.p, cc.p, U, U, k, uk, uc);
        });
    }
    stage_begin(ctx, ST_FILTER);   // the second exchange is booked under "filter" (with the all-gather of the graph stage)
    U = dist_range_repartition(cm, uk, uc, U, k, ukeys, ucounts);
    stage_end(ctx, ST_FILTER);
    return U;
}
