// umi.cuh -- UniqueMerIndex core: every k-mer of every node sequence, canonical key + strand,
// sorted by key, keys occurring exactly once kept, per-node unique / total counts.
// Replaces (reference) UniqueMerIndex{M}(graph) src/GraphIndexes.jl:62-102 and
// _discard_nonunique_mers! :43-60 (intended semantics; the upstream code is broken, see
// unique_mers_per_node is sized by n_nodes here).
#pragma once
#include "pack.cuh"
#include "sort.cuh"

struct gg_umi {
    CtxRef ctx;
    int k = 0, W = 0;
    u64 n_unique = 0, n_nodes = 0;
    u64 *d_keys = nullptr;   // n_unique * W
    u64 *d_pay = nullptr;    // n_unique : strand << 63 | node0 << 32 | pos1
    u64 *d_unique_per_node = nullptr, *d_total_per_node = nullptr;
};

// compaction functor: key + payload (strand, node, position) for every valid window
template <int W>
struct ExtractPayF {
    const u32 *valid;
    const u64 *packed;
    const u64 *offs;  // n_nodes + 1
    u64 n_nodes;
    u64 *out;
    u64 *pay;
    int k;
    __device__ bool pred(u64 p) const { return (valid[p >> 5] >> (31 - (p & 31))) & 1u; }
    __device__ void emit(u64 p, u64 dst) const {
        bool fwc;
        km_store<W>(out, dst, kmer_at<W>(packed, p, k, 1, &fwc));
        // node holding base p: last r with offs[r] <= p
        u64 lo = 0, hi = n_nodes;
        while (lo < hi) {
            u64 mid = lo + ((hi - lo) >> 1);
            if (offs[mid + 1] <= p) lo = mid + 1; else hi = mid;
        }
        u64 pos1 = p - offs[lo] + 1;
        pay[dst] = ((u64)(fwc ? 0 : 1) << 63) | (lo << 32) | (pos1 & 0xFFFFFFFFull);
    }
};
template <int W>
struct UniqueRunF {
    const u64 *keys;
    const u64 *pay;
    u64 n;
    u64 *okeys;
    u64 *opay;
    unsigned long long *unique_per_node;
    __device__ bool pred(u64 i) const {
        Kmer<W> x = km_load<W>(keys, i);
        if (i > 0 && km_eq<W>(x, km_load<W>(keys, i - 1))) return false;
        if (i + 1 < n && km_eq<W>(x, km_load<W>(keys, i + 1))) return false;
        return true;
    }
    __device__ void emit(u64 i, u64 dst) const {
        km_store<W>(okeys, dst, km_load<W>(keys, i));
        u64 p = pay[i];
        opay[dst] = p;
        atomicAdd(&unique_per_node[(p >> 32) & 0x7FFFFFFFull], 1ULL);
    }
};
__global__ void umi_totals_kernel(const u64 *__restrict__ offs, u64 nn, int k, u64 *__restrict__ total) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    u64 len = offs[i + 1] - offs[i];
    total[i] = len >= (u64)k ? len + 1 - (u64)k : 0;
}
__global__ void umi_unpack_kernel(const u64 *__restrict__ pay, u64 n, i64 *__restrict__ node, u32 *__restrict__ pos) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 p = pay[i];
    i64 id = (i64)((p >> 32) & 0x7FFFFFFFull) + 1;
    node[i] = (p >> 63) ? -id : id;
    pos[i] = (u32)p;
}

template <int W>
static gg_umi *build_umi(gg_ctx *ctx, const u8 *d_seq, const u64 *d_offs, u64 nn, u64 nbytes, int k) {
    gg_umi *u = new gg_umi();
    u->ctx = ctx; u->k = k; u->W = W; u->n_nodes = nn;
    try {
        if (nn >= (1ULL << 31)) GG_THROW(GG_ERR_LIMIT, "too many nodes for the index payload");
        u->d_unique_per_node = dalloc<u64>(ctx, nn + 1);
        u->d_total_per_node = dalloc<u64>(ctx, nn + 1);
        CK(cudaMemsetAsync(u->d_unique_per_node, 0, (nn + 1) * sizeof(u64), ctx->stream));
        if (nn) LAUNCH(ctx, umi_totals_kernel, (unsigned)ceil_div(nn, 256), 256, 0, d_offs, nn, k, u->d_total_per_node);
        PackedReads pr;
        pack_reads(ctx, d_seq, d_offs, nn, nbytes, pr);
        DBuf<u32> valid;
        window_valid_mask(ctx, pr, k, valid);
        ExtractPayF<W> ef{valid.p, pr.packed.p, d_offs, nn, nullptr, nullptr, k};
        Compactor<ExtractPayF<W>> cp(ctx, nbytes, "umi_extract");
        u64 n = cp.count(ef);
        DBuf<u64> keys(ctx, n * W), keys_alt(ctx, n * W), pay(ctx, n + 1), pay_alt(ctx, n + 1);
        ef.out = keys.p; ef.pay = pay.p;
        cp.emit(ef);
        u64 *a = keys.p, *b = keys_alt.p, *pa = pay.p, *pb = pay_alt.p;
        std::vector<BitRange> ranges = {{0, 2 * k}};
        radix_sort<W>(ctx, &a, &b, &pa, &pb, n, ranges);
        UniqueRunF<W> uf{a, pa, n, nullptr, nullptr, (unsigned long long *)u->d_unique_per_node};
        Compactor<UniqueRunF<W>> cu(ctx, n, "umi_unique");
        u64 nu = cu.count(uf);
        u->n_unique = nu;
        u->d_keys = dalloc<u64>(ctx, nu * W);
        u->d_pay = dalloc<u64>(ctx, nu);
        uf.okeys = u->d_keys; uf.opay = u->d_pay;
        cu.emit(uf);
        CK(cudaStreamSynchronize(ctx->stream));
        return u;
    } catch (...) {
        dfree(ctx, u->d_keys); dfree(ctx, u->d_pay); dfree(ctx, u->d_unique_per_node); dfree(ctx, u->d_total_per_node);
        delete u;
        throw;
    }
}
