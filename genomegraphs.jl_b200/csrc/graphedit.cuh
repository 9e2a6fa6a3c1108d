// graphedit.cuh -- tip removal on the device-resident graph (SURVEY.md section 8 f, rank 3).
//
// Replaces (reference):
//   remove_tips!           src/remove_tips.jl:2-30      (find tips, remove_node! each, collapse_all_unitigs!(.., 2, true), repeat)
//   remove_node!           src/Graphs.jl:459-466        (remove_link! :487-495 drops every equal entry, order kept)
//   collapse_all_unitigs!  src/Graphs.jl:528-539  =  find_all_unitigs! :831-862  +  join_path!(p, consume = true) :1040-1088
//                          (sequence(path) :1006-1037, check_overlap :991-1003, reverse!(path) :975-988, add_link! :473-476)
//
// The reference walks the nodes one by one, joins one path after the other and pushes links as it goes; the result
// -- ids of the new nodes, their orientation, the ORDER of the link entries at every node -- depends on that sequence.
// Here the same result is computed without a sequence (tests/collapse_model.py is an executable model of this file,
// checked against the sequential restatement oracle/graph_edit.py on fuzzed graphs):
//
//  * oriented nodes (vertex 2(n-1) = node n forwards, +1 = backwards) with succ(v) = the sole forward neighbour when it
//    has no other backward neighbour form disjoint paths and cycles that come in mirror pairs (links made by add_link!
//    are symmetric, a hairpin link appears twice at its node, so no chain is its own mirror image);
//  * pointer jumping (16-byte records {pointer, smallest vertex, vertices, bases}, double buffered) finds the cycles and
//    their smallest vertex; a cycle is cut where the reference's walk starts (its smallest node id, forwards) and the
//    mirror cycle at the mirror place; two more passes give every vertex the prefix / suffix aggregates of its path;
//  * a path is collapsed when it has >= min_nodes nodes; the orientation that holds its smallest node id forwards is the
//    one the reference finds; new node ids = ascending smallest node id (a scan over flags);
//  * bases: every vertex of a chosen path writes its bases minus the overlap with its predecessor at its prefix offset
//    (the overlap is compared with the predecessor's last bases: "Sequences do not overlap." otherwise), one warp per
//    new node decides iscanonical from both ends and reverse-complements in place;
//  * links: every old entry (s, d, dist) maps to (phi(s), phi(d), dist), phi = the new node's first / last end for the
//    two outer ends of a path, identity for kept nodes; entries with an inner end and entries between two ends on the
//    same side of one path disappear; an entry between the first and the last end of one path becomes a self link of
//    the new node.  Entries between two kept nodes stay in place (a scan); the place of every other entry in its row
//    follows from a closed form of the push order of join_path! (rows of
//    kept nodes: old entries in place, then one block per joined path in joining order; rows of new nodes: first-end
//    loop, last-end loop -- each over old entries, then entries pushed by earlier joins -- then blocks of later joins),
//    packed into a 128-bit key and ordered with one stable radix sort.
//
// Limits: symmetric links as add_link! makes them (checked: GG_ERR_ARG otherwise), consume = true, min_nodes >= 2,
// < 2^31 nodes, < 2^32 bases and link entries, < 2^30 entries per node, ACGT bases in collapsed nodes.
#pragma once
#include "graph.cuh"

#define GE_NIL 0xFFFFFFFFu
#define GE_BADTRIM 0x80000000u
enum { GE_E_ASYM = 1, GE_E_SELF = 2, GE_E_ROW = 4, GE_E_OVLEN = 8, GE_E_OVSEQ = 16, GE_E_BASE = 32, GE_E_ROWLEN = 64 };

__device__ __forceinline__ u32 ge_vtx(i64 x) { return x < 0 ? (u32)(2 * (-x - 1) + 1) : (u32)(2 * (x - 1)); }
__device__ __forceinline__ u8 ge_comp(u8 c, u32 *err) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
    }
    atomicOr(err, (u32)GE_E_BASE);
    return c;
}
// base j of vertex v read in v's direction
__device__ __forceinline__ u8 ge_base(const u8 *__restrict__ bases, const u64 *__restrict__ node_off, u32 v, u64 j, u32 *err) {
    const u64 a = node_off[v >> 1], b = node_off[(v >> 1) + 1];
    return (v & 1u) ? ge_comp(bases[b - 1 - j], err) : bases[a + j];
}

// entries with source == -n leave node n forwards (src/Graphs.jl:188), entries with source == n leave it backwards;
// the backward degree of a vertex is the forward degree of its mirror vertex
__global__ void ge_degree_kernel(const u64 *__restrict__ row, const i64 *__restrict__ tri, u64 nn, u32 *__restrict__ outdeg,
                                 u32 *__restrict__ dvtx, i64 *__restrict__ ddist) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const i64 n = (i64)(i + 1);
    u32 fw = 0, bw = 0, dfw = GE_NIL, dbw = GE_NIL;
    i64 sfw = 0, sbw = 0;
    for (u64 q = row[i]; q < row[i + 1]; ++q) {
        const i64 s = tri[3 * q];
        if (s == -n) { if (!fw) { dfw = ge_vtx(tri[3 * q + 1]); sfw = tri[3 * q + 2]; } ++fw; }
        if (s == n) { if (!bw) { dbw = ge_vtx(tri[3 * q + 1]); sbw = tri[3 * q + 2]; } ++bw; }
    }
    outdeg[2 * i] = fw; dvtx[2 * i] = dfw; ddist[2 * i] = sfw;
    outdeg[2 * i + 1] = bw; dvtx[2 * i + 1] = dbw; ddist[2 * i + 1] = sbw;
}
// place (inside the row of |d|) of the entry add_link! pushed together with entry q: the t-th (s, d) of its row belongs
// to the t-th (d, s) of the other row.  Also checks the shape of the link table.  err[1] = longest row.
__global__ void ge_mirror_kernel(const u64 *__restrict__ row, const i64 *__restrict__ tri, u64 ne, u64 nn, u32 *__restrict__ mirror,
                                 u32 *__restrict__ err) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ne) return;
    const i64 s = tri[3 * q], d = tri[3 * q + 1];
    const u64 i = (u64)(s < 0 ? -s : s) - 1, j = (u64)(d < 0 ? -d : d) - 1;
    mirror[q] = 0;
    if (i >= nn || j >= nn || q < row[i] || q >= row[i + 1]) { atomicOr(err, (u32)GE_E_ROW); return; }
    if (row[i + 1] - row[i] >= (1ULL << 30)) { atomicOr(err, (u32)GE_E_ROWLEN); return; }
    if (q == row[i]) atomicMax(err + 1, (u32)(row[i + 1] - row[i]));
    u32 t = 0, tot = 0;
    for (u64 p = row[i]; p < row[i + 1]; ++p)
        if (tri[3 * p] == s && tri[3 * p + 1] == d) { if (p < q) ++t; ++tot; }
    if (s == d && (tot & 1u)) atomicOr(err, (u32)GE_E_SELF);
    u32 c = 0, found = GE_NIL;
    for (u64 p = row[j]; p < row[j + 1]; ++p)
        if (tri[3 * p] == d && tri[3 * p + 1] == s) { if (c == t) found = (u32)(p - row[j]); ++c; }
    if (found == GE_NIL || c != tot) { atomicOr(err, (u32)GE_E_ASYM); return; }
    mirror[q] = found;
}
__global__ void ge_succ_kernel(const u32 *__restrict__ outdeg, const u32 *__restrict__ dvtx, u64 nv, u32 *__restrict__ succ, u32 *__restrict__ pred) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 s = GE_NIL;
    if (outdeg[v] == 1) {
        const u32 d = dvtx[v];
        if (outdeg[d ^ 1u] == 1) s = d;   // length(backward_links(dest)) == 1, src/Graphs.jl:845
    }
    succ[v] = s;
    if (s != GE_NIL) pred[s] = (u32)v;    // succ is injective
}
// record = {pointer, smallest vertex, vertices, bases} of the window [v, pointer)
__global__ void ge_rec_init_kernel(const u32 *__restrict__ ptr, const u32 *__restrict__ wgt, u64 nv, uint4 *__restrict__ rec) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    rec[v] = make_uint4(ptr[v], (u32)v, 1u, wgt ? wgt[v] : 0u);
}
// `active` (optional): set when some pointer has not run out yet after this round
__global__ void ge_jump_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, u64 nv, u32 *__restrict__ active) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    uint4 a = in[v];
    if (a.x != GE_NIL) {
        const uint4 b = in[a.x];
        a.x = b.x;
        a.y = a.y < b.y ? a.y : b.y;
        a.z += b.z;
        a.w += b.w;
        if (active && a.x != GE_NIL) *active = 1u;
    }
    out[v] = a;
}
// a vertex whose pointer never ran out sits on a cycle; the cycle's smallest vertex cuts it: in front of itself when it
// is a forward vertex (the reference's walk starts there, src/Graphs.jl:834-840), behind itself otherwise (mirror cycle)
__global__ void ge_cut_kernel(const uint4 *__restrict__ rec, u64 nv, u32 *__restrict__ succ, u32 *__restrict__ pred) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint4 a = rec[v];
    if (a.x == GE_NIL || a.y != (u32)v) return;
    if ((v & 1u) == 0) {
        const u32 p = pred[v];
        succ[p] = GE_NIL;
        pred[v] = GE_NIL;
    } else {
        const u32 s = succ[v];
        pred[s] = GE_NIL;
        succ[v] = GE_NIL;
    }
}
// bases a vertex adds to its path: its length minus the overlap with its predecessor (dist <= 0: overlap -dist,
// src/Graphs.jl:1017-1028; dist > 0: the whole node)
__global__ void ge_weight_kernel(const u64 *__restrict__ node_off, const u32 *__restrict__ pred, const i64 *__restrict__ ddist, u64 nv,
                                 u32 *__restrict__ trim, u32 *__restrict__ wgt) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const u64 ln = node_off[(v >> 1) + 1] - node_off[v >> 1];
    u32 t = 0;
    const u32 p = pred[v];
    if (p != GE_NIL) {
        const i64 ds = ddist[p];
        if (ds <= 0) {
            const u64 lp = node_off[(p >> 1) + 1] - node_off[p >> 1];
            const u64 o = (u64)(-ds);
            t = (o > ln || o > lp) ? GE_BADTRIM : (u32)o;
        }
    }
    trim[v] = t;
    wgt[v] = (u32)ln - (t == GE_BADTRIM ? 0u : t);
}
// rs / rp: suffix / prefix records of every vertex (paths only by now).  cmin = smallest vertex of the path (GE_NIL when
// the path stays), coff = bases in front of the vertex, flag / ulen per node: this node numbers a new node of ulen bases
__global__ void ge_chain_kernel(const uint4 *__restrict__ rs, const uint4 *__restrict__ rp, const u32 *__restrict__ wgt,
                                const u32 *__restrict__ trim, u64 nv, u32 min_nodes, u32 *__restrict__ cmin, u32 *__restrict__ coff,
                                u32 *__restrict__ flag, u32 *__restrict__ ulen, u32 *__restrict__ err) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint4 a = rs[v], b = rp[v];
    const u32 mn = a.y < b.y ? a.y : b.y;
    const u32 len = a.z + b.z - 1;
    const bool in = len >= min_nodes;
    cmin[v] = in ? mn : GE_NIL;
    coff[v] = b.w - wgt[v];
    if (in && trim[v] == GE_BADTRIM) atomicOr(err, (u32)GE_E_OVLEN);
    if ((v & 1u) == 0) {
        const bool numbers = in && mn == (u32)v;
        flag[v >> 1] = numbers ? 1u : 0u;
        ulen[v >> 1] = numbers ? a.w + b.w - wgt[v] : 0u;
    }
}
// lengths of the nodes of the new graph: collapsed nodes become empty (the reference's empty_node), new nodes follow
__global__ void ge_newlen_kernel(const u64 *__restrict__ node_off, const u32 *__restrict__ cmin, const u32 *__restrict__ flag,
                                 const u32 *__restrict__ uidx, const u32 *__restrict__ ulen, u64 nn, u32 *__restrict__ lens) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    lens[i] = cmin[2 * i] != GE_NIL ? 0u : (u32)(node_off[i + 1] - node_off[i]);
    if (flag[i]) lens[nn + uidx[i]] = ulen[i];
}
// one warp per vertex: kept nodes are copied (forward vertex), vertices of a chosen path write their share of the new node
__global__ void ge_emit_kernel(const u64 *__restrict__ node_off, const u8 *__restrict__ bases, const u32 *__restrict__ cmin,
                               const u32 *__restrict__ coff, const u32 *__restrict__ trim, const u32 *__restrict__ pred,
                               const u32 *__restrict__ uidx, u64 nn, const u64 *__restrict__ new_off, u8 *__restrict__ out, u32 *__restrict__ err) {
    const u64 v = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31u;
    if (v >= 2 * nn) return;
    const u64 i = v >> 1;
    const u64 a = node_off[i], ln = node_off[i + 1] - a;
    const u32 mn = cmin[v];
    if (mn == GE_NIL) {
        if ((v & 1u) == 0) {
            const u64 o = new_off[i];
            for (u64 j = lane; j < ln; j += 32) out[o + j] = bases[a + j];
        }
        return;
    }
    if (mn & 1u) return;   // the mirror image of a chosen path
    u32 t = trim[v];
    if (t == GE_BADTRIM) t = 0;   // reported by ge_chain_kernel
    const u64 o = new_off[nn + uidx[mn >> 1]] + coff[v];
    const u32 p = pred[v];
    const u64 lp = p != GE_NIL ? node_off[(p >> 1) + 1] - node_off[p >> 1] : 0;
    for (u64 j = lane; j < ln; j += 32) {
        const u8 b = ge_base(bases, node_off, (u32)v, j, err);
        if (j < t) {
            if (ge_base(bases, node_off, p, lp - t + j, err) != b) atomicOr(err, (u32)GE_E_OVSEQ);   // check_overlap, src/Graphs.jl:991-1003
        } else {
            out[o + j - t] = b;
        }
    }
}
// one warp per new node: iscanonical from both ends (ties keep the sequence), reverse complement in place otherwise
// (join_path!, src/Graphs.jl:1046-1051)
__global__ void ge_canon_kernel(const u64 *__restrict__ new_off, u64 nn, u64 m, u8 *__restrict__ out, u32 *__restrict__ flip, u32 *__restrict__ err) {
    const u64 u = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31u;
    if (u >= m) return;
    u8 *s = out + new_off[nn + u];
    const u64 L = new_off[nn + u + 1] - new_off[nn + u];
    const u64 half = (L + 1) / 2;   // i <= j
    u32 fl = 0;
    for (u64 base = 0; base < half; base += 32) {
        const u64 i = base + lane;
        u32 st = 0;   // 0 equal, 1 forward smaller, 2 reverse complement smaller
        if (i < half) {
            const u8 f = s[i], r = ge_comp(s[L - 1 - i], err);
            st = f < r ? 1u : (r < f ? 2u : 0u);
        }
        const u32 any = __ballot_sync(FULL_MASK, st != 0);
        if (any) {
            const int first = __ffs(any) - 1;
            fl = __shfl_sync(FULL_MASK, st, first) == 2u ? 1u : 0u;
            break;
        }
    }
    if (lane == 0) flip[u] = fl;
    if (!fl) return;
    __syncwarp();
    for (u64 i = lane; i < half; i += 32) {
        const u64 j = L - 1 - i;
        const u8 x = s[i], y = s[j];
        s[i] = ge_comp(y, err);
        s[j] = ge_comp(x, err);   // i == j: the middle base, complemented once
    }
}

// what an entry's end becomes: 0 kept node, 1 outer end of new node u (side 0: its first end, 1: its last end), 2 inner end
struct GeEnds {
    const u32 *cmin, *pred, *uidx, *flip;
    u64 nn;
    __device__ __forceinline__ int type(i64 x, u32 &u, u32 &side) const {
        const u32 v = ge_vtx(x);
        const u32 mn = cmin[v];
        u = 0; side = 0;
        if (mn == GE_NIL) return 0;
        if (pred[v] != GE_NIL) return 2;
        u = uidx[mn >> 1];
        side = (mn & 1u) ^ flip[u];   // head of the chosen orientation = first end, unless the path was flipped
        return 1;
    }
    __device__ __forceinline__ i64 phi(i64 x, int ty, u32 u, u32 side) const {
        if (ty == 0) return x;
        const i64 id = (i64)(nn + 1 + u);
        return side == 0 ? id : -id;
    }
};
// class A (keep_a): entry between two kept nodes -- it stays in its row and the order among such entries is the old one;
// class B (keep_b): every other surviving entry -- these get a key and are sorted
__global__ void ge_link_flag_kernel(const i64 *__restrict__ tri, u64 ne, GeEnds e, u32 *__restrict__ keep_a, u32 *__restrict__ keep_b) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ne) return;
    u32 us, ss, ud, sd;
    const int ts = e.type(tri[3 * q], us, ss), td = e.type(tri[3 * q + 1], ud, sd);
    u32 k = 1;
    if (ts == 2 || td == 2) k = 0;
    if (ts == 1 && td == 1 && us == ud && ss == sd) k = 0;
    const bool a = ts == 0 && td == 0;
    keep_a[q] = (k && a) ? 1u : 0u;
    keep_b[q] = (k && !a) ? 1u : 0u;
}
// key word 0 = row << 4 | seg << 3 | S << 2 | sub, word 1 = j << 32 | sx << 31 | sy << 30 | q  (see tests/collapse_model.py)
__global__ void ge_link_emit_kernel(const u64 *__restrict__ row, const i64 *__restrict__ tri, const u32 *__restrict__ mirror, u64 ne, GeEnds e,
                                    const u32 *__restrict__ keep_b, const u32 *__restrict__ pos_b, u64 *__restrict__ keys, u64 *__restrict__ pay,
                                    i64 *__restrict__ tmp) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ne || !keep_b[q]) return;
    const i64 s = tri[3 * q], d = tri[3 * q + 1];
    const u64 i = (u64)(s < 0 ? -s : s) - 1;
    const u64 qs = q - row[i], qm = mirror[q];
    u32 us, ss, ud, sd;
    const int ts = e.type(s, us, ss), td = e.type(d, ud, sd);
    const i64 fs = e.phi(s, ts, us, ss), fd = e.phi(d, td, ud, sd);
    const u64 r = (u64)(fs < 0 ? -fs : fs) - 1;
    u64 seg = 0, S = 0, sub = 0, j = 0, sx = 0, sy = 0, qq = 0;
    if (ts == 0) { seg = 1; j = ud; sx = sd; qq = qm; }                 // pushed to a kept node when path ud is joined
    else if (td == 0) { S = ss; qq = qs; }                              // creation loops of the new node, old entries first
    else if (ud == us) { S = 1; sub = 2; j = ss == 0 ? qs : qm; qq = ss == 0 ? 1 : 0; }   // first end -- last end: (-N, N) then (N, -N)
    else if (ud < us) { S = ss; sub = 1; j = ud; sx = sd; qq = qm; }    // entries an earlier join had pushed
    else { seg = 1; j = ud; sx = sd; sy = ss; qq = qs; }                // pushed when the later path ud is joined
    const u64 dst = pos_b[q];
    keys[2 * dst] = (r << 4) | (seg << 3) | (S << 2) | sub;
    keys[2 * dst + 1] = (j << 32) | (sx << 31) | (sy << 30) | qq;
    pay[dst] = dst;
    tmp[3 * dst] = fs; tmp[3 * dst + 1] = fd; tmp[3 * dst + 2] = tri[3 * q + 2];
}
// first_b[r] = index of the first sorted class B entry of row r (r = 0 .. nn2)
__global__ void ge_first_b_kernel(const u64 *__restrict__ keys, u64 nb, u64 nn2, u32 *__restrict__ first_b) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nb) return;
    const u64 r1 = t < nb ? (keys[2 * t] >> 4) : nn2;          // rows (r0 .. r1) start at t
    const u64 r0 = t > 0 ? (keys[2 * (t - 1)] >> 4) + 1 : 0;
    for (u64 r = r0; r <= r1; ++r) first_b[r] = (u32)t;
}
// row r of the new graph starts at (class A entries of the rows before r) + first_b[r]; its class A entries come first
// in their old order, its sorted class B entries follow
struct GePlace {
    const u64 *row;       // old row offsets, nn + 1
    const u32 *pos_a;     // exclusive scan of keep_a, ne
    const u32 *first_b;   // nn2 + 1
    u64 nn, ne, na;
    __device__ __forceinline__ u64 a_before(u64 r) const {
        if (r > nn) return na;
        const u64 p = row[r];
        return p < ne ? (u64)pos_a[p] : na;
    }
};
__global__ void ge_link_place_kernel(GePlace g, const i64 *__restrict__ tri, const u32 *__restrict__ keep_a, const u64 *__restrict__ keys,
                                     const u64 *__restrict__ pay, const i64 *__restrict__ tmp, u64 nb, u64 nn2, u64 *__restrict__ new_row,
                                     i64 *__restrict__ out) {
    const u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x <= nn2) new_row[x] = g.a_before(x) + g.first_b[x];
    if (x < g.ne && keep_a[x]) {
        const i64 s = tri[3 * x];
        const u64 i = (u64)(s < 0 ? -s : s) - 1;
        const u64 dst = (u64)g.first_b[i] + g.pos_a[x];
        out[3 * dst] = s; out[3 * dst + 1] = tri[3 * x + 1]; out[3 * dst + 2] = tri[3 * x + 2];
    }
    if (x < nb) {
        const u64 r = keys[2 * x] >> 4, p = pay[x];
        const u64 dst = g.a_before(r + 1) + x;
        out[3 * dst] = tmp[3 * p]; out[3 * dst + 1] = tmp[3 * p + 1]; out[3 * dst + 2] = tmp[3 * p + 2];
    }
}

// ---- remove_node! for a set of nodes --------------------------------------------------------------------------------
__global__ void ge_mark_ids_kernel(const i64 *__restrict__ ids, u64 n, u32 *__restrict__ dead) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const i64 x = ids[t];
    dead[(u64)(x < 0 ? -x : x) - 1] = 1u;
}
__global__ void ge_dead_len_kernel(const u64 *__restrict__ node_off, const u32 *__restrict__ dead, u64 nn, u32 *__restrict__ lens) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) lens[i] = dead[i] ? 0u : (u32)(node_off[i + 1] - node_off[i]);
}
__global__ void ge_dead_copy_kernel(const u64 *__restrict__ node_off, const u8 *__restrict__ bases, const u64 *__restrict__ new_off, u64 nn,
                                    u8 *__restrict__ out) {
    const u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31u;
    if (i >= nn) return;
    const u64 a = node_off[i], o = new_off[i], ln = new_off[i + 1] - o;   // 0 for a removed node
    for (u64 j = lane; j < ln; j += 32) out[o + j] = bases[a + j];
}
__global__ void ge_dead_flag_kernel(const i64 *__restrict__ tri, const u32 *__restrict__ dead, u64 ne, u32 *__restrict__ keep) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ne) return;
    const i64 s = tri[3 * q], d = tri[3 * q + 1];
    keep[q] = (dead[(u64)(s < 0 ? -s : s) - 1] || dead[(u64)(d < 0 ? -d : d) - 1]) ? 0u : 1u;
}
__global__ void ge_dead_links_kernel(const u64 *__restrict__ row, const i64 *__restrict__ tri, const u32 *__restrict__ keep,
                                     const u32 *__restrict__ pos, u64 ne, u64 nl, u64 nn, u64 *__restrict__ new_row, i64 *__restrict__ out) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= nn) {
        const u64 r = row[q];
        new_row[q] = r < ne ? pos[r] : nl;
    }
    if (q < ne && keep[q]) {
        const u64 p = pos[q];
        out[3 * p] = tri[3 * q]; out[3 * p + 1] = tri[3 * q + 1]; out[3 * p + 2] = tri[3 * q + 2];
    }
}

static void ge_free_graph(gg_ctx *ctx, gg_graph *g) {
    dfree(ctx, g->d_node_off); dfree(ctx, g->d_bases); dfree(ctx, g->d_link_row); dfree(ctx, g->d_link_tri);
    delete g;
}
static gg_graph *ge_clone_graph(gg_ctx *ctx, const gg_graph *g) {
    gg_graph *h = new gg_graph();
    h->ctx = ctx; h->k = g->k; h->n_nodes = g->n_nodes; h->total_bases = g->total_bases; h->n_link_entries = g->n_link_entries;
    try {
        h->d_node_off = dalloc<u64>(ctx, g->n_nodes + 1);
        h->d_link_row = dalloc<u64>(ctx, g->n_nodes + 1);
        h->d_bases = dalloc<u8>(ctx, g->total_bases + 16);
        h->d_link_tri = dalloc<i64>(ctx, 3 * (g->n_link_entries ? g->n_link_entries : 1));
        CK(cudaMemcpyAsync(h->d_node_off, g->d_node_off, (g->n_nodes + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(h->d_link_row, g->d_link_row, (g->n_nodes + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        if (g->total_bases) CK(cudaMemcpyAsync(h->d_bases, g->d_bases, g->total_bases, cudaMemcpyDeviceToDevice, ctx->stream));
        if (g->n_link_entries) CK(cudaMemcpyAsync(h->d_link_tri, g->d_link_tri, 3 * g->n_link_entries * sizeof(i64), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        ge_free_graph(ctx, h);
        throw;
    }
    return h;
}
static void ge_check_limits(const gg_graph *g) {
    if (g->n_nodes >= (1ULL << 31) - 1) GG_THROW(GG_ERR_LIMIT, "%llu nodes exceed the 2^31 limit of the graph editing kernels", (unsigned long long)g->n_nodes);
    if (g->total_bases >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu node bases exceed the 2^32 limit", (unsigned long long)g->total_bases);
    if (g->n_link_entries >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu link entries exceed the 2^32 limit", (unsigned long long)g->n_link_entries);
}

// nodes with dead[i] != 0 are emptied, every entry that touches one is dropped, the order of the others is kept
static gg_graph *ge_remove_dead(gg_ctx *ctx, const gg_graph *g, const u32 *dead) {
    ge_check_limits(g);
    const u64 nn = g->n_nodes, ne = g->n_link_entries;
    gg_graph *h = new gg_graph();
    h->ctx = ctx; h->k = g->k; h->n_nodes = nn;
    try {
        h->d_node_off = dalloc<u64>(ctx, nn + 1);
        h->d_link_row = dalloc<u64>(ctx, nn + 1);
        DBuf<u32> lens(ctx, nn + 1), keep(ctx, ne + 1), pos(ctx, ne + 1);
        DBuf<u64> tot(ctx, 2);
        u64 nb = 0, nl = 0;
        if (nn) {
            LAUNCH(ctx, ge_dead_len_kernel, (unsigned)ceil_div(nn, 256), 256, 0, g->d_node_off, dead, nn, lens.p);
            exclusive_scan_u32(ctx, lens.p, lens.p, nn, tot.p);
            nb = read_scalar<u64>(ctx, tot.p);
        }
        LAUNCH(ctx, widen_offsets_kernel, (unsigned)ceil_div(nn + 1, 256), 256, 0, lens.p, nn, nb, h->d_node_off);
        h->total_bases = nb;
        h->d_bases = dalloc<u8>(ctx, nb + 16);
        if (nn) LAUNCH(ctx, ge_dead_copy_kernel, (unsigned)ceil_div(nn * 32, 256), 256, 0, g->d_node_off, g->d_bases, h->d_node_off, nn, h->d_bases);
        if (ne) {
            LAUNCH(ctx, ge_dead_flag_kernel, (unsigned)ceil_div(ne, 256), 256, 0, g->d_link_tri, dead, ne, keep.p);
            exclusive_scan_u32(ctx, keep.p, pos.p, ne, tot.p + 1);
            nl = read_scalar<u64>(ctx, tot.p + 1);
        }
        h->n_link_entries = nl;
        h->d_link_tri = dalloc<i64>(ctx, 3 * (nl ? nl : 1));
        const u64 span = (ne > nn + 1 ? ne : nn + 1);
        LAUNCH(ctx, ge_dead_links_kernel, (unsigned)ceil_div(span, 256), 256, 0, g->d_link_row, g->d_link_tri, keep.p, pos.p, ne, nl, nn, h->d_link_row, h->d_link_tri);
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        ge_free_graph(ctx, h);
        throw;
    }
    return h;
}

// `rounds` = ceil(log2(vertices)) covers the longest possible chain; every fourth round asks whether any pointer is still
// running and stops when none is (paths of a unitig graph are short; only cycles keep their pointers for ever)
static void ge_jump(gg_ctx *ctx, uint4 **a, uint4 **b, u64 nv, int rounds, u32 *d_active) {
    for (int r = 0; r < rounds; ++r) {
        const bool ask = d_active && (r & 3) == 3 && r + 1 < rounds;
        if (ask) CK(cudaMemsetAsync(d_active, 0, sizeof(u32), ctx->stream));
        prof_bytes(ctx, 48.0 * (double)nv);
        LAUNCH(ctx, ge_jump_kernel, (unsigned)ceil_div(nv, 256), 256, 0, *a, *b, nv, ask ? d_active : (u32 *)nullptr);
        std::swap(*a, *b);
        if (ask && read_scalar<u32>(ctx, d_active) == 0) break;
    }
}
static void ge_raise(u32 e) {
    if (e & GE_E_ROW) GG_THROW(GG_ERR_ARG, "a link entry is not stored at its source node");
    if (e & GE_E_ROWLEN) GG_THROW(GG_ERR_LIMIT, "a node has 2^30 or more link entries");
    if (e & GE_E_ASYM) GG_THROW(GG_ERR_ARG, "links are not symmetric (every (a, b) entry needs its (b, a) entry, src/Graphs.jl:473-476)");
    if (e & GE_E_SELF) GG_THROW(GG_ERR_ARG, "a link from a node end to itself must appear twice (src/Graphs.jl:473-476)");
    if (e & (GE_E_OVLEN | GE_E_OVSEQ)) GG_THROW(GG_ERR_ARG, "Sequences do not overlap.");   // src/Graphs.jl:1023
    if (e & GE_E_BASE) GG_THROW(GG_ERR_ARG, "a collapsed node holds a base other than A, C, G, T");
}

// collapse_all_unitigs!(sg, min_nodes, true): a new graph handle; *n_new = number of nodes appended
static gg_graph *ge_collapse(gg_ctx *ctx, const gg_graph *g, u64 min_nodes, u64 *n_new) {
    if (min_nodes < 2) GG_THROW(GG_ERR_ARG, "min_nodes must be at least 2");
    ge_check_limits(g);
    const u64 nn = g->n_nodes, ne = g->n_link_entries, nv = 2 * nn;
    if (n_new) *n_new = 0;
    if (nn == 0 || ne == 0) return ge_clone_graph(ctx, g);
    const u32 mn32 = min_nodes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)min_nodes;
    DBuf<u32> err(ctx, 4);
    CK(cudaMemsetAsync(err.p, 0, 4 * sizeof(u32), ctx->stream));
    DBuf<u32> outdeg(ctx, nv), dvtx(ctx, nv), mirror(ctx, ne), succ(ctx, nv), pred(ctx, nv);
    DBuf<i64> ddist(ctx, nv);
    LAUNCH(ctx, ge_degree_kernel, (unsigned)ceil_div(nn, 256), 256, 0, g->d_link_row, g->d_link_tri, nn, outdeg.p, dvtx.p, ddist.p);
    LAUNCH(ctx, ge_mirror_kernel, (unsigned)ceil_div(ne, 256), 256, 0, g->d_link_row, g->d_link_tri, ne, nn, mirror.p, err.p);
    u32 herr[2];
    CK(cudaMemcpyAsync(herr, err.p, sizeof(herr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ge_raise(herr[0]);
    const u64 maxrow = herr[1];
    dfill(ctx, pred.p, nv * sizeof(u32), GE_NIL);
    LAUNCH(ctx, ge_succ_kernel, (unsigned)ceil_div(nv, 256), 256, 0, outdeg.p, dvtx.p, nv, succ.p, pred.p);
    int rounds = ilog2_ceil(nv);
    if (rounds < 1) rounds = 1;
    DBuf<uint4> ra(ctx, nv), rb(ctx, nv), rc(ctx, nv);
    uint4 *a = ra.p, *b = rb.p;
    // pass 1: cycles
    LAUNCH(ctx, ge_rec_init_kernel, (unsigned)ceil_div(nv, 256), 256, 0, succ.p, (const u32 *)nullptr, nv, a);
    ge_jump(ctx, &a, &b, nv, rounds, err.p + 2);
    LAUNCH(ctx, ge_cut_kernel, (unsigned)ceil_div(nv, 256), 256, 0, a, nv, succ.p, pred.p);
    // pass 2: prefix and suffix aggregates of the paths
    DBuf<u32> trim(ctx, nv), wgt(ctx, nv);
    LAUNCH(ctx, ge_weight_kernel, (unsigned)ceil_div(nv, 256), 256, 0, g->d_node_off, pred.p, ddist.p, nv, trim.p, wgt.p);
    LAUNCH(ctx, ge_rec_init_kernel, (unsigned)ceil_div(nv, 256), 256, 0, succ.p, wgt.p, nv, a);
    ge_jump(ctx, &a, &b, nv, rounds, err.p + 2);
    uint4 *rs = a;             // suffix records; b is scratch
    uint4 *c = rc.p;
    LAUNCH(ctx, ge_rec_init_kernel, (unsigned)ceil_div(nv, 256), 256, 0, pred.p, wgt.p, nv, b);
    ge_jump(ctx, &b, &c, nv, rounds, err.p + 2);
    uint4 *rp = b;
    DBuf<u32> cmin(ctx, nv), coff(ctx, nv), flag(ctx, nn), uidx(ctx, nn), ulen(ctx, nn);
    DBuf<u64> tot(ctx, 2);
    LAUNCH(ctx, ge_chain_kernel, (unsigned)ceil_div(nv, 256), 256, 0, rs, rp, wgt.p, trim.p, nv, mn32, cmin.p, coff.p, flag.p, ulen.p, err.p);
    exclusive_scan_u32(ctx, flag.p, uidx.p, nn, tot.p);
    const u64 m = read_scalar<u64>(ctx, tot.p);
    if (m == 0) return ge_clone_graph(ctx, g);
    ge_raise(read_scalar<u32>(ctx, err.p));
    const u64 nn2 = nn + m;
    if (nn2 >= (1ULL << 31) - 1) GG_THROW(GG_ERR_LIMIT, "%llu nodes exceed the 2^31 limit of the graph editing kernels", (unsigned long long)nn2);
    gg_graph *h = new gg_graph();
    h->ctx = ctx; h->k = g->k; h->n_nodes = nn2;
    try {
        h->d_node_off = dalloc<u64>(ctx, nn2 + 1);
        h->d_link_row = dalloc<u64>(ctx, nn2 + 1);
        // ---- bases
        DBuf<u32> lens(ctx, nn2 + 1);
        LAUNCH(ctx, ge_newlen_kernel, (unsigned)ceil_div(nn, 256), 256, 0, g->d_node_off, cmin.p, flag.p, uidx.p, ulen.p, nn, lens.p);
        exclusive_scan_u32(ctx, lens.p, lens.p, nn2, tot.p);
        const u64 nb = read_scalar<u64>(ctx, tot.p);
        if (nb >= 0xFFFFFFFFull) GG_THROW(GG_ERR_LIMIT, "%llu node bases exceed the 2^32 limit", (unsigned long long)nb);
        LAUNCH(ctx, widen_offsets_kernel, (unsigned)ceil_div(nn2 + 1, 256), 256, 0, lens.p, nn2, nb, h->d_node_off);
        h->total_bases = nb;
        h->d_bases = dalloc<u8>(ctx, nb + 16);
        prof_bytes(ctx, (double)g->total_bases + (double)nb);
        LAUNCH(ctx, ge_emit_kernel, (unsigned)ceil_div(nv * 32, 256), 256, 0, g->d_node_off, g->d_bases, cmin.p, coff.p, trim.p, pred.p, uidx.p, nn, h->d_node_off,
               h->d_bases, err.p);
        DBuf<u32> flip(ctx, m);
        LAUNCH(ctx, ge_canon_kernel, (unsigned)ceil_div(m * 32, 256), 256, 0, h->d_node_off, nn, m, h->d_bases, flip.p, err.p);
        ge_raise(read_scalar<u32>(ctx, err.p));
        // ---- links
        GeEnds ends{cmin.p, pred.p, uidx.p, flip.p, nn};
        DBuf<u32> keep_a(ctx, ne), keep_b(ctx, ne), pos_a(ctx, ne), pos_b(ctx, ne), first_b(ctx, nn2 + 1);
        LAUNCH(ctx, ge_link_flag_kernel, (unsigned)ceil_div(ne, 256), 256, 0, g->d_link_tri, ne, ends, keep_a.p, keep_b.p);
        exclusive_scan_u32(ctx, keep_a.p, pos_a.p, ne, tot.p);
        exclusive_scan_u32(ctx, keep_b.p, pos_b.p, ne, tot.p + 1);
        u64 hn[2];
        CK(cudaMemcpyAsync(hn, tot.p, sizeof(hn), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const u64 na = hn[0], nb2 = hn[1], nl = na + nb2;
        h->n_link_entries = nl;
        h->d_link_tri = dalloc<i64>(ctx, 3 * (nl ? nl : 1));
        DBuf<u64> keys(ctx, 2 * (nb2 + 1)), keys_alt(ctx, 2 * (nb2 + 1)), pay(ctx, nb2 + 1), pay_alt(ctx, nb2 + 1);
        DBuf<i64> tmp(ctx, 3 * (nb2 + 1));
        u64 *k0 = keys.p, *k1 = keys_alt.p, *p0 = pay.p, *p1 = pay_alt.p;
        if (nb2) {
            LAUNCH(ctx, ge_link_emit_kernel, (unsigned)ceil_div(ne, 256), 256, 0, g->d_link_row, g->d_link_tri, mirror.p, ne, ends, keep_b.p, pos_b.p, keys.p, pay.p, tmp.p);
            const u64 jmax = m > maxrow ? m : maxrow;
            int qbits = ilog2_ceil(maxrow + 2);
            if (qbits > 30) qbits = 30;
            std::vector<BitRange> ranges = {{0, qbits}, {30, 32}, {32, 32 + ilog2_ceil(jmax + 2)}, {64, 64 + 4 + ilog2_ceil(nn2 + 1)}};
            radix_sort<2>(ctx, &k0, &k1, &p0, &p1, nb2, ranges);
        }
        if (nb2) LAUNCH(ctx, ge_first_b_kernel, (unsigned)ceil_div(nb2 + 1, 256), 256, 0, k0, nb2, nn2, first_b.p);
        else dfill(ctx, first_b.p, (nn2 + 1) * sizeof(u32), 0u);
        GePlace place{g->d_link_row, pos_a.p, first_b.p, nn, ne, na};
        u64 span = nn2 + 1;
        if (ne > span) span = ne;
        LAUNCH(ctx, ge_link_place_kernel, (unsigned)ceil_div(span, 256), 256, 0, place, g->d_link_tri, keep_a.p, k0, p0, tmp.p, nb2, nn2, h->d_link_row, h->d_link_tri);
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        ge_free_graph(ctx, h);
        throw;
    }
    if (n_new) *n_new = m;
    return h;
}
