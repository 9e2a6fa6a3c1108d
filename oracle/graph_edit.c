/*
 * graph_edit.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (part of libggoracle.so).
 *
 * Plain C restatement, function by function and sequential like the reference, of the graph editing behind
 * remove_tips! of BioJulia/GenomeGraphs.jl v0.2.0 (the same functions as oracle/graph_edit.py, which stays as the
 * readable second statement; tests compare the two).  It exists so that the CPU path can be TIMED next to the CUDA
 * one on graphs of millions of nodes and so that full-size results can be compared (tests/full_size_tips_check.py).
 *
 *   remove_tips!            src/remove_tips.jl:2-30
 *   find_tip_nodes!         src/Graphs.jl:778-802
 *   remove_node!            src/Graphs.jl:459-466      remove_link!  :487-495  (== on links ignores dist, :183-185)
 *   forward/backward_links  src/Graphs.jl:667-690      (is_forwards_from: l.source == -n, :188)
 *   find_link               src/Graphs.jl:643-651
 *   collapse_all_unitigs!   src/Graphs.jl:528-539      find_all_unitigs!  :831-862   reverse!(path)  :975-988
 *   sequence(path)          src/Graphs.jl:1006-1037    check_overlap  :991-1003
 *   join_path!              src/Graphs.jl:1040-1088    add_node! :418-444   add_link! :473-476
 *
 * PARITY UNPINNED, like the rest of oracle/ (no reference output can be produced in this project); recalled upstream
 * semantics: iscanonical(seq) of BioSequences 2 compares seq[i] with complement(seq[j]) from both ends, ties keep the
 * sequence; A < C < G < T.  The reference allocates a vector per forward_links call; here the same entries are
 * visited in the same order without the copy, except where the reference's copy is a SNAPSHOT that later pushes must
 * not change (remove_node!'s oldlinks, the two loops of join_path!).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct { int64_t s, d, dist; } ge_link;
typedef struct { ge_link *v; uint64_t n, cap; } ge_links;
typedef struct { uint8_t *seq; uint64_t len; int deleted; } ge_node;
typedef struct {
    ge_node *nodes;
    ge_links *links;
    uint64_t n, cap;
    int error; /* 1: "Sequences do not overlap."  2: "This path is not a valid path." */
} ge_graph;

static uint64_t ge_abs(int64_t x) { return (uint64_t)(x < 0 ? -x : x); }
static void ge_links_push(ge_links *l, ge_link x) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 4; l->v = (ge_link *)realloc(l->v, l->cap * sizeof(ge_link)); }
    l->v[l->n++] = x;
}
static ge_links *ge_linksof(ge_graph *g, int64_t n) { return &g->links[ge_abs(n) - 1]; }

/* add_node! src/Graphs.jl:418-444 (takes ownership of seq) */
static int64_t ge_add_node(ge_graph *g, uint8_t *seq, uint64_t len) {
    if (g->n == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 16;
        g->nodes = (ge_node *)realloc(g->nodes, g->cap * sizeof(ge_node));
        g->links = (ge_links *)realloc(g->links, g->cap * sizeof(ge_links));
    }
    g->nodes[g->n].seq = seq; g->nodes[g->n].len = len; g->nodes[g->n].deleted = 0;
    g->links[g->n].v = NULL; g->links[g->n].n = 0; g->links[g->n].cap = 0;
    return (int64_t)(++g->n);
}
/* add_link! src/Graphs.jl:473-476 */
static void ge_add_link(ge_graph *g, int64_t src, int64_t dst, int64_t dist) {
    ge_link a = { src, dst, dist }, b = { dst, src, dist };
    ge_links_push(ge_linksof(g, src), a);
    ge_links_push(ge_linksof(g, dst), b);
}
/* filter!(!isequal(Link(src, dst, 0)), links): order kept, every equal entry goes */
static void ge_filter(ge_links *l, int64_t src, int64_t dst) {
    uint64_t w = 0;
    for (uint64_t i = 0; i < l->n; ++i)
        if (!(l->v[i].s == src && l->v[i].d == dst)) l->v[w++] = l->v[i];
    l->n = w;
}
/* remove_link! src/Graphs.jl:487-495 */
static void ge_remove_link(ge_graph *g, int64_t src, int64_t dst) {
    ge_filter(ge_linksof(g, src), src, dst);
    ge_filter(ge_linksof(g, dst), dst, src);
}
/* remove_node! src/Graphs.jl:459-466 */
static void ge_remove_node(ge_graph *g, int64_t n) {
    ge_links *l = ge_linksof(g, n);
    uint64_t cnt = l->n;
    ge_link *old = (ge_link *)malloc((cnt ? cnt : 1) * sizeof(ge_link));   /* oldlinks = copy(linksof(sg, n)) */
    memcpy(old, l->v, cnt * sizeof(ge_link));
    for (uint64_t i = 0; i < cnt; ++i) ge_remove_link(g, old[i].s, old[i].d);
    free(old);
    ge_node *nd = &g->nodes[ge_abs(n) - 1];
    free(nd->seq);
    nd->seq = NULL; nd->len = 0; nd->deleted = 1;                          /* empty_node(S) */
}
/* length(forward_links(sg, n)) and its first entry; backward_links(sg, n) = forward_links(sg, -n)  (src/Graphs.jl:667-690) */
static uint64_t ge_fw_count(ge_graph *g, int64_t n, ge_link *first) {
    ge_links *l = ge_linksof(g, n);
    uint64_t c = 0;
    for (uint64_t i = 0; i < l->n; ++i)
        if (l->v[i].s == -n) { if (c == 0 && first) *first = l->v[i]; ++c; }
    return c;
}
/* find_tip_nodes! src/Graphs.jl:778-802: ids ascending (a Set upstream); returns the count */
static uint64_t ge_find_tips(ge_graph *g, uint64_t min_size, int64_t *out) {
    uint64_t nt = 0;
    for (uint64_t i = 0; i < g->n; ++i) {
        int64_t n = (int64_t)(i + 1);
        ge_node *nd = &g->nodes[i];
        if (nd->deleted || nd->len > min_size) continue;
        ge_link f, b;
        uint64_t fw = ge_fw_count(g, n, &f), bw = ge_fw_count(g, -n, &b);
        int tip = 0;
        if (fw == 1 && bw == 0 && ge_fw_count(g, -f.d, NULL) > 1) tip = 1;   /* backward_links(destination(first(fwl))) */
        if (fw == 0 && bw == 1 && ge_fw_count(g, -b.d, NULL) > 1) tip = 1;   /* forward_links(-destination(first(bwl))) */
        if (fw == 0 && bw == 0) tip = 1;
        if (tip) out[nt++] = n;
    }
    return nt;
}

typedef struct { int64_t *v; uint64_t n, cap; } ge_path;
static void ge_path_push(ge_path *p, int64_t x) {
    if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 8; p->v = (int64_t *)realloc(p->v, p->cap * sizeof(int64_t)); }
    p->v[p->n++] = x;
}
/* reverse!(path) src/Graphs.jl:975-988 */
static void ge_path_reverse(ge_path *p) {
    for (uint64_t i = 0, j = p->n; i < j; ++i) {
        --j;
        int64_t x = p->v[i], y = p->v[j];
        p->v[i] = -y; p->v[j] = -x;
        if (i == j) p->v[i] = -x;
    }
}
static uint8_t ge_comp(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
/* sequence(sg, n) src/Graphs.jl:592-600, appended to s minus its first `skip` bases */
static void ge_append_node(ge_graph *g, int64_t n, uint64_t skip, uint8_t **s, uint64_t *len, uint64_t *cap) {
    ge_node *nd = &g->nodes[ge_abs(n) - 1];
    if (*len + nd->len + 1 > *cap) { *cap = (*len + nd->len + 1) * 2; *s = (uint8_t *)realloc(*s, *cap); }
    for (uint64_t j = skip; j < nd->len; ++j)
        (*s)[(*len)++] = n > 0 ? nd->seq[j] : ge_comp(nd->seq[nd->len - 1 - j]);
}
static uint8_t ge_node_base(ge_graph *g, int64_t n, uint64_t j) {
    ge_node *nd = &g->nodes[ge_abs(n) - 1];
    return n > 0 ? nd->seq[j] : ge_comp(nd->seq[nd->len - 1 - j]);
}
/* sequence(path) src/Graphs.jl:1006-1037; NULL + g->error on the reference's two error() calls */
static uint8_t *ge_path_sequence(ge_graph *g, const ge_path *p, uint64_t *out_len) {
    uint8_t *s = NULL;
    uint64_t len = 0, cap = 0;
    int64_t pnode = 0;
    for (uint64_t t = 0; t < p->n; ++t) {
        int64_t n = p->v[t];
        uint64_t skip = 0;
        if (pnode != 0) {
            ge_links *l = ge_linksof(g, pnode);            /* find_link(sg, pnode, n) src/Graphs.jl:643-651 */
            const ge_link *lnk = NULL;
            for (uint64_t i = 0; i < l->n && !lnk; ++i)
                if (l->v[i].s == pnode && l->v[i].d == n) lnk = &l->v[i];
            if (!lnk) { g->error = 2; free(s); return NULL; }
            if (lnk->dist <= 0) {
                uint64_t ovl = (uint64_t)(-lnk->dist), nl = g->nodes[ge_abs(n) - 1].len;
                if (ovl > len || ovl > nl) { g->error = 1; free(s); return NULL; }
                for (uint64_t j = 0; j < ovl; ++j)          /* check_overlap src/Graphs.jl:991-1003 */
                    if (s[len - ovl + j] != ge_node_base(g, n, j)) { g->error = 1; free(s); return NULL; }
                skip = ovl;
            }
        }
        ge_append_node(g, n, skip, &s, &len, &cap);
        pnode = -n;
    }
    if (!s) s = (uint8_t *)malloc(1);
    *out_len = len;
    return s;
}
static int ge_iscanonical(const uint8_t *s, uint64_t len) {
    if (len == 0) return 1;
    for (uint64_t i = 0, j = len - 1; i <= j; ++i, --j) {
        uint8_t f = s[i], r = ge_comp(s[j]);
        if (f < r) return 1;
        if (r < f) return 0;
        if (j == 0) break;
    }
    return 1;
}
/* join_path!(p, consume = true) src/Graphs.jl:1040-1088.  stamp[] marks the nodes of the path (pnodes holds +n and -n of
 * every path node, so membership only depends on |n|). */
static int64_t ge_join_path(ge_graph *g, ge_path *p, uint64_t *stamp, uint64_t tag) {
    for (uint64_t t = 0; t < p->n; ++t) stamp[ge_abs(p->v[t]) - 1] = tag;
    uint64_t len = 0;
    uint8_t *ps = ge_path_sequence(g, p, &len);
    if (!ps) return 0;
    if (!ge_iscanonical(ps, len)) {
        for (uint64_t i = 0, j = len; i < j; ++i) {             /* reverse_complement!(ps) */
            --j;
            uint8_t x = ps[i], y = ps[j];
            ps[i] = ge_comp(y); ps[j] = ge_comp(x);
        }
        ge_path_reverse(p);
    }
    int64_t newid = ge_add_node(g, ps, len);
    int64_t first = p->v[0], last = p->v[p->n - 1];
    for (int pass = 0; pass < 2; ++pass) {
        /* backward_links(first(p)) / forward_links(last(p)) are vectors built BEFORE the pushes of their loop */
        int64_t end = pass == 0 ? first : -last;               /* entries with source == end */
        ge_links *l = ge_linksof(g, end);
        uint64_t cnt = 0;
        ge_link *snap = (ge_link *)malloc((l->n ? l->n : 1) * sizeof(ge_link));
        for (uint64_t i = 0; i < l->n; ++i)
            if (l->v[i].s == end) snap[cnt++] = l->v[i];
        for (uint64_t i = 0; i < cnt; ++i) ge_add_link(g, pass == 0 ? newid : -newid, snap[i].d, snap[i].dist);
        free(snap);
    }
    /* consume: for n in pnodes (+ and - of every path node; the outcome does not depend on the order of the Set) */
    for (uint64_t t = 0; t < p->n; ++t)
        for (int sgn = 0; sgn < 2; ++sgn) {
            int64_t n = sgn == 0 ? p->v[t] : -p->v[t];
            int ext = 0;
            ge_links *l = ge_linksof(g, n);
            if (n != last)
                for (uint64_t i = 0; i < l->n; ++i)
                    if (l->v[i].s == -n) { uint64_t a = ge_abs(l->v[i].d); if (a > g->n || stamp[a - 1] != tag) ext = 1; }
            if (n != first)
                for (uint64_t i = 0; i < l->n; ++i)
                    if (l->v[i].s == n) { uint64_t a = ge_abs(l->v[i].d); if (a > g->n || stamp[a - 1] != tag) ext = 1; }
            if (ext) continue;
            ge_remove_node(g, n);
        }
    return newid;
}
/* collapse_all_unitigs!(unitigs, newnodes, sg, min_nodes, true) src/Graphs.jl:528-539 with find_all_unitigs! :831-862;
 * returns the number of new nodes, or -1 with g->error set */
static int64_t ge_collapse(ge_graph *g, uint64_t min_nodes) {
    uint64_t n0 = g->n, np = 0, pcap = 0;
    ge_path *paths = NULL;
    uint8_t *consumed = (uint8_t *)calloc(n0 ? n0 : 1, 1);
    for (uint64_t i = 0; i < n0; ++i) {
        int64_t n = (int64_t)(i + 1);
        if (consumed[i] || g->nodes[i].deleted) continue;
        consumed[i] = 1;
        ge_path p = { NULL, 0, 0 };
        ge_path_push(&p, n);
        for (int pass = 0; pass < 2; ++pass) {
            ge_link f;
            while (ge_fw_count(g, p.v[p.n - 1], &f) == 1) {
                int64_t dest = f.d;
                if (!consumed[ge_abs(dest) - 1] && ge_fw_count(g, -dest, NULL) == 1) {
                    ge_path_push(&p, dest);
                    consumed[ge_abs(dest) - 1] = 1;
                } else {
                    break;
                }
            }
            ge_path_reverse(&p);
        }
        if (p.n >= min_nodes) {
            if (np == pcap) { pcap = pcap ? pcap * 2 : 16; paths = (ge_path *)realloc(paths, pcap * sizeof(ge_path)); }
            paths[np++] = p;
        } else {
            free(p.v);
        }
    }
    free(consumed);
    /* the stamp array must cover the nodes added while joining: at most np of them */
    uint64_t *stamp = (uint64_t *)calloc(n0 + np + 1, sizeof(uint64_t));
    int64_t made = 0;
    for (uint64_t t = 0; t < np; ++t) {
        if (made >= 0 && ge_join_path(g, &paths[t], stamp, t + 1) != 0) ++made;
        else made = -1;
        free(paths[t].v);
    }
    free(paths);
    free(stamp);
    return made;
}

/* ---------------------------------------------- C interface (ctypes) ---------------------------------------------- */
EXPORT void *ggo_edit_graph_new(uint64_t n_nodes, const uint64_t *node_off, const uint8_t *bases, const uint64_t *link_row,
                                const int64_t *tri) {
    ge_graph *g = (ge_graph *)calloc(1, sizeof(ge_graph));
    for (uint64_t i = 0; i < n_nodes; ++i) {
        uint64_t len = node_off[i + 1] - node_off[i];
        uint8_t *s = (uint8_t *)malloc(len ? len : 1);
        memcpy(s, bases + node_off[i], len);
        ge_add_node(g, s, len);
        if (len == 0) { free(g->nodes[i].seq); g->nodes[i].seq = NULL; g->nodes[i].deleted = 1; }   /* empty_node */
    }
    for (uint64_t i = 0; i < n_nodes; ++i)
        for (uint64_t q = link_row[i]; q < link_row[i + 1]; ++q) {
            ge_link l = { tri[3 * q], tri[3 * q + 1], tri[3 * q + 2] };
            ge_links_push(&g->links[i], l);
        }
    return g;
}
EXPORT void ggo_edit_graph_free(void *h) {
    ge_graph *g = (ge_graph *)h;
    if (!g) return;
    for (uint64_t i = 0; i < g->n; ++i) { free(g->nodes[i].seq); free(g->links[i].v); }
    free(g->nodes); free(g->links); free(g);
}
EXPORT void ggo_edit_graph_sizes(void *h, uint64_t *n_nodes, uint64_t *total_bases, uint64_t *n_link_entries) {
    ge_graph *g = (ge_graph *)h;
    uint64_t nb = 0, nl = 0;
    for (uint64_t i = 0; i < g->n; ++i) { nb += g->nodes[i].len; nl += g->links[i].n; }
    *n_nodes = g->n; *total_bases = nb; *n_link_entries = nl;
}
EXPORT void ggo_edit_graph_export(void *h, uint64_t *node_off, uint8_t *bases, uint64_t *link_row, int64_t *tri) {
    ge_graph *g = (ge_graph *)h;
    uint64_t nb = 0, nl = 0;
    for (uint64_t i = 0; i < g->n; ++i) {
        node_off[i] = nb; link_row[i] = nl;
        memcpy(bases + nb, g->nodes[i].seq, g->nodes[i].len);
        nb += g->nodes[i].len;
        for (uint64_t q = 0; q < g->links[i].n; ++q) {
            tri[3 * nl] = g->links[i].v[q].s; tri[3 * nl + 1] = g->links[i].v[q].d; tri[3 * nl + 2] = g->links[i].v[q].dist;
            ++nl;
        }
    }
    node_off[g->n] = nb; link_row[g->n] = nl;
}
/* find_tip_nodes: tips needs room for n_nodes ids; returns the count */
EXPORT uint64_t ggo_edit_find_tips(void *h, uint64_t min_size, int64_t *tips) { return ge_find_tips((ge_graph *)h, min_size, tips); }
/* remove_node! for every id */
EXPORT void ggo_edit_remove_nodes(void *h, const int64_t *ids, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) ge_remove_node((ge_graph *)h, ids[i]);
}
/* collapse_all_unitigs!(sg, min_nodes, true): number of new nodes, or -1 ("Sequences do not overlap.") / -2 ("This path is
 * not a valid path."); the graph is left as the reference leaves it when error() is thrown half way */
EXPORT int64_t ggo_edit_collapse(void *h, uint64_t min_nodes) {
    ge_graph *g = (ge_graph *)h;
    g->error = 0;
    int64_t m = ge_collapse(g, min_nodes);
    return m >= 0 ? m : -(int64_t)g->error;
}
/* remove_tips!(sdg, min_size) src/remove_tips.jl:2-30: 0, or the negative error of the collapse */
EXPORT int64_t ggo_edit_remove_tips(void *h, uint64_t min_size, uint64_t *passes, uint64_t *removed) {
    ge_graph *g = (ge_graph *)h;
    uint64_t pass = 1, rem = 0;
    int64_t *tips = (int64_t *)malloc((g->n ? g->n : 1) * sizeof(int64_t));
    uint64_t nt = ge_find_tips(g, min_size, tips);
    int64_t rc = 0;
    while (nt) {
        for (uint64_t i = 0; i < nt; ++i) ge_remove_node(g, tips[i]);
        rem += nt;
        g->error = 0;
        if (ge_collapse(g, 2) < 0) { rc = -(int64_t)g->error; break; }
        ++pass;
        tips = (int64_t *)realloc(tips, (g->n ? g->n : 1) * sizeof(int64_t));
        nt = ge_find_tips(g, min_size, tips);
    }
    free(tips);
    *passes = pass; *removed = rem;
    return rc;
}
