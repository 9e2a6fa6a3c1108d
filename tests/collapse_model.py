"""Executable model of the data-parallel formulation in genomegraphs.jl_b200/csrc/graphedit.cuh.

Every function below is one kernel (a loop over independent indices) or one primitive (scan, stable sort) of the CUDA
code, with the same record layouts and sort keys, so that the formulation itself -- chains of oriented nodes, cycle
cuts, the closed form of the reference's sequential link push order -- is checked against the sequential restatement
(oracle/graph_edit.py) on the CPU (tests/test_graph_edit.py).  Test infrastructure; the product path is the CUDA code.
"""

NIL = 0xFFFFFFFF
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def vtx(x):
    """oriented node (signed id) -> vertex index"""
    return 2 * (abs(x) - 1) + (1 if x < 0 else 0)


def val(v):
    """vertex index -> signed id"""
    return -((v >> 1) + 1) if v & 1 else (v >> 1) + 1


def oriented_base(nodes, v, j):
    s = nodes[v >> 1]
    return _COMP[s[len(s) - 1 - j]] if v & 1 else s[j]


def csr(links):
    row = [0]
    tri = []
    for ls in links:
        tri.extend(ls)
        row.append(len(tri))
    return row, tri


def remove_nodes(nodes, links, ids):
    """remove_node! for every id: sequences emptied, every entry touching a removed node dropped, order kept"""
    dead = [False] * len(nodes)
    for n in ids:
        dead[abs(n) - 1] = True
    nodes2 = ["" if dead[i] else s for i, s in enumerate(nodes)]
    links2 = [[l for l in ls if not dead[abs(l[0]) - 1] and not dead[abs(l[1]) - 1]] for ls in links]
    return nodes2, links2


def jump(rec, rounds):
    """pointer jumping, double buffered: rec[v] = [ptr, mn, cnt, w] covers the window [v, ptr)"""
    for _ in range(rounds):
        out = []
        for v in range(len(rec)):
            p, mn, cnt, w = rec[v]
            if p != NIL:
                q = rec[p]
                out.append([q[0], min(mn, q[1]), (cnt + q[2]) & 0xFFFFFFFF, (w + q[3]) & 0xFFFFFFFF])
            else:
                out.append([p, mn, cnt, w])
        rec = out
    return rec


def collapse_all_unitigs(nodes, links, min_nodes, stats=None):
    """collapse_all_unitigs!(sg, min_nodes, consume = true): returns (nodes, links, number of new nodes)"""
    assert min_nodes >= 2
    nn = len(nodes)
    nv = 2 * nn
    row, tri = csr(links)
    ne = len(tri)
    # ---- ge_degree_kernel: per node, entries with source == -n (forward) / == n (backward)
    outdeg = [0] * nv
    indeg = [0] * nv
    dest = [0] * nv
    dist = [0] * nv
    for i in range(nn):
        n = i + 1
        for q in range(row[i], row[i + 1]):
            s, d, ds = tri[q]
            if s == -n:
                if outdeg[2 * i] == 0:
                    dest[2 * i], dist[2 * i] = d, ds
                outdeg[2 * i] += 1
                indeg[2 * i + 1] += 1
            if s == n:
                if outdeg[2 * i + 1] == 0:
                    dest[2 * i + 1], dist[2 * i + 1] = d, ds
                outdeg[2 * i + 1] += 1
                indeg[2 * i] += 1
    # ---- ge_mirror_kernel: position of the mirror entry (t-th (s, d) <-> t-th (d, s)); self links come in pairs
    mirror = [0] * ne
    for i in range(nn):
        for q in range(row[i], row[i + 1]):
            s, d, _ = tri[q]
            t = sum(1 for p in range(row[i], q) if tri[p][0] == s and tri[p][1] == d)
            j = abs(d) - 1
            found = None
            c = 0
            for p in range(row[j], row[j + 1]):
                if tri[p][0] == d and tri[p][1] == s:
                    if c == t:
                        found = p - row[j]
                    c += 1
            if found is None:
                raise ValueError("links are not symmetric")
            if s == d:
                tot = sum(1 for p in range(row[i], row[i + 1]) if tri[p][0] == s and tri[p][1] == d)
                if tot & 1:
                    raise ValueError("a self link must appear twice")
            mirror[q] = found
    # ---- ge_succ_kernel
    succ = [NIL] * nv
    for v in range(nv):
        if outdeg[v] == 1 and indeg[vtx(dest[v])] == 1:
            succ[v] = vtx(dest[v])
    pred = [NIL] * nv
    for v in range(nv):
        if succ[v] != NIL:
            pred[succ[v]] = v
    rounds = max(1, (nv - 1).bit_length())
    # ---- pass 1: cycles (ptr never runs out) and their smallest vertex
    rec = jump([[succ[v], v, 1, 0] for v in range(nv)], rounds)
    # ---- ge_cut_kernel
    for v in range(nv):
        if rec[v][0] != NIL and rec[v][1] == v:
            if v & 1 == 0:
                p = pred[v]
                succ[p] = NIL
                pred[v] = NIL
            else:
                s = succ[v]
                pred[s] = NIL
                succ[v] = NIL
    # ---- weights: bases a vertex adds to its path = length - overlap with its predecessor
    trim = [0] * nv
    wgt = [0] * nv
    bad_overlap = False
    for v in range(nv):
        ln = len(nodes[v >> 1])
        t = 0
        if pred[v] != NIL:
            ds = dist[pred[v]]
            if ds <= 0:
                t = -ds
                if t > ln or t > len(nodes[pred[v] >> 1]):
                    bad_overlap = True
                    t = 0
        trim[v] = t
        wgt[v] = ln - t
    # ---- pass 2: suffix and prefix aggregates of every path
    rs = jump([[succ[v], v, 1, wgt[v]] for v in range(nv)], rounds)
    rp = jump([[pred[v], v, 1, wgt[v]] for v in range(nv)], rounds)
    cmin = [min(rs[v][1], rp[v][1]) for v in range(nv)]
    clen = [rs[v][2] + rp[v][2] - 1 for v in range(nv)]
    ctot = [rs[v][3] + rp[v][3] - wgt[v] for v in range(nv)]
    coff = [rp[v][3] - wgt[v] for v in range(nv)]
    inpath = [clen[v] >= min_nodes for v in range(nv)]
    # ---- unitig numbering: ascending smallest node id, orientation with that node forwards
    flag = [1 if (inpath[2 * i] and cmin[2 * i] == 2 * i) else 0 for i in range(nn)]
    uidx = [0] * nn
    m = 0
    for i in range(nn):
        uidx[i] = m
        m += flag[i]
    if m == 0:
        return list(nodes), [list(ls) for ls in links], 0
    if bad_overlap:
        # only an error when the offending pair sits in a collapsed path
        for v in range(nv):
            if inpath[v] and pred[v] != NIL and dist[pred[v]] <= 0:
                t = -dist[pred[v]]
                if t > len(nodes[v >> 1]) or t > len(nodes[pred[v] >> 1]):
                    raise ValueError("Sequences do not overlap.")
    ulen = [0] * m
    for v in range(nv):
        if inpath[v] and cmin[v] == v and v & 1 == 0:
            ulen[uidx[v >> 1]] = ctot[v]
    # ---- ge_emit_kernel: bases of the chosen orientation, overlap check against the predecessor
    useq = [[None] * ulen[u] for u in range(m)]
    for v in range(nv):
        if not (inpath[v] and cmin[v] & 1 == 0):
            continue
        u = uidx[cmin[v] >> 1]
        ln = len(nodes[v >> 1])
        t = trim[v]
        for j in range(ln):
            b = oriented_base(nodes, v, j)
            if j < t:
                p = pred[v]
                lp = len(nodes[p >> 1])
                if oriented_base(nodes, p, lp - t + j) != b:
                    raise ValueError("Sequences do not overlap.")
            else:
                useq[u][coff[v] + j - t] = b
    # ---- ge_canon_kernel: iscanonical from both ends
    flip = [0] * m
    for u in range(m):
        s = useq[u]
        i, j = 0, len(s) - 1
        while i <= j:
            f, r = s[i], _COMP[s[j]]
            if f != r:
                flip[u] = 1 if r < f else 0
                break
            i += 1
            j -= 1
    for u in range(m):
        if flip[u]:
            useq[u] = [_COMP[c] for c in reversed(useq[u])]
    nodes2 = ["" if inpath[2 * i] else nodes[i] for i in range(nn)] + ["".join(s) for s in useq]

    # ---- ends: 0 kept, 1 outward end of unitig (u, side), 2 interior
    def end_type(x):
        v = vtx(x)
        if not inpath[v]:
            return 0, 0, 0
        if pred[v] != NIL:
            return 2, 0, 0
        chosen = cmin[v] & 1 == 0
        u = uidx[cmin[v] >> 1]
        side = (0 if chosen else 1) ^ flip[u]
        return 1, u, side

    def phi(x, ty, u, side):
        if ty == 0:
            return x
        return (nn + 1 + u) if side == 0 else -(nn + 1 + u)

    # ---- ge_link_flag_kernel: class A = entry between two kept nodes (stays in its row, order given), class B = the rest
    # of the surviving entries (get a key and are sorted)
    cls = [0] * ne
    for q in range(ne):
        s, d, _ = tri[q]
        ts, us, ss = end_type(s)
        td, ud, sd = end_type(d)
        if ts == 2 or td == 2:
            continue
        if ts == 1 and td == 1 and us == ud and ss == sd:
            continue                # both at one end of one unitig: pushed to the old node only, gone with it
        cls[q] = 1 if (ts == 0 and td == 0) else 2
    pos_a, pos_b = [0] * ne, [0] * ne
    n_a = n_b = 0
    for q in range(ne):             # two exclusive scans
        pos_a[q], pos_b[q] = n_a, n_b
        n_a += cls[q] == 1
        n_b += cls[q] == 2
    if stats is not None:
        stats.update(entries=ne, class_a=n_a, class_b=n_b, new_nodes=m)
    # ---- ge_link_emit_kernel: keys of the class B entries, as packed in the CUDA code
    items = [None] * n_b
    for i in range(nn):
        for q in range(row[i], row[i + 1]):
            if cls[q] != 2:
                continue
            s, d, ds = tri[q]
            qs = q - row[i]
            ts, us, ss = end_type(s)
            td, ud, sd = end_type(d)
            fs, fd = phi(s, ts, us, ss), phi(d, td, ud, sd)
            r = abs(fs) - 1
            if ts == 0:
                key = (1, 0, 0, ud, sd, 0, mirror[q])        # pushed to a kept node when path ud is joined
            elif td == 0:
                key = (0, ss, 0, 0, 0, 0, qs)                # creation loops of the new node, old entries first
            elif ud == us:
                # first end -- last end of one unitig: the second loop of join_path! meets the entries the first
                # loop has just pushed and links the new node to itself, (-N, N) then (N, -N) per link
                key = (0, 1, 2, qs if ss == 0 else mirror[q], 0, 0, 1 if ss == 0 else 0)
            elif ud < us:
                key = (0, ss, 1, ud, sd, 0, mirror[q])       # entries an earlier join had pushed
            else:
                key = (1, 0, 0, ud, sd, ss, qs)              # pushed when the later path ud is joined
            seg, S, sub, j, sx, sy, qq = key
            assert qq < (1 << 30)
            w0 = (r << 4) | (seg << 3) | (S << 2) | sub
            w1 = (j << 32) | (sx << 31) | (sy << 30) | qq
            items[pos_b[q]] = (w0, w1, (fs, fd, ds))
    items.sort(key=lambda t: (t[0], t[1]))   # stable radix sort on the device
    # ---- ge_link_place_kernel: first_b[r] = first sorted class B entry of row r; class A entries before row r =
    # pos_a[row[r]] (all of them for the new rows); row r starts at a_before(r) + first_b[r], its class A entries
    # come first (kept rows: old entries in place), then its class B entries
    nn2 = nn + m
    first_b = [0] * (nn2 + 1)
    for t in range(n_b + 1):
        r1 = (items[t][0] >> 4) if t < n_b else nn2
        r0 = (items[t - 1][0] >> 4) + 1 if t > 0 else 0
        for r in range(r0, r1 + 1):
            first_b[r] = t

    def a_before(r):
        return pos_a[row[r]] if (r <= nn and row[r] < ne) else n_a
    row2 = [a_before(r) + first_b[r] for r in range(nn2 + 1)]
    tri2 = [None] * (n_a + n_b)
    for q in range(ne):
        if cls[q] == 1:
            i = abs(tri[q][0]) - 1
            tri2[first_b[i] + pos_a[q]] = tri[q]
    for t in range(n_b):
        r = items[t][0] >> 4
        tri2[a_before(r + 1) + t] = items[t][2]
    assert all(x is not None for x in tri2) and row2[nn2] == n_a + n_b
    links2 = [tri2[row2[r]:row2[r + 1]] for r in range(nn2)]
    return nodes2, links2, m


def find_tip_nodes(nodes, links, min_size):
    """tip_kernel of ggb200.cu"""
    out = []
    for i in range(len(nodes)):
        n = i + 1
        ln = len(nodes[i])
        if ln == 0 or ln > min_size:
            continue
        fw = [l for l in links[i] if l[0] == -n]
        bw = [l for l in links[i] if l[0] == n]

        def side(mv):
            return sum(1 for l in links[abs(mv) - 1] if l[0] == mv)
        tip = False
        if len(fw) == 1 and not bw and side(fw[0][1]) > 1:
            tip = True
        if not fw and len(bw) == 1 and side(bw[0][1]) > 1:
            tip = True
        if not fw and not bw:
            tip = True
        if tip:
            out.append(n)
    return out


def remove_tips(nodes, links, min_size):
    """the host loop of gg_graph_remove_tips"""
    passes, removed = 1, 0
    tips = find_tip_nodes(nodes, links, min_size)
    while tips:
        nodes, links = remove_nodes(nodes, links, tips)
        removed += len(tips)
        nodes, links, _ = collapse_all_unitigs(nodes, links, 2)
        passes += 1
        tips = find_tip_nodes(nodes, links, min_size)
    return nodes, links, passes, removed
