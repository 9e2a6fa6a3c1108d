"""CPU restatement of the reference's tip removal (TEST INFRASTRUCTURE -- never imported by the product path).

Function by function after the reference, sequential, pure Python (small graphs only):

    remove_tips!            src/remove_tips.jl:2-30
    find_tip_nodes!         src/Graphs.jl:778-802
    remove_node!            src/Graphs.jl:459-466
    remove_link!            src/Graphs.jl:487-495   (== on links ignores dist, src/Graphs.jl:183-185)
    forward/backward_links  src/Graphs.jl:667-690   (is_forwards_from: l.source == -n, :188)
    find_link               src/Graphs.jl:643-651
    collapse_all_unitigs!   src/Graphs.jl:528-539
    find_all_unitigs!       src/Graphs.jl:831-862
    reverse!(path)          src/Graphs.jl:975-988
    check_overlap, sequence(path)  src/Graphs.jl:991-1037
    join_path!              src/Graphs.jl:1040-1088
    add_node! / add_link!   src/Graphs.jl:418-444, 473-476

A graph is (nodes, links): nodes[i] = sequence of node i + 1 as an ACGT string ('' = deleted node, the reference's
empty_node), links[i] = list of (source, destination, dist) tuples in push order.

Parity unpinned (like the rest of oracle/): no reference output can be produced in this project.  Recalled upstream
semantics used here: `iscanonical(seq)` of BioSequences 2 compares seq[i] with complement(seq[j]) from both ends and
keeps the sequence on a tie; A < C < G < T.
"""

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def iscanonical(s):
    i, j = 0, len(s) - 1
    while i <= j:
        f, r = s[i], _COMP[s[j]]
        if f < r:
            return True
        if r < f:
            return False
        i += 1
        j -= 1
    return True


class Graph:
    def __init__(self, nodes, links):
        self.nodes = list(nodes)
        self.links = [list(map(tuple, ls)) for ls in links]
        assert len(self.nodes) == len(self.links)

    def copy(self):
        return Graph(self.nodes, self.links)

    # src/Graphs.jl:667-690
    def forward_links(self, n):
        return [l for l in self.links[abs(n) - 1] if l[0] == -n]

    def backward_links(self, n):
        return self.forward_links(-n)

    # src/Graphs.jl:643-651
    def find_link(self, src, dst):
        for l in self.links[abs(src) - 1]:
            if l[0] == src and l[1] == dst:
                return l
        return None

    # src/Graphs.jl:418-444
    def add_node(self, seq):
        self.nodes.append(seq)
        self.links.append([])
        return len(self.nodes)

    # src/Graphs.jl:473-476
    def add_link(self, src, dst, dist):
        self.links[abs(src) - 1].append((src, dst, dist))
        self.links[abs(dst) - 1].append((dst, src, dist))

    # src/Graphs.jl:487-495
    def remove_link(self, src, dst):
        a = self.links[abs(src) - 1]
        a[:] = [l for l in a if not (l[0] == src and l[1] == dst)]
        b = self.links[abs(dst) - 1]
        b[:] = [l for l in b if not (l[0] == dst and l[1] == src)]

    # src/Graphs.jl:459-466
    def remove_node(self, n):
        for l in list(self.links[abs(n) - 1]):
            self.remove_link(l[0], l[1])
        self.nodes[abs(n) - 1] = ""

    def sequence(self, n):
        s = self.nodes[abs(n) - 1]
        return revcomp(s) if n < 0 else s

    # src/Graphs.jl:778-802 (the Set is returned sorted)
    def find_tip_nodes(self, min_size):
        out = set()
        for n in range(1, len(self.nodes) + 1):
            nd = self.nodes[n - 1]
            if nd == "" or len(nd) > min_size:
                continue
            fwl, bwl = self.forward_links(n), self.backward_links(n)
            if len(fwl) == 1 and len(bwl) == 0:
                if len(self.backward_links(fwl[0][1])) > 1:
                    out.add(n)
            if len(fwl) == 0 and len(bwl) == 1:
                if len(self.forward_links(-bwl[0][1])) > 1:
                    out.add(n)
            if not fwl and not bwl:
                out.add(n)
        return sorted(out)

    # src/Graphs.jl:831-862
    def find_all_unitigs(self, min_nodes):
        unitigs = []
        consumed = [False] * len(self.nodes)
        for n in range(1, len(self.nodes) + 1):
            if consumed[n - 1] or self.nodes[n - 1] == "":
                continue
            consumed[n - 1] = True
            path = [n]
            for _ in range(2):
                fn = self.forward_links(path[-1])
                while len(fn) == 1:
                    dest = fn[0][1]
                    if not consumed[abs(dest) - 1] and len(self.backward_links(dest)) == 1:
                        path.append(dest)
                        consumed[abs(dest) - 1] = True
                    else:
                        break
                    fn = self.forward_links(path[-1])
                path = [-x for x in reversed(path)]     # reverse!(path), src/Graphs.jl:975-988
            if len(path) >= min_nodes:
                unitigs.append(path)
        return unitigs

    # src/Graphs.jl:1006-1037
    def path_sequence(self, path):
        s = ""
        pnode = 0
        for n in path:
            nseq = self.sequence(n)
            if pnode != 0:
                lnk = self.find_link(pnode, n)
                if lnk is None:
                    raise ValueError("This path is not a valid path.")
                dst = lnk[2]
                if dst <= 0:
                    ovl = -dst
                    if ovl > len(s) or ovl > len(nseq) or s[len(s) - ovl:] != nseq[:ovl]:
                        raise ValueError("Sequences do not overlap.")
                    nseq = nseq[ovl:]
            s += nseq
            pnode = -n
        return s

    # src/Graphs.jl:1040-1088
    def join_path(self, path, consume):
        pnodes = set()
        for n in path:
            pnodes.add(n)
            pnodes.add(-n)
        ps = self.path_sequence(path)
        if not iscanonical(ps):
            ps = revcomp(ps)
            path = [-x for x in reversed(path)]
        newid = self.add_node(ps)
        for l in self.backward_links(path[0]):
            self.add_link(newid, l[1], l[2])
        for l in self.forward_links(path[-1]):
            self.add_link(-newid, l[1], l[2])
        if consume:
            for n in sorted(pnodes):        # a Set upstream; the outcome does not depend on the order
                ext = False
                if n != path[-1]:
                    for l in self.forward_links(n):
                        if l[1] not in pnodes:
                            ext = True
                if n != path[0]:
                    for l in self.backward_links(n):
                        if l[1] not in pnodes:
                            ext = True
                if ext:
                    continue
                self.remove_node(n)
        return newid

    # src/Graphs.jl:528-539
    def collapse_all_unitigs(self, min_nodes, consume):
        unitigs = self.find_all_unitigs(min_nodes)
        return [self.join_path(p, consume) for p in unitigs]

    # src/remove_tips.jl:2-30; returns (passes, tips removed)
    def remove_tips(self, min_size):
        passes = 1
        removed = 0
        tips = self.find_tip_nodes(min_size)
        while tips:
            for t in tips:
                self.remove_node(t)
            removed += len(tips)
            self.collapse_all_unitigs(2, True)
            passes += 1
            tips = self.find_tip_nodes(min_size)
        return passes, removed


class CGraph:
    """The same functions in plain C (oracle/graph_edit.c, part of libggoracle.so): for graphs of millions of nodes,
    where this module's Python loops are too slow, and as the CPU path that is timed next to the CUDA one.  Arrays in
    the layout of the C ABI: node_off (n + 1), bases (ASCII), link_row (n + 1), link_tri (3 per entry)."""

    def __init__(self, node_off, bases, link_row, link_tri):
        import ctypes as C
        import numpy as np
        from . import oracle as O
        self._C, self._np, self._O = C, np, O
        node_off = np.ascontiguousarray(node_off, dtype=np.uint64)
        link_row = np.ascontiguousarray(link_row, dtype=np.uint64)
        bases = np.ascontiguousarray(bases if len(bases) else np.zeros(1, dtype=np.uint8), dtype=np.uint8)
        link_tri = np.ascontiguousarray(link_tri if len(link_tri) else np.zeros(3, dtype=np.int64), dtype=np.int64)
        self.h = O.lib().ggo_edit_graph_new(len(node_off) - 1, O._p(node_off, C.c_uint64), O._p(bases, C.c_uint8), O._p(link_row, C.c_uint64),
                                            O._p(link_tri, C.c_int64))

    @classmethod
    def from_lists(cls, nodes, links):
        import numpy as np
        node_off = np.zeros(len(nodes) + 1, dtype=np.uint64)
        node_off[1:] = np.cumsum([len(x) for x in nodes], dtype=np.uint64)
        link_row = np.zeros(len(nodes) + 1, dtype=np.uint64)
        link_row[1:] = np.cumsum([len(ls) for ls in links], dtype=np.uint64)
        return cls(node_off, np.frombuffer("".join(nodes).encode(), dtype=np.uint8), link_row,
                   np.array([t for ls in links for t in ls], dtype=np.int64).reshape(-1))

    def sizes(self):
        C = self._C
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._O.lib().ggo_edit_graph_sizes(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def arrays(self):
        C, np, O = self._C, self._np, self._O
        nn, nb, nl = self.sizes()
        node_off, link_row = np.zeros(nn + 1, dtype=np.uint64), np.zeros(nn + 1, dtype=np.uint64)
        bases, tri = np.zeros(max(nb, 1), dtype=np.uint8), np.zeros(max(3 * nl, 3), dtype=np.int64)
        O.lib().ggo_edit_graph_export(self.h, O._p(node_off, C.c_uint64), O._p(bases, C.c_uint8), O._p(link_row, C.c_uint64), O._p(tri, C.c_int64))
        return node_off, bases[:nb], link_row, tri[:3 * nl]

    def lists(self):
        node_off, bases, link_row, tri = self.arrays()
        b = bases.tobytes().decode()
        t = tri.reshape(-1, 3)
        nodes = [b[int(node_off[i]):int(node_off[i + 1])] for i in range(len(node_off) - 1)]
        links = [[tuple(int(x) for x in t[j]) for j in range(int(link_row[i]), int(link_row[i + 1]))] for i in range(len(node_off) - 1)]
        return nodes, links

    def find_tip_nodes(self, min_size):
        C, np, O = self._C, self._np, self._O
        out = np.zeros(max(self.sizes()[0], 1), dtype=np.int64)
        n = O.lib().ggo_edit_find_tips(self.h, int(min_size), O._p(out, C.c_int64))
        return out[:n].tolist()

    def remove_nodes(self, ids):
        C, np, O = self._C, self._np, self._O
        a = np.ascontiguousarray(ids if len(ids) else [0], dtype=np.int64)
        O.lib().ggo_edit_remove_nodes(self.h, O._p(a, C.c_int64), len(ids))

    @staticmethod
    def _raise(rc):
        if rc < 0:
            raise ValueError("Sequences do not overlap." if rc == -1 else "This path is not a valid path.")

    def collapse_all_unitigs(self, min_nodes=2):
        rc = self._O.lib().ggo_edit_collapse(self.h, int(min_nodes))
        self._raise(rc)
        return rc

    def remove_tips(self, min_size):
        C = self._C
        p, r = C.c_uint64(), C.c_uint64()
        self._raise(self._O.lib().ggo_edit_remove_tips(self.h, int(min_size), C.byref(p), C.byref(r)))
        return p.value, r.value

    def free(self):
        if self.h:
            self._O.lib().ggo_edit_graph_free(self.h)
            self.h = None

    def __del__(self):
        self.free()
