"""Independent, string-based Python transliteration of the reference algorithm
(/root/reference/src/dbg.jl, src/GraphIndexes.jl) -- pure-Python loops for SMALL cases.

It is deliberately written against Python strings / bisect rather than packed words so
that it shares no code or representation with oracle/gg_oracle.c or the CUDA path; the
tests use it to cross-check the C oracle (the reference itself cannot run here).
Lexicographic order of ACGT strings == unsigned order of the 2-bit words.
"""
import bisect

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(_COMP[c] for c in reversed(s))


def canonical(s):
    r = rc(s)
    return s if s <= r else r


def extract(reads, k, canon=True):
    """each(M, seq) + canonical: windows with a non-ACGT symbol are skipped."""
    out = []
    for r in reads:
        r = r.upper()
        for i in range(len(r) - k + 1):
            w = r[i:i + k]
            if all(c in "ACGT" for c in w):
                out.append(canonical(w) if canon else w)
    return out


def count(kmers):
    """collect -> sort -> run-length collapse."""
    s = sorted(kmers)
    out = []
    for x in s:
        if out and out[-1][0] == x:
            out[-1][1] += 1
        else:
            out.append([x, 1])
    return [(a, b) for a, b in out]


def fw_neighbours(m):      # src/dbg.jl:30-34
    return [m[1:] + b for b in "ACGT"]


def bw_neighbours(m):      # src/dbg.jl:36-42
    return [b + m[:-1] for b in "ACGT"]


def _get_idxs(m, lst, fw):  # src/dbg.jl:70-90
    out = []
    for n in (fw_neighbours(m) if fw else bw_neighbours(m)):
        c = canonical(n)
        i = min(bisect.bisect_left(lst, c), len(lst) - 1)
        if lst[i] == c:
            out.append((n, i))
    return out


def is_end_fw(m, lst):      # src/dbg.jl:57-68
    nxt = _get_idxs(m, lst, True)
    if len(nxt) != 1:
        return True
    return len(_get_idxs(nxt[0][0], lst, False)) != 1


def is_end_bw(m, lst):      # src/dbg.jl:44-55
    nxt = _get_idxs(m, lst, False)
    if len(nxt) != 1:
        return True
    return len(_get_idxs(nxt[0][0], lst, True)) != 1


def seq_canonical(s):
    r = rc(s)
    return s if s <= r else r


def build_unitigs(lst):     # src/dbg.jl:94-185
    lst = sorted(lst)
    used = [False] * len(lst)
    nodes = []
    for si, start in enumerate(lst):
        if used[si]:
            continue
        end_bw = is_end_bw(start, lst)
        end_fw = is_end_fw(start, lst)
        if not end_bw and not end_fw:
            continue
        if end_bw and end_fw:
            s = start
            used[si] = True
        else:
            cur = start
            used[si] = True
            if end_fw:
                cur = rc(start)
                end_fw = end_bw
            s = cur
            while not end_fw:
                fwn = _get_idxs(cur, lst, True)
                cur = fwn[0][0]
                if used[fwn[0][1]]:
                    break
                used[fwn[0][1]] = True
                s += cur[-1]
                end_fw = is_end_fw(cur, lst)
        nodes.append(seq_canonical(s))
    for ni in range(len(lst)):
        if used[ni]:
            continue
        start = lst[ni]
        cur = start
        used[ni] = True
        s = cur
        while True:
            fwn = _get_idxs(cur, lst, True)
            cur = fwn[0][0]
            if cur == start:
                break
            used[fwn[0][1]] = True
            s += cur[-1]
        nodes.append(seq_canonical(s))
    return lst, nodes


def connect(nodes, k):      # src/dbg.jl:190-240 + src/Graphs.jl:473-476
    links = [[] for _ in nodes]
    if len(nodes) <= 1:     # src/dbg.jl:267
        return links
    inn, out = [], []
    for nid, s in enumerate(nodes, start=1):
        first = s[:k - 1]
        if first <= rc(first):
            inn.append((first, nid))
        else:
            out.append((rc(first), nid))
        last = s[len(s) - (k - 1):]
        if last <= rc(last):
            out.append((last, -nid))
        else:
            inn.append((rc(last), -nid))
    inn.sort()
    out.sort()
    nxt = 0
    for m, nid in inn:
        while nxt < len(out) and out[nxt][0] < m:
            nxt += 1
        o = nxt
        while o < len(out) and out[o][0] == m:
            d = out[o][1]
            links[abs(nid) - 1].append((nid, d, -(k - 1)))
            links[abs(d) - 1].append((d, nid, -(k - 1)))
            o += 1
    return links


def dbg(kmers):
    """dbg!(graph, kmers) -> (nodes, links)."""
    if not kmers:
        return [], []
    k = len(kmers[0])
    _, nodes = build_unitigs(list(kmers))
    return nodes, connect(nodes, k)


def unique_mer_index(nodes, k):  # src/GraphIndexes.jl:62-102 (intended semantics)
    recs = []
    total = []
    for nid, s in enumerate(nodes, start=1):
        total.append(max(0, len(s) - k + 1))
        for i in range(len(s) - k + 1):
            w = s[i:i + k]
            if not all(c in "ACGT" for c in w):
                continue
            r = rc(w)
            recs.append((w, nid, i + 1) if w <= r else (r, -nid, i + 1))
    recs.sort(key=lambda x: x[0])
    uniq = []
    i = 0
    while i < len(recs):
        j = i + 1
        while j < len(recs) and recs[j][0] == recs[i][0]:
            j += 1
        if j - i == 1:
            uniq.append(recs[i])
        i = j
    upn = [0] * len(nodes)
    for _, nid, _ in uniq:
        upn[abs(nid) - 1] += 1
    return uniq, upn, total
