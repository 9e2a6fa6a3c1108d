"""Writes tests/golden/tips_golden.json: tip-removal cases (input graph, min_size) with the results of the sequential
restatement of the reference (oracle/graph_edit.py): find_tip_nodes, collapse_all_unitigs!(sg, 2, true) on the input,
remove_tips!(sg, min_size).  oracle/make_ref_golden.jl recomputes every expected value with the real package.

    python tests/golden/make_tips_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import graph_edit as E  # noqa: E402
from oracle import oracle as O      # noqa: E402
import test_graph_edit as T         # noqa: E402  (the graph generators of the fuzz tests)


def case(name, nodes, links, min_size):
    g = E.Graph(nodes, links)
    out = {"name": name, "min_size": min_size, "nodes": list(nodes), "links": [[list(t) for t in ls] for ls in links],
           "tips": g.find_tip_nodes(min_size)}
    c = E.Graph(nodes, links)
    try:
        new = c.collapse_all_unitigs(2, True)
        out["collapse"] = {"new_nodes": new, "nodes": c.nodes, "links": [[list(t) for t in ls] for ls in c.links]}
    except ValueError as e:
        out["collapse"] = {"error": str(e)}
    r = E.Graph(nodes, links)
    try:
        passes, removed = r.remove_tips(min_size)
        out["remove_tips"] = {"passes": passes, "removed": removed, "nodes": r.nodes, "links": [[list(t) for t in ls] for ls in r.links]}
    except ValueError as e:
        out["remove_tips"] = {"error": str(e)}
    return out


def main():
    cases = []
    # T1: the fork of tests/test_graph_edit.py::test_restatement_tip_rules (checked by hand against src/Graphs.jl:778-802)
    g = E.Graph(["AAAC", "AACGGG", "AACT", "CGGGT", "TTTTT"], [[] for _ in range(5)])
    g.add_link(-1, 2, -3)
    g.add_link(-1, 3, -3)
    g.add_link(-2, 4, -4)
    cases.append(case("T1 fork with one tip and an isolated node", g.nodes, g.links, 4))
    rnd = random.Random(77)
    for i in range(8):
        nodes, links = T._chain_graph(rnd)
        cases.append(case("chains %d (paths and circles in pieces, k = 5)" % i, nodes, links, 7))
    for i in range(8):
        nodes, links = T._random_graph(rnd, rnd.randint(3, 12), rnd.randint(2, 14), dup=0.2)
        cases.append(case("random links %d (duplicates, self links, deleted nodes)" % i, nodes, links, 6))
    for i, (k, err, ms) in enumerate([(5, 30000, 8), (7, 20000, 12), (9, 20000, 16), (15, 10000, 45)]):
        seq, offs = O.synth_reads(40 + i, 300 if k < 15 else 1500, 120 if k < 15 else 500, 30 if k < 15 else 60, True, 80 if k < 15 else 150, err)
        keys, cnt = O.count_kmers(seq, offs, k)
        og = O.dbg(O.filter_counts(keys, cnt, k, 1), k)
        cases.append(case("dbg graph of reads with errors, k = %d" % k, og.nodes(), og.links(), ms))
    out = {"generator": "tests/golden/make_tips_golden.py (oracle/graph_edit.py, sequential restatement of src/remove_tips.jl + src/Graphs.jl)",
           "cases": cases}
    with open(os.path.join(HERE, "tips_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(len(cases), "cases;", sum(1 for c in cases if "error" not in c["remove_tips"] and c["remove_tips"]["removed"]), "with tips removed;",
          sum(1 for c in cases if "error" in c["collapse"]), "collapse errors")


if __name__ == "__main__":
    main()
