"""Long fuzz campaign of the tip-removal statements against each other (not collected by pytest; CPU only):
sequential Python restatement (oracle/graph_edit.py) = C restatement (oracle/graph_edit.c) = executable model of the CUDA
formulation (tests/collapse_model.py) on random-link graphs, cut-up paths and circles at k = 5 and k = 3, min_nodes 2..4,
plus remove_tips!.     python tests/fuzz_graph_edit.py [first_seed] [n_seeds]      (12 seeds: 18,000 collapses, ~3 min)"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import collapse_model as M   # noqa: E402
import test_graph_edit as T  # noqa: E402

E = T.E


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    n = errs = 0
    for seed in range(first, first + count):
        rnd = random.Random(seed)
        for trial in range(500):
            kind = trial % 3
            if kind == 0:
                nodes, links = T._random_graph(rnd, rnd.randint(1, 24), rnd.randint(0, 30), dup=0.3)
            else:
                nodes, links = T._chain_graph(rnd, k=5 if kind == 1 else 3)
            for mn in (2, 3, 4):
                try:
                    exp = T._seq_collapse(nodes, links, mn)
                except ValueError:
                    errs += 1
                    for f in (lambda: M.collapse_all_unitigs(nodes, links, mn), lambda: E.CGraph.from_lists(nodes, links).collapse_all_unitigs(mn)):
                        try:
                            f()
                        except ValueError:
                            continue
                        raise AssertionError(("no error", seed, trial, mn))
                    continue
                got = M.collapse_all_unitigs(nodes, links, mn)
                assert got[0] == exp[0] and got[1] == exp[1] and got[2] == exp[2], (seed, trial, mn, "model")
                c = E.CGraph.from_lists(nodes, links)
                assert c.collapse_all_unitigs(mn) == exp[2] and c.lists() == (exp[0], exp[1]), (seed, trial, mn, "C")
                n += 1
            g = E.Graph(nodes, links)
            try:
                pr = g.remove_tips(8)
            except ValueError:
                continue
            n2, l2, p2, r2 = M.remove_tips(nodes, links, 8)
            assert (p2, r2) == pr and n2 == g.nodes and l2 == g.links, (seed, trial, "tips model")
            c = E.CGraph.from_lists(nodes, links)
            assert c.remove_tips(8) == pr and c.lists() == (g.nodes, g.links), (seed, trial, "tips C")
    print("ok: %d collapses equal in all three statements, %d error cases agreed on" % (n, errs))


if __name__ == "__main__":
    main()
