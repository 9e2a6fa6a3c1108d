"""Times gg_graph_remove_tips (remove_tips!, src/remove_tips.jl:2-30) on device-built graphs of error-rich reads.
Prints one JSON line per case.  Checks that no tip of that size is left and that two runs give the same checksum.
    python scripts/time_remove_tips.py [genome_len] [coverage]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    cov = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    pkg = G.build()
    N = pkg._native
    L = N.lib()
    ctx = pkg.Context(0)
    H = ctx.h
    rl, k = 150, 31
    nreads = genome * cov // rl // 2 * 2
    nbytes = nreads * rl
    d_seq, d_offs = C.c_void_p(), C.c_void_p()
    N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq)))
    N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs)))
    N.check(H, L.gg_synth_reads_range_dev(H, 2, genome, 0, nreads, rl, 1, 500, 10000, 0, d_seq))
    import numpy as np
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(rl)
    N.check(H, L.gg_memcpy_h2d(H, d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))

    def sizes(gh):
        nn, nb, nl, kk = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        N.check(H, L.gg_graph_sizes(gh, C.byref(nn), C.byref(nb), C.byref(nl), C.byref(kk)))
        return nn.value, nb.value, nl.value

    for min_freq, min_size in ((1, 60), (2, 200)):
        gh = C.c_void_p()
        N.check(H, L.gg_dbg_from_reads_dev(H, d_seq, d_offs, nreads, nbytes, k, min_freq, C.byref(gh)))
        before = sizes(gh)
        ms, sums, res = [], [], None
        for _ in range(3):
            out, p, r = C.c_void_p(), C.c_uint64(), C.c_uint64()
            L.gg_ctx_sync(H)
            t0 = time.perf_counter()
            N.check(H, L.gg_graph_remove_tips(gh, min_size, C.byref(out), C.byref(p), C.byref(r)))
            L.gg_ctx_sync(H)
            ms.append(1e3 * (time.perf_counter() - t0))
            cs = C.c_uint64()
            N.check(H, L.gg_graph_checksum(out, C.byref(cs)))
            sums.append(cs.value)
            nt = C.c_uint64(0)
            N.check(H, L.gg_graph_find_tip_nodes(out, min_size, None, C.byref(nt)))
            res = dict(after=sizes(out), passes=p.value, tips_removed=r.value, tips_left=nt.value)
            L.gg_graph_free(out)
        L.gg_graph_free(gh)
        print(json.dumps({"workload": "%d bp uniform genome, %dx 150 bp paired, 1%% substitutions, k=31, min-count %d" % (genome, cov, min_freq),
                          "min_size": min_size, "nodes_bases_links_before": before, "nodes_bases_links_after": res["after"],
                          "passes": res["passes"], "tips_removed": res["tips_removed"], "tips_left": res["tips_left"],
                          "ms": [round(x, 2) for x in ms], "checksum_stable": len(set(sums)) == 1}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
