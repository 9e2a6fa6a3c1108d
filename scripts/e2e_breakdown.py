"""Where the e2e step of bench.py spends its time (run on the GPU box)."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G
pkg = G.build(); N = pkg._native; L = N.lib(); ctx = pkg.Context(0); H = ctx.h
glen, rl, k = 4641652, 150, 31
nreads = 2 * int(round(50 * glen / (2.0 * rl))); nbytes = nreads * rl
d_seq, d_offs, h_seq = C.c_void_p(), C.c_void_p(), C.c_void_p()
N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq))); N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs)))
N.check(H, L.gg_synth_reads_dev(H, 1, glen, nreads, rl, 1, 700, 1000, 0, d_seq))
h_offs = C.c_void_p()
N.check(H, L.gg_host_alloc(H, 8 * (nreads + 1), C.byref(h_offs)))   # pinned, as bench.py does
offs = np.frombuffer((C.c_uint8 * (8 * (nreads + 1))).from_address(h_offs.value), dtype=np.uint64)
offs[:] = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(rl)
N.check(H, L.gg_memcpy_h2d(H, d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))
N.check(H, L.gg_host_alloc(H, nbytes + 64, C.byref(h_seq))); N.check(H, L.gg_memcpy_d2h(H, h_seq, d_seq, nbytes))
def T(f, n=5):
    f(); t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e3
print("H2D reads (pinned, %d MB): %.2f ms" % (nbytes >> 20, T(lambda: L.gg_memcpy_h2d(H, d_seq, h_seq, nbytes))))
print("H2D offsets (pinned, %d MB): %.2f ms" % (offs.nbytes >> 20, T(lambda: L.gg_memcpy_h2d(H, d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))))
def dev():
    gh = C.c_void_p(); N.check(H, L.gg_dbg_from_reads_dev(H, d_seq, d_offs, nreads, nbytes, k, 2, C.byref(gh))); L.gg_graph_free(gh)
print("device pipeline: %.2f ms" % T(dev))
sizes = {}
def host(export=True):
    gh = C.c_void_p(); N.check(H, L.gg_dbg_from_reads(H, h_seq, offs.ctypes.data_as(C.c_void_p), nreads, k, 2, C.byref(gh)))
    if export:
        nn, nb, nl, kk = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        L.gg_graph_sizes(gh, C.byref(nn), C.byref(nb), C.byref(nl), C.byref(kk))
        if "o" not in sizes:
            sizes["o"] = (np.zeros(nn.value + 1, np.uint64), np.zeros(nb.value, np.uint8), np.zeros(nn.value + 1, np.uint64), np.zeros(3 * nl.value, np.int64))
        a, b, c, d = sizes["o"]
        N.check(H, L.gg_graph_export(gh, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p)))
    L.gg_graph_free(gh)
print("host call without export: %.2f ms" % T(lambda: host(False)))
print("host call + export (bench e2e): %.2f ms" % T(host))
