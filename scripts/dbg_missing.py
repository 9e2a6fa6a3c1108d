import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as G
from oracle import oracle as O
pkg = G.build()
ctx = pkg.Context(0)
k = 21
seq, offs = pkg.synth_reads(seed=600 + k, genome_len=40000, nreads=8000, read_len=150, paired=True, frag_len=400, err_ppm=5000, n_ppm=200)
okeys, ocnt = O.count_kmers(seq, offs, k, threads=4)
for env in ({"GG_SKM_TEST_OVERFLOW": "3", "GG_SKM_BIN": "512"}, {"GG_SKM_TEST_OVERFLOW": "3", "GG_SKM_BIN": "512", "GG_SKM_VARIANT": "5"}, {"GG_SKM_BIN": "512"}, {"GG_SKM_TEST_OVERFLOW": "3"}):
    for kk in ("GG_SKM_TEST_OVERFLOW", "GG_SKM_BIN", "GG_SKM_VARIANT"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    for rep in range(3):
        c = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((seq, offs))
        keys, c32, _ = c.export()
        c.free()
        a, b = set(keys[:, 0].tolist()), set(okeys[:, 0].tolist())
        miss = sorted(b - a); extra = sorted(a - b)
        print(env, "rep", rep, "gpu", len(keys), "oracle", len(okeys), "missing", len(miss), "extra", len(extra), "dups", len(keys) - len(a),
              "missing top8:", sorted(set(m >> 34 for m in miss))[:10], flush=True)
