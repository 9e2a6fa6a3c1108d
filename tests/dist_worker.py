"""Worker of the multi-rank tests (launched with torch.distributed.run, one process per rank).

mode cpu : gloo, no GPU.  The host-side sharding logic of the multi-GPU path (read shards,
           12-bit histogram, owner ranges from gg_dist_splitters, exchange, concatenation in
           rank order) is exercised with the CPU oracle standing in for the kernels.
mode gpu : one GPU per rank.  gg_dist_count_kmers_dev / gg_dist_dbg_from_reads_dev over NCCL,
           checked on rank 0 against the oracle run on all the reads.
Prints "DIST_OK <mode>" on rank 0 when everything matched.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [  # seed, genome_len, nreads, read_len, frag_len, err_ppm, k, min_freq
    (5, 30000, 6000, 100, 300, 3000, 31, 2),
    (6, 8000, 3001, 80, 200, 0, 21, 1),
    (7, 500, 64, 50, 120, 0, 4, 1),      # 2k < 12: fewer histogram bins than ranks may own
    (8, 20000, 4000, 150, 400, 5000, 32, 3),
    (9, 20000, 3000, 150, 400, 3000, 63, 2),     # two-word keys
    (10, 15000, 2000, 150, 400, 2000, 127, 1),   # four-word keys
]


def main():
    import torch.distributed as dist
    mode = sys.argv[1]
    pkg = G.load_package()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    D = pkg.dist
    ctx = comm = None
    if mode == "gpu":
        ctx = pkg.Context(int(os.environ.get("LOCAL_RANK", rank)))
        comm = D.Comm.from_torch(ctx)
    else:
        token = D.broadcast_bytes(os.urandom(128) if rank == 0 else None)
        toks = [None] * world
        dist.all_gather_object(toks, token)
        assert all(t == toks[0] for t in toks) and len(token) == 128
    for seed, glen, nreads, rl, fl, err, k, min_freq in CASES:
        first, cnt = D.shard_range(nreads, rank, world, 2)
        seq, offs = pkg.synth_reads(seed, glen, cnt, rl, True, fl, err, first_read=first)
        if mode == "cpu" and k > 32:
            continue   # the CPU emulation of the sharding logic works on one-word keys
        if mode == "cpu":
            km = O.extract(seq, offs, k)[:, 0]
            hb = min(12, 2 * k)
            hist = np.bincount((km >> np.uint64(2 * k - hb)).astype(np.int64), minlength=1 << hb).astype(np.uint64)
            hists = [None] * world
            dist.all_gather_object(hists, hist)
            fb = D.splitters(np.sum(hists, axis=0), world)
            dig = (km >> np.uint64(2 * k - hb)).astype(np.int64)
            outgoing = [km[(dig >= fb[p]) & (dig < fb[p + 1])] for p in range(world)]
            everything = [None] * world
            dist.all_gather_object(everything, outgoing)
            mine = np.concatenate([everything[p][rank] for p in range(world)])
            keys, counts = np.unique(mine, return_counts=True)
            keep = counts >= min_freq
            part = (keys[keep].reshape(-1, 1), counts[keep].astype(np.uint32))
            # ---- the super-k-mer route's TWO exchanges, emulated on the same shard: (1) every canonical k-mer goes to the rank
            # that owns its bin (any function of the k-mer stands in for the minimizer digit: equal k-mers meet in one bin; bins
            # are owned in equal shares), the owner counts and filters; (2) the kept pairs, a hash subset of the table, are
            # range-partitioned on a 14-bit digit histogram with gg_dist_splitters and merged.  Must equal `part`.
            nb0 = 64
            binof = ((km * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(58)).astype(np.int64)   # 0 .. 63
            owner_first = [q * nb0 // world for q in range(world + 1)]
            out1 = [km[(binof >= owner_first[p]) & (binof < owner_first[p + 1])] for p in range(world)]
            ex1 = [None] * world
            dist.all_gather_object(ex1, out1)
            got = np.concatenate([ex1[p][rank] for p in range(world)])
            k2, c2 = np.unique(got, return_counts=True)
            k2, c2 = k2[c2 >= min_freq], c2[c2 >= min_freq].astype(np.uint32)
            tb = min(14, 2 * k)
            dig2 = (k2 >> np.uint64(2 * k - tb)).astype(np.int64)
            h2 = np.bincount(dig2, minlength=1 << tb).astype(np.uint64)
            hs2 = [None] * world
            dist.all_gather_object(hs2, h2)
            fb2 = D.splitters(np.sum(hs2, axis=0), world)
            out2 = [(k2[(dig2 >= fb2[p]) & (dig2 < fb2[p + 1])], c2[(dig2 >= fb2[p]) & (dig2 < fb2[p + 1])]) for p in range(world)]
            ex2 = [None] * world
            dist.all_gather_object(ex2, out2)
            rk = np.concatenate([ex2[p][rank][0] for p in range(world)])
            rc = np.concatenate([ex2[p][rank][1] for p in range(world)])
            order = np.argsort(rk, kind="stable")
            part2 = (rk[order].reshape(-1, 1), rc[order])
        else:
            c = D.dist_mem(k, pkg.CANONICAL, comm, min_freq)((seq, offs))
            kk, c32, _ = c.export()
            part = (kk, c32)
            c.free()
            ga = D.dist_dbg_from_reads(comm, (seq, offs), k, min_freq)
            graph = (ga.node_strings(), ga.link_lists())
        parts = [None] * world
        dist.all_gather_object(parts, part)
        parts2 = None
        if mode == "cpu":
            parts2 = [None] * world
            dist.all_gather_object(parts2, part2)
        graphs = [None] * world
        if mode == "gpu":
            dist.all_gather_object(graphs, graph)
        if rank == 0:
            aseq, aoffs = pkg.synth_reads(seed, glen, nreads, rl, True, fl, err)
            okeys, ocnt = O.count_kmers(aseq, aoffs, k)
            okeys, ocnt = O.filter_counts(okeys, ocnt, k, min_freq), ocnt[ocnt >= min_freq]
            gk = np.concatenate([p[0] for p in parts])
            gc = np.concatenate([p[1] for p in parts])
            assert np.array_equal(gk, okeys), "k-mers differ (case seed %d)" % seed
            assert np.array_equal(gc, ocnt), "counts differ (case seed %d)" % seed
            if parts2 is not None:   # minimizer-bin partition + range repartition: the same global table, cut at (possibly) other keys
                assert np.array_equal(np.concatenate([p[0] for p in parts2]), okeys), "two-exchange k-mers differ (case seed %d)" % seed
                assert np.array_equal(np.concatenate([p[1] for p in parts2]), ocnt), "two-exchange counts differ (case seed %d)" % seed
            if mode == "gpu":
                og = O.dbg(okeys, k)
                for g in graphs:
                    assert g[0] == og.nodes() and g[1] == og.links(), "graph differs (case seed %d)" % seed
    dist.barrier()
    if comm is not None:
        comm.close()
        ctx.close()
    if rank == 0:
        print("DIST_OK", mode, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
