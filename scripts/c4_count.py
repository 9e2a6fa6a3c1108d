#!/usr/bin/env python
"""Config 4's counting half at full size: 3 Gbp genome, 30x 150 bp reads (7.2e10 31-mer instances), min-count 2, sharded over
the GPUs of one box (torchrun, one rank per GPU).  Reads are generated on the device per shard; gg_dist_count_kmers_dev counts
in rounds over slices of the minimizer digit.  The graph stage does not run at this size (32-bit vertex ids).
Prints one JSON line on rank 0.  usage: torchrun --nproc-per-node 8 scripts/c4_count.py [genome_len] [coverage]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    genome_len = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
    cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    rl, k, min_freq, seed, frag = 150, 31, 2, 3, 700
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = G.load_package()
    N, L = pkg._native, pkg._native.lib()
    ctx = pkg.Context(local)
    H = ctx.h
    comm = pkg.dist.Comm.from_torch(ctx)
    nreads_total = 2 * int(round(cov * genome_len / (2.0 * rl)))
    first, nreads = pkg.dist.shard_range(nreads_total, rank, world, 2)
    nbytes = nreads * rl
    d_seq, d_offs = C.c_void_p(), C.c_void_p()
    N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq)))
    N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs)))
    t0 = time.perf_counter()
    N.check(H, L.gg_synth_reads_range_dev(H, seed, genome_len, first, nreads, rl, 1, frag, 10000, 0, d_seq))
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(rl)
    N.check(H, L.gg_memcpy_h2d(H, d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))
    del offs
    L.gg_ctx_sync(H)
    t_gen = time.perf_counter() - t0
    times, kept, ng = [], 0, C.c_uint64()
    spec = np.zeros(256, dtype=np.uint64)
    for it in range(2):   # first call fills the memory pool
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        ch = C.c_void_p()
        N.check(H, L.gg_dist_count_kmers_dev(comm.h, d_seq, d_offs, nreads, nbytes, k, 1, min_freq, C.byref(ch), C.byref(ng)))
        L.gg_ctx_sync(H); torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
        nu = C.c_uint64()
        L.gg_counts_sizes(ch, C.byref(nu), None, None, None)
        N.check(H, L.gg_counts_spectrum(ch, spec.ctypes.data_as(N.u64p)))
        L.gg_counts_free(ch)
        kept = nu.value
    t = torch.tensor([kept] + [int(x) for x in spec], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    tt = torch.tensor(times, device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_inst = nreads_total * (rl - k + 1)
        sp = t[1:].tolist()
        out = {"what": "config 4 counting half: reads -> canonical 31-mers -> counts -> min-count filter (sorted table, range-partitioned over the ranks)",
               "genome_len": genome_len, "coverage": cov, "reads": nreads_total, "kmer_instances": n_inst, "n_gpus": world,
               "instances_seen_equal_reads_times_windows": ng.value == n_inst, "kept_kmers": int(t[0].item()),
               "seconds_first_call": float(tt[0].item()), "seconds": float(tt[1].item()), "kmers_per_s": n_inst / float(tt[1].item()),
               "read_generation_s": t_gen, "spectrum_mode_above_3": int(np.argmax(sp[4:]) + 4), "spectrum_head": sp[:40]}
        print(json.dumps(out), flush=True)
    L.gg_dev_free(H, d_seq); L.gg_dev_free(H, d_offs)
    comm.close(); ctx.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
