#!/usr/bin/env python
"""bench.py -- DBG build throughput (input k-mers/sec) on B200, contract per the task brief.

A "step" is one pass of the whole hot path over one batch of synthetic reads:
reads (ASCII) -> 2-bit pack -> canonical k-mers -> sort + run-length counts -> min-count
filter -> unitig nodes -> (k-1)-overlap links.  Workload = BASELINE.json configs[1]
(C2): 4,641,652 bp uniform genome, 50x 150 bp paired reads (fragment
700, 0.1 % substitutions), k = 31, min-count 2.

  value  : whole-job input k-mers/s with the read buffers already resident in HBM
  e2e    : same metric through the host-buffer C-ABI call (gg_dbg_from_reads) + full export of
           the graph to host memory, H2D / D2H inside the timed region
  roofline / cpu_baseline : see DESIGN.md, "Measurement"

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port":
the Julia reference cannot run here) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

WORKLOADS = {
    # name: genome_len, coverage, read_len, frag_len, err_ppm, k, min_freq
    "c2_ecoli50x_k31": dict(genome_len=4641652, cov=50, read_len=150, frag_len=700, err_ppm=1000, k=31, min_freq=2),
    "c2_ecoli50x_k63": dict(genome_len=4641652, cov=50, read_len=150, frag_len=700, err_ppm=1000, k=63, min_freq=2),
    "c3_100mbp30x_k31": dict(genome_len=100000000, cov=30, read_len=150, frag_len=700, err_ppm=10000, k=31, min_freq=2),
    "c5_100mbp30x_k127": dict(genome_len=100000000, cov=30, read_len=150, frag_len=700, err_ppm=10000, k=127, min_freq=2),
    "tiny": dict(genome_len=200000, cov=30, read_len=150, frag_len=500, err_ppm=2000, k=31, min_freq=2),
}

# algorithmic (compulsory) bytes per LAUNCH of the kernels that can dominate ("Kernels"):
# every input element read once, every output element written once.  N = k-mer instances,
# nb = input bases, U = distinct k-mers kept by the min-count filter, W = key bytes.
def _stream_bytes(nb):  # packed words (8 B / 32 bases) + window-valid masks (4 B / 32 bases)
    return 12.0 * ((nb + 31) // 32)


ALG_BYTES = {
    "msd_hist_kernel": lambda N, nb, U, W: _stream_bytes(nb),
    "msd_scatter_kernel": lambda N, nb, U, W: _stream_bytes(nb) + W * N,
    "msd_hist1_kernel": lambda N, nb, U, W: W * N,
    "msd_scatter1_kernel": lambda N, nb, U, W: 2.0 * W * N,
    "msd_leaf_kernel": lambda N, nb, U, W: W * N + (W + 4.0) * U,
    "msd_gather_kernel": lambda N, nb, U, W: 2.0 * (W + 4.0) * U,
    # multi-word keys (k > 32): same formulas with W = 16 / 32 bytes
    "msdw_hist_kernel": lambda N, nb, U, W: _stream_bytes(nb),
    "msdw_scatter_kernel": lambda N, nb, U, W: _stream_bytes(nb) + W * N,
    "msdw_hist1_kernel": lambda N, nb, U, W: W * N,
    "msdw_scatter1_kernel": lambda N, nb, U, W: 2.0 * W * N,
    "msdw_leaf_kernel": lambda N, nb, U, W: W * N + (W + 4.0) * U,
    "msdw_gather_kernel": lambda N, nb, U, W: 2.0 * (W + 4.0) * U,
    "radix_scatter_kernel": lambda N, nb, U, W: 2.0 * W * N,   # per pass
    "radix_hist_kernel": lambda N, nb, U, W: 1.0 * W * N,
    "extract_emit_kernel": lambda N, nb, U, W: _stream_bytes(nb) + W * N,
    "pack_kernel": lambda N, nb, U, W: 1.375 * nb,             # 1 B ASCII in, 2 bits + 1 bad bit out
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
# captures of the default workload (profiles/README.md); null where no capture exists
NCU_TRAFFIC = {  # profiles/r1_e_ncu_full_<kernel>.txt, workload c2_ecoli50x_k31 on one GPU
    "msd_leaf_kernel": 1.486635e9 + 91.601664e6,
    "msd_scatter_kernel": 87.484672e6 + 1.428851e9,
    "msd_scatter1_kernel": 1.485387e9 + 1.438692e9,
}


def nreads_for(w):
    return 2 * int(round(w["cov"] * w["genome_len"] / (2.0 * w["read_len"])))


def kmers_per_read(w):
    return max(0, w["read_len"] - w["k"] + 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """index of the next sample: brackets the timed region (the sampler itself is started before
        the warm-up so that nvidia-smi's start-up never lands inside a timed region)"""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        window = self.lines[first:last] or self.lines[max(0, first - 1):]  # a region shorter than one period: nearest sample
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, w, rank, world):
    """CPU arm: the oracle (C restatement of the reference algorithm) on a bounded sample."""
    if rank != 0:
        return
    from oracle import oracle as O
    pkg = G.load_package()
    frac = args.cpu_sample_frac
    glen = max(int(w["genome_len"] * frac), 4 * w["frag_len"])
    ws = dict(w, genome_len=glen)
    nreads = nreads_for(ws)
    seq, offs = pkg.synth_reads(1, glen, nreads, w["read_len"], True, w["frag_len"], w["err_ppm"])
    n_inst = nreads * kmers_per_read(w)
    threads = O.lib().ggo_max_threads()

    def step():
        keys, cnt = O.count_kmers(seq, offs, w["k"], threads=threads)
        g = O.dbg(O.filter_counts(keys, cnt, w["k"], w["min_freq"]), w["k"])
        return g.n_nodes

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n_inst * args.steps / dt
    sample = "genome %d bp (%.3g of the workload's), %d reads, %d k-mer instances per step; counting with %d OpenMP threads, unitig walk sequential" % (
        glen, frac, nreads, n_inst, threads)
    out = {
        "impl": "reference", "metric": "DBG build: input k-mers/sec", "value": v, "unit": "k-mers/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64" if w["k"] <= 32 else "u128" if w["k"] <= 64 else "u256", "data": "synthetic",
        "config": {"workload": args.workload, **w, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the Julia algorithm (oracle/), not Julia itself; JULIA_NUM_THREADS n/a (the reference is single-threaded)"},
        "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_ecoli50x_k31", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-frac", type=float, default=0.25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    pkg = G.build() if rank == 0 or world == 1 else None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        if pkg is None:
            pkg = G.load_package()
    N = pkg._native
    L = N.lib()
    ctx = pkg.Context(local_rank)
    H = ctx.h
    k, min_freq, rl = w["k"], w["min_freq"], w["read_len"]
    W = 8 * L.gg_kmer_words(k)
    nreads = nreads_for(w)
    nbytes = nreads * rl
    n_inst = nreads * kmers_per_read(w)
    # Weak scaling: ONE job whose size grows with the GPU count -- genome world x genome_len, world x nreads
    # reads; rank r owns reads [r * nreads, (r + 1) * nreads) and (after the exchange) one key range.
    seed = 1
    genome_total = w["genome_len"] * world
    comm = pkg.dist.Comm.from_torch(ctx) if world > 1 else None

    # ---- inputs: generated on the device, kept resident in HBM; pinned host copy for the e2e leg
    d_seq, d_offs, h_seq = C.c_void_p(), C.c_void_p(), C.c_void_p()
    N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq)))
    N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs)))
    N.check(H, L.gg_synth_reads_range_dev(H, seed, genome_total, rank * nreads, nreads, rl, 1, w["frag_len"], w["err_ppm"], 0, d_seq))
    pinned = []

    def pinned_array(count, dtype):
        """numpy view of page-locked host memory (gg_host_alloc): every host buffer of the e2e leg is pinned"""
        p = C.c_void_p()
        nbytes_ = max(int(count), 1) * np.dtype(dtype).itemsize
        N.check(H, L.gg_host_alloc(H, nbytes_, C.byref(p)))
        pinned.append(p)
        return np.frombuffer((C.c_uint8 * nbytes_).from_address(p.value), dtype=dtype, count=max(int(count), 1))

    offs = pinned_array(nreads + 1, np.uint64)
    offs[:] = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(rl)
    N.check(H, L.gg_memcpy_h2d(H, d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))
    N.check(H, L.gg_host_alloc(H, nbytes + 64, C.byref(h_seq)))
    N.check(H, L.gg_memcpy_d2h(H, h_seq, d_seq, nbytes))

    def barrier():
        L.gg_ctx_sync(H)
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def step_dev():
        gh = C.c_void_p()
        if comm is None:
            N.check(H, L.gg_dbg_from_reads_dev(H, d_seq, d_offs, nreads, nbytes, k, min_freq, C.byref(gh)))
        else:  # distributed count (one NCCL exchange) + all-gather of the kept k-mers + graph stage
            N.check(H, L.gg_dist_dbg_from_reads_dev(comm.h, d_seq, d_offs, nreads, nbytes, k, min_freq, C.byref(gh), None))
        return gh

    def graph_sizes(gh):
        nn, nb, nl, kk = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        N.check(H, L.gg_graph_sizes(gh, C.byref(nn), C.byref(nb), C.byref(nl), C.byref(kk)))
        return nn.value, nb.value, nl.value

    sampler = ClockSampler(local_rank)
    if rank == 0:   # one sampler per job: concurrent nvidia-smi pollers contend on the driver and stall CUDA calls
        sampler.start()

    # ---- warm-up (also fills the memory pool) ----
    launches = 0
    sizes = None
    for _ in range(args.warmup):
        gh = step_dev()
        sizes = graph_sizes(gh)
        L.gg_graph_free(gh)
        launches = ctx.last_timings()[1]

    # ---- timed: K steps, inputs resident in HBM (232 MB of reads > 126 MB L2) ----
    barrier()
    s_first = sampler.mark()
    N.check(H, L.gg_ctx_event_record(H, 0))
    t0 = time.perf_counter()
    stage_ms = None
    for _ in range(args.steps):
        gh = step_dev()
        L.gg_graph_free(gh)
    N.check(H, L.gg_ctx_event_record(H, 1))
    barrier()
    wall = time.perf_counter() - t0
    ev_ms = C.c_float()
    N.check(H, L.gg_ctx_event_elapsed_ms(H, 0, 1, C.byref(ev_ms)))
    stage_ms, launches = ctx.last_timings()
    dev_s = ev_ms.value / 1e3

    # ---- e2e: host buffers in, full graph out, every step ----
    nn, nb, nl = sizes
    out_off = pinned_array(nn + 1, np.uint64); out_bases = pinned_array(max(nb, 1), np.uint8)
    out_row = pinned_array(nn + 1, np.uint64); out_tri = pinned_array(max(3 * nl, 1), np.int64)

    d_seq2, d_offs2 = C.c_void_p(), C.c_void_p()
    if comm is not None:
        N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq2)))
        N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs2)))

    def step_e2e():
        gh = C.c_void_p()
        if comm is None:
            N.check(H, L.gg_dbg_from_reads(H, h_seq, offs.ctypes.data_as(C.c_void_p), nreads, k, min_freq, C.byref(gh)))
        else:  # host buffers of this rank's shard -> device, distributed build, graph back to the host
            N.check(H, L.gg_memcpy_h2d(H, d_seq2, h_seq, nbytes))
            N.check(H, L.gg_memcpy_h2d(H, d_offs2, offs.ctypes.data_as(C.c_void_p), offs.nbytes))
            N.check(H, L.gg_dist_dbg_from_reads_dev(comm.h, d_seq2, d_offs2, nreads, nbytes, k, min_freq, C.byref(gh), None))
        N.check(H, L.gg_graph_export(gh, out_off.ctypes.data_as(C.c_void_p), out_bases.ctypes.data_as(C.c_void_p),
                                     out_row.ctypes.data_as(C.c_void_p), out_tri.ctypes.data_as(C.c_void_p)))
        L.gg_graph_free(gh)

    step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t1
    clocks = sampler.stop(s_first, sampler.mark())

    # ---- max over ranks / sum of work ----
    if dist is not None:
        import torch
        t = torch.tensor([dev_s, e2e_s, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, wall = (float(x) for x in t.tolist())
    total_inst = n_inst * world

    # ---- roofline of the dominant kernel: one extra, untimed, profiled step ----
    roofline, top, allk = None, [], []
    L.gg_ctx_set_profile(H, 1)   # every rank runs the step (it is collective for N > 1); rank 0 reports its own kernels
    gh = step_dev()
    L.gg_graph_free(gh)
    L.gg_ctx_set_profile(H, 0)
    prof_text = L.gg_ctx_kernel_profile(H).decode()
    # distinct k-mers (all / kept by the filter) of this rank's key range: the writes of the counting kernels
    ch = C.c_void_p()
    if comm is None:
        N.check(H, L.gg_count_kmers_dev(H, d_seq, d_offs, nreads, nbytes, k, 1, C.byref(ch)))
    else:
        N.check(H, L.gg_dist_count_kmers_dev(comm.h, d_seq, d_offs, nreads, nbytes, k, 1, 1, C.byref(ch), None))
    nu_all = C.c_uint64()
    N.check(H, L.gg_counts_sizes(ch, C.byref(nu_all), None, None, None))
    N.check(H, L.gg_counts_filter(ch, min_freq, 0))
    nu_kept = C.c_uint64()
    N.check(H, L.gg_counts_sizes(ch, C.byref(nu_kept), None, None, None))
    L.gg_counts_free(ch)
    if rank == 0:
        prof = []
        for ln in prof_text.splitlines():
            name, cnt, ms = ln.split("\t")
            prof.append((name.split("<")[0].strip("( "), int(cnt), float(ms)))
        agg = {}
        for name, cnt, ms in prof:
            a = agg.setdefault(name, [0, 0.0]); a[0] += cnt; a[1] += ms
        tot_ms = sum(v[1] for v in agg.values())
        allk = sorted(agg.items(), key=lambda kv: -kv[1][1])
        top = allk[:8]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
        distinct = {"all": nu_all.value, "kept": nu_kept.value}

        def model(kname, kcnt, kms):
            f = ALG_BYTES.get(kname)
            if f is None:
                return None
            by = f(float(n_inst), float(nbytes), float(nu_kept.value), float(W))
            gbs = by / (kms / kcnt / 1e3) / 1e9
            return by, gbs

        name, (cnt, ms) = top[0]
        m = model(name, cnt, ms)
        if m is not None:
            roofline = {"bound": "hbm", "kernel": name, "achieved": m[1], "peak": peak, "unit": "GB/s", "frac": m[1] / peak,
                        "traffic": NCU_TRAFFIC.get(name) if (args.workload == "c2_ecoli50x_k31" and world == 1) else None, "launches_per_step": cnt, "avg_launch_ms": ms / cnt, "share_of_kernel_time": ms / tot_ms,
                        "alg_bytes_per_launch": m[0], "peak_source": peak_src}
        else:
            roofline = {"bound": "hbm", "kernel": name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                        "launches_per_step": cnt, "avg_launch_ms": ms / cnt, "share_of_kernel_time": ms / tot_ms, "peak_source": peak_src,
                        "note": "latency-bound pointer chasing; no streaming byte model"}
        kernel_rooflines = []
        for kname, (kcnt, kms) in top:
            m = model(kname, kcnt, kms)
            if m is not None:
                kernel_rooflines.append({"kernel": kname, "alg_bytes_per_launch": m[0], "avg_launch_ms": kms / kcnt,
                                         "achieved_gbs": round(m[1], 1), "frac": round(m[1] / peak, 4)})
        roofline["distinct_kmers"] = distinct
        roofline["other_kernels"] = kernel_rooflines

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        frac = args.cpu_sample_frac
        glen = int(w["genome_len"] * frac)
        ws = dict(w, genome_len=glen)
        nr = nreads_for(ws)
        sseq, soffs = pkg.synth_reads(1, glen, nr, rl, True, w["frag_len"], w["err_ppm"])
        tc = time.perf_counter()
        keys, cnt = O.count_kmers(sseq, soffs, k, threads=1)
        og = O.dbg(O.filter_counts(keys, cnt, k, min_freq), k)
        tc = time.perf_counter() - tc
        cpu = {"value": nr * kmers_per_read(w) / tc, "unit": "k-mers/s", "cores": 1, "kind": "port",
               "sample": "same workload at %.3g of the genome length (%d bp, %d reads, %d k-mer instances), 1 thread as the reference runs; %.1f s" % (
                   frac, glen, nr, nr * kmers_per_read(w), tc),
               "host_cpus": os.cpu_count(), "nodes": og.n_nodes,
               "note": "C restatement of the Julia algorithm (oracle/), not Julia itself; JULIA_NUM_THREADS n/a"}

    if rank == 0:
        h2d = nbytes + offs.nbytes
        d2h = out_off.nbytes + nb + out_row.nbytes + 3 * nl * 8
        out = {
            "metric": "DBG build: input k-mers/sec", "value": total_inst * args.steps / dev_s, "unit": "k-mers/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64" if k <= 32 else "u128" if k <= 64 else "u256", "data": "synthetic",
            "config": {"workload": args.workload, **w, "reads_per_gpu": nreads, "kmer_instances_per_gpu": n_inst,
                       "input_bytes_per_gpu": nbytes, "l2_policy": "inputs (%.0f MB of reads + %.0f MB of k-mers per step) larger than the 126 MB L2" % (nbytes / 1e6, n_inst * W / 1e6),
                       "genome_len_total": genome_total,
                       "parallelism": ("1 GPU" if world == 1 else "1 process per GPU; reads sharded by index, k-mers range-partitioned over the ranks "
                                       "with one NCCL send/recv exchange, kept k-mers all-gathered, graph stage replicated"),
                       "graph": {"nodes": nn, "bases": nb, "link_entries": nl}},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "stage_ms_last_step": stage_ms,
            "hbm_gbs_effective": None,
            "e2e": {"value": total_inst * args.steps / e2e_s, "unit": "k-mers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clocks,
            "roofline": roofline,
            "top_kernels_ms": [{"kernel": n_, "launches": v[0], "ms": round(v[1], 4)} for n_, v in top],
            "all_kernels_ms": {n_: [v[0], round(v[1], 4)] for n_, v in allk},
            "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)

    L.gg_dev_free(H, d_seq); L.gg_dev_free(H, d_offs); L.gg_host_free(H, h_seq)
    del offs, out_off, out_bases, out_row, out_tri
    for p in pinned:
        L.gg_host_free(H, p)
    if comm is not None:
        L.gg_dev_free(H, d_seq2); L.gg_dev_free(H, d_offs2)
        comm.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
