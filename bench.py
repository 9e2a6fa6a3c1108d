#!/usr/bin/env python
"""bench.py -- DBG build throughput (input k-mers/sec) on B200, contract per the task brief.

A "step" is one pass of the whole hot path over one batch of synthetic reads:
reads (ASCII) -> 2-bit pack -> canonical k-mers -> counts -> min-count filter -> unitig nodes ->
(k-1)-overlap links.  Default workload = BASELINE.json configs[2] (C3), the
largest configuration that fits one GPU: 100 Mbp uniform genome, 30x 150 bp paired reads with 1 %
substitutions, k = 31, min-count 2 (2.4e9 k-mer instances).  `--gpus N` runs the SAME job on N GPUs
(reads sharded by index: strong scaling, as config 3 is written); `--scaling weak` grows the genome
with N instead.

  value  : whole-job input k-mers/s with the read buffers already resident in HBM
  e2e    : same metric through the host-buffer C-ABI call (gg_dbg_from_reads) + full export of
           the graph to host memory, H2D / D2H inside the timed region
  roofline / cpu_baseline / parity_checked : see DESIGN.md, "Measurement"

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port": the
Julia reference cannot run here) on a bounded sample of the same workload; it loads nothing but
oracle/ (the read generator included).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: genome_len, coverage, read_len, frag_len, err_ppm, k, min_freq, seed, cpu_frac (share of the genome the CPU legs build)
    "c2_ecoli50x_k31": dict(genome_len=4641652, cov=50, read_len=150, frag_len=700, err_ppm=1000, k=31, min_freq=2, seed=1, cpu_frac=0.25),
    "c2_ecoli50x_k63": dict(genome_len=4641652, cov=50, read_len=150, frag_len=700, err_ppm=1000, k=63, min_freq=2, seed=1, cpu_frac=0.25),
    "c3_100mbp30x_k31": dict(genome_len=100000000, cov=30, read_len=150, frag_len=700, err_ppm=10000, k=31, min_freq=2, seed=2, cpu_frac=0.03),
    "c5_100mbp30x_k127": dict(genome_len=100000000, cov=30, read_len=150, frag_len=700, err_ppm=10000, k=127, min_freq=2, seed=2, cpu_frac=0.03),
    "tiny": dict(genome_len=200000, cov=30, read_len=150, frag_len=500, err_ppm=2000, k=31, min_freq=2, seed=1, cpu_frac=1.0),
}
DEFAULT_WORKLOAD = "c3_100mbp30x_k31"


def nreads_for(w, genome_len=None):
    return 2 * int(round(w["cov"] * (genome_len or w["genome_len"]) / (2.0 * w["read_len"])))


def kmers_per_read(w):
    return max(0, w["read_len"] - w["k"] + 1)


def shard_range(n, rank, world, mult=2):
    """reads [first, first + count) of rank `rank` (pairs stay together); same rule as genomegraphs.jl_b200/dist.py"""
    units = n // mult
    lo = (units * rank // world) * mult
    hi = (units * (rank + 1) // world) * mult if rank + 1 < world else n
    return lo, hi - lo


def dtype_of(k):
    return "u64" if k <= 32 else "u128" if k <= 64 else "u256"


def base_config(args, w, world):
    """identical in both arms (ours / reference): the workload and how it is spread over the GPUs"""
    weak = args.scaling == "weak"
    genome_total = w["genome_len"] * (world if weak else 1)
    nreads_total = nreads_for(w) * (world if weak else 1)
    wl = {k: v for k, v in w.items() if k != "cpu_frac"}
    return {"workload": args.workload, **wl, "n_gpus": world, "scaling": args.scaling if world > 1 else "single",
            "genome_len_total": genome_total, "reads_total": nreads_total, "kmer_instances_total": nreads_total * kmers_per_read(w),
            "input_bytes_total": nreads_total * w["read_len"],
            "l2_policy": "inputs (%.0f MB of reads per GPU per step) larger than the 126 MB L2; no flush needed" % (nreads_total * w["read_len"] / world / 1e6),
            "parallelism": ("1 GPU" if world == 1 else
                            "1 process per GPU; reads sharded by index; super-k-mers partitioned by minimizer hash and stored into the owner "
                            "GPU's buffer over NVLink (bulk asynchronous copies); kept k-mers range-partitioned in a second exchange; unitig stage "
                            "sharded by vertex range over the all-gathered table; every rank returns the whole graph")}


def cpu_sample(w):
    """the bounded sample of the workload the CPU legs build: the same generator parameters on a shorter genome"""
    glen = max(int(w["genome_len"] * w["cpu_frac"]), 4 * w["frag_len"])
    return glen, nreads_for(w, glen)


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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, w, rank, world):
    """CPU arm: the oracle (C restatement of the reference algorithm) on a bounded sample.  Loads oracle/ only."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()   # stated explicitly: torch.distributed.run exports OMP_NUM_THREADS=1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    glen, nreads = cpu_sample(w)
    seq, offs = O.synth_reads(w["seed"], glen, nreads, w["read_len"], True, w["frag_len"], w["err_ppm"], threads=threads)
    n_inst = nreads * kmers_per_read(w)

    def step():
        keys, cnt = O.count_kmers(seq, offs, w["k"], threads=threads)
        g = O.dbg(O.filter_counts(keys, cnt, w["k"], w["min_freq"], u8=True), w["k"])
        return g.n_nodes

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n_inst * args.steps / dt
    sample = ("each step = the workload's generator on a %d bp genome (%.3g of the workload's %d bp), %d reads, %d k-mer instances; "
              "counting with %d OpenMP threads (set explicitly), unitig walk sequential as in the reference" % (
                  glen, glen / w["genome_len"], w["genome_len"], nreads, n_inst, threads))
    out = {
        "impl": "reference", "metric": "DBG build: input k-mers/sec", "value": v, "unit": "k-mers/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.scaling == "weak" else "strong", "vs_baseline": None, "dtype": dtype_of(w["k"]), "data": "synthetic",
        "config": base_config(args, w, world),
        "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the Julia algorithm (oracle/), not Julia itself; JULIA_NUM_THREADS n/a (the reference is single-threaded)"},
        "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def csrc_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "genomegraphs.jl_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# the files that hold the kernels of the counting + graph stages (the ones an ncu traffic table lists) and the host code
# that launches them; ggb200.cu (ABI entry points) and the files of other routes are left out, so that adding an entry
# point next to the path does not void a capture of kernels that did not change
TRAFFIC_FILES = ("common.cuh", "prims.cuh", "pack.cuh", "sort.cuh", "skm.cuh", "graph.cuh", "graphdist.cuh")


def kernel_files_sha(read=None):
    """per-file hashes of TRAFFIC_FILES (read: name -> bytes, default: the working tree)"""
    d = os.path.join(ROOT, "genomegraphs.jl_b200", "csrc")
    read = read or (lambda f: open(os.path.join(d, f), "rb").read())
    return {f: hashlib.sha256(read(f)).hexdigest()[:16] for f in TRAFFIC_FILES}


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from an `ncu --set full` capture (profiles/ncu_traffic.json,
    written by scripts/ncu_traffic.py): used only while every file that defines or launches the captured kernels
    (TRAFFIC_FILES) is byte-identical to the build the capture saw; else None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    if t.get("workload") != workload:
        return None
    if t.get("files") != kernel_files_sha() and t.get("csrc_sha") != csrc_sha():
        return None
    return t.get("kernels", {}).get(kernel)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the one-GPU rebuild of the whole job that checks the N-GPU result")
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

    import __graft_entry__ as G
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
    k, min_freq, rl, seed = w["k"], w["min_freq"], w["read_len"], w["seed"]
    weak = args.scaling == "weak" and world > 1
    genome_total = w["genome_len"] * (world if weak else 1)
    nreads_total = nreads_for(w) * (world if weak else 1)
    first_read, nreads = shard_range(nreads_total, rank, world, 2)
    nbytes = nreads * rl
    n_inst_total = nreads_total * kmers_per_read(w)
    comm = pkg.dist.Comm.from_torch(ctx) if world > 1 else None

    # ---- inputs: generated on the device, kept resident in HBM; pinned host copy for the e2e leg
    d_seq, d_offs, h_seq = C.c_void_p(), C.c_void_p(), C.c_void_p()
    N.check(H, L.gg_dev_alloc(H, nbytes + 64, C.byref(d_seq)))
    N.check(H, L.gg_dev_alloc(H, 8 * (nreads + 1), C.byref(d_offs)))
    N.check(H, L.gg_synth_reads_range_dev(H, seed, genome_total, first_read, nreads, rl, 1, w["frag_len"], w["err_ppm"], 0, d_seq))
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
        else:
            N.check(H, L.gg_dist_dbg_from_reads_dev(comm.h, d_seq, d_offs, nreads, nbytes, k, min_freq, C.byref(gh), None))
        return gh

    def graph_sizes(gh):
        nn, nb, nl, kk = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        N.check(H, L.gg_graph_sizes(gh, C.byref(nn), C.byref(nb), C.byref(nl), C.byref(kk)))
        return nn.value, nb.value, nl.value

    def graph_checksum(gh):
        cs = C.c_uint64()
        N.check(H, L.gg_graph_checksum(gh, C.byref(cs)))
        return cs.value

    sampler = ClockSampler(local_rank)
    if rank == 0:   # one sampler per job: concurrent nvidia-smi pollers contend on the driver and stall CUDA calls
        sampler.start()

    # ---- warm-up (also fills the memory pool) ----
    launches = 0
    sizes = checksum = None
    for _ in range(args.warmup):
        gh = step_dev()
        sizes, checksum = graph_sizes(gh), graph_checksum(gh)
        L.gg_graph_free(gh)
        launches = ctx.last_timings()[1]

    # ---- timed: K steps, inputs resident in HBM ----
    barrier()
    s_first = sampler.mark()
    N.check(H, L.gg_ctx_event_record(H, 0))
    t0 = time.perf_counter()
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

    # ---- roofline of the dominant kernel: one extra, untimed, profiled step ----
    roofline, top, allk = None, [], []
    L.gg_ctx_set_profile(H, 1)   # every rank runs the step (it is collective for N > 1); rank 0 reports its own kernels
    gh = step_dev()
    L.gg_graph_free(gh)
    L.gg_ctx_set_profile(H, 0)
    prof_text = L.gg_ctx_kernel_profile(H).decode()
    if rank == 0:
        agg = {}
        for ln in prof_text.splitlines():
            f = ln.split("\t")
            name = f[0].split("<")[0].strip("( ")
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += int(f[1]); a[1] += float(f[2]); a[2] += float(f[3]) if len(f) > 3 else 0.0
        tot_ms = sum(v[1] for v in agg.values())
        allk = sorted(agg.items(), key=lambda kv: -kv[1][1])
        top = allk[:8]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")

        def line(kname, kcnt, kms, kby):
            d = {"kernel": kname, "launches_per_step": kcnt, "avg_launch_ms": kms / kcnt, "share_of_kernel_time": kms / tot_ms}
            if kby > 0:   # algorithmic bytes of the launches themselves (library-side model)
                gbs = kby / (kms / 1e3) / 1e9
                d.update({"alg_bytes_per_launch": kby / kcnt, "achieved": gbs, "frac": gbs / peak})
            else:
                d.update({"alg_bytes_per_launch": None, "achieved": None, "frac": None,
                          "note": "latency-bound (random access / pointer chasing) or bookkeeping: no streaming byte model"})
            return d

        name, (cnt, ms, by) = top[0]
        roofline = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, **line(name, cnt, ms, by),
                    "traffic": ncu_traffic(name, args.workload) if world == 1 else None,
                    "other_kernels": [line(n_, *v) for n_, v in top[1:] if v[2] > 0]}
        # the whole path against the pipeline-independent floor of SURVEY 8(d): ASCII bases in, kept (k-mer, count) out
        Wb = 8 * L.gg_kmer_words(k)
        roofline["path"] = {"floor_bytes_per_step_per_gpu": 1.0 * nbytes, "note": "floor = every input base read once as ASCII (SURVEY 8d, + (W+4) per kept k-mer)",
                            "floor_gbs_at_value": nbytes / (dev_s / args.steps) / 1e9, "key_bytes": Wb}

    # ---- parity of the benchmarked result + CPU baseline beside it ----
    parity = {"ok": True}
    cpu = None
    if world > 1 and not args.no_parity and n_inst_total >= 2 ** 32 - 1:
        # the whole job exceeds what ONE GPU builds (2^32 - 1 instances): size-independent properties instead -- every instance
        # was seen, and the unitigs hold every kept k-mer exactly once: bases - (k - 1) * nodes = kept k-mers of the count table
        import torch
        ch, ng = C.c_void_p(), C.c_uint64()
        N.check(H, L.gg_dist_count_kmers_dev(comm.h, d_seq, d_offs, nreads, nbytes, k, 1, min_freq, C.byref(ch), C.byref(ng)))
        nu = C.c_uint64()
        L.gg_counts_sizes(ch, C.byref(nu), None, None, None)
        L.gg_counts_free(ch)
        t = torch.tensor([nu.value], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        kept = int(t.item())
        props = {"instances_seen_equal_reads_times_windows": ng.value == n_inst_total,
                 "kmers_in_unitigs_equal_kept_kmers": sizes[1] - (k - 1) * sizes[0] == kept, "kept_kmers": kept}
        parity.update({"n_gpu_vs_1_gpu": "skipped: %d instances exceed the one-GPU limit" % n_inst_total, "properties": props})
        parity["ok"] = parity["ok"] and props["instances_seen_equal_reads_times_windows"] and props["kmers_in_unitigs_equal_kept_kmers"]
        barrier()
    elif world > 1 and not args.no_parity:
        # the same job once more on ONE GPU (rank 0, outside every timed region): checksum + sizes must equal the N-GPU result
        if rank == 0:
            da, do = C.c_void_p(), C.c_void_p()
            nb_all = nreads_total * rl
            N.check(H, L.gg_dev_alloc(H, nb_all + 64, C.byref(da)))
            N.check(H, L.gg_dev_alloc(H, 8 * (nreads_total + 1), C.byref(do)))
            N.check(H, L.gg_synth_reads_range_dev(H, seed, genome_total, 0, nreads_total, rl, 1, w["frag_len"], w["err_ppm"], 0, da))
            oa = np.arange(nreads_total + 1, dtype=np.uint64) * np.uint64(rl)
            N.check(H, L.gg_memcpy_h2d(H, do, oa.ctypes.data_as(C.c_void_p), oa.nbytes))
            g1 = C.c_void_p()
            N.check(H, L.gg_dbg_from_reads_dev(H, da, do, nreads_total, nb_all, k, min_freq, C.byref(g1)))
            s1, c1 = graph_sizes(g1), graph_checksum(g1)
            L.gg_graph_free(g1); L.gg_dev_free(H, da); L.gg_dev_free(H, do)
            parity.update({"n_gpu_vs_1_gpu": {"sizes_equal": s1 == sizes, "checksum_equal": c1 == checksum, "sizes": list(sizes), "checksum": "%016x" % checksum}})
            parity["ok"] = parity["ok"] and s1 == sizes and c1 == checksum
        barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        glen, nr = cpu_sample(w)
        sseq, soffs = O.synth_reads(seed, glen, nr, rl, True, w["frag_len"], w["err_ppm"], threads=host_threads())
        tc = time.perf_counter()
        okeys, ocnt = O.count_kmers(sseq, soffs, k, threads=1)
        okept = O.filter_counts(okeys, ocnt, k, min_freq, u8=True)
        og = O.dbg(okept, k)
        tc = time.perf_counter() - tc
        cpu = {"value": nr * kmers_per_read(w) / tc, "unit": "k-mers/s", "cores": 1, "kind": "port",
               "sample": "the workload's generator on a %d bp genome (%.3g of the workload's), %d reads, %d k-mer instances, 1 thread as the reference runs; %.1f s" % (
                   glen, glen / w["genome_len"], nr, nr * kmers_per_read(w), tc),
               "host_cpus": os.cpu_count(), "nodes": og.n_nodes,
               "note": "C restatement of the Julia algorithm (oracle/), not Julia itself; JULIA_NUM_THREADS n/a"}
        # the GPU path on the SAME bytes, through the host-buffer C-ABI calls: everything must be bit-equal
        counts = pkg.serial_mem(k, pkg.CANONICAL, ctx=ctx)((sseq, soffs))
        gk, gc, _ = counts.export()
        ga = pkg.dbg_from_reads((sseq, soffs), k, min_freq, ctx=ctx)
        counts.free()
        chk = {"kmers": bool(np.array_equal(gk, okeys)), "counts": bool(np.array_equal(gc, ocnt)),
               "nodes": bool(np.array_equal(ga.bases, og.bases) and np.array_equal(ga.node_offsets.astype(np.int64), og.node_off)),
               "links": bool(np.array_equal(ga.link_triples, og.link_tri) and np.array_equal(ga.link_row_offsets.astype(np.int64), og.link_row))}
        parity.update({"gpu_vs_oracle_on_cpu_sample": chk, "distinct_kmers": int(len(okeys)), "kept_kmers": int(len(okept)), "nodes": int(og.n_nodes),
                       "link_entries": int(len(og.link_tri) // 3)})
        parity["ok"] = parity["ok"] and all(chk.values())

    if rank == 0:
        h2d = nbytes + offs.nbytes
        d2h = out_off.nbytes + nb + out_row.nbytes + 3 * nl * 8
        out = {
            "metric": "DBG build: input k-mers/sec", "value": n_inst_total * args.steps / dev_s, "unit": "k-mers/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": dtype_of(k), "data": "synthetic",
            "config": base_config(args, w, world),
            "graph": {"nodes": nn, "bases": nb, "link_entries": nl, "checksum": "%016x" % checksum},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "stage_ms_last_step": stage_ms,
            "e2e": {"value": n_inst_total * args.steps / e2e_s, "unit": "k-mers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "per": "rank 0's shard (every rank copies its own shard in and the graph out)"},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clocks,
            "roofline": roofline,
            "top_kernels_ms": [{"kernel": n_, "launches": v[0], "ms": round(v[1], 4)} for n_, v in top],
            "all_kernels_ms": {n_: [v[0], round(v[1], 4)] for n_, v in allk},
            "cpu_baseline": cpu,
            "parity_checked": parity,
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
    if not parity["ok"]:
        sys.stderr.write("PARITY FAILURE: %s\n" % json.dumps(parity))
        sys.exit(3)


if __name__ == "__main__":
    main()
