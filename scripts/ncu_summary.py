#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + hottest source lines."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, vals = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__average_warp_latency_issue_stalled_barrier", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_drain_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct"]
for r in vals:
    print("== kernel:", r[h.index("Kernel Name")][:100])
    for w in want:
        if w in h:
            i = h.index(w)
            print("  %-70s %s %s" % (w, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if len(r) > 6 and r[0] == "Line No"), None)
if hi is None:
    print("no source rows"); sys.exit(0)
hh = rows[hi]
ci = hh.index("# Samples"); ii = hh.index("Instructions Executed")
stalls = [(i, c) for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[0].isdigit():
        continue
    a = agg.setdefault(int(r[0]), {"src": r[1], "samples": 0, "inst": 0, "st": collections.Counter()})
    try:
        a["samples"] += int(r[ci] or 0); a["inst"] += int(r[ii] or 0)
        for i, c in stalls:
            a["st"][c] += int(r[i] or 0)
    except ValueError:
        pass
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
allst = collections.Counter()
for a in agg.values():
    allst.update(a["st"])
print("== stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / tot) for k, v in allst.most_common(8)))
print("== hottest source lines (%% of %d stall samples | %% of %d warp-instructions)" % (tot, toti))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top = ",".join("%s:%d" % (k[6:], 100 * v // max(a["samples"], 1)) for k, v in a["st"].most_common(2))
    print("  %5.1f%% %5.1f%%  L%-4d %-28s %s" % (100.0 * a["samples"] / tot, 100.0 * a["inst"] / toti, ln, top, a["src"].strip()[:110]))
