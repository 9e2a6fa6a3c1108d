#!/usr/bin/env python
"""Per-kernel hot source lines of an .ncu-rep: scripts/ncu_lines.py <rep> <launch-index> [topn]"""
import csv, subprocess, sys, io, collections
rep, idx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if len(r) > 6 and r[0] == "Line No"), None)
hh = rows[hi]
ci = hh.index("# Samples"); ii = hh.index("Instructions Executed")
ti = hh.index("Thread Instructions Executed") if "Thread Instructions Executed" in hh else None
stalls = [(i, c) for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[0].isdigit():
        continue
    a = agg.setdefault(int(r[0]), {"src": r[1], "samples": 0, "inst": 0, "tinst": 0, "st": collections.Counter()})
    try:
        a["samples"] += int(r[ci] or 0); a["inst"] += int(r[ii] or 0)
        if ti is not None: a["tinst"] += int(r[ti] or 0)
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
print("== lines by warp-instructions (%% of %d stall samples | %% of %d warp-instructions | lanes)" % (tot, toti))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:topn]:
    top = ",".join("%s:%d" % (k[6:], 100 * v // max(a["samples"], 1)) for k, v in a["st"].most_common(2))
    print("  %5.1f%% %5.1f%% %5.1f L%-4d %-28s %s" % (100.0 * a["samples"] / tot, 100.0 * a["inst"] / toti, a["tinst"] / max(a["inst"], 1), ln, top, a["src"].strip()[:100]))
