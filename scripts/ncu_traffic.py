#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of THIS build: dram__bytes_read.sum + dram__bytes_write.sum per
launch of every captured kernel (bench.py's roofline.traffic reads it while the files that define and launch the captured kernels are unchanged and the workload matches).
usage: scripts/ncu_traffic.py <rep> <workload> [more reps of the same build]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
workload = sys.argv[2]
out = {"csrc_sha": bench.csrc_sha(), "files": bench.kernel_files_sha(), "workload": workload, "kernels": {}, "source": [os.path.basename(r) for r in [sys.argv[1]] + sys.argv[3:]]}
acc = {}
for rep in [sys.argv[1]] + sys.argv[3:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    def col(r, name):
        i = h.index(name)
        v = float(r[i])
        u = units[i].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").strip()
        acc.setdefault(name, []).append(col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum"))
for k, v in acc.items():
    out["kernels"][k] = max(v)   # (several launches of one kernel: the full-size one)
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
