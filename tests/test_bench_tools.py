"""bench.py's bookkeeping that can be checked without a GPU: the DRAM-traffic table (profiles/ncu_traffic.json) is only
used while the files that define and launch the captured kernels are the ones the ncu capture saw."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_table_is_bound_to_the_kernel_files(monkeypatch):
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert set(t["files"]) == set(bench.TRAFFIC_FILES)
    if t["files"] == bench.kernel_files_sha():
        v = bench.ncu_traffic("skm_leaf_kernel", t["workload"])
        assert isinstance(v, float) and v > 1e9
        assert bench.ncu_traffic("skm_leaf_kernel", "some other workload") is None
    else:   # kernels changed since the capture: the bench line must say traffic = null
        assert bench.ncu_traffic("skm_leaf_kernel", t["workload"]) is None
    changed = dict(bench.kernel_files_sha(), **{"skm.cuh": "0" * 16})
    monkeypatch.setattr(bench, "kernel_files_sha", lambda read=None: changed)
    monkeypatch.setattr(bench, "csrc_sha", lambda: "f" * 16)
    assert bench.ncu_traffic("skm_leaf_kernel", t["workload"]) is None
