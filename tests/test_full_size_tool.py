"""The full-size parity tool (tests/full_size_check.py) itself: slicing the key space, in one pass or several, must not change
the oracle's result, and the numpy port of gg_graph_checksum must agree with the committed GPU bench lines it was validated on
(profiles/r2_f_fullsize_*.log keep those runs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "full_size_check.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_slicing_and_passes_do_not_change_the_result():
    a = _run("tiny")
    b = _run("tiny", "-", "1", "3")
    assert a["oracle"] == b["oracle"] and a["oracle"]["nodes"] > 0 and a["oracle"]["link_entries"] > 0


def test_committed_full_size_logs_say_equal():
    logs = [f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("r2_f_fullsize_") and f.endswith(".log")]
    assert len(logs) >= 4
    for f in logs:
        d = json.loads(open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1])
        assert d["equal"] is True, f
        g = json.loads(open(os.path.join(ROOT, d["gpu"]["file"])).read().strip().splitlines()[-1])["graph"]
        assert (g["nodes"], g["bases"], g["link_entries"], g["checksum"]) == tuple(d["oracle"][x] for x in ("nodes", "bases", "link_entries", "checksum")), f


def test_full_size_tips_tool_and_committed_log():
    """tests/full_size_tips_check.py at a small size, and the committed full-size log says CPU = GPU"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "full_size_tips_check.py"), "100000", "30", "/nonexistent"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    assert len(lines) == 2 and all(x["cpu"]["tips_left"] == 0 and x["cpu"]["tips_removed"] > 0 for x in lines)
    log = os.path.join(ROOT, "profiles", "r2_g_fullsize_remove_tips.log")
    if os.path.exists(log):
        for x in (json.loads(y) for y in open(log).read().strip().splitlines() if y.startswith("{")):
            assert x["equal"] is True
            g = [json.loads(y) for y in open(os.path.join(ROOT, x["gpu"]["file"])).read().strip().splitlines()]
            m = next(y for y in g if y["workload"] == x["workload"])
            assert m["nodes_bases_links_after"] == x["cpu"]["nodes_bases_links_after"] and m["tips_removed"] == x["cpu"]["tips_removed"]
