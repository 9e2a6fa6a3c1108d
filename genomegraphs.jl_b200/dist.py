"""Multi-GPU host layer: one process per GPU.

Plays the role of KmerAnalysis `dist_mem` (re-exported by the reference at src/GenomeGraphs.jl:46-47): every rank
counts on ITS shard of the reads.  For 17 <= k <= 32 the super-k-mers are partitioned by minimizer bin and stored into
the owner GPU's buffer over NVLink (bulk asynchronous copies), each rank counts its bins, and a second exchange
range-partitions the kept k-mers, so that every rank ends up owning a contiguous key range of the global sorted
(k-mer, count) table; other k partition W-word keys by range directly.  The graph stage is sharded by vertex range.
All exchanges run inside libggb200.so (csrc/dist.cuh, csrc/graphdist.cuh: NCCL + CUDA IPC peer memory);
`torch.distributed` is only the host channel that ships the 128-byte NCCL unique id (any backend, gloo included).
"""
import ctypes as C

import numpy as np

from . import _native as N
from .host import CANONICAL, Context, MerCounts, _as_reads, _export_graph


def shard_range(n_items, rank, world, multiple=1):
    """[first, first + count) of rank's contiguous shard; boundaries are multiples of `multiple`
    (2 keeps read pairs together).  The shards tile [0, n_items) exactly."""
    n_units = n_items // multiple
    lo = (n_units * rank) // world * multiple
    hi = (n_units * (rank + 1)) // world * multiple
    if rank == world - 1:
        hi = n_items
    return lo, hi - lo


def splitters(hist, world):
    """Owner ranges of histogram buckets (gg_dist_splitters): rank q owns [first[q], first[q+1])."""
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    first = np.zeros(world + 1, dtype=np.uint32)
    code = N.lib().gg_dist_splitters(h.ctypes.data_as(C.c_void_p), len(h), int(world), first.ctypes.data_as(C.c_void_p))
    if code != N.GG_OK:
        raise ValueError("bad splitter arguments")
    return first


def broadcast_bytes(payload, src=0, group=None):
    """Ship a small byte string from rank `src` to every rank over torch.distributed."""
    import torch.distributed as dist
    box = [payload if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return bytes(box[0])


class Comm:
    """gg_comm: NCCL communicator of the ranks that build one graph together."""

    def __init__(self, ctx, rank, world, unique_id):
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id has 128 bytes")
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        N.check(ctx.h, N.lib().gg_comm_create(ctx.h, buf, int(rank), int(world), C.byref(h)))
        self.ctx, self.h, self.rank, self.world = ctx, h, int(rank), int(world)

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        if N.lib().gg_comm_unique_id(buf) != N.GG_OK:
            raise N.GGError(N.GG_ERR_CUDA, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    @classmethod
    def from_torch(cls, ctx, group=None):
        """Bootstrap over an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = broadcast_bytes(cls.unique_id() if rank == 0 else None, 0, group)
        return cls(ctx, rank, world, uid)

    def close(self):
        if self.h:
            N.lib().gg_comm_destroy(self.h)
            self.h = None


def dist_count_dev(comm, d_seq, d_offs, nreads, nbytes, k, canonical=CANONICAL, min_freq=1):
    """Device buffers of this rank's reads -> (MerCounts slice of this rank, global #instances)."""
    h, ng = C.c_void_p(), C.c_uint64()
    N.check(comm.ctx.h, N.lib().gg_dist_count_kmers_dev(comm.h, d_seq, d_offs, int(nreads), int(nbytes), int(k), int(canonical),
                                                        int(min_freq), C.byref(h), C.byref(ng)))
    return MerCounts(comm.ctx, h), ng.value


class _DeviceReads:
    """host read buffers staged on the device for the *_dev entry points"""

    def __init__(self, ctx, reads):
        seq, offs = _as_reads(reads)
        L = N.lib()
        self.ctx, self.nreads, self.nbytes = ctx, len(offs) - 1, int(offs[-1])
        self.d_seq, self.d_offs = C.c_void_p(), C.c_void_p()
        N.check(ctx.h, L.gg_dev_alloc(ctx.h, self.nbytes + 64, C.byref(self.d_seq)))
        N.check(ctx.h, L.gg_dev_alloc(ctx.h, offs.nbytes, C.byref(self.d_offs)))
        if self.nbytes:
            N.check(ctx.h, L.gg_memcpy_h2d(ctx.h, self.d_seq, seq.ctypes.data_as(C.c_void_p), self.nbytes))
        N.check(ctx.h, L.gg_memcpy_h2d(ctx.h, self.d_offs, offs.ctypes.data_as(C.c_void_p), offs.nbytes))

    def free(self):
        L = N.lib()
        L.gg_dev_free(self.ctx.h, self.d_seq)
        L.gg_dev_free(self.ctx.h, self.d_offs)


def dist_mem(k, mode=CANONICAL, comm=None, min_freq=1):
    """`dist_mem(DNAMer{k}, CANONICAL)`-style counter over a communicator: counter(reads_of_this_rank)
    -> MerCounts holding this rank's key range of the global table."""
    def counter(reads):
        dr = _DeviceReads(comm.ctx, reads)
        try:
            return dist_count_dev(comm, dr.d_seq, dr.d_offs, dr.nreads, dr.nbytes, k, mode, min_freq)[0]
        finally:
            dr.free()
    return counter


def counts_allgather(comm, counts):
    h = C.c_void_p()
    N.check(comm.ctx.h, N.lib().gg_counts_allgather(comm.h, counts.h, C.byref(h)))
    return MerCounts(comm.ctx, h)


def dist_dbg_from_reads(comm, reads, k, min_freq):
    """reads of this rank -> the whole graph (GraphArrays) on every rank."""
    dr = _DeviceReads(comm.ctx, reads)
    gh, ng = C.c_void_p(), C.c_uint64()
    try:
        N.check(comm.ctx.h, N.lib().gg_dist_dbg_from_reads_dev(comm.h, dr.d_seq, dr.d_offs, dr.nreads, dr.nbytes, int(k), int(min_freq),
                                                               C.byref(gh), C.byref(ng)))
        return _export_graph(comm.ctx, gh)
    finally:
        if gh:
            N.lib().gg_graph_free(gh)
        dr.free()
