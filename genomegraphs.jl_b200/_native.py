"""ctypes binding of libggb200.so (include/ggb200.h).  No torch types cross this boundary.

The library is built in-tree (genomegraphs.jl_b200/libggb200.so) by `build()`; importing
the package never falls back to a CPU implementation: without the shared library or
without a CUDA device every compute entry point raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libggb200.so")
CSRC = os.path.join(HERE, "csrc")

GG_OK, GG_ERR_ARG, GG_ERR_CUDA, GG_ERR_OOM, GG_ERR_LIMIT = 0, -1, -2, -3, -4

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/ggb200.h one to one
SIGNATURES = {
    "gg_version": (C.c_int, []),
    "gg_kmer_words": (C.c_int, [C.c_int]),
    "gg_ctx_create": (C.c_int, [C.c_int, vpp]),
    "gg_ctx_destroy": (None, [vp]),
    "gg_last_error": (C.c_char_p, [vp]),
    "gg_host_alloc": (C.c_int, [vp, C.c_uint64, vpp]),
    "gg_host_free": (None, [vp, vp]),
    "gg_dev_alloc": (C.c_int, [vp, C.c_uint64, vpp]),
    "gg_dev_free": (None, [vp, vp]),
    "gg_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "gg_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "gg_ctx_last_timings": (C.c_int, [vp, f32p, u64p]),
    "gg_ctx_event_record": (C.c_int, [vp, C.c_int]),
    "gg_ctx_event_elapsed_ms": (C.c_int, [vp, C.c_int, C.c_int, f32p]),
    "gg_ctx_sync": (C.c_int, [vp]),
    "gg_ctx_set_profile": (C.c_int, [vp, C.c_int]),
    "gg_ctx_kernel_profile": (C.c_char_p, [vp]),
    "gg_count_kmers": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vpp]),
    "gg_count_kmers_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vpp]),
    "gg_counts_from_kmers": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, vpp]),
    "gg_counts_sizes": (C.c_int, [vp, u64p, u64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gg_counts_filter": (C.c_int, [vp, C.c_uint64, C.c_int]),
    "gg_counts_export": (C.c_int, [vp, vp, vp, vp]),
    "gg_counts_spectrum": (C.c_int, [vp, u64p]),
    "gg_counts_spectrum_cn": (C.c_int, [vp, vp, u64p]),
    "gg_counts_free": (None, [vp]),
    "gg_count_kmers_4bit": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vpp]),
    "gg_dbg_from_reads_4bit": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_uint64, vpp]),
    "gg_graph_find_tip_nodes": (C.c_int, [vp, C.c_uint64, vp, u64p]),
    "gg_graph_remove_nodes": (C.c_int, [vp, vp, C.c_uint64, vpp]),
    "gg_graph_collapse_all_unitigs": (C.c_int, [vp, C.c_uint64, vpp, u64p]),
    "gg_graph_remove_tips": (C.c_int, [vp, C.c_uint64, vpp, u64p, u64p]),
    "gg_graph_write_gfa1": (C.c_int, [vp, C.c_char_p]),
    "gg_gfa1_write": (C.c_int, [C.c_char_p, C.c_uint64, vp, vp, vp, vp]),
    "gg_graph_load_gfa1": (C.c_int, [vp, C.c_char_p, C.c_char_p, vpp]),
    "gg_graph_from_arrays": (C.c_int, [vp, C.c_int, C.c_uint64, vp, vp, vp, vp, vpp]),
    "gg_extract_kmers": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp, u64p]),
    "gg_dbg_build": (C.c_int, [vp, vp, vpp]),
    "gg_dbg_from_reads": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_uint64, vpp]),
    "gg_dbg_from_reads_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, vpp]),
    "gg_graph_sizes": (C.c_int, [vp, u64p, u64p, u64p, C.POINTER(C.c_int)]),
    "gg_graph_export": (C.c_int, [vp, vp, vp, vp, vp]),
    "gg_graph_checksum": (C.c_int, [vp, u64p]),
    "gg_graph_free": (None, [vp]),
    "gg_unique_mer_index": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, vpp]),
    "gg_umi_sizes": (C.c_int, [vp, u64p, u64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gg_umi_export": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "gg_umi_free": (None, [vp]),
    "gg_synth_reads_host": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "gg_synth_reads_dev": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "gg_synth_genome_host": (C.c_int, [C.c_uint64, C.c_uint64, vp]),
    "gg_synth_reads_range_dev": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "gg_synth_reads_range_host": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "gg_comm_unique_id": (C.c_int, [vp]),
    "gg_comm_create": (C.c_int, [vp, vp, C.c_int, C.c_int, vpp]),
    "gg_comm_destroy": (None, [vp]),
    "gg_dist_splitters": (C.c_int, [vp, C.c_uint32, C.c_int, vp]),
    "gg_dist_count_kmers_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, vpp, u64p]),
    "gg_counts_allgather": (C.c_int, [vp, vp, vpp]),
    "gg_dist_dbg_from_reads_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, vpp, u64p]),
}

_LIB = None


def build(force=False):
    """Compile libggb200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ggb200.h")]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "../libggb200.so"] + (["-B"] if force else []))
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("libggb200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class GGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libggb200 error %d: %s" % (code, msg))
        self.code = code


def check(ctx_handle, code):
    if code != GG_OK:
        msg = lib().gg_last_error(ctx_handle).decode(errors="replace") if ctx_handle else ""
        if code == GG_ERR_ARG:
            raise ValueError(msg or "bad argument")  # reference: ArgumentError
        raise GGError(code, msg)
