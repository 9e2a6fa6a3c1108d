"""genomegraphs.jl_b200 -- B200-native (sm_100a) de Bruijn graph construction path of
GenomeGraphs.jl: hand-written CUDA kernels behind a C ABI (libggb200.so), a Python host
mirror of the Julia call surface (host.py) and the Julia host files (julia/).

The directory name carries a dot, so import it through `__graft_entry__.load_package()`
(registered in sys.modules as `genomegraphs_b200`).
"""
from . import _native  # noqa: F401
from . import dist  # noqa: F401
from .host import *  # noqa: F401,F403
from .host import (CANONICAL, NONCANONICAL, Context, DeviceGraph, GraphArrays, GraphStrandPosition, MerCounts, SDGLink, SDGNode,  # noqa: F401
                   SequenceDistanceGraph, UniqueMerIndex, WorkSpace, dbg_, dbg_arrays, dbg_from_reads, empty_graph, extract_kmers, freq,
                   mer, pack_reads, read_sequences, serial_mem, spectra, synth_genome, synth_reads,
                   count_kmers_4bit, dbg_from_reads_4bit, pack_4bit, remove_tips_, spectra_cn, write_gfa1_arrays, PairedReads, read_datastore, FwRv, RvFw)
