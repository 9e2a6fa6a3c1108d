# GenomeGraphs.jl host layer for the B200-native DBG construction path.
#
# Same module name, exports and call surface as the reference package (reference
# src/GenomeGraphs.jl:3-90) for the path reads -> k-mer counts -> dbg! -> graph; everything
# computational is a `ccall` into libggb200.so (hand-written CUDA for sm_100a, no CUDA.jl,
# no CPU fallback).  NOTE: Julia is not available in the build environment, so these files
# are syntax-reviewed only; the same C ABI is exercised end to end through the Python mirror
# (genomegraphs.jl_b200/host.py) by tests/.
module GenomeGraphs

export
    # BioSequences re-exports used on this path (reference src/GenomeGraphs.jl:25-34)
    DNAAlphabet, BioSequence, LongSequence, AbstractMer, Mer, BigMer, DNAMer, DNAKmer, BigDNAMer, BigDNAKmer,
    # KmerAnalysis re-exports (reference src/GenomeGraphs.jl:39-48)
    mer, freq, Canonical, NonCanonical, CANONICAL, NONCANONICAL, serial_mem, spectra,
    empty_graph,
    WorkSpace, add_mer_counts!, mer_counts,
    dbg, dbg!, remove_tips!,
    # device-side counter (drop-in for serial_mem(M, CANONICAL))
    gpu_mem, GPUContext, DeviceGraph, load_from_gfa1

using BioSequences
using KmerAnalysis
import KmerAnalysis: MerCount, mer, freq, Canonical, NonCanonical, CANONICAL, NONCANONICAL, serial_mem, spectra

include("libggb200.jl")    # ccall bindings of include/ggb200.h
include("Graphs.jl")       # SequenceDistanceGraph, SDGNode, SDGLink, add_node!, add_link!
include("GraphIndexes.jl") # UniqueMerIndex
include("counting.jl")     # gpu_mem: device-side serial_mem

"Make an empty graph (reference src/GenomeGraphs.jl:85)."
empty_graph(::Type{T}) where {T<:LongSequence} = Graphs.SequenceDistanceGraph{T}()

include("dbg.jl")
include("workspace/WorkSpace.jl")

end # module GenomeGraphs
