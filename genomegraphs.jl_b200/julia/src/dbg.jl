# dbg! on the GPU: same methods, argument meaning, mutation and error behaviour as the
# reference (src/dbg.jl:261-272 and :274-278).  build_unitigs! (:94-185) and
# connect_unitigs_by_overlaps! (:224-240) run as CUDA kernels behind gg_dbg_build; node ids,
# node orientation and the order of every node's link vector come back exactly as the
# reference's sequential walk would have produced them.

const GRAPH_TYPE = Graphs.SequenceDistanceGraph{LongSequence{DNAAlphabet{4}}}

# k-mer value -> W UInt64 words, most significant word first (include/ggb200.h).  W comes from the LIBRARY
# (gg_kmer_words(k): 1 word for k <= 32, 2 for k <= 64), not from the Julia type: BigMer{A,K} with K <= 32 is legal
# in the reference and still travels as one word.
@inline kmer_value(m::AbstractMer) = UInt128(BioSequences.encoded_data(m))
words_per_mer(::Type{M}) where {M<:AbstractMer} = LibGG.kmer_words(BioSequences.ksize(M))

function pack_kmers(kmers::AbstractVector{M}) where {M<:AbstractMer}
    w = words_per_mer(M)
    w in (1, 2) || throw(ArgumentError("k-mers of this length are not representable as Mer / BigMer"))
    words = Vector{UInt64}(undef, w * length(kmers))
    for (i, m) in enumerate(kmers)
        x = kmer_value(m)
        if w == 1
            @inbounds words[i] = UInt64(x & typemax(UInt64))
        else
            @inbounds words[2i - 1] = UInt64(x >> 64)
            @inbounds words[2i] = UInt64(x & typemax(UInt64))
        end
    end
    return words
end

"word `i` (1-based k-mer index) of an exported table of `w`-word k-mers -> `M`"
@inline function mer_from_words(::Type{M}, words::Vector{UInt64}, i::Int, w::Int) where {M<:AbstractMer}
    x = w == 1 ? UInt128(words[i]) : ((UInt128(words[2i - 1]) << 64) | UInt128(words[2i]))
    return M <: BigMer ? M(x) : M(UInt64(x))
end

@inline function dbg_message(x::AbstractVector{M}) where {M<:AbstractMer}
    return string("Constructing a compressed de-bruijn graph from ", length(x), ' ', BioSequences.ksize(M), "-mers")
end

# Append the exported nodes and links to the graph.  Like build_unitigs! (src/dbg.jl:156, :179) the nodes are pushed
# behind whatever the graph already holds, so on a non-empty graph the new ids (and the ids inside the new links) are
# shifted by the number of nodes that were there.  One deviation, stated: the reference's link pass
# (src/dbg.jl:224-240) runs over ALL nodes of the graph and would also link old nodes with new ones (and re-add the old
# links a second time); here only the links among the new nodes are made -- dbg! is meant for an empty graph
# (`empty_graph`, docs/src/man/assembly.md:165-190), which is the case in which both agree exactly.
function fill_graph!(sdg::GRAPH_TYPE, node_off, bases, link_row, tri)
    nn = length(node_off) - 1
    base_id = Graphs.n_nodes(sdg)
    base_id == 0 || @warn "dbg! on a non-empty graph: links between the existing nodes and the new unitigs are not created"
    sizehint!(Graphs.nodes(sdg), base_id + nn)
    for i in 1:nn
        lo, hi = Int(node_off[i]) + 1, Int(node_off[i + 1])
        Graphs.add_node!(sdg, LongSequence{DNAAlphabet{4}}(String(bases[lo:hi])))
    end
    shift(id) = id > 0 ? id + base_id : id - base_id
    for i in 1:nn
        lv = Graphs.links(sdg)[base_id + i]
        for j in (Int(link_row[i]) + 1):Int(link_row[i + 1])
            push!(lv, Graphs.SDGLink(shift(tri[3j - 2]), shift(tri[3j - 1]), tri[3j]))
        end
    end
    return sdg
end

function dbg_from_handle!(sdg::GRAPH_TYPE, ctx::LibGG.Context, counts::Ptr{Cvoid})
    g = LibGG.dbg_build(ctx, counts)
    try
        fill_graph!(sdg, LibGG.graph_export(ctx, g)...)
    finally
        LibGG.graph_free(g)
    end
    return sdg
end

"`dbg!(sdg, kmers)`: canonical, duplicate-free k-mers in any order (sorted on the device if needed)."
function dbg!(sdg::GRAPH_TYPE, kmers; ctx::LibGG.Context = LibGG.default_context())
    if !(eltype(kmers) <: AbstractMer)
        throw(ArgumentError("Didn't provide a collection of kmers for the `kmers` parameter did ya?"))
    end
    @info dbg_message(kmers)
    M = eltype(kmers)
    # build_unitigs! sorts the CALLER's vector when it is not sorted (src/dbg.jl:98-100): same visible side effect
    # (the device would sort its own copy otherwise, leaving the caller's order unchanged)
    if kmers isa AbstractVector && !issorted(kmers)
        sort!(kmers)
    end
    c = LibGG.counts_from_kmers(ctx, pack_kmers(kmers), length(kmers), BioSequences.ksize(M))
    try
        dbg_from_handle!(sdg, ctx, c)
    finally
        LibGG.counts_free(c)
    end
    @info "Done"
    return sdg
end

"`dbg!(sg, counted_kmers, min_freq)`: filters the caller's vector in place, like the reference."
function dbg!(sg::GRAPH_TYPE, counted_kmers::Vector{<:MerCount}, min_freq::Integer; ctx::LibGG.Context = LibGG.default_context())
    filter!(x -> freq(x) >= min_freq, counted_kmers)
    merlist = [mer(x) for x in counted_kmers]
    dbg!(sg, merlist; ctx = ctx)
end

"Counts that never left the GPU (see `gpu_mem`): filter and build on the device."
function dbg!(sg::GRAPH_TYPE, counted::DeviceMerCounts, min_freq::Integer)
    LibGG.counts_filter!(counted.ctx, counted.h, min_freq, true)   # UInt8-saturated `freq`, as MerCount
    @info string("Constructing a compressed de-bruijn graph from ", length(counted), ' ', counted.k, "-mers")
    dbg_from_handle!(sg, counted.ctx, counted.h)
    @info "Done"
    return sg
end

"Non-mutating form (exported but undefined in the reference, src/GenomeGraphs.jl:63)."
dbg(kmers; kwargs...) = dbg!(empty_graph(LongSequence{DNAAlphabet{4}}), kmers; kwargs...)
dbg(counted, min_freq::Integer; kwargs...) = dbg!(empty_graph(LongSequence{DNAAlphabet{4}}), counted, min_freq; kwargs...)


# ---- the graph kept on the device: persistence and the tip finder ----------------------------------
"A `gg_graph` handle (device-resident nodes + links)."
mutable struct DeviceGraph
    ctx::LibGG.Context
    h::Ptr{Cvoid}
    function DeviceGraph(ctx, h)
        g = new(ctx, h)
        finalizer(x -> (x.h == C_NULL || LibGG.graph_free(x.h); x.h = C_NULL), g)
        return g
    end
end
"Upload a host graph (deleted nodes travel as empty sequences)."
function DeviceGraph(sg::GRAPH_TYPE, k::Integer; ctx::LibGG.Context = LibGG.default_context())
    buf = IOBuffer(); node_off = UInt64[0]; link_row = UInt64[0]; tri = Int64[]
    for (i, n) in enumerate(Graphs.nodes(sg))
        n.deleted || print(buf, n.seq)
        push!(node_off, UInt64(position(buf)))
        for l in Graphs.links(sg)[i]
            push!(tri, l.source, l.destination, l.dist)
        end
        push!(link_row, UInt64(length(tri) ÷ 3))
    end
    return DeviceGraph(ctx, LibGG.graph_from_arrays(ctx, k, node_off, take!(buf), link_row, tri))
end
"`write_to_gfa1(sg, filename)` (src/Graphs.jl:892-926): `filename.gfa` + `filename.fasta`, same bytes."
Graphs.write_to_gfa1(g::DeviceGraph, filename::String) = LibGG.graph_write_gfa1(g.ctx, g.h, filename)
"`load_from_gfa1!` is unfinished upstream (src/Graphs.jl:943-958); this loads what `write_to_gfa1` writes."
load_from_gfa1(gfafile::AbstractString, fafile::AbstractString; ctx::LibGG.Context = LibGG.default_context()) =
    DeviceGraph(ctx, LibGG.graph_load_gfa1(ctx, gfafile, fafile))
"`find_tip_nodes(sg, min_size)` (src/Graphs.jl:778-816) on the device graph; ids ascending."
Graphs.find_tip_nodes(g::DeviceGraph, min_size::Integer) = Set{Graphs.NodeID}(LibGG.graph_find_tip_nodes(g.ctx, g.h, min_size))

# ---- tip removal (reference src/remove_tips.jl:2-30) --------------------------------------------------------------
# The device entry points return a NEW handle; the `!` methods below keep the reference's mutating contract by
# swapping the handle inside the DeviceGraph (or by writing the result back into the host graph).
function swap_handle!(g::DeviceGraph, h::Ptr{Cvoid})
    LibGG.graph_free(g.h)
    g.h = h
    return g
end
"`remove_node!(sg, n)` (src/Graphs.jl:459-466); a vector of ids removes them all in one pass."
Graphs.remove_node!(g::DeviceGraph, n::Graphs.NodeID) = swap_handle!(g, LibGG.graph_remove_nodes(g.ctx, g.h, Int64[n]))
Graphs.remove_node!(g::DeviceGraph, ns::AbstractVector{<:Integer}) = swap_handle!(g, LibGG.graph_remove_nodes(g.ctx, g.h, collect(Int64, ns)))
"`collapse_all_unitigs!(sg, min_nodes, consume)` (src/Graphs.jl:528-539); only `consume = true` runs on the device."
function Graphs.collapse_all_unitigs!(g::DeviceGraph, min_nodes::Integer, consume::Bool)
    consume || throw(ArgumentError("collapse_all_unitigs! on a DeviceGraph needs consume = true"))
    h, _ = LibGG.graph_collapse_all_unitigs(g.ctx, g.h, min_nodes)
    swap_handle!(g, h)
    return nothing
end
"`remove_tips!(sdg, min_size)` (src/remove_tips.jl:2-30) on the device-resident graph."
function remove_tips!(g::DeviceGraph, min_size::Integer)
    @info "Beginning tip removal process"
    h, passes, ntips = LibGG.graph_remove_tips(g.ctx, g.h, min_size)
    swap_handle!(g, h)
    @info string("Removed ", ntips, " tips")
    @info string("Finished tip removal process in ", passes, " passes")
    return g
end
"The reference's signature: the host graph is uploaded, edited on the device and replaced in place."
function remove_tips!(sdg::GRAPH_TYPE, min_size::Integer; ctx::LibGG.Context = LibGG.default_context())
    g = remove_tips!(DeviceGraph(sdg, 0; ctx = ctx), min_size)
    node_off, bases, link_row, tri = LibGG.graph_export(g.ctx, g.h)
    empty!(Graphs.nodes(sdg)); empty!(Graphs.links(sdg))
    fill_graph!(sdg, node_off, bases, link_row, tri)
    for (i, n) in enumerate(Graphs.nodes(sdg))      # an emptied node is the reference's empty_node (deleted = true)
        isempty(n.seq) && (Graphs.nodes(sdg)[i] = Graphs.SDGNode{LongSequence{DNAAlphabet{4}}}(n.seq, true))
    end
    return sdg
end

"Device graph -> the reference's graph type."
function Base.convert(::Type{GRAPH_TYPE}, g::DeviceGraph)
    return fill_graph!(empty_graph(LongSequence{DNAAlphabet{4}}), LibGG.graph_export(g.ctx, g.h)...)
end
