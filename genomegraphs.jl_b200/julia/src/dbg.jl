# dbg! on the GPU: same methods, argument meaning, mutation and error behaviour as the
# reference (src/dbg.jl:261-272 and :274-278).  build_unitigs! (:94-185) and
# connect_unitigs_by_overlaps! (:224-240) run as CUDA kernels behind gg_dbg_build; node ids,
# node orientation and the order of every node's link vector come back exactly as the
# reference's sequential walk would have produced them.

const GRAPH_TYPE = Graphs.SequenceDistanceGraph{LongSequence{DNAAlphabet{4}}}

# k-mer value -> W little-endian UInt64 words, most significant word first (include/ggb200.h)
@inline kmer_words!(dst::Vector{UInt64}, i::Int, m::Mer{A,K}) where {A,K} =
    (@inbounds dst[i] = UInt64(BioSequences.encoded_data(m)); nothing)
@inline function kmer_words!(dst::Vector{UInt64}, i::Int, m::BigMer{A,K}) where {A,K}
    x = UInt128(BioSequences.encoded_data(m))
    @inbounds dst[2i - 1] = UInt64(x >> 64)
    @inbounds dst[2i] = UInt64(x & typemax(UInt64))
    nothing
end
words_per_mer(::Type{<:Mer}) = 1
words_per_mer(::Type{<:BigMer}) = 2

function pack_kmers(kmers::AbstractVector{M}) where {M<:AbstractMer}
    w = words_per_mer(M)
    words = Vector{UInt64}(undef, w * length(kmers))
    for (i, m) in enumerate(kmers)
        kmer_words!(words, i, m)
    end
    return words
end

@inline function dbg_message(x::AbstractVector{M}) where {M<:AbstractMer}
    return string("Constructing a compressed de-bruijn graph from ", length(x), ' ', BioSequences.ksize(M), "-mers")
end

# fill the (empty) graph from the exported arrays
function fill_graph!(sdg::GRAPH_TYPE, node_off, bases, link_row, tri)
    nn = length(node_off) - 1
    base_id = Graphs.n_nodes(sdg)
    base_id == 0 || throw(ArgumentError("dbg! expects an empty graph (link ids are absolute)"))
    sizehint!(Graphs.nodes(sdg), nn)
    for i in 1:nn
        lo, hi = Int(node_off[i]) + 1, Int(node_off[i + 1])
        Graphs.add_node!(sdg, LongSequence{DNAAlphabet{4}}(String(bases[lo:hi])))
    end
    for i in 1:nn
        lv = Graphs.links(sdg)[i]
        for j in (Int(link_row[i]) + 1):Int(link_row[i + 1])
            push!(lv, Graphs.SDGLink(tri[3j - 2], tri[3j - 1], tri[3j]))
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
