# A live WorkSpace for the construction path.  In the reference v0.2.0 this file is not
# included by the module and names types that exist nowhere (src/workspace/WorkSpace.jl:2-17,
# src/GenomeGraphs.jl:53-58,75); only the field / accessor names are kept: a graph plus named
# k-mer count stores, here device-resident.
mutable struct WorkSpace
    sdg::Graphs.SequenceDistanceGraph{LongSequence{DNAAlphabet{4}}}
    mer_count_stores::Dict{Symbol,DeviceMerCounts}
    ctx::LibGG.Context
end
WorkSpace(; ctx::LibGG.Context = LibGG.default_context()) =
    WorkSpace(Graphs.SequenceDistanceGraph{LongSequence{DNAAlphabet{4}}}(), Dict{Symbol,DeviceMerCounts}(), ctx)

graph(ws::WorkSpace) = ws.sdg

"Count the k-mers of `reads` on the GPU and keep the table in the workspace under `name`."
function add_mer_counts!(ws::WorkSpace, ::Type{M}, mode, name::Symbol, reads) where {M<:AbstractMer}
    ws.mer_count_stores[name] = gpu_mem(M, mode; ctx = ws.ctx)(reads)
    return ws
end

function mer_counts(ws::WorkSpace, name::Symbol)
    haskey(ws.mer_count_stores, name) || error(string("WorkSpace has no mer count datastore datastore called ", name))
    return ws.mer_count_stores[name]
end

"`dbg!(ws, name, min_freq)`: build the workspace graph from a stored count table."
dbg!(ws::WorkSpace, name::Symbol, min_freq::Integer) = (dbg!(ws.sdg, mer_counts(ws, name), min_freq); ws)

"`remove_tips!(ws, min_size)`: the WorkSpace form the reference has commented out (src/remove_tips.jl:1-5, :28-29)."
remove_tips!(ws::WorkSpace, min_size::Integer) = (remove_tips!(ws.sdg, min_size; ctx = ws.ctx); ws)

function Base.show(io::IO, ws::WorkSpace)
    println(io, "Graph Genome workspace")
    println(io, " Sequence distance graph (", Graphs.n_nodes(ws.sdg), " nodes)")
    println(io, " Mer count stores: ", length(ws.mer_count_stores))
    for (nm, c) in ws.mer_count_stores
        println(io, "  ", nm, ": ", length(c), " distinct ", c.k, "-mers")
    end
end
