# UniqueMerIndex built on the GPU (reference src/GraphIndexes.jl:14-17, :27-41, :62-102; the
# constructor there references unimported names and mis-sizes unique_mers_per_node -- this file
# implements the documented intent: one entry per node).
module GraphIndexes

export UniqueMerIndex, GraphStrandPosition, find_unique_mer, isunmappable

using BioSequences
import ..Graphs
import ..LibGG

struct GraphStrandPosition
    node::Graphs.NodeID   # sign = strand of the canonical k-mer relative to the node
    position::UInt32      # 1-based start of the k-mer in the node sequence
end
Base.show(io::IO, x::GraphStrandPosition) = print(io, "(Node: ", x.node, ", Base position: ", x.position, ')')

struct UniqueMerIndex{M<:AbstractMer}
    mer_to_graphposition::Dict{M,GraphStrandPosition}
    unique_mers_per_node::Vector{UInt64}
    total_mers_per_node::Vector{UInt64}
end

isunmappable(idx::UniqueMerIndex, nodeid::Graphs.NodeID) = idx.unique_mers_per_node[abs(nodeid)] == 0

function find_unique_mer(idx::UniqueMerIndex{M}, mer::M) where {M<:AbstractMer}
    pos = get(idx.mer_to_graphposition, mer, GraphStrandPosition(0, 0))
    return pos != GraphStrandPosition(0, 0), pos
end

# (the library picks the word count from k, not from the Julia type: BigMer{A,K} with K <= 32 is one word)
function mer_from_words(::Type{M}, words, i, w) where {M<:AbstractMer}
    x = w == 1 ? UInt128(words[i]) : ((UInt128(words[2i - 1]) << 64) | UInt128(words[2i]))
    return M <: BigMer ? M(x) : M(UInt64(x))
end

function UniqueMerIndex{M}(sg::Graphs.SequenceDistanceGraph; ctx::LibGG.Context = LibGG.default_context()) where {M<:AbstractMer}
    buf = IOBuffer()
    offs = UInt64[0]
    for n in Graphs.nodes(sg)
        print(buf, n.seq)
        push!(offs, UInt64(position(buf)))
    end
    words, node, pos, upn, tpn, w = LibGG.unique_mer_index(ctx, take!(buf), offs, BioSequences.ksize(M))
    d = Dict{M,GraphStrandPosition}()
    sizehint!(d, length(node))
    for i in eachindex(node)
        d[mer_from_words(M, words, i, w)] = GraphStrandPosition(node[i], pos[i])
    end
    return UniqueMerIndex{M}(d, upn, tpn)
end

end # module GraphIndexes
