# Output container of the DBG construction path: the same types, field names, field order and
# sign conventions as the reference's Graphs submodule (reference src/Graphs.jl:44, :65-68,
# :163-167, :198, :232-243, :263-268, :396-399, :418-444, :473-476, :567, :592-600), so that
# graphs built on the GPU can be handed to the reference's downstream code (remove_tips!,
# write_to_gfa1, ...) unchanged.  Only what the construction path touches is defined here; the
# reference's graph-editing functions are out of scope (DESIGN.md, "Out of scope").
module Graphs

export SequenceDistanceGraph, SDG, SDGNode, SDGLink, SequenceDistanceGraphLink, NodeID,
    nodes, links, node, n_nodes, each_node_id, linksof, sequence, add_node!, add_link!,
    source, destination, distance, is_forwards_from, is_backwards_from, write_to_gfa1, find_tip_nodes

using BioSequences

const NodeID = Int64

struct SDGNode{S<:BioSequence}
    seq::S
    deleted::Bool
end

struct SequenceDistanceGraphLink
    source::NodeID       # +n: start ("+" end) of node n, -n: its end
    destination::NodeID
    dist::Int64          # -(k-1) for DBG overlaps
end
const SDGLink = SequenceDistanceGraphLink
SequenceDistanceGraphLink(src::NodeID, dst::NodeID) = SequenceDistanceGraphLink(src, dst, 0)

source(l::SDGLink) = l.source
destination(l::SDGLink) = l.destination
distance(l::SDGLink) = l.dist
# equality ignores the distance, as the reference does (src/Graphs.jl:183-185)
Base.:(==)(x::SDGLink, y::SDGLink) = x.source == y.source && x.destination == y.destination
is_forwards_from(l::SDGLink, n::NodeID) = l.source == -n
is_backwards_from(l::SDGLink, n::NodeID) = l.source == n

const LinksVecVec = Vector{Vector{SDGLink}}

struct SequenceDistanceGraph{S<:BioSequence}
    nodes::Vector{SDGNode{S}}
    links::LinksVecVec
end
const SDG = SequenceDistanceGraph
SequenceDistanceGraph{S}() where {S<:BioSequence} = SequenceDistanceGraph{S}(Vector{SDGNode{S}}(), LinksVecVec())

Base.summary(io::IO, sdg::SDG) = print(io, "Sequence distance graph (", n_nodes(sdg), " nodes)")
Base.show(io::IO, sdg::SDG) = println(io, summary(sdg))

@inline nodes(sg::SDG) = sg.nodes
@inline links(sg::SDG) = sg.links
@inline n_nodes(sg::SDG) = length(sg.nodes)
@inline each_node_id(sg::SDG) = eachindex(sg.nodes)
@inline name(sg::SDG) = :sdg

# bad ids are logged, not thrown (reference src/Graphs.jl:263-268)
@inline function check_node_id(sg::SDG, i::NodeID)
    0 < abs(i) <= n_nodes(sg) && return true
    @error "Sequence graph has no node with ID of $i"
end

@inline function node(sg::SDG, n::NodeID)
    check_node_id(sg, n)
    return @inbounds sg.nodes[abs(n)]
end
@inline function linksof(sg::SDG, n::NodeID)
    check_node_id(sg, n)
    return @inbounds sg.links[abs(n)]
end

"Node sequence; a negative id gives the reverse complement (reference src/Graphs.jl:592-600)."
function sequence(sg::SDG, n::NodeID)
    s = copy(node(sg, n).seq)
    n < 0 && reverse_complement!(s)
    return s
end

function add_node!(sg::SDG{S}, n::SDGNode{S}) where {S<:BioSequence}
    push!(sg.nodes, n)
    push!(sg.links, Vector{SDGLink}())
    return length(sg.nodes)
end
add_node!(sg::SDG{S}, seq::BioSequence) where {S<:BioSequence} = add_node!(sg, SDGNode{S}(convert(S, seq), false))

"Both directions are stored: one entry under each node (reference src/Graphs.jl:473-476)."
function add_link!(sg::SDG, src::NodeID, dst::NodeID, dist::Int)
    push!(linksof(sg, src), SDGLink(src, dst, dist))
    push!(linksof(sg, dst), SDGLink(dst, src, dist))
end

# generic functions the device-graph methods of ../dbg.jl extend (reference src/Graphs.jl:892, :814)
function write_to_gfa1 end
function find_tip_nodes end
function remove_node! end            # reference src/Graphs.jl:459
function collapse_all_unitigs! end   # reference src/Graphs.jl:528

end # module Graphs
