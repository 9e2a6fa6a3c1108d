# Thin `ccall` layer over libggb200.so (include/ggb200.h).  One GPUContext = one GPU.
module LibGG

using Libdl

const LIB = Ref{String}(get(ENV, "GGB200_LIB", joinpath(@__DIR__, "..", "..", "libggb200.so")))

# The library is opened at first use and every entry point is called through its address:
# `ccall((:sym, lib), ...)` needs a CONSTANT library expression (Julia 1.0, the oldest version in
# the reference's CI matrix, rejects anything else), a function pointer works everywhere.
const HANDLE = Ref{Ptr{Cvoid}}(C_NULL)
const SYMS = Dict{Symbol,Ptr{Cvoid}}()
function fptr(s::Symbol)
    p = get(SYMS, s, C_NULL)
    if p == C_NULL
        HANDLE[] == C_NULL && (HANDLE[] = Libdl.dlopen(LIB[]))
        p = Libdl.dlsym(HANDLE[], s)
        SYMS[s] = p
    end
    return p
end

const GG_OK = Cint(0)
const GG_ERR_ARG = Cint(-1)

struct GGError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::GGError) = print(io, "libggb200 error ", e.code, ": ", e.msg)

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall(fptr(:gg_ctx_create), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
        rc == GG_OK || throw(GGError(rc, "no usable CUDA device $device (libggb200 has no CPU fallback)"))
        ctx = new(out[])
        # finalizers run in no particular order: the library keeps a context alive until its last handle is freed
        # (gg_ctx_destroy only marks it while gg_counts / gg_graph / gg_umi objects exist)
        finalizer(c -> (c.h == C_NULL || ccall(fptr(:gg_ctx_destroy), Cvoid, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        return ctx
    end
end

const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT[] === nothing && (DEFAULT[] = Context(0)); DEFAULT[]::Context)

function check(ctx::Context, rc::Cint)
    rc == GG_OK && return nothing
    msg = unsafe_string(ccall(fptr(:gg_last_error), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == GG_ERR_ARG && throw(ArgumentError(msg))   # reference: ArgumentError (src/dbg.jl:262-264)
    throw(GGError(rc, msg))
end

kmer_words(k::Integer) = Int(ccall(fptr(:gg_kmer_words), Cint, (Cint,), k))

# ---- counts ---------------------------------------------------------------------------------
function count_kmers(ctx::Context, seq::Vector{UInt8}, offs::Vector{UInt64}, k::Integer, canonical::Bool)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve seq offs begin
        check(ctx, ccall(fptr(:gg_count_kmers), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, seq, offs, length(offs) - 1, k, canonical, out))
    end
    return out[]
end

function counts_from_kmers(ctx::Context, words::Vector{UInt64}, n::Integer, k::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve words begin
        check(ctx, ccall(fptr(:gg_counts_from_kmers), Cint,
                         (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt32}, UInt64, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, words, C_NULL, n, k, out))
    end
    return out[]
end

function counts_sizes(c::Ptr{Cvoid})
    nu = Ref{UInt64}(0); ni = Ref{UInt64}(0); k = Ref{Cint}(0); w = Ref{Cint}(0)
    ccall(fptr(:gg_counts_sizes), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}, Ref{Cint}), c, nu, ni, k, w)
    return Int(nu[]), Int(ni[]), Int(k[]), Int(w[])
end

counts_filter!(ctx::Context, c::Ptr{Cvoid}, min_freq::Integer, u8sat::Bool) =
    check(ctx, ccall(fptr(:gg_counts_filter), Cint, (Ptr{Cvoid}, UInt64, Cint), c, min_freq, u8sat))

function counts_export(ctx::Context, c::Ptr{Cvoid})
    nu, _, _, w = counts_sizes(c)
    words = Vector{UInt64}(undef, nu * w)
    c8 = Vector{UInt8}(undef, nu)
    GC.@preserve words c8 begin
        check(ctx, ccall(fptr(:gg_counts_export), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt32}), c, words, c8, C_NULL))
    end
    return words, c8, w
end

counts_free(c::Ptr{Cvoid}) = ccall(fptr(:gg_counts_free), Cvoid, (Ptr{Cvoid},), c)

# ---- graph ----------------------------------------------------------------------------------
function dbg_build(ctx::Context, c::Ptr{Cvoid})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(fptr(:gg_dbg_build), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, c, out))
    return out[]
end

function dbg_from_reads(ctx::Context, seq::Vector{UInt8}, offs::Vector{UInt64}, k::Integer, min_freq::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve seq offs begin
        check(ctx, ccall(fptr(:gg_dbg_from_reads), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, UInt64, Ref{Ptr{Cvoid}}),
                         ctx.h, seq, offs, length(offs) - 1, k, min_freq, out))
    end
    return out[]
end

"(node_offsets, bases_ascii, link_row_offsets, link_triples) of a gg_graph"
function graph_export(ctx::Context, g::Ptr{Cvoid})
    nn = Ref{UInt64}(0); nb = Ref{UInt64}(0); nl = Ref{UInt64}(0); k = Ref{Cint}(0)
    check(ctx, ccall(fptr(:gg_graph_sizes), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}), g, nn, nb, nl, k))
    node_off = Vector{UInt64}(undef, nn[] + 1)
    bases = Vector{UInt8}(undef, nb[])
    link_row = Vector{UInt64}(undef, nn[] + 1)
    tri = Vector{Int64}(undef, 3 * nl[])
    GC.@preserve node_off bases link_row tri begin
        check(ctx, ccall(fptr(:gg_graph_export), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}),
                         g, node_off, bases, link_row, tri))
    end
    return node_off, bases, link_row, tri
end

graph_free(g::Ptr{Cvoid}) = ccall(fptr(:gg_graph_free), Cvoid, (Ptr{Cvoid},), g)

# ---- rows next to the path: spectra-cn, 4-bit input, persistence, tips ------------------------
function counts_spectrum(ctx::Context, c::Ptr{Cvoid})
    h = zeros(UInt64, 256)
    GC.@preserve h check(ctx, ccall(fptr(:gg_counts_spectrum), Cint, (Ptr{Cvoid}, Ptr{UInt64}), c, h))
    return h
end
function counts_spectrum_cn(ctx::Context, a::Ptr{Cvoid}, b::Ptr{Cvoid})
    h = zeros(UInt64, 65536)
    GC.@preserve h check(ctx, ccall(fptr(:gg_counts_spectrum_cn), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}), a, b, h))
    return permutedims(reshape(h, 256, 256))   # [count in a + 1, count in b + 1]
end
"reads in the 4-bit encoding of LongSequence{DNAAlphabet{4}}: 16 bases per word, offsets in bases"
function dbg_from_reads_4bit(ctx::Context, data4::Vector{UInt64}, offs::Vector{UInt64}, k::Integer, min_freq::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve data4 offs begin
        check(ctx, ccall(fptr(:gg_dbg_from_reads_4bit), Cint,
                         (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Cint, UInt64, Ref{Ptr{Cvoid}}),
                         ctx.h, data4, offs, length(offs) - 1, k, min_freq, out))
    end
    return out[]
end
graph_write_gfa1(ctx::Context, g::Ptr{Cvoid}, filename::AbstractString) =
    check(ctx, ccall(fptr(:gg_graph_write_gfa1), Cint, (Ptr{Cvoid}, Cstring), g, filename))
function graph_load_gfa1(ctx::Context, gfafile::AbstractString, fafile::AbstractString)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(fptr(:gg_graph_load_gfa1), Cint, (Ptr{Cvoid}, Cstring, Cstring, Ref{Ptr{Cvoid}}), ctx.h, gfafile, fafile, out))
    return out[]
end
function graph_from_arrays(ctx::Context, k::Integer, node_off::Vector{UInt64}, bases::Vector{UInt8}, link_row::Vector{UInt64}, tri::Vector{Int64})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve node_off bases link_row tri begin
        check(ctx, ccall(fptr(:gg_graph_from_arrays), Cint,
                         (Ptr{Cvoid}, Cint, UInt64, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                         ctx.h, k, length(node_off) - 1, node_off, bases, link_row, tri, out))
    end
    return out[]
end
function graph_find_tip_nodes(ctx::Context, g::Ptr{Cvoid}, min_size::Integer)
    n = Ref{UInt64}(0)
    check(ctx, ccall(fptr(:gg_graph_find_tip_nodes), Cint, (Ptr{Cvoid}, UInt64, Ptr{Int64}, Ref{UInt64}), g, min_size, C_NULL, n))
    tips = Vector{Int64}(undef, n[])
    GC.@preserve tips check(ctx, ccall(fptr(:gg_graph_find_tip_nodes), Cint, (Ptr{Cvoid}, UInt64, Ptr{Int64}, Ref{UInt64}), g, min_size, tips, n))
    return tips
end

"remove_node! for every id (src/Graphs.jl:459-466): a new graph handle"
function graph_remove_nodes(ctx::Context, g::Ptr{Cvoid}, ids::Vector{Int64})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve ids check(ctx, ccall(fptr(:gg_graph_remove_nodes), Cint, (Ptr{Cvoid}, Ptr{Int64}, UInt64, Ref{Ptr{Cvoid}}), g, ids, length(ids), out))
    return out[]
end
"collapse_all_unitigs!(sg, min_nodes, true) (src/Graphs.jl:528-539): (new graph handle, number of new nodes)"
function graph_collapse_all_unitigs(ctx::Context, g::Ptr{Cvoid}, min_nodes::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL); m = Ref{UInt64}(0)
    check(ctx, ccall(fptr(:gg_graph_collapse_all_unitigs), Cint, (Ptr{Cvoid}, UInt64, Ref{Ptr{Cvoid}}, Ref{UInt64}), g, min_nodes, out, m))
    return out[], Int(m[])
end
"remove_tips!(sdg, min_size) (src/remove_tips.jl:2-30): (new graph handle, passes, tips removed)"
function graph_remove_tips(ctx::Context, g::Ptr{Cvoid}, min_size::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL); p = Ref{UInt64}(0); r = Ref{UInt64}(0)
    check(ctx, ccall(fptr(:gg_graph_remove_tips), Cint, (Ptr{Cvoid}, UInt64, Ref{Ptr{Cvoid}}, Ref{UInt64}, Ref{UInt64}), g, min_size, out, p, r))
    return out[], Int(p[]), Int(r[])
end

# ---- unique mer index -----------------------------------------------------------------------
function unique_mer_index(ctx::Context, bases::Vector{UInt8}, offs::Vector{UInt64}, k::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bases offs begin
        check(ctx, ccall(fptr(:gg_unique_mer_index), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, bases, offs, length(offs) - 1, k, out))
    end
    u = out[]
    nu = Ref{UInt64}(0); nn = Ref{UInt64}(0); kk = Ref{Cint}(0); w = Ref{Cint}(0)
    ccall(fptr(:gg_umi_sizes), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}, Ref{Cint}), u, nu, nn, kk, w)
    words = Vector{UInt64}(undef, nu[] * w[]); node = Vector{Int64}(undef, nu[]); pos = Vector{UInt32}(undef, nu[])
    upn = Vector{UInt64}(undef, nn[]); tpn = Vector{UInt64}(undef, nn[])
    GC.@preserve words node pos upn tpn begin
        check(ctx, ccall(fptr(:gg_umi_export), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int64}, Ptr{UInt32}, Ptr{UInt64}, Ptr{UInt64}),
                         u, words, node, pos, upn, tpn))
    end
    ccall(fptr(:gg_umi_free), Cvoid, (Ptr{Cvoid},), u)
    return words, node, pos, upn, tpn, Int(w[])
end

end # module LibGG

const GPUContext = LibGG.Context
