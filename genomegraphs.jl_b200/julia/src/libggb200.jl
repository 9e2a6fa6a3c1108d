# Thin `ccall` layer over libggb200.so (include/ggb200.h).  One GPUContext = one GPU.
module LibGG

using Libdl

const LIB = Ref{String}(get(ENV, "GGB200_LIB", joinpath(@__DIR__, "..", "..", "libggb200.so")))

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
        rc = ccall((:gg_ctx_create, LIB[]), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
        rc == GG_OK || throw(GGError(rc, "no usable CUDA device $device (libggb200 has no CPU fallback)"))
        ctx = new(out[])
        finalizer(c -> (c.h == C_NULL || ccall((:gg_ctx_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        return ctx
    end
end

const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT[] === nothing && (DEFAULT[] = Context(0)); DEFAULT[]::Context)

function check(ctx::Context, rc::Cint)
    rc == GG_OK && return nothing
    msg = unsafe_string(ccall((:gg_last_error, LIB[]), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == GG_ERR_ARG && throw(ArgumentError(msg))   # reference: ArgumentError (src/dbg.jl:262-264)
    throw(GGError(rc, msg))
end

kmer_words(k::Integer) = Int(ccall((:gg_kmer_words, LIB[]), Cint, (Cint,), k))

# ---- counts ---------------------------------------------------------------------------------
function count_kmers(ctx::Context, seq::Vector{UInt8}, offs::Vector{UInt64}, k::Integer, canonical::Bool)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve seq offs begin
        check(ctx, ccall((:gg_count_kmers, LIB[]), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, seq, offs, length(offs) - 1, k, canonical, out))
    end
    return out[]
end

function counts_from_kmers(ctx::Context, words::Vector{UInt64}, n::Integer, k::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve words begin
        check(ctx, ccall((:gg_counts_from_kmers, LIB[]), Cint,
                         (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt32}, UInt64, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, words, C_NULL, n, k, out))
    end
    return out[]
end

function counts_sizes(c::Ptr{Cvoid})
    nu = Ref{UInt64}(0); ni = Ref{UInt64}(0); k = Ref{Cint}(0); w = Ref{Cint}(0)
    ccall((:gg_counts_sizes, LIB[]), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}, Ref{Cint}), c, nu, ni, k, w)
    return Int(nu[]), Int(ni[]), Int(k[]), Int(w[])
end

counts_filter!(ctx::Context, c::Ptr{Cvoid}, min_freq::Integer, u8sat::Bool) =
    check(ctx, ccall((:gg_counts_filter, LIB[]), Cint, (Ptr{Cvoid}, UInt64, Cint), c, min_freq, u8sat))

function counts_export(ctx::Context, c::Ptr{Cvoid})
    nu, _, _, w = counts_sizes(c)
    words = Vector{UInt64}(undef, nu * w)
    c8 = Vector{UInt8}(undef, nu)
    GC.@preserve words c8 begin
        check(ctx, ccall((:gg_counts_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt32}), c, words, c8, C_NULL))
    end
    return words, c8, w
end

counts_free(c::Ptr{Cvoid}) = ccall((:gg_counts_free, LIB[]), Cvoid, (Ptr{Cvoid},), c)

# ---- graph ----------------------------------------------------------------------------------
function dbg_build(ctx::Context, c::Ptr{Cvoid})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:gg_dbg_build, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, c, out))
    return out[]
end

function dbg_from_reads(ctx::Context, seq::Vector{UInt8}, offs::Vector{UInt64}, k::Integer, min_freq::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve seq offs begin
        check(ctx, ccall((:gg_dbg_from_reads, LIB[]), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, UInt64, Ref{Ptr{Cvoid}}),
                         ctx.h, seq, offs, length(offs) - 1, k, min_freq, out))
    end
    return out[]
end

"(node_offsets, bases_ascii, link_row_offsets, link_triples) of a gg_graph"
function graph_export(ctx::Context, g::Ptr{Cvoid})
    nn = Ref{UInt64}(0); nb = Ref{UInt64}(0); nl = Ref{UInt64}(0); k = Ref{Cint}(0)
    check(ctx, ccall((:gg_graph_sizes, LIB[]), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}), g, nn, nb, nl, k))
    node_off = Vector{UInt64}(undef, nn[] + 1)
    bases = Vector{UInt8}(undef, nb[])
    link_row = Vector{UInt64}(undef, nn[] + 1)
    tri = Vector{Int64}(undef, 3 * nl[])
    GC.@preserve node_off bases link_row tri begin
        check(ctx, ccall((:gg_graph_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}),
                         g, node_off, bases, link_row, tri))
    end
    return node_off, bases, link_row, tri
end

graph_free(g::Ptr{Cvoid}) = ccall((:gg_graph_free, LIB[]), Cvoid, (Ptr{Cvoid},), g)

# ---- unique mer index -----------------------------------------------------------------------
function unique_mer_index(ctx::Context, bases::Vector{UInt8}, offs::Vector{UInt64}, k::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bases offs begin
        check(ctx, ccall((:gg_unique_mer_index, LIB[]), Cint,
                         (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, UInt64, Cint, Ref{Ptr{Cvoid}}),
                         ctx.h, bases, offs, length(offs) - 1, k, out))
    end
    u = out[]
    nu = Ref{UInt64}(0); nn = Ref{UInt64}(0); kk = Ref{Cint}(0); w = Ref{Cint}(0)
    ccall((:gg_umi_sizes, LIB[]), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{Cint}, Ref{Cint}), u, nu, nn, kk, w)
    words = Vector{UInt64}(undef, nu[] * w[]); node = Vector{Int64}(undef, nu[]); pos = Vector{UInt32}(undef, nu[])
    upn = Vector{UInt64}(undef, nn[]); tpn = Vector{UInt64}(undef, nn[])
    GC.@preserve words node pos upn tpn begin
        check(ctx, ccall((:gg_umi_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int64}, Ptr{UInt32}, Ptr{UInt64}, Ptr{UInt64}),
                         u, words, node, pos, upn, tpn))
    end
    ccall((:gg_umi_free, LIB[]), Cvoid, (Ptr{Cvoid},), u)
    return words, node, pos, upn, tpn, Int(w[])
end

end # module LibGG

const GPUContext = LibGG.Context
