# `gpu_mem(M, CANONICAL)`: drop-in for KmerAnalysis `serial_mem(M, CANONICAL)` (reference
# call site docs/src/man/assembly.md:126-128) that counts on the GPU.  The counter accepts
# anything that iterates read sequences (a ReadDatastore buffer, a Vector of LongSequence, ...);
# `collect` turns the device table into the reference's `Vector{MerCount{M}}`.

mutable struct DeviceMerCounts{M<:AbstractMer}
    ctx::LibGG.Context
    h::Ptr{Cvoid}
    k::Int
    function DeviceMerCounts{M}(ctx, h, k) where {M}
        c = new{M}(ctx, h, k)
        finalizer(x -> (x.h == C_NULL || LibGG.counts_free(x.h); x.h = C_NULL), c)
        return c
    end
end
Base.length(c::DeviceMerCounts) = LibGG.counts_sizes(c.h)[1]

# reads -> concatenated ASCII bytes + offsets (the layout gg_count_kmers takes)
function concat_reads(reads)
    buf = IOBuffer()
    offs = UInt64[0]
    for r in reads
        print(buf, r)                 # LongSequence prints its bases; N and IUPAC codes are skipped on the device
        push!(offs, UInt64(position(buf)))
    end
    return take!(buf), offs
end

struct GPUCounter{M<:AbstractMer,C}
    ctx::LibGG.Context
end
gpu_mem(::Type{M}, mode::C; ctx::LibGG.Context = LibGG.default_context()) where {M<:AbstractMer,C} = GPUCounter{M,C}(ctx)

function (gc::GPUCounter{M,C})(reads) where {M,C}
    seq, offs = concat_reads(reads)
    k = BioSequences.ksize(M)
    h = LibGG.count_kmers(gc.ctx, seq, offs, k, C === typeof(CANONICAL))
    return DeviceMerCounts{M}(gc.ctx, h, k)
end

"Device table -> `Vector{MerCount{M}}` (sorted by k-mer, UInt8-saturated counts)."
function Base.collect(c::DeviceMerCounts{M}) where {M<:AbstractMer}
    words, c8, w = LibGG.counts_export(c.ctx, c.h)
    return [MerCount{M}(mer_from_words(M, words, i, w), c8[i]) for i in eachindex(c8)]
end

"`spectra(counts)` (docs/src/man/assembly.md:159) and spectra-cn `spectra(reads_kcount, graph_kcount)` (:221-228) on the device."
KmerAnalysis.spectra(c::DeviceMerCounts) = LibGG.counts_spectrum(c.ctx, c.h)
KmerAnalysis.spectra(a::DeviceMerCounts, b::DeviceMerCounts) = LibGG.counts_spectrum_cn(a.ctx, a.h, b.h)
