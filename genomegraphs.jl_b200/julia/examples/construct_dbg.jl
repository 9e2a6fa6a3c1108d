### The reference's documented recipe (docs/src/man/assembly.md:77-216) with the counting and
### the graph build on the GPU.  Not executed in the build environment (no Julia there).
using GenomeGraphs, BioSequences, FASTX

reads = LongSequence{DNAAlphabet{4}}[]
for f in ARGS
    open(FASTQ.Reader, f) do r
        for record in r
            push!(reads, FASTQ.sequence(LongSequence{DNAAlphabet{4}}, record))
        end
    end
end

counts = gpu_mem(DNAMer{31}, CANONICAL)(reads)          # was: serial_mem(DNAMer{31}, CANONICAL)(ds)
graph = dbg!(empty_graph(LongSequence{DNAAlphabet{4}}), counts, 2)
println(graph)
