#!/usr/bin/env julia
# make_ref_golden.jl -- pins oracle/ (and the CUDA path) to the REAL reference.
#
# Test infrastructure.  This image has no Julia, so the committed vectors under tests/golden/ come from
# tests/pyref.py and from hand traces (tests/golden/hand_traced.py); DESIGN.md therefore says "parity
# unpinned".  On any machine with Julia >= 1.0 this script recomputes every expected value of those
# fixtures with the reference package itself and reports each difference:
#
#   julia --project=/path/to/env oracle/make_ref_golden.jl /path/to/GenomeGraphs.jl [--write]
#
# The environment needs GenomeGraphs' own dependencies at its compat bounds (Project.toml:12-17:
# BioSequences 2, KmerAnalysis 0.4.4, ReadDatastores 0.4, FASTX 1.1) plus JSON.  The reference is
# `Pkg.develop`ed from the given path, so the code that runs is /root/reference's src/ unmodified.
#
# What is called (reference sites):
#   graph cases : GenomeGraphs.dbg!(empty_graph(LongSequence{DNAAlphabet{4}}), kmers)      src/dbg.jl:261-272
#   read cases  : serial_mem(M, CANONICAL)(reads) (docs/src/man/assembly.md:126-128), then
#                 dbg!(sg, counted_kmers, min_freq) src/dbg.jl:274-278, then
#                 GraphIndexes.UniqueMerIndex{M}(sg) src/GraphIndexes.jl:62-102
#   tip cases   : Graphs.find_tip_nodes(sg, min_size) src/Graphs.jl:778-816, Graphs.collapse_all_unitigs!(unitigs, newnodes,
#                 sg, 2, true) src/Graphs.jl:528-539 and remove_tips!(sg, min_size) src/remove_tips.jl:2-30 on the graphs of
#                 tests/golden/tips_golden.json (rebuilt node by node, link vectors filled in their stored order)
#                 -> tests/golden/tips_golden_ref.json (--write replaces tips_golden.json)
# Output: tests/golden/dbg_golden_ref.json (same schema as dbg_golden.json, "generator" names this
# script and the package versions) and a line per case: OK / DIFF <field>.  With --write the committed
# dbg_golden.json is replaced, after which `pytest tests/test_golden.py` checks the oracle and the CUDA
# path against reference-made vectors and the "unpinned" caveat can be dropped.
#
# KmerAnalysis' counters take a ReadDatastore or any iterable of sequences [upstream, recalled]; the
# reads of a case are handed over as a Vector{LongSequence{DNAAlphabet{4}}} (the alphabet the
# reference's read_datastore builds, src/GenomeGraphs.jl:79-83).  Reads with lower-case letters are
# upper-cased by the LongSequence constructor, N stays N (4-bit alphabet).

import Pkg
using JSON

length(ARGS) >= 1 || error("usage: make_ref_golden.jl /path/to/GenomeGraphs.jl [--write]")
const REF = abspath(ARGS[1])
const WRITE = "--write" in ARGS
Pkg.develop(Pkg.PackageSpec(path = REF))
using BioSequences, KmerAnalysis, GenomeGraphs
const GG = GenomeGraphs

const HERE = dirname(abspath(@__FILE__))
const GOLD = joinpath(dirname(HERE), "tests", "golden")

mertype(k::Int) = k <= 32 ? DNAMer{k} : BigDNAMer{k}   # src/dbg.jl:187-188: Mer / BigMer are what the path supports

function graph_arrays(sg)
    nodes = String[string(GG.Graphs.sequence(sg, i)) for i in GG.Graphs.each_node_id(sg)]
    links = [[[GG.Graphs.source(l), GG.Graphs.destination(l), GG.Graphs.distance(l)] for l in GG.Graphs.linksof(sg, i)]
             for i in GG.Graphs.each_node_id(sg)]
    return nodes, links
end

function run_graph_case(c)
    k = c["k"]
    k > 64 && return nothing          # the reference stops at BigMer (k <= 64); longer keys are this repo's generalisation
    M = mertype(k)
    kmers = M[M(LongSequence{DNAAlphabet{2}}(s)) for s in c["kmers"]]
    sg = empty_graph(LongSequence{DNAAlphabet{4}})
    dbg!(sg, kmers)
    nodes, links = graph_arrays(sg)
    return Dict("name" => c["name"], "k" => k, "kmers" => c["kmers"], "nodes" => nodes, "links" => links)
end

function run_reads_case(c)
    k = c["k"]
    k > 64 && return nothing
    M = mertype(k)
    reads = [LongSequence{DNAAlphabet{4}}(r) for r in c["reads"]]
    counted = serial_mem(M, CANONICAL)(reads)
    counts = [[string(mer(x)), Int(freq(x))] for x in counted]
    sg = empty_graph(LongSequence{DNAAlphabet{4}})
    dbg!(sg, counted, c["min_freq"])
    nodes, links = graph_arrays(sg)
    idx = GG.GraphIndexes.UniqueMerIndex{M}(sg)
    uniq = sort!([[string(m), Int(p.node), Int(p.position)] for (m, p) in idx.mer_to_graphposition], by = x -> x[1])
    return Dict("name" => c["name"], "k" => k, "min_freq" => c["min_freq"], "reads" => c["reads"], "counts" => counts,
                "nodes" => nodes, "links" => links,
                "umi" => Dict("unique" => uniq, "unique_per_node" => Int.(idx.unique_mers_per_node),
                              "total_per_node" => Int.(idx.total_mers_per_node)))
end

# ---- tip removal ------------------------------------------------------------------------------------------------
const SEQ4 = LongSequence{DNAAlphabet{4}}
function graph_from_case(nodes, links)
    sg = empty_graph(SEQ4)
    for s in nodes
        GG.Graphs.add_node!(sg, SEQ4(s))
    end
    for (i, s) in enumerate(nodes)     # an empty sequence is a removed node: the reference's empty_node (deleted = true)
        isempty(s) && (GG.Graphs.nodes(sg)[i] = GG.Graphs.empty_node(SEQ4))
    end
    for (i, ls) in enumerate(links), l in ls
        push!(GG.Graphs.linksof(sg, i), GG.Graphs.SequenceDistanceGraphLink(l[1], l[2], l[3]))
    end
    return sg
end
function edited_arrays(sg)   # deleted nodes print as "" (their sequence is empty)
    nodes = String[string(GG.Graphs.node(sg, i).seq) for i in GG.Graphs.each_node_id(sg)]
    links = [[[GG.Graphs.source(l), GG.Graphs.destination(l), GG.Graphs.distance(l)] for l in GG.Graphs.linksof(sg, i)]
             for i in GG.Graphs.each_node_id(sg)]
    return nodes, links
end
function run_tips_case(c)
    r = Dict{String,Any}("name" => c["name"], "min_size" => c["min_size"], "nodes" => c["nodes"], "links" => c["links"])
    sg = graph_from_case(c["nodes"], c["links"])
    r["tips"] = sort!(collect(GG.Graphs.find_tip_nodes(sg, c["min_size"])))
    try
        utgs = Vector{GG.Graphs.SequenceDistanceGraphPath{typeof(sg)}}()
        newnodes = Vector{GG.Graphs.NodeID}()
        GG.Graphs.collapse_all_unitigs!(utgs, newnodes, sg, 2, true)
        nodes, links = edited_arrays(sg)
        r["collapse"] = Dict("new_nodes" => newnodes, "nodes" => nodes, "links" => links)
    catch e
        r["collapse"] = Dict("error" => sprint(showerror, e))
    end
    sg = graph_from_case(c["nodes"], c["links"])
    try
        remove_tips!(sg, c["min_size"])
        nodes, links = edited_arrays(sg)
        # passes / removed are only logged by the reference (src/remove_tips.jl:11-27): keep the committed numbers
        r["remove_tips"] = Dict("passes" => c["remove_tips"]["passes"], "removed" => c["remove_tips"]["removed"], "nodes" => nodes, "links" => links)
    catch e
        r["remove_tips"] = Dict("error" => sprint(showerror, e))
    end
    return r
end

function diff_fields(a, b, fields)
    bad = String[]
    for f in fields
        JSON.json(a[f]) == JSON.json(b[f]) || push!(bad, f)
    end
    return bad
end

function main()
    committed = JSON.parsefile(joinpath(GOLD, "dbg_golden.json"))
    out = Dict{String,Any}("generator" => string("oracle/make_ref_golden.jl: GenomeGraphs ", Pkg.project().dependencies["GenomeGraphs"],
                                                  " from ", REF, ", Julia ", VERSION), "graphs" => Any[], "reads" => Any[])
    ndiff = 0
    for c in committed["graphs"]
        r = run_graph_case(c)
        if r === nothing
            push!(out["graphs"], c); println("SKIP  ", c["name"], " (k > 64)"); continue
        end
        bad = diff_fields(c, r, ["nodes", "links"])
        println(isempty(bad) ? "OK    " : "DIFF  ", c["name"], isempty(bad) ? "" : string("  ", bad))
        ndiff += !isempty(bad)
        push!(out["graphs"], r)
    end
    # the hand-traced vectors (tests/golden/hand_traced.py) are part of dbg_golden.json's graph list under their G* names
    for c in committed["reads"]
        r = run_reads_case(c)
        if r === nothing
            push!(out["reads"], c); println("SKIP  ", c["name"], " (k > 64)"); continue
        end
        bad = diff_fields(c, r, ["counts", "nodes", "links"])
        ubad = diff_fields(c["umi"], r["umi"], ["unique", "unique_per_node", "total_per_node"])
        append!(bad, ["umi." * f for f in ubad])
        println(isempty(bad) ? "OK    " : "DIFF  ", c["name"], isempty(bad) ? "" : string("  ", bad))
        ndiff += !isempty(bad)
        push!(out["reads"], r)
    end
    open(joinpath(GOLD, "dbg_golden_ref.json"), "w") do io
        JSON.print(io, out)
    end
    WRITE && cp(joinpath(GOLD, "dbg_golden_ref.json"), joinpath(GOLD, "dbg_golden.json"); force = true)
    tips = JSON.parsefile(joinpath(GOLD, "tips_golden.json"))
    tout = Dict{String,Any}("generator" => out["generator"], "cases" => Any[])
    for c in tips["cases"]
        r = run_tips_case(c)
        bad = diff_fields(c, r, ["tips"])
        for part in ("collapse", "remove_tips")
            if haskey(c[part], "error") || haskey(r[part], "error")
                haskey(c[part], "error") == haskey(r[part], "error") || push!(bad, part * ".error")
            else
                append!(bad, [part * "." * f for f in diff_fields(c[part], r[part], ["nodes", "links"])])
            end
        end
        println(isempty(bad) ? "OK    " : "DIFF  ", c["name"], isempty(bad) ? "" : string("  ", bad))
        ndiff += !isempty(bad)
        push!(tout["cases"], r)
    end
    open(joinpath(GOLD, "tips_golden_ref.json"), "w") do io
        JSON.print(io, tout)
    end
    WRITE && cp(joinpath(GOLD, "tips_golden_ref.json"), joinpath(GOLD, "tips_golden.json"); force = true)
    println(ndiff == 0 ? "all cases equal the committed vectors: the oracle is pinned to the reference" :
            string(ndiff, " case(s) differ from the committed vectors"))
    exit(ndiff == 0 ? 0 : 1)
end

main()
