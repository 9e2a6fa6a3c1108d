/*
 * ggb200.h -- C ABI of libggb200.so: the B200-native (sm_100a) de Bruijn graph
 * construction path of GenomeGraphs.jl.
 *
 * The reference (BioJulia/GenomeGraphs.jl v0.2.0) has no FFI / plugin interface: its
 * boundary is a Julia call surface.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference repository); the Julia host
 * files under genomegraphs.jl_b200/julia/ bind these with `ccall`, the Python host
 * mirror (genomegraphs.jl_b200/host.py) with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *  - every function returns GG_OK (0) or a negative gg_status; nothing throws across the ABI;
 *    gg_last_error(ctx) gives the message of the last failure on that context.
 *  - opaque handles (gg_counts, gg_graph, gg_umi) own DEVICE memory; the caller allocates
 *    host export buffers after the matching *_sizes call (Julia: GC.@preserve).
 *  - one context = one GPU = one private CUDA stream; calls on a context are blocking and
 *    must be serialised by the caller (the reference is single-threaded as well).
 *  - there is NO CPU fallback: without a CUDA device gg_ctx_create fails with GG_ERR_CUDA.
 *
 * k-mer words: W = 1 (k <= 32), 2 (k <= 64), 4 (k <= 128) little-endian uint64 per k-mer,
 * MOST significant word first; 2 bits per base (A0 C1 G2 T3), first base in the most
 * significant of the 2k used bits, value right-aligned (BioSequences Mer / BigMer layout:
 * W=1 is the UInt64 of a Mer, W=2 the (hi, lo) halves of a BigMer's UInt128).
 */
#ifndef GGB200_H
#define GGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gg_ctx gg_ctx;
typedef struct gg_counts gg_counts; /* sorted unique k-mers + counts (Vector{MerCount{M}}) */
typedef struct gg_graph gg_graph;   /* SequenceDistanceGraph: unitig nodes + links */
typedef struct gg_umi gg_umi;       /* UniqueMerIndex */

typedef enum {
    GG_OK = 0,
    GG_ERR_ARG = -1,     /* bad argument (reference: ArgumentError, src/dbg.jl:262-264) */
    GG_ERR_CUDA = -2,    /* CUDA runtime failure / no device */
    GG_ERR_OOM = -3,     /* device or host allocation failed */
    GG_ERR_LIMIT = -4    /* size exceeds an implementation limit */
} gg_status;

#define GG_MAX_K 128

int gg_version(void);
/* words per k-mer for k (0 if k is outside 1..GG_MAX_K) */
int gg_kmer_words(int k);

/* ---- context ------------------------------------------------------------------------ */
int gg_ctx_create(int device, gg_ctx **out);
void gg_ctx_destroy(gg_ctx *ctx);
const char *gg_last_error(gg_ctx *ctx);
/* pinned host memory for fast host<->device copies of read buffers */
int gg_host_alloc(gg_ctx *ctx, uint64_t bytes, void **out);
void gg_host_free(gg_ctx *ctx, void *p);
/* plain device buffers (benchmarks keep inputs resident in HBM with these) */
int gg_dev_alloc(gg_ctx *ctx, uint64_t bytes, void **out);
void gg_dev_free(gg_ctx *ctx, void *p);
int gg_memcpy_h2d(gg_ctx *ctx, void *dst_dev, const void *src_host, uint64_t bytes);
int gg_memcpy_d2h(gg_ctx *ctx, void *dst_host, const void *src_dev, uint64_t bytes);
/* device time (ms, CUDA events on the context's stream) of the stages of the last pipeline
 * call: [0] pack, [1] extract, [2] sort, [3] count/RLE, [4] filter, [5] neighbour lookup,
 * [6] unitigs, [7] links, [8] total; and the number of kernel launches it issued. */
int gg_ctx_last_timings(gg_ctx *ctx, float ms[9], uint64_t *kernel_launches);
/* measurement helpers: CUDA events on the context's (launching) stream, slots 0..7 */
int gg_ctx_event_record(gg_ctx *ctx, int slot);
int gg_ctx_event_elapsed_ms(gg_ctx *ctx, int slot_a, int slot_b, float *ms);
int gg_ctx_sync(gg_ctx *ctx);
/* per-kernel profile of the next pipeline calls (events around every launch; slows the call
 * down, never enabled inside a timed region).  gg_ctx_kernel_profile returns
 * "kernel\tlaunches\ttotal_ms\talgorithmic_bytes\n" lines for the last pipeline call (bytes = sum over the
 * launches of the compulsory traffic of each launch's own element counts; 0 where no streaming model applies). */
int gg_ctx_set_profile(gg_ctx *ctx, int enabled);
const char *gg_ctx_kernel_profile(gg_ctx *ctx);

/* ---- reads -> k-mer counts ------------------------------------------------------------
 * Replaces `serial_mem(DNAMer{K}, CANONICAL)(reads)` of KmerAnalysis.jl (re-exported at
 * src/GenomeGraphs.jl:39-48; call site docs/src/man/assembly.md:126-128): every window of
 * k bases of every read (windows holding a non-ACGT symbol are skipped), canonicalised when
 * canonical != 0, sorted ascending, run-length collapsed.
 * seq_bytes: the reads' bases (ASCII, any case) concatenated; read i is
 * seq_bytes[read_offsets[i] .. read_offsets[i+1]) ; read_offsets has nreads + 1 entries.
 * gg_count_kmers takes HOST buffers (copies inside the call); gg_count_kmers_dev takes
 * DEVICE buffers of the same layout (nbytes = read_offsets[nreads]). */
int gg_count_kmers(gg_ctx *ctx, const uint8_t *seq_bytes, const uint64_t *read_offsets,
                   uint64_t nreads, int k, int canonical, gg_counts **out);
int gg_count_kmers_dev(gg_ctx *ctx, const uint8_t *d_seq_bytes, const uint64_t *d_read_offsets,
                       uint64_t nreads, uint64_t nbytes, int k, int canonical, gg_counts **out);
/* wrap an existing k-mer vector (host, any order, canonical and duplicate-free as
 * src/dbg.jl:261 requires); counts may be NULL (then all 1).  Sorted on the device when
 * needed (src/dbg.jl:98-100). */
int gg_counts_from_kmers(gg_ctx *ctx, const uint64_t *kmer_words, const uint32_t *counts,
                         uint64_t n, int k, gg_counts **out);
int gg_counts_sizes(gg_counts *c, uint64_t *n_unique, uint64_t *n_instances, int *k, int *words);
/* `filter!(x -> freq(x) >= min_freq, counts)` (src/dbg.jl:275): in place, order kept.
 * u8_saturated != 0 compares the MerCount UInt8 view min(count, 255). */
int gg_counts_filter(gg_counts *c, uint64_t min_freq, int u8_saturated);
/* any of the three output pointers may be NULL */
int gg_counts_export(gg_counts *c, uint64_t *kmer_words, uint8_t *counts_u8_saturated, uint32_t *counts_u32);
/* `spectra(counts)` of KmerAnalysis.jl (docs/src/man/assembly.md:159): 256-bin histogram of
 * min(count, 255) */
int gg_counts_spectrum(gg_counts *c, uint64_t hist[256]);
/* spectra-cn, `spectra(reads_kcount, graph_kcount)` (docs/src/man/assembly.md:221-228): joint histogram of the
 * UInt8-saturated counts of every k-mer of two tables (same context, same k); hist[a_count * 256 + b_count], a
 * k-mer missing from one table counts 0 there. */
int gg_counts_spectrum_cn(gg_counts *a, gg_counts *b, uint64_t hist[65536]);
void gg_counts_free(gg_counts *c);

/* staged entry points (tests / measurement) */
/* k-mers of all windows in read order (each(M, seq) + canonical, src/GraphIndexes.jl:78-82).
 * Call with out_words == NULL to get *n_kmers first. */
int gg_extract_kmers(gg_ctx *ctx, const uint8_t *seq_bytes, const uint64_t *read_offsets,
                     uint64_t nreads, int k, int canonical, uint64_t *out_words, uint64_t *n_kmers);

/* ---- k-mers -> SequenceDistanceGraph ---------------------------------------------------
 * Replaces `dbg!(sdg, kmers)` (src/dbg.jl:261-272: build_unitigs! :94-185 +
 * connect_unitigs_by_overlaps! :224-240) on the k-mers currently held by `counts`
 * (after gg_counts_filter this is `dbg!(sg, counted_kmers, min_freq)`, src/dbg.jl:274-278).
 * Node ids, node orientation (canonical!) and per-node link order equal the reference's. */
int gg_dbg_build(gg_ctx *ctx, gg_counts *counts, gg_graph **out);
/* reads -> graph in one call (count, filter, build); HOST / DEVICE input variants */
int gg_dbg_from_reads(gg_ctx *ctx, const uint8_t *seq_bytes, const uint64_t *read_offsets,
                      uint64_t nreads, int k, uint64_t min_freq, gg_graph **out);
int gg_dbg_from_reads_dev(gg_ctx *ctx, const uint8_t *d_seq_bytes, const uint64_t *d_read_offsets,
                          uint64_t nreads, uint64_t nbytes, int k, uint64_t min_freq, gg_graph **out);
/* 4-bit packed host input -- the encoding of the reference's read store (LongSequence{DNAAlphabet{4}},
 * src/GenomeGraphs.jl:79-83): 16 bases per uint64, base i of the concatenated reads in bits 4(i%16) .. +3 of word
 * i/16, A = 1, C = 2, G = 4, T = 8, anything else ambiguous (its windows are skipped).  read_offsets are in BASES.
 * Half the host-to-device bytes of the ASCII entry points, same results. */
int gg_count_kmers_4bit(gg_ctx *ctx, const uint64_t *data4, const uint64_t *read_offsets, uint64_t nreads, int k,
                        int canonical, gg_counts **out);
int gg_dbg_from_reads_4bit(gg_ctx *ctx, const uint64_t *data4, const uint64_t *read_offsets, uint64_t nreads, int k,
                           uint64_t min_freq, gg_graph **out);
int gg_graph_sizes(gg_graph *g, uint64_t *n_nodes, uint64_t *total_bases, uint64_t *n_link_entries, int *k);
/* node i (id i+1, src/Graphs.jl:418-422) is bases_ascii[node_offsets[i] .. node_offsets[i+1]);
 * links of node i (src/Graphs.jl:198,473-476) are the (source, destination, dist) int64
 * triples link_triples[3*link_row_offsets[i] .. 3*link_row_offsets[i+1]) in push order. */
int gg_graph_export(gg_graph *g, uint64_t *node_offsets, uint8_t *bases_ascii,
                    uint64_t *link_row_offsets, int64_t *link_triples);
/* small summary read back without exporting: sum over nodes of an order-independent hash */
int gg_graph_checksum(gg_graph *g, uint64_t *checksum);
/* find_tip_nodes(sg, min_size) (src/Graphs.jl:778-816) on the device-resident graph: ids of the tip nodes, ascending
 * (the reference returns a Set).  Call with tips == NULL to get *n_tips first; *n_tips = capacity on input. */
int gg_graph_find_tip_nodes(gg_graph *g, uint64_t min_size, int64_t *tips, uint64_t *n_tips);
/* Tip removal on the device-resident graph.  Each call returns a NEW handle and leaves `g` as it is; a removed /
 * collapsed node keeps its id and becomes empty (zero bases, no links: the reference's empty_node), new nodes are
 * appended.  Node ids, orientation (canonical) and the order of the link entries at every node equal the reference's.
 * Links must be symmetric as add_link! (src/Graphs.jl:473-476) makes them (GG_ERR_ARG otherwise).
 *   gg_graph_remove_nodes          remove_node! (src/Graphs.jl:459-466) for every id; +n and -n both name node n
 *   gg_graph_collapse_all_unitigs  collapse_all_unitigs!(sg, min_nodes, consume = true) (src/Graphs.jl:528-539:
 *                                  find_all_unitigs! :831-862, join_path! :1040-1088); min_nodes >= 2; an overlap
 *                                  (dist <= 0) that the sequences do not have is GG_ERR_ARG ("Sequences do not
 *                                  overlap.", src/Graphs.jl:1023)
 *   gg_graph_remove_tips           remove_tips!(sdg, min_size) (src/remove_tips.jl:2-30); *n_passes as in its log
 *                                  (1 = no tip found), *n_tips_removed over all passes */
int gg_graph_remove_nodes(gg_graph *g, const int64_t *ids, uint64_t n_ids, gg_graph **out);
int gg_graph_collapse_all_unitigs(gg_graph *g, uint64_t min_nodes, gg_graph **out, uint64_t *n_new_nodes);
int gg_graph_remove_tips(gg_graph *g, uint64_t min_size, gg_graph **out, uint64_t *n_passes, uint64_t *n_tips_removed);
/* persistence: write_to_gfa1 (src/Graphs.jl:892-926), <filename>.gfa + <filename>.fasta in the reference's format,
 * from a graph handle or straight from exported arrays (host only, no context needed); and a loader for that pair of
 * files (load_from_gfa1! is unfinished upstream, src/Graphs.jl:943-958): nodes by their seq<id> names, one add_link!
 * (src/Graphs.jl:473-476) per L line. */
int gg_graph_write_gfa1(gg_graph *g, const char *filename);
int gg_gfa1_write(const char *filename, uint64_t n_nodes, const uint64_t *node_offsets, const uint8_t *bases_ascii,
                  const uint64_t *link_row_offsets, const int64_t *link_triples);
int gg_graph_load_gfa1(gg_ctx *ctx, const char *gfa_file, const char *fasta_file, gg_graph **out);
int gg_graph_from_arrays(gg_ctx *ctx, int k, uint64_t n_nodes, const uint64_t *node_offsets, const uint8_t *bases_ascii,
                         const uint64_t *link_row_offsets, const int64_t *link_triples, gg_graph **out);
void gg_graph_free(gg_graph *g);

/* ---- UniqueMerIndex ---------------------------------------------------------------------
 * Replaces `UniqueMerIndex{M}(graph)` (src/GraphIndexes.jl:62-102, intended semantics: the
 * unique_mers_per_node vector has n_nodes entries).  Node sequences are passed as ASCII. */
int gg_unique_mer_index(gg_ctx *ctx, const uint8_t *node_bases_ascii, const uint64_t *node_offsets,
                        uint64_t n_nodes, int k, gg_umi **out);
int gg_umi_sizes(gg_umi *u, uint64_t *n_unique, uint64_t *n_nodes, int *k, int *words);
/* keys sorted ascending; node = +-id (sign = strand, src/GraphIndexes.jl:80), pos 1-based */
int gg_umi_export(gg_umi *u, uint64_t *kmer_words, int64_t *node, uint32_t *pos,
                  uint64_t *unique_mers_per_node, uint64_t *total_mers_per_node);
void gg_umi_free(gg_umi *u);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------
 * Replaces the role of KmerAnalysis `dist_mem` (re-export src/GenomeGraphs.jl:46-47, prose
 * docs/src/man/assembly.md:108-117: workers count read ranges, sorted count vectors are
 * merged).  Every rank passes ITS shard of the reads.  17 <= k <= 32: the super-k-mers are
 * partitioned by minimizer bin and stored into the owner GPU's buffer over NVLink (bulk
 * asynchronous copies; grouped ncclSend / ncclRecv without peer mapping), every rank counts
 * its bins -- in rounds over slices of the minimizer digit for jobs beyond ~7.7e9 instances --
 * and a second exchange range-partitions the kept k-mers.  Other k <= 128: W-word keys are
 * range-partitioned on the all-gathered level-0 histogram, the level-0 scatter writing
 * straight into the owners' buffers.  Either way rank r returns the r-th contiguous slice of
 * the global sorted table, so the concatenation over ranks equals the single-GPU
 * gg_count_kmers result.  A communicator spans at most 8 ranks (the GPUs of one box);
 * non-canonical counting at k = 32 is not covered (GG_ERR_LIMIT).  A failure on one rank is
 * agreed on before the next collective: every rank returns an error, none is left waiting.
 * Bootstrap: rank 0 calls gg_comm_unique_id and ships the bytes to the other ranks by any
 * host channel (the Python host uses torch.distributed, Julia would use Distributed / MPI);
 * every rank then calls gg_comm_create on its own context.  Collective calls must be made
 * by all ranks in the same order. */
typedef struct gg_comm gg_comm;
#define GG_COMM_ID_BYTES 128
int gg_comm_unique_id(uint8_t id[GG_COMM_ID_BYTES]);
int gg_comm_create(gg_ctx *ctx, const uint8_t id[GG_COMM_ID_BYTES], int rank, int world, gg_comm **out);
void gg_comm_destroy(gg_comm *comm);
/* owner ranges of the histogram buckets: rank q owns [first[q], first[q + 1]); first has
 * world + 1 entries.  Pure host function (no GPU needed). */
int gg_dist_splitters(const uint64_t *hist, uint32_t n_buckets, int world, uint32_t *first);
/* min_freq > 1 drops rarer k-mers on the owner (every k-mer has exactly one owner, so the
 * counts are complete).  *n_instances_global = k-mer windows over all ranks. */
int gg_dist_count_kmers_dev(gg_comm *comm, const uint8_t *d_seq_bytes, const uint64_t *d_read_offsets, uint64_t nreads,
                            uint64_t nbytes, int k, int canonical, uint64_t min_freq, gg_counts **out,
                            uint64_t *n_instances_global);
/* slices of all ranks -> the whole sorted table on every rank */
int gg_counts_allgather(gg_comm *comm, gg_counts *local, gg_counts **full);
/* reads -> graph: distributed count + filter, all-gather of the kept k-mers, unitig stage
 * sharded by vertex range (lookups, successor pointers, list ranking, emission on the owner
 * of each vertex; bitmaps / node table / packed bases summed over the ranks), link stage on
 * every rank.  Each rank returns the same, whole graph.  At most 2^30 kept k-mers. */
int gg_dist_dbg_from_reads_dev(gg_comm *comm, const uint8_t *d_seq_bytes, const uint64_t *d_read_offsets, uint64_t nreads,
                               uint64_t nbytes, int k, uint64_t min_freq, gg_graph **out, uint64_t *n_instances_global);

/* ---- deterministic synthetic reads -----------------------------------
 * Counter-based generator (splitmix64 keyed by seed and index): the host and device
 * variants produce identical bytes.  Genome of genome_len i.i.d. uniform bases; nreads
 * reads of read_len bases; paired != 0: reads 2j / 2j+1 are the forward start and the
 * reverse-complemented end of fragment j (FwRv) of frag_len bases; substitution errors with
 * probability err_ppm / 1e6 per base; n_ppm / 1e6 of the bases become 'N'.
 * Output: nreads * read_len ASCII bytes (read i at i * read_len). */
int gg_synth_reads_host(uint64_t seed, uint64_t genome_len, uint64_t nreads, uint32_t read_len,
                        int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *out);
int gg_synth_reads_dev(gg_ctx *ctx, uint64_t seed, uint64_t genome_len, uint64_t nreads, uint32_t read_len,
                       int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *d_out);
/* reads first_read .. first_read + nreads of the same job (a rank's shard); first_read must be
 * even when paired */
int gg_synth_reads_range_dev(gg_ctx *ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t nreads,
                             uint32_t read_len, int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm,
                             uint8_t *d_out);
int gg_synth_reads_range_host(uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t nreads, uint32_t read_len,
                              int paired, uint32_t frag_len, uint32_t err_ppm, uint32_t n_ppm, uint8_t *out);
int gg_synth_genome_host(uint64_t seed, uint64_t genome_len, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* GGB200_H */
