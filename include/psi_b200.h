/*
 * psi_b200.h -- C-ABI of libpsi_b200.so: the B200-native fully-sensitive seed
 * finding path of PSI (cartoonist/psi) behind plain pointers and sizes.
 *
 * The reference has no FFI layer: its boundary for this path is the C++
 * template class psi::SeedFinder as driven by find_seeds() of the CLI
 * (reference src/psikt.cpp:83-212).  Every entry point below names the
 * reference interface it stands behind (file:line relative to the reference
 * tree).  The C++17 mirror of that class lives in psi_b200/include/psi/ and
 * calls only these functions; INTEGRATION.md shows the binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every function returns PSI_B200_OK (0) or a negative error code; the
 *     message is available from psi_b200_last_error() (per context) or
 *     psi_b200_global_error() (for calls without a context);
 *   - no exceptions, no STL, no torch types cross this boundary;
 *   - host pointers are caller-owned and only read/written during the call;
 *   - one context per GPU, calls on one context are serialised by the caller;
 *   - all graph arrays are indexed by node RANK (0-based position in the
 *     reference graph's rank order, i.e. gum rank - 1);
 *   - there is NO CPU fallback: device entry points fail with
 *     PSI_B200_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PSI_B200_H
#define PSI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSI_B200_OK            0
#define PSI_B200_ERR_ARG      -1   /* invalid argument (reference: std::runtime_error) */
#define PSI_B200_ERR_CUDA     -2   /* CUDA runtime / no device */
#define PSI_B200_ERR_NOMEM    -3
#define PSI_B200_ERR_IO       -4   /* reference: `false` from load/save calls */
#define PSI_B200_ERR_STATE    -5   /* call order (e.g. seeds_all before submit_chunk) */
#define PSI_B200_ERR_OVERFLOW -6   /* device buffer exhausted after regrow attempts */

#define PSI_B200_MAX_SEED_LEN 32   /* a seed is one 2-bit packed 64-bit word */

/* ===================================================================== *
 *  Host side: graph, reads, genome-wide paths (no GPU needed)
 * ===================================================================== */

/* Flattened sequence graph; stands behind gum::SeqGraph<Succinct> as loaded by
 * gum::util::load(graph, file, sort=true) (src/psikt.cpp:249-251).  Node
 * ranks, internal ids and out-edge order reproduce gum's. */
typedef struct psi_b200_graph psi_b200_graph;

typedef struct {
  uint64_t        n_nodes;
  uint64_t        n_edges;
  uint64_t        n_bases;
  uint64_t        n_paths;     /* embedded (reference) paths */
  const uint64_t* seq_start;   /* n_nodes+1 */
  const char*     seq;         /* n_bases, ASCII */
  const uint64_t* row_ptr;     /* n_nodes+1 */
  const uint32_t* col;         /* n_edges successor ranks */
  const uint64_t* internal_id; /* n_nodes: gum Succinct id written by psikt (src/psikt.cpp:172-181) */
  const uint64_t* coord_id;    /* n_nodes: external id, graph.coordinate_id(id) */
} psi_b200_graph_view;

/* gum::util::load(graph, fname, sort) for GFA1 (S/L/P) and GFA2 (S/E/O)
 * (gum/gfa_utils.hpp:541-554). */
int  psi_b200_graph_load_gfa(const char* path, int sort, psi_b200_graph** out);
/* The same for vg's protobuf graph files (`.vg`: libvgio's BGZF / gzip stream of Graph chunks; gum/io_utils.hpp:77-80,
 * vg_utils.hpp:377-396): nodes, edges and embedded paths of all chunks in file order, mappings in rank order, then the
 * same node ordering as for GFA.  No protobuf library is involved. */
int  psi_b200_graph_load_vg(const char* path, int sort, psi_b200_graph** out);
/* gum::util::load(graph, fname, sort): by file name, `*.vg` is read as vg, anything else as GFA (gum/io_utils.hpp:66-80). */
int  psi_b200_graph_load(const char* path, int sort, psi_b200_graph** out);
/* Build from flat arrays (synthetic graphs).  ids[] are external ids; when
 * sort != 0 ranks are re-ordered exactly like the GFA loader would.  Embedded
 * paths: n_paths, path_ptr[n_paths+1], path_nodes = indices into the INPUT
 * node order. */
int  psi_b200_graph_from_arrays(uint64_t n_nodes, const uint64_t* ids,
                                const uint64_t* seq_start, const char* seq,
                                const uint64_t* row_ptr, const uint32_t* col,
                                uint64_t n_paths, const uint64_t* path_ptr,
                                const uint32_t* path_nodes, int sort,
                                psi_b200_graph** out);
void psi_b200_graph_free(psi_b200_graph* g);
int  psi_b200_graph_get_view(const psi_b200_graph* g, psi_b200_graph_view* view);
/* Embedded path i: name and node ranks (gum for_each_path / path(id)). */
int  psi_b200_graph_path(const psi_b200_graph* g, uint64_t i, const char** name,
                         const uint32_t** nodes, uint64_t* n_nodes);
/* Write the graph as GFA1 (used to hand synthetic graphs to the reference). */
int  psi_b200_graph_write_gfa(const psi_b200_graph* g, const char* path);

/* Genome-wide path selection; stands behind SeedFinder::pick_paths
 * (seed_finder.hpp:1138-1167) + Haplotyper / least_covered_adjacent
 * (graph_iter.hpp:537-1005, graph.hpp:216-287).  For every embedded path of the
 * graph, n walks are started at its first node and extended to a sink, always
 * taking a least-covered successor; ties are broken by a seeded xorshift (the
 * reference uses std::random_device, so its paths differ run to run; the seed
 * SET does not depend on the choice).  Fails with PSI_B200_ERR_ARG
 * when the graph embeds no path (seed_finder.hpp:1145-1147). */
typedef struct psi_b200_pathset psi_b200_pathset;
typedef struct {
  uint64_t        n_paths;
  const uint64_t* path_ptr;   /* n_paths+1 */
  const uint32_t* nodes;      /* node ranks */
  const uint32_t* head_off;   /* n_paths: bases trimmed from first node */
  const uint32_t* tail_trim;  /* n_paths: bases trimmed from last node */
} psi_b200_pathset_view;
int  psi_b200_pick_paths(const psi_b200_graph* g, unsigned n, int patched,
                         unsigned context, uint64_t seed, psi_b200_pathset** out);
/* The path set of a path index SAVED BY THE REFERENCE (psikt -I <prefix>): reads `<prefix>_paths` as written by
 * PathIndex::save_paths_set / PathSet::serialize / Path<graph, Compact>::serialize (pathindex.hpp:313-332,
 * pathset.hpp:261-273, path_base.hpp:552-560: coordinate node ids as sdsl enc_vector<elias_delta>, left, right, node
 * breaks).  Together with `<prefix>_loci_e<step>l<k>` (same format in both builds) a saved index is shared: the
 * device index is rebuilt from the paths.  PSI_B200_ERR_IO when the file is missing, malformed or belongs to another
 * graph.  *context receives the patch context stored in the file. */
int  psi_b200_pathset_load_reference(const psi_b200_graph* g, const char* paths_file, psi_b200_pathset** out, uint64_t* context);
void psi_b200_pathset_free(psi_b200_pathset* p);
int  psi_b200_pathset_get_view(const psi_b200_pathset* p, psi_b200_pathset_view* view);

/* Read chunking; stands behind readRecords(Records&, SeqStreamIn&, n)
 * (sequence.hpp:1608-1624) over kseq++: FASTQ/FASTA, plain or gzip. */
typedef struct psi_b200_reader psi_b200_reader;
typedef struct {
  uint64_t        n_reads;
  uint64_t        first_read_id; /* Records::rec_offset = records read before (sequence.hpp:1616) */
  const uint64_t* read_ptr;      /* n_reads+1 */
  const char*     bases;
  const uint64_t* name_ptr;      /* n_reads+1 */
  const char*     names;
} psi_b200_chunk_view;
int  psi_b200_reader_open(const char* path, psi_b200_reader** out);
/* Loads up to max_reads records (0 = all) into the reader's (pinned when a GPU
 * is present) chunk buffer.  view->n_reads == 0 at end of input.  The view stays valid until the call after the
 * NEXT one on this reader (two chunk buffers alternate, so that chunk i+1 can be parsed while chunk i is in flight). */
int  psi_b200_reader_next(psi_b200_reader* r, uint64_t max_reads, psi_b200_chunk_view* view);
void psi_b200_reader_close(psi_b200_reader* r);

/* A read chunk as 2-bit words: what Records<Dna5QStringSet> holds after readRecords (sequence.hpp:1608-1624), packed
 * the way the device consumes it -- 4x fewer bytes over PCIe than the characters.
 *   words   32 bases per 64-bit word, base i of the chunk at bits [2 (i % 32), +2) of word i / 32 (A, C, G, T = 0..3,
 *           either case; anything else is stored as 0 and listed in exc), reads back to back in read order;
 *           n_words = n_bases / 32 + 2 (the kernels read one word past the last base);
 *   exc     ascending base positions (within the chunk) of the characters outside A/C/G/T: a seed covering one
 *           never matches;
 *   read_len  != 0: every read has exactly this many bases and read_ptr may be NULL. */
typedef struct {
  uint64_t        n_reads;
  uint64_t        first_read_id;
  uint64_t        n_bases;
  uint32_t        read_len;
  uint32_t        reserved;
  const uint64_t* read_ptr;      /* n_reads+1 base offsets, or NULL when read_len != 0 */
  const uint64_t* words;
  const uint64_t* exc;
  uint64_t        n_exc;
  const uint64_t* name_ptr;      /* reader only; may be NULL */
  const char*     names;
} psi_b200_packed_chunk;
/* readRecords straight into 2-bit words (same record semantics as psi_b200_reader_next; same buffer lifetime). */
int  psi_b200_reader_next_packed(psi_b200_reader* r, uint64_t max_reads, psi_b200_packed_chunk* chunk);
/* Packs n_bases characters into words[n_bases / 32 + 2]; positions of characters outside A/C/G/T go to exc (at most
 * exc_cap of them are stored; *n_exc is their true number).  Host code, any thread. */
int  psi_b200_pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, uint64_t* exc, uint64_t exc_cap,
                         uint64_t* n_exc);

const char* psi_b200_global_error(void);

/* ===================================================================== *
 *  Device side: one context per GPU
 * ===================================================================== */

typedef struct psi_b200_ctx psi_b200_ctx;

/* SeedFinder(graph, seed_len, ...) (seed_finder.hpp:930-942); seed_len <= 32. */
int  psi_b200_create(int device, unsigned seed_len, psi_b200_ctx** out);
/* A further pipeline on the same GPU that SHARES the parent's resident graph,
 * path index and starting loci (one copy in HBM) and owns its stream, chunk
 * and result buffers.  Stands behind the reference's "one finder, several
 * threads, one chunk each" use (SeedFinderStats is keyed by thread id,
 * seed_finder.hpp:386-399,486-493); with two pipelines the upload of chunk i+1
 * overlaps the kernels and the download of chunk i.  The graph/paths/loci must
 * be final before the first fork; set_paths/find_loci/set_loci on a shared
 * index fail with PSI_B200_ERR_STATE.  Forks may outlive the parent. */
int  psi_b200_fork(psi_b200_ctx* parent, psi_b200_ctx** out);
void psi_b200_destroy(psi_b200_ctx* ctx);
const char* psi_b200_last_error(const psi_b200_ctx* ctx);
/* Run all work of this context on the given cudaStream_t (default: a private
 * non-blocking stream). */
int  psi_b200_set_stream(psi_b200_ctx* ctx, void* cuda_stream);
int  psi_b200_sync(psi_b200_ctx* ctx);
/* Options.  Set before set_paths:
 *   "gocc_threshold"    -r of the CLI / gocc_threshold of SeedFinder(graph, len, gocc_threshold, ...) (seed_finder.hpp:930-942):
 *                       seeds_on_paths skips every k-mer that occurs more often than this in the text of the indexed paths
 *                       (index_iter.hpp:842-848; 0 = no threshold).  seeds_off_paths is unaffected, so hits at starting loci
 *                       survive.  Needs the off-path walks in the index (PSI_B200_ERR_ARG from find_loci / set_loci otherwise).
 * Set before set_paths / find_loci / set_loci:
 *   "code_by_rank"      1: index entries carry (node rank, offset) even when (node id, offset) would fit (test hook);
 *   "offpath_mode"      0 auto (default), 1 walk the graph from the starting loci for every chunk
 *                       (the reference's scheme), 2 always materialise those walks into the index;
 *   "offpath_max_pairs" auto mode materialises when the walks number at most this (default 2^28);
 *   "build_group_windows" set_paths indexes the paths in groups of at most this many k-windows (32 B of scratch each;
 *                       default 0 = what the free device memory allows, at most 2^30) and merges the groups' distinct pairs;
 *   "build_slices"      set_paths builds the index in this many slices of the k-mer space (1, 4, 16, 64 or 256; 0 = auto:
 *                       one slice unless the graph may hold more than ~3 * 2^30 distinct (k-mer, locus) pairs).  A sliced
 *                       build has no 2^32 limit on the number of pairs, releases its sorted pairs once the starting loci
 *                       are set (find_loci / set_loci once; set_paths again before changing them) and does not serve a
 *                       gocc threshold;
 *   "index_slack"       extra doublings of the index's bucket count: fewer full buckets (slow-path probes) for twice
 *                       the memory each; -1 (default) = 1 for 16-byte slots (k > ~27) while the index stays small, else 0.
 * Set before create_distance_index:
 *   "dindex_mode"       0 auto (default), 1 keep no rows (every query enumerates), 2 always materialise the rows;
 *   "dindex_max_bytes"  auto mode materialises when the rows take at most this (0 = half of the free device memory);
 *   "dindex_list_cap"   (node, distance) states a warp holds in shared memory (power of two, 64..1024; default 256):
 *                       nodes / queries with more are served from a global scratch region (test hook).
 * Any time (they select among kernels that produce the same records):
 *   "fused"             1 (default): when the index answers the requested phases by itself, a chunk is ONE kernel
 *                       (seeding + probe + records); 0: separate seeding / probe / resolve kernels;
 *   "blocking_sync"     1: seeds_all / fetch wait on a blocking event (the host thread sleeps) instead of spinning in
 *                       cudaStreamSynchronize -- for more pipelines than host cores per GPU; default 0;
 *   "timers"            0: no CUDA-event records around the kernels of a step (default 1);
 *   "seeding_mode", "resolve_items", "resolve_ctas": variants of the separate kernels. */
int  psi_b200_set_option(psi_b200_ctx* ctx, const char* name, long long value);

/* The graph the finder borrows (seed_finder.hpp:1747), flattened: CSR
 * adjacency + concatenated labels (gum layout graph_traits_succinct.hpp:27-57,
 * accessors digraph_succinct.hpp:595-610,722-731, seqgraph_succinct.hpp:194-199).
 * node_id[r] is the id reported in seed records. */
int  psi_b200_set_graph(psi_b200_ctx* ctx, uint64_t n_nodes,
                        const uint64_t* seq_start, const char* seq,
                        const uint64_t* row_ptr, const uint32_t* col,
                        const uint64_t* node_id);

/* index_paths() -> PathIndex::create_index (seed_finder.hpp:1169-1176,
 * pathindex.hpp:235-243, fmindex.hpp:257-271): builds the GPU-resident path
 * index (distinct (k-mer, locus) pairs of all path windows, bucketised hash)
 * from the picked paths. */
int  psi_b200_set_paths(psi_b200_ctx* ctx, uint64_t n_paths,
                        const uint64_t* path_ptr, const uint32_t* path_nodes,
                        const uint32_t* head_off, const uint32_t* tail_trim);

/* add_uncovered_loci(step) (seed_finder.hpp:1481-1541): starting loci = graph
 * positions with at least one k-walk whose (k-mer, locus) is not in the path
 * index.  Computed on the device. */
int  psi_b200_find_loci(psi_b200_ctx* ctx, unsigned step, uint64_t* n_loci);
/* get_starting_loci() (seed_finder.hpp:957) / loci load (:1637-1679). */
int  psi_b200_get_loci(psi_b200_ctx* ctx, uint32_t* node_rank, uint32_t* offset,
                       uint64_t cap, uint64_t* n_loci);
int  psi_b200_set_loci(psi_b200_ctx* ctx, uint64_t n_loci,
                       const uint32_t* node_rank, const uint32_t* offset);

/* get_seeds(seeds, chunk, distance) + index_reads(seeds)
 * (seed_finder.hpp:1089-1109, sequence.hpp:1688-1745): uploads one read chunk,
 * packs its seeds (offsets 0,d,2d,.. while off+k <= len) into 2-bit k-mers and
 * prepares the device read index.  distance == 0 means seed_len
 * (src/psikt.cpp:469).  Asynchronous on the context's stream. */
int  psi_b200_submit_chunk(psi_b200_ctx* ctx, uint64_t n_reads,
                           const uint64_t* read_ptr, const char* bases,
                           uint64_t first_read_id, unsigned distance);
/* Same with read_ptr/bases already resident in device memory.  The kernels read the chunk in aligned 32-bit words:
 * d_bases must be readable up to the next multiple of 4 bytes past n_bases (true of any cudaMalloc'ed buffer). */
int  psi_b200_submit_chunk_device(psi_b200_ctx* ctx, uint64_t n_reads,
                                  const uint64_t* d_read_ptr, const char* d_bases,
                                  uint64_t n_bases, uint64_t first_read_id,
                                  unsigned distance);

/* The same for a chunk of 2-bit words (psi_b200_packed_chunk): host memory (on_device == 0; copied asynchronously, see
 * "buffer lifetime" below) or device memory (on_device != 0; words must be readable for n_bases / 32 + 2 words). */
int  psi_b200_submit_chunk_packed(psi_b200_ctx* ctx, const psi_b200_packed_chunk* chunk, unsigned distance, int on_device);
/* Buffer lifetime: submit_chunk / submit_chunk_packed queue asynchronous copies FROM the caller's host buffers on the
 * context's stream and return; the buffers must stay valid and unchanged until the chunk's seeds_all / wait has
 * returned (or psi_b200_sync).  psi_b200_reader_next* alternates two buffers for exactly this reason. */

#define PSI_B200_ON_PATHS   1u   /* seeds_on_paths  (seed_finder.hpp:1426-1457) */
#define PSI_B200_OFF_PATHS  2u   /* seeds_off_paths (seed_finder.hpp:1703-1722) */
#define PSI_B200_ALL        3u   /* seeds_all       (seed_finder.hpp:1724-1732) */
#define PSI_B200_SORTED     4u   /* additionally sort records canonically on the device */
#define PSI_B200_NO_RESOLVE 8u   /* keep compact device records only (benchmark of the probe alone) */
#define PSI_B200_COMPACT   16u   /* resolve into 4 x u32 records (psi_b200_fetch32): same fields, half the bytes over PCIe;
                                    PSI_B200_ERR_ARG when a node id or a read id of the chunk does not fit 32 bits */

#define PSI_B200_DENSE     32u   /* per-seed results instead of per-hit records (psi_b200_fetch_dense): a u32 node id
                                    and a u16 / u32 node offset per seed, in seed order; 6 (or 8) bytes per seed over
                                    PCIe, read id / offset implied.  Needs the off-path walks in the index (offpath_mode
                                    0 or 2) and ids below 2^32 - 1; excludes SORTED, NO_RESOLVE, COMPACT */

#define PSI_B200_DENSE5    64u   /* PSI_B200_DENSE in 5 bytes per seed: the locus as one number e = (node id << off_bits) | node
                                    offset, its low 32 bits in the first plane (u32) and, in a second plane of one byte per
                                    seed, bits 32..38 of e with the off-path flag in bit 7; no hit = 0xffffffff / 0xff.  For
                                    graphs whose e stays below 2^39 - 1 (psi_b200_dense5_layout tells, and gives off_bits);
                                    PSI_B200_ERR_ARG otherwise.  The extra list is the same as PSI_B200_DENSE's */

/* Finds the seeds of the submitted chunk.  The result is the SET of hits
 * (each (read, offset, node, offset) once), resident in device
 * memory; *n_hits is its size (synchronises the stream).  Record order: grouped by blocks of 256
 * consecutive reads (fused route) or 512 consecutive seeds (separate kernels), blocks in completion
 * order, the records of seeds with several loci appended at the end; PSI_B200_SORTED gives the exact
 * canonical order.  The seeding kernels run here, not in submit_chunk: device buffers handed to
 * psi_b200_submit_chunk_device must stay valid until the last seeds_all of the chunk has returned. */
int  psi_b200_seeds_all(psi_b200_ctx* ctx, unsigned flags, uint64_t* n_hits);

/* The same, split in two so that one host thread can keep several contexts (psi_b200_fork) busy: seeds_all_async
 * queues the step on the context's stream and returns; psi_b200_wait blocks until it is done, repeats it with larger
 * device buffers if one overflowed, and reports the counts.  Between the two only psi_b200_fetch_dense_async may be
 * called on the context.  Steps the fused kernel cannot serve (walk mode, SORTED, NO_RESOLVE) run to completion inside
 * seeds_all_async; wait then only reports. */
int  psi_b200_seeds_all_async(psi_b200_ctx* ctx, unsigned flags);
int  psi_b200_wait(psi_b200_ctx* ctx, uint64_t* n_hits);

/* Results of a PSI_B200_DENSE step: two planes in seed order.  Seed s of the chunk (seeds in read order; read r of a
 * chunk of equal-length reads owns seeds r * per_read .. with per_read = (len - k) / d + 1, offset (s % per_read) * d;
 * sequence.hpp:1712) has
 *     ids[s]  (u32)  node id, 0xffffffff = no hit;
 *     offs[s] (u16 when no node label is longer than 32 768 bases, else u32; psi_b200_dense_layout tells) node offset,
 *             top bit set when the hit was found off the indexed paths (seeds_off_paths).
 * (After a PSI_B200_DENSE5 step the second plane is the one-byte plane described at the flag: off_bytes = 1.)
 * `dense` receives ids[n_seeds] immediately followed by offs[n_seeds]: n_seeds * (4 + off_bytes) bytes; cap_seeds is
 * the number of seeds the buffer has room for.  Seeds whose k-mer occurs at several loci have their further hits in
 * `extra`: 4 x u32 {node_id, node_off, read_id, read_off | off-path << 31} each.  n_hits of the step = dense hits +
 * n_extra.  fetch_dense copies after a completed step; fetch_dense_async queues the copies behind a step in flight
 * (the buffers -- pinned, for the copy to be asynchronous -- are valid after psi_b200_wait; chunks with reads of
 * several lengths make it wait for the step's seed count first).  When more than cap_extra extra records exist the
 * first cap_extra are delivered and *n_extra tells (fetch again with more room). */
int  psi_b200_fetch_dense(psi_b200_ctx* ctx, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra,
                          uint64_t* n_seeds, uint64_t* n_extra);
int  psi_b200_fetch_dense_async(psi_b200_ctx* ctx, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra);
/* Width in bytes (2 or 4) of the node-offset plane for the graph of this context. */
int  psi_b200_dense_layout(psi_b200_ctx* ctx, unsigned* off_bytes);
/* PSI_B200_DENSE5: off_bits of the entries (bits of the longest label - 1, fixed when the index is built) and whether
 * the current index can deliver them. */
int  psi_b200_dense5_layout(psi_b200_ctx* ctx, unsigned* off_bits, int* available);
/* Counts of the last completed dense step. */
int  psi_b200_dense_counts(psi_b200_ctx* ctx, uint64_t* n_seeds, uint64_t* n_extra);

/* Copies the records of the last seeds_all to the host in the reference
 * CLI's byte layout (src/psikt.cpp:172-181, seed.hpp:32-46): per hit 4 x u64
 * {node_id, node_offset, read_id, read_offset}. */
int  psi_b200_fetch(psi_b200_ctx* ctx, uint64_t* hits, uint64_t cap, uint64_t* n_hits);
/* The same records after a seeds_all with PSI_B200_COMPACT: per hit 4 x u32 {node_id, node_offset, read_id,
 * read_offset} -- the fields of Seed<> (seed.hpp:32-46) a caller widens when it builds the callback argument or
 * writes the CLI's 4 x size_t (src/psikt.cpp:172-181).  PSI_B200_ERR_STATE when the last seeds_all was not compact
 * (and psi_b200_fetch fails likewise after a compact one). */
int  psi_b200_fetch32(psi_b200_ctx* ctx, uint32_t* hits, uint64_t cap, uint64_t* n_hits);
/* Per record of the last (unsorted) seeds_all: 1 = found on an indexed path (seeds_on_paths), 2 = found only by
 * an off-path walk (seeds_off_paths); lets a caller route hits to the two callbacks of
 * seeds_all(reads, index, traverser, callback1, callback2) (seed_finder.hpp:1734-1743). */
int  psi_b200_fetch_kinds(psi_b200_ctx* ctx, uint8_t* kinds, uint64_t cap, uint64_t* n_hits);
/* Device pointer to the same records (valid until the next seeds_all). */
int  psi_b200_fetch_device(psi_b200_ctx* ctx, const uint64_t** d_hits, uint64_t* n_hits);

/* ---- MEM mode: SeedFinder::seeds_on_paths(sequence, callback) -> find_mems (seed_finder.hpp:1459-1479,
 * index_iter.hpp:854-906) ----
 * build_mem_index: the index those queries need -- the sorted suffix table of the text of the given paths (the same
 * arrays as psi_b200_set_paths; stands behind PathIndex::create_index, pathindex.hpp:235-243).  Before the first fork.
 * find_mems: the reference's scan over every read of the submitted chunk (any submit_chunk* call): extend
 * read[start : start + len] while it occurs in the path text; once len >= seed length and it occurs at most
 * gocc_threshold times ("gocc_threshold" option, 0 = no limit), report all its occurrences and restart behind it; a
 * read stops after max_mem raw hits (0 = no limit).  The result is the SET of hits, 6 x u64 each:
 * {node_id, node_off, read_id, read_off, match_len, gocc}; gocc = occurrences in the path text, counted per path like
 * the reference's.  Reads longer than 65 535 bases are not supported. */
int  psi_b200_build_mem_index(psi_b200_ctx* ctx, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                              const uint32_t* head_off, const uint32_t* tail_trim);
int  psi_b200_find_mems(psi_b200_ctx* ctx, unsigned max_mem, uint64_t* n_hits);
int  psi_b200_fetch_mems(psi_b200_ctx* ctx, uint64_t* hits, uint64_t cap, uint64_t* n_hits);

/* ---- Paired-end distance verification: SeedFinder::create_distance_index / verify_distance (seed_finder.hpp:1193-1265,
 * 1300-1317; DiVerG's distance index, ext/diverg/include/diverg/dindex.hpp:767-914) ----
 * create_distance_index: prepares the queries for the window dmin <= l <= dmax (characters walked from the first locus
 * to the second).  As in the reference nothing is built when dmin == 0 or dmax < dmin, and verify_distance then fails
 * with PSI_B200_ERR_STATE.  The device keeps, per node, the (node, distance) pairs reachable inside the window (8 bytes
 * each; "dindex_max_bytes" option, default half of the free device memory) -- or nothing at all when they would not
 * fit ("dindex_mode" 1), in which case every query enumerates the walks from its first locus.  Before the first fork.
 * verify_distance: n queries of 4 x u32 {rank of v, offset in v, rank of u, offset in u}; ok[i] = 1 when the reference's
 * verify_distance(v, o, u, p) holds: v != u and some walk of dmin..dmax characters leads from (v, o) to (u, p), or v == u
 * and dmin <= p - o <= dmax (seed_finder.hpp:1306-1309).  Loci that do not exist answer 0.  on_device != 0: pairs and ok
 * are device pointers (pairs 16-byte aligned).  Synchronous. */
int  psi_b200_create_distance_index(psi_b200_ctx* ctx, unsigned dmin, unsigned dmax);
int  psi_b200_verify_distance(psi_b200_ctx* ctx, uint64_t n, const uint32_t* pairs, uint8_t* ok, int on_device);

/* Pinned host memory for chunk / result buffers. */
int  psi_b200_host_alloc(void** p, size_t bytes);
void psi_b200_host_free(void* p);

/* SeedFinderStats counters / timers (seed_finder.hpp:50-726) for the last chunk
 * and index build; times are CUDA-event milliseconds on the context's stream. */
typedef struct {
  uint64_t n_nodes, n_edges, n_bases;
  uint64_t n_path_bases;       /* indexed path text length */
  uint64_t n_index_entries;    /* distinct (k-mer, locus) pairs in the device index */
  uint64_t n_index_kmers;      /* distinct k-mers */
  uint64_t index_bytes;        /* device bytes of the index */
  uint64_t index_buckets;      /* 128-byte buckets (one DRAM line each) */
  uint32_t index_slot_bytes;   /* 8 or 16 */
  uint32_t index_stash_used;   /* keys that found MAX_DISP + 1 lines full */
  uint64_t n_offpath_entries;  /* (k-mer, locus) pairs of the materialised off-path walks (0 in walk mode) */
  uint64_t n_offpath_walks;    /* k-walks from the starting loci that are not on an indexed path */
  uint32_t offpath_mode;       /* 1 = walk the graph per chunk, 2 = walks materialised into the index */
  uint32_t fused;              /* 1 = the last seeds_all ran the fused one-pass kernel (seeding + probe + records) */
  uint64_t n_loci;
  uint64_t n_reads, n_seeds;   /* last chunk */
  uint64_t n_hits_on, n_hits_off, n_hits;
  uint64_t n_walks;            /* k-walks completed by the off-path kernel */
  uint64_t n_on_probe_sectors; /* seeds whose probe needed more than the home line (locus lists, displaced keys) */
  float ms_index_build, ms_find_loci;
  float ms_h2d, ms_pack, ms_read_index, ms_on, ms_off, ms_resolve, ms_sort, ms_d2h;
  uint32_t launches;           /* kernels of this library launched since create / reset */
  float ms_probe;              /* the seeds_on_paths probe kernel alone (ms_on also covers the slow-queue kernel) */
  /* sums over the fused steps completed since create / reset (CUDA events on the context's stream; "timers" option) */
  double ms_probe_sum, ms_on_sum;
  uint64_t timed_steps;
  uint64_t n_gocc_dropped;     /* on-path entries the gocc threshold removed from the index */
  uint32_t code_by_rank;       /* 1: index entries carry (node rank, offset), 0: (node id, offset) -- no gather per hit */
  uint32_t code_off_bits;
  uint64_t n_dindex_entries;   /* (node, distance) pairs of the distance index (counted even when they are not kept) */
  uint64_t dindex_bytes;       /* device bytes of the materialised rows */
  uint32_t dindex_mode;        /* 0 none, 1 queries enumerate, 2 rows materialised */
  float ms_dindex_build;
  uint32_t index_build_slices; /* slices of the k-mer space the index was built in (1: one-shot build) */
  uint32_t reserved0;
} psi_b200_counters_t;
int  psi_b200_counters(psi_b200_ctx* ctx, psi_b200_counters_t* out);
int  psi_b200_reset_counters(psi_b200_ctx* ctx);

/* Library build info: "psi_b200 <version> sm_100a". */
const char* psi_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSI_B200_H */
