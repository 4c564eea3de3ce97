// context.hpp -- per-GPU state of libpsi_b200 (host-side view).
#ifndef PSI_B200_DEVICE_CONTEXT_HPP
#define PSI_B200_DEVICE_CONTEXT_HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../../include/psi_b200.h"
#include "common.cuh"

namespace psi_b200 {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StateError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct OverflowError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define PSI_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      throw ::psi_b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// Grow-only device buffer.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  // Ensure room for n elements; contents are NOT preserved on growth.
  void ensure(size_t n, double slack = 1.0)
  {
    if (n <= cap) return;
    size_t want = (size_t)((double)n * slack) + 16;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    PSI_CUDA(cudaMalloc((void**)&p, want * sizeof(T)));
    cap = want;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// Node record used by the graph walkers: one 16-byte load per visited node.
struct alignas(16) NodeRec {
  uint32_t seq_start;  // global position of the first base
  uint32_t seq_len;
  uint32_t edge_start; // into col[]
  uint32_t outdeg;
};

// Position -> node in two 16-byte gathers: rank16[pos >> 6] holds the node-start bits of 64 positions and the
// number of node starts before them; node_res[rank] holds what a record needs from the node.
struct alignas(16) Rank16 {
  uint64_t bits;     // bit i: a node starts at position 64 * w + i
  uint32_t prefix;   // node starts before position 64 * w
  uint32_t pad;
};
struct alignas(16) NodeRes {
  uint32_t seq_start;
  uint32_t pad;
  uint64_t node_id;
};

// A compact hit: (seed index within the chunk, global graph position).
struct alignas(8) Hit {
  uint32_t seed;
  uint32_t gpos;
};

// A seed the fused kernel's one-line probe could not settle (fused.cu): everything the slow kernel needs.
struct alignas(16) SlowItem {
  uint64_t kmer;
  uint32_t read;   // read index within the chunk
  uint32_t off;    // offset of the seed in the read
  uint32_t seed;   // index of the seed within the chunk (dense results)
  uint32_t pad;
};

struct HostTable {
  dev::KmerTable view{};
  DevBuf<char> slots;
  DevBuf<dev::Slot16> stash;
  DevBuf<uint32_t> stash_used;
  uint64_t n_lines = 0;
};

// Everything that is immutable once built -- the flattened graph, the path
// index and the starting loci -- lives in one object that forked contexts share
// (psi_b200_fork): one resident copy per GPU, any number of chunk pipelines.
struct Shared {
  // ---- graph ----
  bool has_graph = false;
  uint32_t n_nodes = 0, n_edges = 0;
  uint64_t n_bases = 0;
  bool graph_has_n = false;
  DevBuf<NodeRec> node_rec;      // n_nodes + 1 (sentinel with seq_start = n_bases)
  DevBuf<uint32_t> col;
  DevBuf<uint64_t> seq2;         // 2-bit labels
  DevBuf<uint32_t> nmask;        // 1 bit per base: not A/C/G/T
  DevBuf<uint64_t> node_id;
  uint64_t max_node_id = 0;      // PSI_B200_COMPACT records need it below 2^32
  uint32_t max_node_len = 0;
  // locus codes carried by single-locus index entries (common.cuh): (id or rank) << code_off_bits | offset
  uint32_t code_off_bits = 0;
  bool code_by_rank = false;
  DevBuf<uint32_t> pos2node;     // node rank containing position (i << POS2NODE_SHIFT)
  // rank16 and node_res live back to back in one allocation (gathered by locus-list entries and walker hits only:
  // single-locus index entries carry their locus code)
  DevBuf<char> gather_pool;
  Rank16* rank16 = nullptr;      // unused when the graph has zero-length nodes (then pos2node is used)
  NodeRes* node_res = nullptr;
  size_t gather_pool_bytes = 0;
  bool has_rank16 = false;

  // ---- index: (k-mer -> loci) for every k-window of the indexed paths and, when the
  // off-path walks have been materialised, for every k-walk from a starting locus ----
  bool has_index = false;        // paths have been indexed
  bool has_table = false;        // `index` is allocated and probeable
  bool offpath_indexed = false;  // the k-walks from the starting loci are IN the table (no per-chunk walk needed)
  HostTable index;
  DevBuf<uint32_t> multi;        // [n_on, n_total, on-path loci..., off-path loci...] per k-mer with several loci
  DevBuf<uint64_t> on_kmer;      // distinct on-path (k-mer, locus) pairs, sorted: kept to re-merge when the loci change
  DevBuf<uint32_t> on_gpos;
  uint64_t n_on_pairs = 0, n_off_pairs = 0;
  // Sliced build (graphs whose distinct pairs outgrow 32-bit counts, index_build.cu): the distinct on-path pairs are
  // kept per slice of the k-mer space (a k-mer belongs to the slice its first slice_bits / 2 bases spell), each slice
  // sorted by (k-mer, locus); on_kmer / on_gpos stay empty.  The slices are released once the final table stands.
  struct PairSlice {
    DevBuf<uint64_t> kmer;
    DevBuf<uint32_t> gpos;
    uint64_t n = 0, n_kmers = 0;
  };
  std::vector<std::unique_ptr<PairSlice>> slices;
  uint32_t slice_bits = 0;
  bool pairs_released = false;

  // ---- gocc threshold (-r): the picked paths stay on the device so that the k-mers' occurrence counts can be taken
  // when the final table is built ----
  uint32_t gocc_threshold = 0;
  bool paths_kept = false, table_filtered = false;
  DevBuf<uint64_t> path_ptr;
  DevBuf<uint32_t> path_nodes, path_head, path_tail;
  uint64_t n_paths = 0, n_path_entries = 0;

  // ---- MEM mode (mems.cu): the suffix table of the path text ----
  bool has_mem_index = false;
  uint64_t mem_n = 0;              // text positions in the table
  uint32_t mem_paths_n = 0;
  DevBuf<uint64_t> mem_path_ptr, mem_entry_start, mem_key;
  DevBuf<uint32_t> mem_nodes, mem_head, mem_tail, mem_gpos, mem_ent, mem_pstart;
  DevBuf<uint8_t> mem_vlen;

  // ---- paired-end distance index (distance.cu): per node, the (node, distance) pairs inside the window ----
  bool has_dindex = false;         // create_distance_index was given a constructible window
  bool dist_rows = false;          // the rows are materialised (else queries enumerate)
  uint32_t dist_dmin = 0, dist_dmax = 0;
  DevBuf<uint32_t> dist_row_len;                              // build time only
  DevBuf<uint4> dist_packed;                                  // per node {row start lo, hi, row length, label length}
  DevBuf<unsigned long long> dist_row_start, dist_entries;   // entry = node rank << 32 | distance between first characters
  uint64_t n_dist_entries = 0;

  // ---- starting loci ----
  uint64_t n_loci = 0;
  DevBuf<uint32_t> loci_node, loci_off;
};

struct Ctx {
  int device = 0;
  unsigned k = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::string error;
  psi_b200_counters_t counters{};
  cudaEvent_t ev[24]{};
  cudaEvent_t ev_sync = nullptr;   // cudaEventBlockingSync: lets a host thread sleep while it waits for its chunk
  int ev_state[12]{};   // 0 never recorded, 1 started, 2 start+stop recorded

  std::shared_ptr<Shared> sh = std::make_shared<Shared>();
  static constexpr uint32_t POS2NODE_SHIFT = 6;

  // ---- current chunk ----
  bool has_chunk = false;
  bool chunk_seeded = false;     // seed_kmer / seed_valid / seed_read / seed_first hold the chunk's seeds (engine_seed_chunk)
  bool chunk_indexed = false;
  uint64_t n_reads = 0, n_read_bases = 0, first_read_id = 0, n_seeds_cap = 0;
  unsigned distance = 0;
  DevBuf<char> bases;            // owned copy for host submissions
  DevBuf<uint64_t> read_ptr;
  const char* d_bases = nullptr; // points to `bases` or to caller's device memory
  const uint64_t* d_read_ptr = nullptr;   // null: every read has `read_len` bases (packed chunks only)
  // a chunk submitted as 2-bit words (psi_b200_submit_chunk_packed): d_words replaces d_bases
  bool chunk_packed = false;
  uint32_t read_len = 0;         // != 0: all reads of the chunk have this length
  DevBuf<uint64_t> words;        // owned copy for host submissions
  DevBuf<uint64_t> exc;
  const uint64_t* d_words = nullptr;
  const uint64_t* d_exc = nullptr;   // sorted base positions (within the chunk) of the characters outside A/C/G/T
  uint64_t n_exc = 0;
  DevBuf<uint64_t> reads2;       // the chunk's bases, 2 bits each
  DevBuf<uint32_t> reads_n;      // 1 bit per base: not A/C/G/T
  DevBuf<uint32_t> cta_first;    // per CTA of the seeding kernels: first seed, then the raw counts
  DevBuf<uint32_t> seed_first;   // n_reads + 1: first seed of each read (exclusive scan)
  DevBuf<uint32_t> seed_read;    // per seed: local read index
  DevBuf<uint64_t> seed_kmer;
  DevBuf<uint8_t> seed_valid;    // per seed: 1 = only A/C/G/T inside
  DevBuf<uint32_t> seed_next;    // chain links of the read index
  HostTable read_index;
  DevBuf<uint32_t> filter_bits;  // walk mode: bitmap of the seeds' prefixes (see ReadIndexSink)
  uint32_t filter_pfx = 0;       // bases of a prefix (min(k, FILTER_PFX))
  DevBuf<char> scan_tmp;
  uint64_t* h_pinned = nullptr;  // small pinned scratch for async counter read-back

  // ---- results ----
  DevBuf<uint64_t> seed_hit;     // per seed: locus code found by the probe
  DevBuf<uint8_t> seed_kind;     // per seed: 0 none, 1 on an indexed path, 2 off-path, 3 queued for the slow kernel
  DevBuf<uint32_t> slow_queue;   // seeds the one-line probe could not settle
  DevBuf<SlowItem> slow_items;   // the same for the fused kernel
  DevBuf<Hit> hits;              // overflow hit list: locus lists, walker hits
  DevBuf<uint8_t> hit_kind;
  DevBuf<uint64_t> sorted_code;  // compacted hits (PSI_B200_SORTED / NO_RESOLVE): locus code and seed index
  DevBuf<uint32_t> sorted_seed;
  DevBuf<uint8_t> rec_kind;      // per record: 1 on-path, 2 off-path
  bool kinds_valid = false;
  DevBuf<unsigned long long> dedup;  // off-path (chain head, gpos) set
  DevBuf<uint64_t> records;      // 4 x u64 per hit, reference layout
  DevBuf<unsigned long long> dev_counters;  // see DC_* below
  uint64_t n_hits = 0;
  bool records_valid = false;
  bool records_compact = false;  // records hold 4 x u32 per hit (PSI_B200_COMPACT)
  // dense results (PSI_B200_DENSE): `records` holds two planes in seed order -- node ids (u32, NIL32 = no hit) and node
  // offsets with the off-path flag in the top bit (u16 or u32); further hits of seeds with several loci are 4 x u32
  // records in `extra`
  bool records_dense = false;
  uint32_t dense_off_bytes = 4;      // width of the second plane: 2 when no node is longer than 32 768 bases, 1 for DENSE5
  uint64_t dense_off_plane = 0;      // byte offset of the node-offset plane inside `records`
  DevBuf<uint32_t> extra;
  uint64_t n_dense_seeds = 0, n_extra = 0;
  // ---- MEM mode results ----
  DevBuf<uint4> mem_raw;             // raw hits of find_mems_kernel (16 bytes each)
  DevBuf<uint64_t> mem_records;      // 6 x u64 per hit: node_id, node_off, read_id, read_off, match_len, gocc
  bool mem_valid = false;
  uint64_t n_mems = 0, n_mems_raw = 0;
  // ---- distance queries (distance.cu) ----
  DevBuf<uint32_t> dist_q, dist_queue;
  DevBuf<uint8_t> dist_out;
  DevBuf<unsigned long long> dist_counters, dist_big;
  // ---- a step in flight (psi_b200_seeds_all_async .. psi_b200_wait) ----
  bool pending = false;
  int pending_out_kind = 0;               // 0 = 4 x u64 records, 1 = 4 x u32 records, 2 = dense, 3 = dense, 5 bytes per seed
  unsigned pending_probe_mode = 0;
  uint64_t pending_out_cap_want = 0;      // records: capacity learnt from an overflowed step
  int pending_attempts = 0;
  void* pending_dense_dst = nullptr;      // host destinations of the asynchronous dense fetch, if one was queued
  void* pending_extra_dst = nullptr;
  uint64_t pending_dense_cap = 0, pending_extra_cap = 0, pending_extra_copied = 0;
  uint32_t spill_items = 4096;   // per-warp global spill of the walker
  DevBuf<char> walk_spill;

  // ---- options (psi_b200_set_option) ----
  int opt_build_slices = 0;                    // set_paths: slices of the k-mer space the index is built in (0 auto, else a power of 4 up to 256)
  uint64_t opt_build_group_windows = 0;        // set_paths: path windows materialised at a time (0: from the free memory, <= 2^30)
  int opt_index_slack = -1;                    // extra doublings of the path index's bucket count (-1 auto: 1 for 16-byte slots)
  int opt_blocking_sync = 0;                   // 1: wait for a chunk on a blocking event (thread sleeps) instead of spinning
  uint32_t opt_gocc_threshold = 0;             // seeds_on_paths skips k-mers with more occurrences in the path text (0: none)
  int opt_code_by_rank = 0;                    // 1: locus codes carry the node rank even when the ids would fit (tests)
  int opt_timers = 1;                          // 0: no CUDA-event records around the kernels of a step
  int opt_fused = 1;                           // 1: index-mode steps run the fused one-pass kernel (fused.cu)
  int opt_seeding_mode = 0;                     // 0 seeds straight from the ASCII chunk, 1 via a 2-bit copy of the reads
  int opt_resolve_items = 2;                   // items per thread of the resolve kernel (2 or 4)
  int opt_resolve_ctas = 6;                    // resident CTAs per SM the resolve kernel is compiled for (5 or 6)
  int opt_dindex_mode = 0;                     // distance index: 0 auto, 1 never materialise the rows (queries enumerate), 2 always
  uint32_t opt_dindex_list_cap = 256;          // (node, distance) states a warp keeps in shared memory before the global scratch serves it (test hook)
  uint64_t opt_dindex_max_bytes = 0;           // auto: materialise when the rows take at most this (0: half of the free memory)
  int opt_offpath_mode = 0;                    // 0 auto, 1 walk per chunk, 2 always materialise
  uint64_t opt_offpath_max_pairs = 1ull << 28; // auto: materialise when the k-walks number at most this
};

// indices into Ctx::dev_counters
enum {
  DC_HITS = 0,        // records written by the compaction
  DC_WALKS = 1,       // completed k-walks (off-path kernel)
  DC_ERR = 2,         // bit 0: walker stack overflow, bit 1: table overflow, bit 2: dedup overflow
  DC_HITS_ON = 3,     // records that are on-path hits
  DC_WORK = 4,        // dynamic work counter of the walkers
  DC_SEEDS = 5,       // total seeds of the chunk
  DC_AUX = 6,
  DC_AUX2 = 7,
  DC_OVF = 8,         // entries of the overflow hit list
  DC_SLOW = 9,        // seeds queued for seeds_slow_kernel
  DC_EXTRA = 10,      // dense results: records in the extra list
  DC_COUNT = 12
};

}  // namespace psi_b200
#endif
