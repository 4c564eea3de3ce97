// engine.hpp -- host entry points of the device engine (one per C-ABI call).
#ifndef PSI_B200_DEVICE_ENGINE_HPP
#define PSI_B200_DEVICE_ENGINE_HPP

#include "context.hpp"

#include <nvtx3/nvToolsExt.h>

namespace psi_b200 {

Ctx* engine_create(int device, unsigned seed_len);
Ctx* engine_fork(Ctx& parent);
void engine_destroy(Ctx* ctx);

void engine_set_graph(Ctx& c, uint64_t n_nodes, const uint64_t* seq_start, const char* seq,
                      const uint64_t* row_ptr, const uint32_t* col, const uint64_t* node_id);
void engine_build_index(Ctx& c, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                        const uint32_t* head_off, const uint32_t* tail_trim);
void engine_find_loci(Ctx& c, unsigned step);
void engine_set_loci(Ctx& c, uint64_t n, const uint32_t* node, const uint32_t* off);
void engine_get_loci(Ctx& c, uint32_t* node, uint32_t* off, uint64_t cap);
void engine_submit_chunk(Ctx& c, uint64_t n_reads, const uint64_t* read_ptr, const char* bases,
                         uint64_t n_bases, uint64_t first_read_id, unsigned distance, bool on_device);
void engine_seed_chunk(Ctx& c);   // the seeding kernels of the separate-kernel route (idempotent per chunk)
void engine_seeds(Ctx& c, unsigned flags);
// out_kind: 0 = 4 x u64 records, 1 = 4 x u32 records, 2 = dense per-seed results, 3 = dense results of 5 bytes (DENSE5)
void engine_seeds_fused(Ctx& c, unsigned probe_mode, int out_kind);
void engine_seeds_fused_async(Ctx& c, unsigned probe_mode, int out_kind);
void engine_seeds_async(Ctx& c, unsigned flags);   // queues the step when the fused route serves it, else runs it synchronously
void engine_wait(Ctx& c);
void engine_fetch_dense(Ctx& c, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra);
void engine_fetch_dense_async(Ctx& c, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra);
void engine_build_mem_index(Ctx& c, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                            const uint32_t* head_off, const uint32_t* tail_trim);
void engine_find_mems(Ctx& c, unsigned max_mem);
void engine_fetch_mems(Ctx& c, uint64_t* hits, uint64_t cap);
void engine_create_distance_index(Ctx& c, unsigned dmin, unsigned dmax);
void engine_verify_distance(Ctx& c, uint64_t n, const uint32_t* pairs, uint8_t* out, bool on_device);
void engine_submit_chunk_packed(Ctx& c, const psi_b200_packed_chunk& chunk, unsigned distance, bool on_device);
void engine_set_option(Ctx& c, const char* name, long long value);
void engine_fetch(Ctx& c, void* hits, uint64_t cap, bool compact);
void engine_fetch_kinds(Ctx& c, uint8_t* kinds, uint64_t cap);

// ---- helpers shared by the .cu files ----

// Size a KmerTable for n_keys distinct k-mers of 2k = kbits bits; allocates and
// clears the slots.  slack_bits: extra doublings of the line count (-1: auto).
// min_payload_bits: what a slot's payload must be able to hold (8-byte slots offer 32 + 27 - rem_bits, 16-byte slots 62).
void table_alloc(Ctx& c, HostTable& t, uint64_t n_keys, uint32_t kbits, uint64_t stash_slots, int slack_bits = 0,
                 uint32_t min_payload_bits = 32);
// payload bits 8-byte slots would offer for this many keys (0: 8-byte slots are not possible)
uint32_t table_fmt8_payload_bits(uint64_t n_keys, uint32_t kbits, int slack_bits);
void table_clear(Ctx& c, HostTable& t);

// CUDA-event timer slots (Ctx::ev holds a start/stop pair per slot)
enum { T_INDEX = 0, T_LOCI = 1, T_H2D = 2, T_PACK = 3, T_READ_INDEX = 4, T_ON = 5, T_OFF = 6, T_RESOLVE = 7,
       T_SORT = 8, T_D2H = 9, T_USER = 10, T_PROBE = 11, T_COUNT = 12 };

// NVTX range names: the reference's timer names (seed_finder.hpp:427-456) where a phase has one
inline const char* phase_name(int slot)
{
  static const char* const names[T_COUNT] = { "index-paths", "find-uncovered", "seeding (upload)", "seeding", "index-reads", "seeds-on-paths",
                                              "seeds-off-path", "seeds-on-paths (records)", "sort-seeds", "fetch-seeds", "user", "seeds-on-paths (probe)" };
  return slot >= 0 && slot < T_COUNT ? names[slot] : "psi_b200";
}

// A phase of a step: a pair of CUDA events on the context's stream (device time, read by psi_b200_counters) and an
// NVTX range on the host thread (what nsys / ncu --nvtx group by).
struct PhaseTimer {
  Ctx& c;
  int i;
  bool open = true;
  PhaseTimer(Ctx& ctx, int slot) : c(ctx), i(slot)
  {
    nvtxRangePushA(phase_name(slot));
    cudaEventRecord(c.ev[2 * i], c.stream);
    c.ev_state[i] = 1;
  }
  PhaseTimer(const PhaseTimer&) = delete;
  PhaseTimer& operator=(const PhaseTimer&) = delete;
  ~PhaseTimer() { if (open) nvtxRangePop(); }
  void stop()
  {
    cudaEventRecord(c.ev[2 * i + 1], c.stream);
    c.ev_state[i] = 2;
    if (open) { nvtxRangePop(); open = false; }
  }
  // valid after the stream has been synchronised
  float ms() const { return timer_ms(c, i); }
  static float timer_ms(Ctx& c, int i)
  {
    float v = 0;
    if (c.ev_state[i] != 2) return 0.f;
    if (cudaEventElapsedTime(&v, c.ev[2 * i], c.ev[2 * i + 1]) != cudaSuccess) { (void)cudaGetLastError(); return 0.f; }
    return v;
  }
};

// Wait until everything queued on the context's stream is done.  Spinning (cudaStreamSynchronize) reacts fastest; with
// more pipelines than host cores per GPU the spinners starve each other, so "blocking_sync" lets the thread sleep.
inline void ctx_wait(Ctx& c)
{
  if (c.opt_blocking_sync && c.ev_sync) {
    PSI_CUDA(cudaEventRecord(c.ev_sync, c.stream));
    PSI_CUDA(cudaEventSynchronize(c.ev_sync));
  }
  else PSI_CUDA(cudaStreamSynchronize(c.stream));
}

// PSI_B200_DENSE5: the entry (node id << code_off_bits | offset) must stay below the 39-bit "no hit" pattern
inline bool dense5_available(const Shared& sh)
{
  if (!sh.has_table || sh.code_off_bits >= 32) return false;
  const unsigned __int128 top = ((unsigned __int128)sh.max_node_id + 1) << sh.code_off_bits;
  return top <= (unsigned __int128)0x7fffffffffull;
}

inline unsigned grid_for(uint64_t items, unsigned block, unsigned items_per_thread = 1)
{
  uint64_t per_block = (uint64_t)block * items_per_thread;
  uint64_t g = (items + per_block - 1) / per_block;
  if (g == 0) g = 1;
  if (g > 0x7fffffffull) g = 0x7fffffffull;
  return (unsigned)g;
}

}  // namespace psi_b200
#endif
