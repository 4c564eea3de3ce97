// capi.cu -- the device half of the extern "C" surface of libpsi_b200.so (include/psi_b200.h); the host half
// (graphs, paths, reads) is capi_host.cpp.  Every entry point converts C++ exceptions into error codes; nothing else
// happens here.
#include <cstring>
#include <exception>
#include <new>
#include <string>

#include "../../include/psi_b200.h"
#include "capi_common.hpp"
#include "device/engine.hpp"

using namespace psi_b200;

struct psi_b200_ctx { Ctx* c; };

#undef HOST_GUARD
#define g_error (::psi_b200::capi_global_error())

namespace {

int translate(std::string& err)
{
  try { throw; }
  catch (const ArgError& e) { err = e.what(); return PSI_B200_ERR_ARG; }
  catch (const StateError& e) { err = e.what(); return PSI_B200_ERR_STATE; }
  catch (const CudaError& e) { err = e.what(); return PSI_B200_ERR_CUDA; }
  catch (const OverflowError& e) { err = e.what(); return PSI_B200_ERR_OVERFLOW; }
  catch (const std::bad_alloc&) { err = "out of host memory"; return PSI_B200_ERR_NOMEM; }
  catch (const std::exception& e) { err = e.what(); return PSI_B200_ERR_ARG; }
  catch (...) { err = "unknown error"; return PSI_B200_ERR_ARG; }
}

#define CTX_GUARD(ctx, body)                                        \
  if (!(ctx) || !(ctx)->c) { g_error = "null context"; return PSI_B200_ERR_ARG; } \
  try { body; return PSI_B200_OK; }                                 \
  catch (...) { return translate((ctx)->c->error); }

}  // namespace

extern "C" {

const char* psi_b200_version(void) { return "psi_b200 0.2.0 sm_100a"; }

/* ----------------------------------------------------------- device -- */

int psi_b200_create(int device, unsigned seed_len, psi_b200_ctx** out)
{
  if (!out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  try {
    Ctx* c = engine_create(device, seed_len);
    *out = new psi_b200_ctx{ c };
    return PSI_B200_OK;
  }
  catch (...) { return translate(g_error); }
}

int psi_b200_fork(psi_b200_ctx* parent, psi_b200_ctx** out)
{
  if (!out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  CTX_GUARD(parent, {
    Ctx* c = engine_fork(*parent->c);
    *out = new psi_b200_ctx{ c };
  })
}

void psi_b200_destroy(psi_b200_ctx* ctx)
{
  if (!ctx) return;
  engine_destroy(ctx->c);
  delete ctx;
}

const char* psi_b200_last_error(const psi_b200_ctx* ctx)
{
  if (!ctx || !ctx->c) return g_error.c_str();
  return ctx->c->error.c_str();
}

int psi_b200_set_stream(psi_b200_ctx* ctx, void* cuda_stream)
{
  CTX_GUARD(ctx, {
    Ctx& c = *ctx->c;
    PSI_CUDA(cudaSetDevice(c.device));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (c.own_stream) { cudaStreamDestroy(c.stream); c.own_stream = false; }
    c.stream = (cudaStream_t)cuda_stream;
  })
}

int psi_b200_set_option(psi_b200_ctx* ctx, const char* name, long long value)
{
  CTX_GUARD(ctx, engine_set_option(*ctx->c, name, value))
}

int psi_b200_sync(psi_b200_ctx* ctx)
{
  CTX_GUARD(ctx, {
    PSI_CUDA(cudaSetDevice(ctx->c->device));
    PSI_CUDA(cudaStreamSynchronize(ctx->c->stream));
  })
}

int psi_b200_set_graph(psi_b200_ctx* ctx, uint64_t n_nodes, const uint64_t* seq_start, const char* seq,
                       const uint64_t* row_ptr, const uint32_t* col, const uint64_t* node_id)
{
  CTX_GUARD(ctx, engine_set_graph(*ctx->c, n_nodes, seq_start, seq, row_ptr, col, node_id))
}

int psi_b200_set_paths(psi_b200_ctx* ctx, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                       const uint32_t* head_off, const uint32_t* tail_trim)
{
  CTX_GUARD(ctx, engine_build_index(*ctx->c, n_paths, path_ptr, path_nodes, head_off, tail_trim))
}

int psi_b200_find_loci(psi_b200_ctx* ctx, unsigned step, uint64_t* n_loci)
{
  CTX_GUARD(ctx, {
    engine_find_loci(*ctx->c, step);
    if (n_loci) *n_loci = ctx->c->sh->n_loci;
  })
}

int psi_b200_get_loci(psi_b200_ctx* ctx, uint32_t* node_rank, uint32_t* offset, uint64_t cap, uint64_t* n_loci)
{
  CTX_GUARD(ctx, {
    if (n_loci) *n_loci = ctx->c->sh->n_loci;
    if (cap) engine_get_loci(*ctx->c, node_rank, offset, cap);
  })
}

int psi_b200_set_loci(psi_b200_ctx* ctx, uint64_t n_loci, const uint32_t* node_rank, const uint32_t* offset)
{
  CTX_GUARD(ctx, engine_set_loci(*ctx->c, n_loci, node_rank, offset))
}

int psi_b200_submit_chunk(psi_b200_ctx* ctx, uint64_t n_reads, const uint64_t* read_ptr, const char* bases,
                          uint64_t first_read_id, unsigned distance)
{
  CTX_GUARD(ctx, engine_submit_chunk(*ctx->c, n_reads, read_ptr, bases, 0, first_read_id, distance, false))
}

int psi_b200_submit_chunk_device(psi_b200_ctx* ctx, uint64_t n_reads, const uint64_t* d_read_ptr, const char* d_bases,
                                 uint64_t n_bases, uint64_t first_read_id, unsigned distance)
{
  CTX_GUARD(ctx, engine_submit_chunk(*ctx->c, n_reads, d_read_ptr, d_bases, n_bases, first_read_id, distance, true))
}

int psi_b200_submit_chunk_packed(psi_b200_ctx* ctx, const psi_b200_packed_chunk* chunk, unsigned distance, int on_device)
{
  CTX_GUARD(ctx, {
    if (!chunk) throw ArgError("submit_chunk_packed: null chunk");
    engine_submit_chunk_packed(*ctx->c, *chunk, distance, on_device != 0);
  })
}

int psi_b200_seeds_all_async(psi_b200_ctx* ctx, unsigned flags)
{
  CTX_GUARD(ctx, {
    if (!(flags & PSI_B200_ALL)) throw ArgError("seeds_all: neither ON_PATHS nor OFF_PATHS requested");
    engine_seeds_async(*ctx->c, flags);
  })
}

int psi_b200_wait(psi_b200_ctx* ctx, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    engine_wait(*ctx->c);
    if (n_hits) *n_hits = ctx->c->n_hits;
  })
}

int psi_b200_fetch_dense(psi_b200_ctx* ctx, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra,
                         uint64_t* n_seeds, uint64_t* n_extra)
{
  CTX_GUARD(ctx, {
    engine_fetch_dense(*ctx->c, dense, cap_seeds, extra, cap_extra);
    if (n_seeds) *n_seeds = ctx->c->n_dense_seeds;
    if (n_extra) *n_extra = ctx->c->n_extra;
  })
}

int psi_b200_fetch_dense_async(psi_b200_ctx* ctx, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra)
{
  CTX_GUARD(ctx, engine_fetch_dense_async(*ctx->c, dense, cap_seeds, extra, cap_extra))
}

int psi_b200_dense_layout(psi_b200_ctx* ctx, unsigned* off_bytes)
{
  CTX_GUARD(ctx, {
    if (!ctx->c->sh->has_graph) throw StateError("dense_layout: no graph");
    if (off_bytes) *off_bytes = ctx->c->sh->max_node_len <= 32768u ? 2u : 4u;
  })
}

int psi_b200_dense5_layout(psi_b200_ctx* ctx, unsigned* off_bits, int* available)
{
  CTX_GUARD(ctx, {
    if (!ctx->c->sh->has_graph) throw StateError("dense5_layout: no graph");
    if (off_bits) *off_bits = ctx->c->sh->code_off_bits;
    if (available) *available = dense5_available(*ctx->c->sh) ? 1 : 0;
  })
}

int psi_b200_dense_counts(psi_b200_ctx* ctx, uint64_t* n_seeds, uint64_t* n_extra)
{
  CTX_GUARD(ctx, {
    if (ctx->c->pending) throw StateError("dense_counts: a step is in flight on this context (call psi_b200_wait first)");
    if (!ctx->c->records_valid || !ctx->c->records_dense) throw StateError("dense_counts: the last seeds_all was not run with PSI_B200_DENSE");
    if (n_seeds) *n_seeds = ctx->c->n_dense_seeds;
    if (n_extra) *n_extra = ctx->c->n_extra;
  })
}

int psi_b200_build_mem_index(psi_b200_ctx* ctx, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                             const uint32_t* head_off, const uint32_t* tail_trim)
{
  CTX_GUARD(ctx, engine_build_mem_index(*ctx->c, n_paths, path_ptr, path_nodes, head_off, tail_trim))
}

int psi_b200_find_mems(psi_b200_ctx* ctx, unsigned max_mem, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    engine_find_mems(*ctx->c, max_mem);
    if (n_hits) *n_hits = ctx->c->n_mems;
  })
}

int psi_b200_fetch_mems(psi_b200_ctx* ctx, uint64_t* hits, uint64_t cap, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (n_hits) *n_hits = ctx->c->n_mems;
    if (cap && !hits) throw ArgError("fetch_mems: null buffer");
    if (cap) engine_fetch_mems(*ctx->c, hits, cap);
  })
}

int psi_b200_create_distance_index(psi_b200_ctx* ctx, unsigned dmin, unsigned dmax)
{
  CTX_GUARD(ctx, engine_create_distance_index(*ctx->c, dmin, dmax))
}

int psi_b200_verify_distance(psi_b200_ctx* ctx, uint64_t n, const uint32_t* pairs, uint8_t* ok, int on_device)
{
  CTX_GUARD(ctx, engine_verify_distance(*ctx->c, n, pairs, ok, on_device != 0))
}

int psi_b200_seeds_all(psi_b200_ctx* ctx, unsigned flags, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (!(flags & PSI_B200_ALL)) throw ArgError("seeds_all: neither ON_PATHS nor OFF_PATHS requested");
    engine_seeds(*ctx->c, flags);
    if (n_hits) *n_hits = ctx->c->n_hits;
  })
}

int psi_b200_fetch(psi_b200_ctx* ctx, uint64_t* hits, uint64_t cap, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (ctx->c->pending) throw StateError("fetch: a step is in flight on this context (call psi_b200_wait first)");
    if (n_hits) *n_hits = ctx->c->n_hits;
    if (cap && !hits) throw ArgError("fetch: null buffer");
    if (cap) engine_fetch(*ctx->c, hits, cap, false);
  })
}

int psi_b200_fetch32(psi_b200_ctx* ctx, uint32_t* hits, uint64_t cap, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (ctx->c->pending) throw StateError("fetch32: a step is in flight on this context (call psi_b200_wait first)");
    if (n_hits) *n_hits = ctx->c->n_hits;
    if (cap && !hits) throw ArgError("fetch32: null buffer");
    if (cap) engine_fetch(*ctx->c, hits, cap, true);
  })
}

int psi_b200_fetch_kinds(psi_b200_ctx* ctx, uint8_t* kinds, uint64_t cap, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (ctx->c->pending) throw StateError("fetch_kinds: a step is in flight on this context (call psi_b200_wait first)");
    if (n_hits) *n_hits = ctx->c->n_hits;
    if (cap && !kinds) throw ArgError("fetch_kinds: null buffer");
    if (cap) engine_fetch_kinds(*ctx->c, kinds, cap);
  })
}

int psi_b200_fetch_device(psi_b200_ctx* ctx, const uint64_t** d_hits, uint64_t* n_hits)
{
  CTX_GUARD(ctx, {
    if (ctx->c->pending) throw StateError("fetch_device: a step is in flight on this context (call psi_b200_wait first)");
    if (!ctx->c->records_valid) throw StateError("fetch_device: no resolved seed records");
    if (d_hits) *d_hits = ctx->c->records.p;
    if (n_hits) *n_hits = ctx->c->n_hits;
  })
}

int psi_b200_host_alloc(void** p, size_t bytes)
{
  if (!p) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { g_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); (void)cudaGetLastError(); *p = nullptr; return PSI_B200_ERR_CUDA; }
  return PSI_B200_OK;
}

void psi_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

int psi_b200_counters(psi_b200_ctx* ctx, psi_b200_counters_t* out)
{
  CTX_GUARD(ctx, {
    if (!out) throw ArgError("counters: null argument");
    Ctx& c = *ctx->c;
    PSI_CUDA(cudaSetDevice(c.device));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    c.counters.ms_h2d = PhaseTimer::timer_ms(c, T_H2D);
    c.counters.ms_pack = PhaseTimer::timer_ms(c, T_PACK);
    c.counters.ms_read_index = PhaseTimer::timer_ms(c, T_READ_INDEX);
    c.counters.ms_on = PhaseTimer::timer_ms(c, T_ON);
    c.counters.ms_probe = PhaseTimer::timer_ms(c, T_PROBE);
    c.counters.ms_off = PhaseTimer::timer_ms(c, T_OFF);
    c.counters.ms_resolve = PhaseTimer::timer_ms(c, T_RESOLVE);
    c.counters.ms_sort = PhaseTimer::timer_ms(c, T_SORT);
    c.counters.ms_d2h = PhaseTimer::timer_ms(c, T_D2H);
    c.counters.code_by_rank = c.sh->code_by_rank ? 1u : 0u;
    c.counters.code_off_bits = c.sh->code_off_bits;
    *out = c.counters;
  })
}

int psi_b200_reset_counters(psi_b200_ctx* ctx)
{
  CTX_GUARD(ctx, {
    ctx->c->counters.launches = 0;
    ctx->c->counters.ms_probe_sum = ctx->c->counters.ms_on_sum = 0.0;
    ctx->c->counters.timed_steps = 0;
  })
}

}  // extern "C"
