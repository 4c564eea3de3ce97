// graph.cu -- context lifecycle, graph upload/flattening, table allocation.
//
// The graph crosses the C-ABI as CSR + concatenated ASCII labels (what gum's
// node_sequence / for_each_edges_out expose, reference
// gum/seqgraph_succinct.hpp:194-199, gum/digraph_succinct.hpp:595-610,722-731)
// and is re-laid out for the walkers: one 16-byte NodeRec per node, 2-bit
// labels, an N bitmap and a sampled position->node table.
#include "engine.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace psi_b200 {

using namespace dev;

// One thread per 32-base word: ASCII -> 2-bit + N bitmap.
__global__ void __launch_bounds__(256)
pack_labels_kernel(const char* __restrict__ seq, uint64_t n_bases, uint64_t* __restrict__ seq2,
                   uint32_t* __restrict__ nmask, unsigned long long* __restrict__ n_count)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_words = (n_bases + 31) >> 5;
  if (w >= n_words) return;
  const uint64_t b0 = w << 5;
  uint64_t bits = 0;
  uint32_t mask = 0;
  const uint32_t cnt = (uint32_t)min((uint64_t)32, n_bases - b0);
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint32_t c = base_code((unsigned char)seq[b0 + i]);
    if (c > 3) mask |= 1u << i;
    else bits |= (uint64_t)c << (2 * i);
  }
  seq2[w] = bits;
  nmask[w] = mask;
  if (mask) atomicAdd(n_count, (unsigned long long)__popc(mask));
}

// One thread per node: sampled position -> node table.
__global__ void __launch_bounds__(256)
pos2node_kernel(const NodeRec* __restrict__ rec, uint32_t n_nodes, uint32_t* __restrict__ pos2node, uint32_t shift)
{
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const NodeRec r = rec[v];
  if (r.seq_len == 0) return;
  const uint64_t step = 1ull << shift;
  uint64_t p = ((uint64_t)r.seq_start + step - 1) & ~(step - 1);
  for (; p < (uint64_t)r.seq_start + r.seq_len; p += step) pos2node[p >> shift] = v;
}

Ctx* engine_create(int device, unsigned seed_len)
{
  if (seed_len == 0 || seed_len > PSI_B200_MAX_SEED_LEN)
    throw ArgError("seed length must be in [1, 32]");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw CudaError(std::string("no CUDA device available (no CPU fallback exists): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) throw ArgError("invalid device ordinal");
  PSI_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PSI_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) throw CudaError("libpsi_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
  // owned by a guard until fully constructed: a failing CUDA call below must not leak the stream, events or buffers
  struct Guard {
    Ctx* c;
    ~Guard() { if (c) engine_destroy(c); }
  } guard{ new Ctx() };
  Ctx* c = guard.c;
  c->device = device;
  c->k = seed_len;
  c->sm_count = prop.multiProcessorCount;
  PSI_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  for (auto& ev : c->ev) PSI_CUDA(cudaEventCreate(&ev));
  PSI_CUDA(cudaEventCreateWithFlags(&c->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
  c->dev_counters.ensure(DC_COUNT);
  PSI_CUDA(cudaMemsetAsync(c->dev_counters.p, 0, DC_COUNT * sizeof(unsigned long long), c->stream));
  PSI_CUDA(cudaHostAlloc((void**)&c->h_pinned, (2 * DC_COUNT + 8) * sizeof(uint64_t), cudaHostAllocDefault));
  guard.c = nullptr;
  return c;
}

// A second pipeline on the same GPU: shares the resident graph, path index and
// starting loci, owns its stream, chunk buffers and result buffers.  This is how
// one SeedFinder is driven from several host threads on different chunks (the
// reference's stats are thread-aware for exactly that use, seed_finder.hpp:386-399)
// and how H2D of chunk i+1 overlaps the kernels and D2H of chunk i.
Ctx* engine_fork(Ctx& parent)
{
  PSI_CUDA(cudaSetDevice(parent.device));
  PSI_CUDA(cudaStreamSynchronize(parent.stream));
  Ctx* c = engine_create(parent.device, parent.k);
  c->sh = parent.sh;
  const psi_b200_counters_t& pc = parent.counters;
  c->counters.n_nodes = pc.n_nodes; c->counters.n_edges = pc.n_edges; c->counters.n_bases = pc.n_bases;
  c->counters.n_path_bases = pc.n_path_bases; c->counters.n_index_entries = pc.n_index_entries;
  c->counters.n_index_kmers = pc.n_index_kmers; c->counters.index_bytes = pc.index_bytes;
  c->counters.index_buckets = pc.index_buckets; c->counters.index_slot_bytes = pc.index_slot_bytes;
  c->counters.n_loci = pc.n_loci; c->counters.ms_index_build = pc.ms_index_build; c->counters.ms_find_loci = pc.ms_find_loci;
  c->counters.n_offpath_entries = pc.n_offpath_entries; c->counters.index_stash_used = pc.index_stash_used;
  c->counters.offpath_mode = pc.offpath_mode; c->counters.n_offpath_walks = pc.n_offpath_walks;
  c->counters.n_dindex_entries = pc.n_dindex_entries; c->counters.dindex_bytes = pc.dindex_bytes;
  c->counters.dindex_mode = pc.dindex_mode; c->counters.ms_dindex_build = pc.ms_dindex_build;
  c->spill_items = parent.spill_items;
  c->opt_offpath_mode = parent.opt_offpath_mode;
  c->opt_seeding_mode = parent.opt_seeding_mode;
  c->opt_fused = parent.opt_fused;
  c->opt_timers = parent.opt_timers;
  c->opt_code_by_rank = parent.opt_code_by_rank;
  c->opt_gocc_threshold = parent.opt_gocc_threshold;
  c->opt_blocking_sync = parent.opt_blocking_sync;
  c->opt_index_slack = parent.opt_index_slack;
  c->opt_resolve_items = parent.opt_resolve_items;
  c->opt_resolve_ctas = parent.opt_resolve_ctas;
  c->opt_offpath_max_pairs = parent.opt_offpath_max_pairs;
  c->opt_dindex_mode = parent.opt_dindex_mode;
  c->opt_build_slices = parent.opt_build_slices;
  c->counters.index_build_slices = pc.index_build_slices;
  c->opt_dindex_list_cap = parent.opt_dindex_list_cap;
  c->opt_dindex_max_bytes = parent.opt_dindex_max_bytes;
  return c;
}

void engine_destroy(Ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  if (c->ev_sync) cudaEventDestroy(c->ev_sync);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

void engine_set_graph(Ctx& c, uint64_t n_nodes, const uint64_t* seq_start, const char* seq,
                      const uint64_t* row_ptr, const uint32_t* col, const uint64_t* node_id)
{
  if (n_nodes == 0 || !seq_start || !seq || !row_ptr || !node_id) throw ArgError("set_graph: null or empty graph");
  if (n_nodes >= 0xfffffff0ull) throw ArgError("set_graph: too many nodes");
  const uint64_t n_bases = seq_start[n_nodes];
  const uint64_t n_edges = row_ptr[n_nodes];
  if (n_bases >= 0xfffffff0ull) throw ArgError("set_graph: graphs above 4 Gbp are not supported (32-bit loci)");
  if (n_edges >= 0xfffffff0ull) throw ArgError("set_graph: too many edges");
  if (n_edges && !col) throw ArgError("set_graph: null adjacency");
  PSI_CUDA(cudaSetDevice(c.device));
  if (c.sh.use_count() > 1) c.sh = std::make_shared<Shared>();  // forks keep the graph they were given

  std::vector<NodeRec> rec(n_nodes + 1);
  for (uint64_t v = 0; v < n_nodes; ++v) {
    if (seq_start[v + 1] < seq_start[v] || row_ptr[v + 1] < row_ptr[v]) throw ArgError("set_graph: offsets not monotone");
    rec[v].seq_start = (uint32_t)seq_start[v];
    rec[v].seq_len = (uint32_t)(seq_start[v + 1] - seq_start[v]);
    rec[v].edge_start = (uint32_t)row_ptr[v];
    rec[v].outdeg = (uint32_t)(row_ptr[v + 1] - row_ptr[v]);
  }
  rec[n_nodes] = NodeRec{ (uint32_t)n_bases, 0, (uint32_t)n_edges, 0 };
  // rank structure for position -> node (zero-length nodes would share a start bit: fall back to pos2node then)
  bool zero_len = false;
  uint64_t max_id = 0;
  uint32_t max_len = 0;
  std::vector<Rank16> rank16((n_bases >> 6) + 2, Rank16{ 0, 0, 0 });
  std::vector<NodeRes> node_res(n_nodes + 1);
  for (uint64_t v = 0; v < n_nodes; ++v) {
    node_res[v] = NodeRes{ rec[v].seq_start, 0, node_id[v] };
    if (node_id[v] > max_id) max_id = node_id[v];
    if (rec[v].seq_len > max_len) max_len = rec[v].seq_len;
    if (rec[v].seq_len == 0) { zero_len = true; continue; }
    rank16[rec[v].seq_start >> 6].bits |= 1ull << (rec[v].seq_start & 63u);
  }
  node_res[n_nodes] = NodeRes{ (uint32_t)n_bases, 0, 0 };
  c.sh->max_node_id = max_id;
  c.sh->max_node_len = max_len;
  {
    uint32_t run = 0;
    for (auto& r : rank16) { r.prefix = run; run += (uint32_t)__builtin_popcountll(r.bits); }
  }
  for (uint64_t e = 0; e < n_edges; ++e)
    if (col[e] >= n_nodes) throw ArgError("set_graph: successor rank out of range");

  c.sh->n_nodes = (uint32_t)n_nodes;
  c.sh->n_edges = (uint32_t)n_edges;
  c.sh->n_bases = n_bases;
  c.sh->node_rec.ensure(n_nodes + 1);
  c.sh->col.ensure(n_edges + 1);
  c.sh->node_id.ensure(n_nodes);
  const uint64_t n_words = (n_bases + 31) >> 5;
  c.sh->seq2.ensure(n_words + 2);
  c.sh->nmask.ensure(n_words + 2);
  c.sh->pos2node.ensure((n_bases >> Ctx::POS2NODE_SHIFT) + 2);
  {
    const size_t rank_bytes = rank16.size() * sizeof(Rank16), res_bytes = node_res.size() * sizeof(NodeRes);
    c.sh->gather_pool.ensure(rank_bytes + res_bytes + 256);
    c.sh->rank16 = reinterpret_cast<Rank16*>(c.sh->gather_pool.p);
    c.sh->node_res = reinterpret_cast<NodeRes*>(c.sh->gather_pool.p + rank_bytes);
    c.sh->gather_pool_bytes = rank_bytes + res_bytes;
    c.sh->has_rank16 = !zero_len;
    PSI_CUDA(cudaMemcpyAsync(c.sh->rank16, rank16.data(), rank_bytes, cudaMemcpyHostToDevice, c.stream));
    PSI_CUDA(cudaMemcpyAsync(c.sh->node_res, node_res.data(), res_bytes, cudaMemcpyHostToDevice, c.stream));
  }

  DevBuf<char> ascii;
  ascii.ensure(n_bases + 1);
  PSI_CUDA(cudaMemcpyAsync(ascii.p, seq, n_bases, cudaMemcpyHostToDevice, c.stream));
  PSI_CUDA(cudaMemcpyAsync(c.sh->node_rec.p, rec.data(), (n_nodes + 1) * sizeof(NodeRec), cudaMemcpyHostToDevice, c.stream));
  if (n_edges) PSI_CUDA(cudaMemcpyAsync(c.sh->col.p, col, n_edges * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  PSI_CUDA(cudaMemcpyAsync(c.sh->node_id.p, node_id, n_nodes * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
  PSI_CUDA(cudaMemsetAsync(c.sh->seq2.p, 0, (n_words + 2) * sizeof(uint64_t), c.stream));
  PSI_CUDA(cudaMemsetAsync(c.sh->nmask.p, 0, (n_words + 2) * sizeof(uint32_t), c.stream));
  PSI_CUDA(cudaMemsetAsync(c.sh->pos2node.p, 0, ((n_bases >> Ctx::POS2NODE_SHIFT) + 2) * sizeof(uint32_t), c.stream));
  PSI_CUDA(cudaMemsetAsync(c.dev_counters.p + DC_AUX, 0, sizeof(unsigned long long), c.stream));
  if (n_words) {
    pack_labels_kernel<<<grid_for(n_words, 256), 256, 0, c.stream>>>(ascii.p, n_bases, c.sh->seq2.p, c.sh->nmask.p,
                                                                     c.dev_counters.p + DC_AUX);
    ++c.counters.launches;
  }
  pos2node_kernel<<<grid_for(n_nodes, 256), 256, 0, c.stream>>>(c.sh->node_rec.p, c.sh->n_nodes, c.sh->pos2node.p, Ctx::POS2NODE_SHIFT);
  ++c.counters.launches;
  PSI_CUDA(cudaGetLastError());
  unsigned long long n_count = 0;
  PSI_CUDA(cudaMemcpyAsync(&n_count, c.dev_counters.p + DC_AUX, sizeof(n_count), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.sh->graph_has_n = n_count != 0;
  c.sh->has_graph = true;
  c.sh->has_index = false;
  c.sh->has_table = false;
  c.sh->offpath_indexed = false;
  c.sh->n_on_pairs = c.sh->n_off_pairs = 0;
  c.sh->n_loci = 0;
  c.counters.n_nodes = n_nodes;
  c.counters.n_edges = n_edges;
  c.counters.n_bases = n_bases;
}

// ------------------------------------------------------------- tables --

static uint32_t ceil_log2(uint64_t x)
{
  uint32_t b = 0;
  while ((1ull << b) < x) ++b;
  return b;
}

struct TablePlan {
  uint32_t fmt, line_bits, rem_bits;
};

// A bucket is a 128-byte line.  Size for at most ~10 keys per 16-slot line
// (fmt 8) or ~5 per 8-slot line (fmt 16): with a power-of-two line count the
// mean load is 0.31-0.63, about 3 % of the lines overflow into their successor
// at the upper end, and a lookup still reads exactly one line in ~97 % of the cases.
// 8-byte slots need the k-mer's remainder to fit 27 bits AND the payload to fit the 32 + 27 - rem_bits left over.
static TablePlan table_plan(uint64_t n_keys, uint32_t kbits, int& slack_bits, uint32_t min_payload_bits)
{
  TablePlan p{ 16, ceil_log2((n_keys + 4) / 5 + 1), 0 };
  {
    uint32_t lb = ceil_log2((n_keys + 9) / 10 + 1);
    if (lb > kbits) lb = kbits;               // tiny k: at most one possible key per line
    // extra line bits an explicit slack will add below also shorten the remainder
    uint32_t lb_final = lb;
    for (int i = 0; i < (slack_bits > 0 ? slack_bits : 0); ++i) if (lb_final < kbits) ++lb_final;
    if (kbits - lb <= 27 && 32u + 27u - (kbits - lb_final) >= min_payload_bits) { p.fmt = 8; p.line_bits = lb; p.rem_bits = kbits - lb; }
  }
  // slack: 2^slack_bits times the lines.  Halving the load makes a full home line (and with it the slow path of the
  // probe) a rarity -- 1 % -> 0.001 % of the lines at 16 slots, 6 % -> 0.1 % at 8 slots -- for twice the memory.
  // Measured (r02c): 16-slot lines (chr22 shape, k = 20) gain 1.3 % per step, not worth 1 GB; 8-slot lines (MHC shape,
  // k = 32) gain 13 % and the probe kernel alone halves its time.  Hence auto (-1): one extra bit for 16-byte slots
  // while the table stays below 1/16 of the device memory (11 GB on a 180 GB B200), none for 8-byte slots.
  if (slack_bits < 0) {
    size_t free_b = 0, total_b = 0;
    slack_bits = 0;
    if (p.fmt == 16 && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (256ull << p.line_bits) <= total_b / 16 &&
        (256ull << p.line_bits) <= free_b / 4)
      slack_bits = 1;
  }
  for (int i = 0; i < slack_bits; ++i) {
    if (p.fmt == 8 && (p.rem_bits == 0 || p.line_bits + 1 > kbits)) break;
    ++p.line_bits;
    if (p.fmt == 8) --p.rem_bits;
  }
  return p;
}

uint32_t table_fmt8_payload_bits(uint64_t n_keys, uint32_t kbits, int slack_bits)
{
  const TablePlan p = table_plan(n_keys, kbits, slack_bits, 32);
  return p.fmt == 8 ? 32u + 27u - p.rem_bits : 0u;
}

void table_alloc(Ctx& c, HostTable& t, uint64_t n_keys, uint32_t kbits, uint64_t stash_slots, int slack_bits, uint32_t min_payload_bits)
{
  if (min_payload_bits > 62) throw ArgError("table: payload wider than 62 bits");
  const TablePlan p = table_plan(n_keys, kbits, slack_bits, min_payload_bits);
  const uint32_t fmt = p.fmt, line_bits = p.line_bits, rem_bits = p.rem_bits;
  if (line_bits > 32) throw ArgError("table too large");   // line indices travel as 32-bit values
  const uint64_t n_lines = 1ull << line_bits;
  t.n_lines = n_lines;
  t.slots.ensure(n_lines * 128);
  uint64_t ss = 1024;
  while (ss < stash_slots) ss <<= 1;
  t.stash.ensure(ss);
  t.stash_used.ensure(1);
  t.view.slots = t.slots.p;
  t.view.stash = t.stash.p;
  t.view.stash_used = t.stash_used.p;
  t.view.line_bits = line_bits;
  t.view.kbits = kbits;
  t.view.rem_bits = rem_bits;
  t.view.pay_hi = fmt == 8 ? 27u - rem_bits : 0u;
  t.view.fmt = fmt;
  t.view.stash_mask = (uint32_t)(ss - 1);
  t.view.stash_nonempty = 1;  // until the build has been checked
  table_clear(c, t);
}

void table_clear(Ctx& c, HostTable& t)
{
  PSI_CUDA(cudaMemsetAsync(t.slots.p, 0xff, t.n_lines * 128, c.stream));
  PSI_CUDA(cudaMemsetAsync(t.stash.p, 0xff, ((size_t)t.view.stash_mask + 1) * sizeof(Slot16), c.stream));
  PSI_CUDA(cudaMemsetAsync(t.stash_used.p, 0, sizeof(uint32_t), c.stream));
}

}  // namespace psi_b200
