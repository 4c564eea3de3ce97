// distance.cu -- paired-end distance verification on the device.
//
// Stands behind SeedFinder::create_distance_index / verify_distance of the reference (include/psi/seed_finder.hpp:
// 1193-1265,1300-1317).  There the index is DiVerG's boolean matrix D = A^dmin (A + I)^(dmax - dmin) over the
// CHARACTER-level adjacency A of the graph (ext/diverg/include/diverg/dindex.hpp:767-914), range-compressed per row,
// and a query is D(charorder(v) + o, charorder(u) + p): "is there a walk of dmin <= l <= dmax characters from locus
// (v, o) to locus (u, p)" (v != u; inside one node only the offsets are compared, seed_finder.hpp:1306-1309).
//
// Here the same relation is kept per NODE instead of per character.  All characters of a node share their
// continuations, so with S(v) = { (x, s) : some walk leaves v and reaches the first character of x, s characters after
// the first character of v } the answer is
//       exists (u, s) in S(v) with  dmin <= s - o + p <= dmax.
// S(v) is enumerated by one warp per node: a frontier of (node, distance) states with a visited set (any graph, cycles
// included; the distance bound ends it), both in shared memory, falling back to a global scratch region for the nodes
// whose state count does not fit.  The rows are kept in HBM ("materialised": 8 bytes per entry, entries that can never
// satisfy the window are dropped) and a query is a scan of row(v) by 8 lanes -- or, when the rows would not fit the
// budget, nothing is kept and a query runs the same enumeration from (v, o) with early exit ("walk" mode).
#include <algorithm>
#include <cstring>
#include <vector>

#include "engine.hpp"
#include "walker.cuh"

namespace psi_b200 {

using namespace dev;

namespace {

constexpr unsigned long long DS_EMPTY = ~0ull;
constexpr uint32_t DS_WARPS = 4;            // warps per CTA of the enumerating kernels
constexpr uint32_t DS_BIG_LIST = 1u << 16;  // states per warp in the global scratch region
constexpr uint32_t ROW_OVERFLOW = 0xffffffffu;

// one warp's frontier: list[0 .. tail) are the states reached so far (node << 32 | distance), set[] is an open-addressing
// hash set of the same keys with twice the capacity
struct DistScratch {
  unsigned long long* list;
  unsigned long long* set;
  uint32_t list_cap;
  uint32_t set_mask;
};

__device__ __forceinline__ uint32_t ds_hash(unsigned long long key)
{
  key ^= key >> 29;
  key *= 0x9e3779b97f4a7c15ull;
  return (uint32_t)(key >> 32);
}

struct NoTarget {
  __device__ __forceinline__ bool hit(uint32_t, uint32_t) const { return false; }
};
// locus (u, p) under the window [dmin, dmax] for distances counted from the query's own start locus
struct Target {
  uint32_t u, p, dmin, dmax;
  __device__ __forceinline__ bool hit(uint32_t x, uint32_t s) const
  {
    const unsigned long long l = (unsigned long long)s + p;
    return x == u && l >= dmin && l <= dmax;
  }
};

// One batch of the enumeration: every lane offers `deg` successors (col[es], col[es + stride], ...) reached at distance
// s_child.  New states are appended to the list with one ballot per round.  Returns false on overflow.
template <class TTarget>
__device__ __forceinline__ bool ds_push_children(const GraphView& g, const DistScratch& sc, uint32_t lane, uint32_t es,
                                                 uint32_t deg, uint32_t stride, uint32_t s_child, uint32_t bound,
                                                 const TTarget& target, uint32_t& tail, bool& hit)
{
  const uint32_t maxdeg = __reduce_max_sync(0xffffffffu, deg);
  for (uint32_t j = 0; j < maxdeg; ++j) {
    bool isnew = false;
    unsigned long long key = 0;
    if (j < deg && s_child <= bound) {
      const uint32_t x = __ldg(g.col + es + (size_t)j * stride);
      key = ((unsigned long long)x << 32) | s_child;
      hit = hit || target.hit(x, s_child);
      uint32_t slot = ds_hash(key) & sc.set_mask;
      for (;;) {
        const unsigned long long old = atomicCAS(sc.set + slot, DS_EMPTY, key);
        if (old == DS_EMPTY) { isnew = true; break; }
        if (old == key) break;
        slot = (slot + 1) & sc.set_mask;
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, isnew);
    const uint32_t add = __popc(m);
    if (tail + add > sc.list_cap) return false;
    if (isnew) sc.list[tail + __popc(m & ((1u << lane) - 1u))] = key;
    tail += add;
    __syncwarp();
  }
  return true;
}

// Enumerates the states reachable from node v when its successors are reached at distance s0 and distances beyond
// `bound` are of no interest.  Warp-cooperative; returns the number of states (list[0 .. n)), or ROW_OVERFLOW when they
// do not fit the scratch.  With a target the enumeration stops at the first state that satisfies it (found = true).
template <class TTarget>
__device__ uint32_t ds_enumerate(const GraphView& g, uint32_t v, uint32_t s0, uint32_t bound, const DistScratch& sc,
                                 uint32_t lane, const TTarget& target, bool& found)
{
  for (uint32_t i = lane; i <= sc.set_mask; i += 32) sc.set[i] = DS_EMPTY;
  __syncwarp();
  uint32_t tail = 0, head = 0;
  bool hit = false;
  found = false;
  const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(g.rec + v));   // {seq_start, seq_len, edge_start, outdeg}
  {
    const uint32_t deg = r0.w > lane ? (r0.w - lane + 31u) / 32u : 0u;
    if (!ds_push_children(g, sc, lane, r0.z + lane, deg, 32u, s0, bound, target, tail, hit)) return ROW_OVERFLOW;
  }
  if (__any_sync(0xffffffffu, hit)) { found = true; return tail; }
  while (head < tail) {
    const uint32_t idx = head + lane;
    uint32_t es = 0, deg = 0, s_child = 0;
    if (idx < tail) {
      const unsigned long long key = sc.list[idx];
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(g.rec + (uint32_t)(key >> 32)));
      const unsigned long long s = (unsigned long long)(uint32_t)key + r.y;
      if (s <= bound) { es = r.z; deg = r.w; s_child = (uint32_t)s; }
    }
    head = min(tail, head + 32u);
    if (!ds_push_children(g, sc, lane, es, deg, 1u, s_child, bound, target, tail, hit)) return ROW_OVERFLOW;
    if (__any_sync(0xffffffffu, hit)) { found = true; return tail; }
  }
  return tail;
}

__device__ __forceinline__ DistScratch ds_scratch(unsigned long long* smem, unsigned long long* big, uint32_t list_cap,
                                                  uint32_t warp, uint32_t warps_per_cta)
{
  DistScratch sc;
  unsigned long long* base = big ? big + ((size_t)blockIdx.x * warps_per_cta + warp) * 3u * list_cap
                                 : smem + (size_t)warp * 3u * list_cap;
  sc.list = base;
  sc.set = base + list_cap;
  sc.list_cap = list_cap;
  sc.set_mask = 2u * list_cap - 1u;
  return sc;
}

// ---- building the rows ----
// FILL = 0: row_len[v] = entries of row(v) (nodes that overflow the scratch are queued and marked);
// FILL = 1: the entries are written behind an atomic reservation and row_start[v] is set.
// items == nullptr: all nodes; else the queued ones.  counters: [0] work claim, [1] queue length, [2] entries
// reserved / counted, [3] error flag.
template <int FILL>
__global__ void __launch_bounds__(DS_WARPS * 32)
dist_rows_kernel(GraphView g, const uint32_t* __restrict__ items, uint32_t n_items, uint32_t dmin, uint32_t dmax,
                 uint32_t list_cap, unsigned long long* big, uint32_t* __restrict__ row_len,
                 unsigned long long* __restrict__ row_start, unsigned long long* __restrict__ entries,
                 uint32_t* __restrict__ queue, unsigned long long* counters)
{
  extern __shared__ unsigned long long ds_smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const DistScratch sc = ds_scratch(ds_smem, big, list_cap, warp, DS_WARPS);
  for (;;) {
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(counters + 0, 1ull);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    if (claim >= n_items) break;
    const uint32_t v = items ? items[claim] : (uint32_t)claim;
    if (!items && FILL && row_len[v] == ROW_OVERFLOW) continue;   // served by the pass over the queue
    if (!items && FILL && row_len[v] == 0) { if (lane == 0) row_start[v] = 0; continue; }
    const uint32_t len_v = __ldg(&g.rec[v].seq_len);
    const uint32_t bound = dmax + (len_v ? len_v - 1u : 0u);
    bool found;
    const uint32_t n = ds_enumerate(g, v, len_v, bound, sc, lane, NoTarget{}, found);
    if (n == ROW_OVERFLOW) {
      if (lane == 0) {
        if (big) atomicExch(counters + 3, 1ull);       // does not fit the global scratch either
        else if (!FILL) { row_len[v] = ROW_OVERFLOW; queue[atomicAdd(counters + 1, 1ull)] = v; }
      }
      continue;
    }
    // keep the states that can satisfy the window for some pair of offsets: s + |x| - 1 >= dmin
    uint32_t kept = 0;
    unsigned long long base = 0;
    if (FILL) {
      if (lane == 0) base = atomicAdd(counters + 2, (unsigned long long)row_len[v]);
      base = __shfl_sync(0xffffffffu, base, 0);
    }
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool keep = false;
      unsigned long long key = 0;
      if (i < n) {
        key = sc.list[i];
        const uint32_t len_x = __ldg(&g.rec[(uint32_t)(key >> 32)].seq_len);
        keep = len_x != 0 && (unsigned long long)(uint32_t)key + len_x - 1u >= dmin;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (FILL && keep) entries[base + kept + __popc(m & ((1u << lane) - 1u))] = key;
      kept += __popc(m);
    }
    if (lane == 0) {
      if (FILL) row_start[v] = base;
      else { row_len[v] = kept; atomicAdd(counters + 2, (unsigned long long)kept); }
    }
    __syncwarp();
  }
}

// ---- queries ----
struct alignas(16) DistQuery { uint32_t v, o, u, p; };

__device__ __forceinline__ bool dist_same_node(const DistQuery& q, uint32_t dmin, uint32_t dmax)
{
  return q.o <= q.p && q.p - q.o >= dmin && q.p - q.o <= dmax;   // seed_finder.hpp:1306-1309
}

// both loci must exist: ranks below the node count, offsets inside the labels (anything else is answered "no")
__device__ __forceinline__ bool dist_valid(const GraphView& g, const DistQuery& q)
{
  return q.v < g.n_nodes && q.u < g.n_nodes && q.o < __ldg(&g.rec[q.v].seq_len) && q.p < __ldg(&g.rec[q.u].seq_len);
}

// One 16-byte record per node for the queries: where its row starts, how long it is, and the node's label length (so that
// a query costs one gather for v and one for u besides its row).
__global__ void __launch_bounds__(256)
dist_pack_rows_kernel(GraphView g, const uint32_t* __restrict__ row_len, const unsigned long long* __restrict__ row_start,
                      uint4* __restrict__ rows)
{
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n_nodes) return;
  const unsigned long long st = row_start[v];
  rows[v] = make_uint4((uint32_t)st, (uint32_t)(st >> 32), row_len[v], g.rec[v].seq_len);
}

// against the materialised rows: 8 lanes scan row(v)
__global__ void __launch_bounds__(256)
dist_query_rows_kernel(const DistQuery* __restrict__ queries, uint64_t n, uint32_t n_nodes, uint32_t dmin, uint32_t dmax,
                       const uint4* __restrict__ rows, const unsigned long long* __restrict__ entries, uint8_t* __restrict__ out)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t qi = t >> 3;
  const uint32_t sub = threadIdx.x & 7u;
  bool ok = false;
  if (qi < n) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(queries + qi));
    const DistQuery q{ raw.x, raw.y, raw.z, raw.w };
    if (q.v < n_nodes && q.u < n_nodes) {
      const uint4 rv = __ldg(rows + q.v);
      const uint32_t len_u = q.u == q.v ? rv.w : __ldg(&rows[q.u].w);
      if (q.o < rv.w && q.p < len_u) {          // both loci exist
        if (q.v == q.u) ok = dist_same_node(q, dmin, dmax);
        else {
          const long long lo = (long long)dmin + q.o - q.p, hi = (long long)dmax + q.o - q.p;
          const unsigned long long* row = entries + (((unsigned long long)rv.y << 32) | rv.x);
          for (uint32_t i = sub; i < rv.z; i += 8) {
            const unsigned long long e = __ldg(row + i);
            const long long s = (long long)(uint32_t)e;
            ok = ok || ((uint32_t)(e >> 32) == q.u && s >= lo && s <= hi);
          }
        }
      }
    }
  }
  uint32_t any = ok ? 1u : 0u;
  any |= __shfl_xor_sync(0xffffffffu, any, 1);
  any |= __shfl_xor_sync(0xffffffffu, any, 2);
  any |= __shfl_xor_sync(0xffffffffu, any, 4);
  if (qi < n && sub == 0) out[qi] = (uint8_t)any;
}

// without rows: a warp enumerates from (v, o) until it meets (u, p) inside the window.  Queries that overflow the shared
// scratch are queued for the pass with the global one.  counters as above.
__global__ void __launch_bounds__(DS_WARPS * 32)
dist_query_walk_kernel(GraphView g, const DistQuery* __restrict__ queries, const uint32_t* __restrict__ items, uint64_t n_items,
                       uint32_t dmin, uint32_t dmax, uint32_t list_cap, unsigned long long* big, uint8_t* __restrict__ out,
                       uint32_t* __restrict__ queue, unsigned long long* counters)
{
  extern __shared__ unsigned long long ds_smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const DistScratch sc = ds_scratch(ds_smem, big, list_cap, warp, DS_WARPS);
  for (;;) {
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(counters + 0, 1ull);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    if (claim >= n_items) break;
    const uint64_t qi = items ? items[claim] : claim;
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(queries + qi));
    const DistQuery q{ raw.x, raw.y, raw.z, raw.w };
    bool ok = false;
    if (dist_valid(g, q)) {
      if (q.v == q.u) ok = dist_same_node(q, dmin, dmax);
      else {
        const uint32_t len_v = __ldg(&g.rec[q.v].seq_len);
        {
          bool found;
          const uint32_t n = ds_enumerate(g, q.v, len_v - q.o, dmax, sc, lane, Target{ q.u, q.p, dmin, dmax }, found);
          if (n == ROW_OVERFLOW) {
            if (lane == 0) {
              if (big) atomicExch(counters + 3, 1ull);
              else queue[atomicAdd(counters + 1, 1ull)] = (uint32_t)qi;
            }
            __syncwarp();
            continue;
          }
          ok = found;
        }
      }
    }
    if (lane == 0) out[qi] = ok ? 1 : 0;
    __syncwarp();
  }
}

size_t ds_smem_bytes(uint32_t list_cap) { return (size_t)DS_WARPS * 3u * list_cap * sizeof(unsigned long long); }

// persistent grid: as many CTAs as fit, a few per SM
unsigned ds_grid(const Ctx& c, uint32_t list_cap)
{
  const size_t per_cta = ds_smem_bytes(list_cap);
  unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (200u << 10) / per_cta));
  return (unsigned)c.sm_count * per_sm;
}

template <class K>
void ds_allow_smem(K kernel, size_t bytes)
{
  PSI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// the global scratch of the overflow passes: 2 warps per SM
unsigned ds_big_grid(const Ctx& c) { return (unsigned)c.sm_count / 2u + 1u; }

}  // namespace

void engine_create_distance_index(Ctx& c, unsigned dmin, unsigned dmax)
{
  Shared& sh = *c.sh;
  if (!sh.has_graph) throw StateError("create_distance_index: no graph (psi_b200_set_graph first)");
  if (c.sh.use_count() > 1) throw StateError("create_distance_index: the index is shared with forked contexts");
  if (c.pending) throw StateError("create_distance_index: a step is in flight on this context (call psi_b200_wait first)");
  PSI_CUDA(cudaSetDevice(c.device));
  sh.has_dindex = false;
  sh.dist_rows = false;
  sh.dist_row_len.release(); sh.dist_row_start.release(); sh.dist_entries.release(); sh.dist_packed.release();
  sh.n_dist_entries = 0;
  c.counters.n_dindex_entries = 0; c.counters.dindex_bytes = 0; c.counters.dindex_mode = 0; c.counters.ms_dindex_build = 0;
  if (dmin == 0 || dmax < dmin) return;   // "not constructible" (seed_finder.hpp:1198): queries stay unavailable
  if ((uint64_t)dmax + sh.max_node_len >= 0x7fffffffull) throw ArgError("create_distance_index: distance window beyond 2^31");
  sh.dist_dmin = dmin;
  sh.dist_dmax = dmax;
  sh.has_dindex = true;
  nvtxRangePushA("index-distances");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop_range;
  if (c.opt_dindex_mode == 1 || sh.n_nodes == 0) { c.counters.dindex_mode = 1; return; }   // walk mode: nothing to build

  const GraphView g = make_graph_view(c);
  const uint32_t n = sh.n_nodes;
  cudaEvent_t e0, e1;
  PSI_CUDA(cudaEventCreate(&e0)); PSI_CUDA(cudaEventCreate(&e1));
  struct Ev { cudaEvent_t a, b; ~Ev() { cudaEventDestroy(a); cudaEventDestroy(b); } } evs{ e0, e1 };
  PSI_CUDA(cudaEventRecord(e0, c.stream));
  DevBuf<uint32_t> queue;
  DevBuf<unsigned long long> counters, big;
  queue.ensure(n);
  counters.ensure(4);
  sh.dist_row_len.ensure(n);
  sh.dist_row_start.ensure(n);
  unsigned long long h_cnt[4] = { 0, 0, 0, 0 };
  uint32_t list_cap = c.opt_dindex_list_cap;
  // pass 1: row lengths.  Retried once with more shared memory per warp when many nodes do not fit.
  for (int attempt = 0; attempt < 2; ++attempt) {
    PSI_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned long long), c.stream));
    ds_allow_smem(dist_rows_kernel<0>, ds_smem_bytes(list_cap));
    dist_rows_kernel<0><<<ds_grid(c, list_cap), DS_WARPS * 32, ds_smem_bytes(list_cap), c.stream>>>(
        g, nullptr, n, dmin, dmax, list_cap, nullptr, sh.dist_row_len.p, sh.dist_row_start.p, nullptr, queue.p, counters.p);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (h_cnt[1] <= n / 16 || list_cap != 256) break;   // only the default is widened; an explicit capacity is kept
    list_cap = 1024;
  }
  const uint64_t n_queued = h_cnt[1];
  uint64_t total = h_cnt[2];
  if (n_queued) {
    big.ensure((size_t)ds_big_grid(c) * DS_WARPS * 3u * DS_BIG_LIST);
    const unsigned long long z[4] = { 0, 0, total, 0 };
    PSI_CUDA(cudaMemcpyAsync(counters.p, z, sizeof z, cudaMemcpyHostToDevice, c.stream));
    dist_rows_kernel<0><<<ds_big_grid(c), DS_WARPS * 32, 0, c.stream>>>(g, queue.p, (uint32_t)n_queued, dmin, dmax, DS_BIG_LIST, big.p,
                                                                       sh.dist_row_len.p, sh.dist_row_start.p, nullptr, nullptr, counters.p);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (h_cnt[3]) {
      sh.has_dindex = false;
      throw OverflowError("create_distance_index: a node reaches more than 65 536 (node, distance) states inside the window");
    }
    total = h_cnt[2];
  }
  // materialise when the rows fit the budget, else answer queries by enumeration
  size_t free_b = 0, total_b = 0;
  PSI_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t budget = c.opt_dindex_max_bytes ? c.opt_dindex_max_bytes : (uint64_t)free_b / 2;
  if (c.opt_dindex_mode != 2 && total * sizeof(unsigned long long) > budget) {
    sh.dist_row_len.release(); sh.dist_row_start.release();
    c.counters.dindex_mode = 1;
    c.counters.n_dindex_entries = total;
    return;
  }
  sh.dist_entries.ensure(total + 1);
  // pass 2: the entries.  The nodes pass 1 queued overflow the shared scratch again (the state count of a node is
  // deterministic) and are skipped by the first kernel; the second one serves them from the queue.
  {
    const unsigned long long z[4] = { 0, 0, 0, 0 };
    PSI_CUDA(cudaMemcpyAsync(counters.p, z, sizeof z, cudaMemcpyHostToDevice, c.stream));
    ds_allow_smem(dist_rows_kernel<1>, ds_smem_bytes(list_cap));
    dist_rows_kernel<1><<<ds_grid(c, list_cap), DS_WARPS * 32, ds_smem_bytes(list_cap), c.stream>>>(
        g, nullptr, n, dmin, dmax, list_cap, nullptr, sh.dist_row_len.p, sh.dist_row_start.p, sh.dist_entries.p, nullptr, counters.p);
    ++c.counters.launches;
    if (n_queued) {
      PSI_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long), c.stream));   // the work claim only
      dist_rows_kernel<1><<<ds_big_grid(c), DS_WARPS * 32, 0, c.stream>>>(g, queue.p, (uint32_t)n_queued, dmin, dmax, DS_BIG_LIST, big.p,
                                                                         sh.dist_row_len.p, sh.dist_row_start.p, sh.dist_entries.p, nullptr, counters.p);
      ++c.counters.launches;
    }
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaEventRecord(e1, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (h_cnt[2] != total || h_cnt[3]) {
      sh.has_dindex = false;
      throw CudaError("create_distance_index: the two passes disagree on the number of entries");
    }
  }
  sh.dist_packed.ensure(n);
  dist_pack_rows_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(g, sh.dist_row_len.p, sh.dist_row_start.p, reinterpret_cast<uint4*>(sh.dist_packed.p));
  ++c.counters.launches;
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  sh.dist_row_len.release(); sh.dist_row_start.release();
  sh.dist_rows = true;
  sh.n_dist_entries = total;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) (void)cudaGetLastError();
  c.counters.ms_dindex_build = ms;
  c.counters.dindex_mode = 2;
  c.counters.n_dindex_entries = total;
  c.counters.dindex_bytes = sh.dist_entries.bytes() + sh.dist_packed.bytes();
}


// verify_distance for n queries {v rank, v offset, u rank, u offset}; out[i] = 1 when the loci comply with the window.
void engine_verify_distance(Ctx& c, uint64_t n, const uint32_t* pairs, uint8_t* out, bool on_device)
{
  Shared& sh = *c.sh;
  if (!sh.has_dindex) throw StateError("verify_distance: no distance index (psi_b200_create_distance_index first)");
  if (c.pending) throw StateError("verify_distance: a step is in flight on this context (call psi_b200_wait first)");
  if (n == 0) return;
  if (!pairs || !out) throw ArgError("verify_distance: null argument");
  if (n >= 0xffffffffull) throw ArgError("verify_distance: more than 2^32 - 1 queries in one call");
  if (on_device && (reinterpret_cast<uintptr_t>(pairs) & 15u)) throw ArgError("verify_distance: device pairs must be 16-byte aligned");
  PSI_CUDA(cudaSetDevice(c.device));
  nvtxRangePushA("query-dindex");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop_range;
  const GraphView g = make_graph_view(c);
  const DistQuery* d_q = reinterpret_cast<const DistQuery*>(pairs);
  uint8_t* d_out = out;
  if (!on_device) {
    c.dist_q.ensure(4 * n);
    c.dist_out.ensure(n);
    PSI_CUDA(cudaMemcpyAsync(c.dist_q.p, pairs, n * sizeof(DistQuery), cudaMemcpyHostToDevice, c.stream));
    d_q = reinterpret_cast<const DistQuery*>(c.dist_q.p);
    d_out = c.dist_out.p;
  }
  if (sh.dist_rows) {
    dist_query_rows_kernel<<<grid_for(n * 8, 256), 256, 0, c.stream>>>(d_q, n, sh.n_nodes, sh.dist_dmin, sh.dist_dmax,
                                                                       reinterpret_cast<const uint4*>(sh.dist_packed.p), sh.dist_entries.p, d_out);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
  }
  else {
    const uint32_t list_cap = c.opt_dindex_list_cap;
    c.dist_queue.ensure(n);
    c.dist_counters.ensure(4);
    unsigned long long h_cnt[4] = { 0, 0, 0, 0 };
    PSI_CUDA(cudaMemsetAsync(c.dist_counters.p, 0, 4 * sizeof(unsigned long long), c.stream));
    ds_allow_smem(dist_query_walk_kernel, ds_smem_bytes(list_cap));
    const unsigned grid = (unsigned)std::min<uint64_t>(ds_grid(c, list_cap), (n + DS_WARPS - 1) / DS_WARPS);
    dist_query_walk_kernel<<<grid, DS_WARPS * 32, ds_smem_bytes(list_cap), c.stream>>>(g, d_q, nullptr, n, sh.dist_dmin, sh.dist_dmax, list_cap,
                                                                                       nullptr, d_out, c.dist_queue.p, c.dist_counters.p);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(h_cnt, c.dist_counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (h_cnt[1]) {
      c.dist_big.ensure((size_t)ds_big_grid(c) * DS_WARPS * 3u * DS_BIG_LIST);
      PSI_CUDA(cudaMemsetAsync(c.dist_counters.p, 0, 4 * sizeof(unsigned long long), c.stream));
      dist_query_walk_kernel<<<ds_big_grid(c), DS_WARPS * 32, 0, c.stream>>>(g, d_q, c.dist_queue.p, h_cnt[1], sh.dist_dmin, sh.dist_dmax,
                                                                            DS_BIG_LIST, c.dist_big.p, d_out, nullptr, c.dist_counters.p);
      ++c.counters.launches;
      PSI_CUDA(cudaGetLastError());
      PSI_CUDA(cudaMemcpyAsync(h_cnt, c.dist_counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream));
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      if (h_cnt[3]) throw OverflowError("verify_distance: a query reaches more than 65 536 (node, distance) states inside the window");
    }
  }
  if (!on_device) PSI_CUDA(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace psi_b200
