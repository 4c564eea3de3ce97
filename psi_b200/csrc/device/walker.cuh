// walker.cuh -- warp-cooperative enumeration of all k-long forward walks of
// the graph that start at a set of loci.
//
// Stands behind TraverserBFS::{run,filter,compute,advance} of the reference
// (include/psi/traverser_bfs.hpp:71-161): a state walks the label of its node,
// 'N' kills it (:124), at the node end it fans out over EVERY out-edge (link
// type ignored, forward strand only, :146-160), a node without successors kills
// it (:140-143), and at depth k the walk is reported together with its START
// locus (:96-110).  The reference keeps the frontier in a std::vector of states
// per locus and descends a suffix tree of the read seeds base by base; here
//   - a warp owns a LIFO frontier of packed states in shared memory (spilling to
//     a private global region), pops up to 32 states per round (one per lane),
//     extends each to its node end with 2-bit word extraction, and pushes the
//     children of all lanes with one warp prefix sum over the out-degrees;
//   - starting loci are claimed 32 at a time from a global counter whenever the
//     frontier runs low, so lanes stay busy across loci with few walks;
//   - the match against the read seeds is one hash probe of the packed k-mer at
//     depth k (exact-match formulation) instead of k tree descents.
#ifndef PSI_B200_DEVICE_WALKER_CUH
#define PSI_B200_DEVICE_WALKER_CUH

#include "common.cuh"
#include "context.hpp"

namespace psi_b200 {
namespace dev {

struct alignas(8) WalkItem {
  uint64_t kmer;
  uint32_t origin;  // global position of the start locus
  uint32_t node;
  uint32_t off;
  uint32_t depth;
};

struct GraphView {
  const NodeRec* rec;
  const uint32_t* col;
  const uint64_t* seq2;
  const uint32_t* nmask;
  const uint32_t* pos2node;
  const Rank16* rank16;      // null: use pos2node
  const NodeRes* node_res;
  const uint64_t* node_id;   // the id reported in seed records, by rank
  uint32_t n_nodes;
  uint32_t pos2node_shift;
  uint64_t n_bases;
  uint32_t has_n;
  // locus codes of the path index (common.cuh): (id or rank) << code_off_bits | offset in the node
  uint32_t code_off_bits;
  uint32_t code_by_rank;
};

inline GraphView make_graph_view(const Ctx& c)
{
  GraphView g;
  g.rec = c.sh->node_rec.p;
  g.col = c.sh->col.p;
  g.seq2 = c.sh->seq2.p;
  g.nmask = c.sh->nmask.p;
  g.pos2node = c.sh->pos2node.p;
  g.rank16 = c.sh->has_rank16 ? c.sh->rank16 : nullptr;
  g.node_res = c.sh->node_res;
  g.node_id = c.sh->node_id.p;
  g.code_off_bits = c.sh->code_off_bits;
  g.code_by_rank = c.sh->code_by_rank ? 1u : 0u;
  g.n_nodes = c.sh->n_nodes;
  g.pos2node_shift = Ctx::POS2NODE_SHIFT;
  g.n_bases = c.sh->n_bases;
  g.has_n = c.sh->graph_has_n ? 1u : 0u;
  return g;
}

// node containing global position g
__device__ __forceinline__ uint32_t node_of_pos(const GraphView& g, uint32_t pos)
{
  uint32_t v = __ldg(g.pos2node + (pos >> g.pos2node_shift));
  while (__ldg(&g.rec[v + 1].seq_start) <= pos) ++v;
  return v;
}

// global position -> (node rank, offset in the node): two dependent 16-byte gathers (rank16, node_res), or the sampled
// pos2node table when the graph has zero-length nodes
__device__ __forceinline__ void rank_of_pos(const GraphView& g, uint32_t gpos, uint32_t& v, uint32_t& off)
{
  if (g.rank16) {
    const uint4 rw = __ldg(reinterpret_cast<const uint4*>(g.rank16 + (gpos >> 6)));
    const uint64_t bits = ((uint64_t)rw.y << 32) | rw.x;
    v = rw.z + (uint32_t)__popcll(bits & (~0ull >> (63u - (gpos & 63u)))) - 1u;
    off = gpos - __ldg(&g.node_res[v].seq_start);
  }
  else {
    v = node_of_pos(g, gpos);
    off = gpos - __ldg(&g.rec[v].seq_start);
  }
}

// locus code of a global position (build time, walkers, locus lists: never on the one-line probe path)
__device__ __forceinline__ uint64_t code_of_gpos(const GraphView& g, uint32_t gpos)
{
  uint32_t v, off;
  rank_of_pos(g, gpos, v, off);
  const uint64_t hi = g.code_by_rank ? (uint64_t)v : __ldg(g.node_id + v);
  return (hi << g.code_off_bits) | off;
}

// locus code -> the graph half of a seed record; by-id codes need no memory access at all
__device__ __forceinline__ void decode_code(const GraphView& g, uint64_t code, uint64_t& id, uint64_t& off)
{
  off = code & low_mask64(g.code_off_bits);
  const uint64_t hi = code >> g.code_off_bits;
  id = g.code_by_rank ? __ldg(g.node_id + hi) : hi;
}

// membership of (kmer, gpos) among the ON-PATH entries of the index.
// multi list layout: [n_on, n_total, on-path loci (sorted)..., off-path loci (sorted)...]
__device__ __forceinline__ bool index_contains(const GraphView& g, const KmerTable& t, const uint32_t* __restrict__ multi,
                                               uint64_t kmer, uint32_t gpos)
{
  Found f;
  if (!table_find_any(t, kmer, f)) return false;
  if (!(f.flags & FLAG_MULTI)) return !(f.flags & FLAG_OFF) && f.payload == code_of_gpos(g, gpos);
  const uint32_t cnt = __ldg(multi + f.payload);
  uint32_t lo = 0, hi = cnt;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t v = __ldg(multi + f.payload + 2 + mid);
    if (v == gpos) return true;
    if (v < gpos) lo = mid + 1; else hi = mid;
  }
  return false;
}

constexpr int WALK_WARPS = 4;          // warps per CTA
constexpr int WALK_SMEM_ITEMS = 96;    // frontier slots per warp in shared memory

struct WalkStack {
  WalkItem* smem;      // this warp's shared slots
  WalkItem* spill;     // this warp's global slots
  uint32_t spill_items;
  __device__ __forceinline__ WalkItem& at(uint32_t i) const
  {
    return i < (uint32_t)WALK_SMEM_ITEMS ? smem[i] : spill[i - WALK_SMEM_ITEMS];
  }
};

// Source: yields the initial state of work item `idx` (false = nothing to do).
// Sink:   skip(origin) -> drop a state early; prune(kmer, d0, d1) -> drop a state whose depth just went from d0 to d1 < k
//         (the reference prunes a walk as soon as its prefix is no prefix of a read seed, traverser_bfs.hpp:124-131);
//         complete(kmer, origin) at depth k.
template <class Source, class Sink>
__device__ void walk_all(const GraphView& g, uint32_t k, uint64_t n_items, unsigned long long* work_counter,
                         WalkItem* smem_all, WalkItem* spill_all, uint32_t spill_items,
                         unsigned long long* err_flag, Source& source, Sink& sink)
{
  const uint32_t lane = lane_id();
  const uint32_t warp_in_cta = threadIdx.x >> 5;
  const uint64_t warp_global = (uint64_t)blockIdx.x * WALK_WARPS + warp_in_cta;
  WalkStack st{ smem_all + warp_in_cta * WALK_SMEM_ITEMS, spill_all + warp_global * spill_items, spill_items };
  const uint32_t capacity = WALK_SMEM_ITEMS + spill_items;

  uint32_t top = 0;
  bool more = true;
  while (true) {
    // ---- refill from the work queue ----
    if (top < 32 && more) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(work_counter, 32ull);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= n_items) more = false;
      else {
        WalkItem it;
        const uint64_t idx = base + lane;
        const bool ok = idx < n_items && source.init(idx, it);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok) st.at(top + __popc(m & ((1u << lane) - 1u))) = it;
        top += __popc(m);
        __syncwarp();
        if (top < 32) continue;  // try to fill a whole round first
      }
    }
    if (top == 0) break;

    // ---- pop one state per lane ----
    const uint32_t n = top < 32 ? top : 32;
    bool alive = lane < n;
    WalkItem it;
    if (alive) it = st.at(top - 1 - lane);
    top -= n;
    __syncwarp();

    uint32_t n_children = 0;
    NodeRec r{ 0, 0, 0, 0 };
    if (alive && sink.skip(it.origin)) alive = false;
    if (alive) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(g.rec + it.node));
      r.seq_start = raw.x; r.seq_len = raw.y; r.edge_start = raw.z; r.outdeg = raw.w;
      const uint32_t avail = r.seq_len - it.off;
      const uint32_t want = k - it.depth;
      const uint32_t take = avail < want ? avail : want;
      if (take) {
        const uint64_t p = (uint64_t)r.seq_start + it.off;
        if (g.has_n && extract_nmask(g.nmask, p, take)) alive = false;
        else {
          it.kmer |= extract_bases(g.seq2, p, take) << (2u * it.depth);
          it.depth += take;
          if (it.depth < k && sink.prune(it.kmer, it.depth - take, it.depth)) alive = false;
        }
      }
      if (alive) {
        if (it.depth == k) { sink.complete(it.kmer, it.origin); alive = false; }
        else n_children = r.outdeg;  // node end reached; 0 successors kills the state
      }
    }

    // ---- push the children of all lanes ----
    uint32_t incl = n_children;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
      if (top + total > capacity) {
        if (lane == 0) atomicOr(err_flag, 1ull);
        break;  // result is invalid; the host retries with a larger spill area
      }
      uint32_t dst = top + incl - n_children;
      for (uint32_t c = 0; c < n_children; ++c) {
        WalkItem ch;
        ch.kmer = it.kmer;
        ch.origin = it.origin;
        ch.node = __ldg(g.col + r.edge_start + c);
        ch.off = 0;
        ch.depth = it.depth;
        st.at(dst + c) = ch;
      }
      top += total;
    }
    __syncwarp();
  }
  sink.finish();
}

// Work source of the walkers: one item per starting locus.
struct LociSource {
  GraphView g;
  const uint32_t* node;
  const uint32_t* off;
  __device__ bool init(uint64_t idx, WalkItem& it) const
  {
    const uint32_t v = __ldg(node + idx), o = __ldg(off + idx);
    if (v >= g.n_nodes) return false;
    const NodeRec r = g.rec[v];
    if (o >= r.seq_len) return false;
    it.kmer = 0; it.origin = r.seq_start + o; it.node = v; it.off = o; it.depth = 0;
    return true;
  }
};

}  // namespace dev
}  // namespace psi_b200
#endif
