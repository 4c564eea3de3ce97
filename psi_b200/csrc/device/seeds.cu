// seeds.cu -- the query kernels: seeds_on_paths (K3), seeds_off_paths (K5) and
// the resolve step that turns compact hits into the reference's seed records.
#include "engine.hpp"
#include "walker.cuh"

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

namespace psi_b200 {

using namespace dev;

void engine_index_chunk(Ctx& c);

// ================================================================== K3 ==
//
// Stands behind SeedFinder::seeds_on_paths -> kmer_exact_matches +
// _add_occurrences (reference include/psi/seed_finder.hpp:1426-1457,
// index_iter.hpp:808-852,728-746): every k-mer present in both the path index
// and the chunk's seeds yields |path occurrences| x |seed occurrences| hits.
// The reference co-traverses an FM-index and a suffix tree in lexicographic
// order (2 wavelet-tree ranks + one tree descent per character, then <= 32 LF
// steps per located occurrence); here each seed is ONE probe:
//   thread = ITEMS seeds; packed k-mer (coalesced 8 B) -> home bucket ->
//   one 32-byte sector load (LDG.256, all ITEMS loads in flight before the
//   first compare) -> tag compare -> warp-aggregated reservation of output
//   slots (one atomic per warp for 32 x ITEMS probes) -> 8-byte compact hits.
// Because the index holds DISTINCT (k-mer, locus) pairs and a seed index is
// unique, the hits of this kernel are already a set.

template <int FMT>
__device__ __forceinline__ int eval_sector(const uint64_t (&v)[4], const Home& h, uint64_t kmer,
                                           uint32_t& payload, bool& multi)
{
  // 1 = found, 0 = definitely absent (bucket has a free slot), -1 = bucket full, look further
  bool has_empty = false;
  if (FMT == 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (v[j] == EMPTY8) has_empty = true;
      else if ((v[j] >> 33) == h.tag) { payload = (uint32_t)v[j]; multi = (v[j] >> 32) & 1u; return 1; }
    }
  }
  else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t fl = (uint32_t)(v[2 * j + 1] >> 32);
      if (fl == NIL32) has_empty = true;
      else if (v[2 * j] == kmer) { payload = (uint32_t)v[2 * j + 1]; multi = fl == 1; return 1; }
    }
  }
  return has_empty ? 0 : -1;
}

template <int FMT, int ITEMS, bool COUNT_SECTORS>
__global__ void __launch_bounds__(256)
seeds_on_paths_kernel(KmerTable t, const uint32_t* __restrict__ multi,
                      const uint64_t* __restrict__ seed_kmer, const uint32_t* __restrict__ seed_valid,
                      const unsigned long long* __restrict__ n_seeds_p,
                      Hit* __restrict__ hits, uint64_t hits_cap, unsigned long long* __restrict__ hit_count,
                      unsigned long long* __restrict__ sector_count)
{
  const uint32_t n_seeds = (uint32_t)*n_seeds_p;
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
  if (blockIdx.x * (256u * ITEMS) >= n_seeds) return;   // whole CTA out of range

  uint64_t km[ITEMS];
  bool ok[ITEMS];
  Home hm[ITEMS];
  uint64_t v[ITEMS][4];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    ok[i] = s < n_seeds && ((__ldg(seed_valid + (s >> 5)) >> (s & 31u)) & 1u);
    km[i] = ok[i] ? __ldg(seed_kmer + s) : 0;
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (ok[i]) {
      hm[i] = home_of<FMT>(t, km[i]);
      ld_sector_nc((const char*)t.slots + hm[i].line * 128u + (hm[i].sec << 5), v[i]);
    }
  }
  uint32_t payload[ITEMS];
  uint32_t count[ITEMS];
  bool is_multi[ITEMS];
  uint32_t total = 0, sectors = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    count[i] = 0;
    is_multi[i] = false;
    payload[i] = 0;
    if (ok[i]) {
      ++sectors;
      int st = eval_sector<FMT>(v[i], hm[i], km[i], payload[i], is_multi[i]);
      if (st < 0) {  // home bucket full: the other three buckets of the line, then the stash
        bool found = false;
        const char* line = (const char*)t.slots + hm[i].line * 128u;
        for (uint32_t j = 1; j < 4 && st < 0; ++j) {
          uint64_t w[4];
          ld_sector_nc(line + (((hm[i].sec + j) & 3u) << 5), w);
          ++sectors;
          st = eval_sector<FMT>(w, hm[i], km[i], payload[i], is_multi[i]);
        }
        if (st < 0) found = stash_find(t, km[i], payload[i], is_multi[i]);
        else found = st == 1;
        st = found ? 1 : 0;
      }
      if (st == 1) count[i] = is_multi[i] ? __ldg(multi + payload[i]) : 1u;
    }
    total += count[i];
  }
  // one reservation per warp for all its probes
  uint64_t slot = warp_reserve(hit_count, total);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (!count[i]) continue;
    const uint32_t s = base + i * 256u;
    if (!is_multi[i]) {
      if (slot < hits_cap) hits[slot] = Hit{ s, payload[i] };
      ++slot;
    }
    else {
      for (uint32_t j = 0; j < count[i]; ++j, ++slot)
        if (slot < hits_cap) hits[slot] = Hit{ s, __ldg(multi + payload[i] + 1 + j) };
    }
  }
  if (COUNT_SECTORS) {
#pragma unroll
    for (int d = 16; d; d >>= 1) sectors += __shfl_xor_sync(0xffffffffu, sectors, d);
    if (lane_id() == 0) atomicAdd(sector_count, (unsigned long long)sectors);
  }
}

// ================================================================== K5 ==
//
// Stands behind SeedFinder::seeds_off_paths + TraverserBFS (reference
// seed_finder.hpp:1703-1722, traverser_bfs.hpp:71-161); see walker.cuh.  At
// depth k the packed k-mer is probed in the chunk's read index; every seed on
// the chain is a hit at the START locus.  Two filters keep the output a set:
//   - (k-mer, locus) already in the path index -> reported by K3, skip;
//   - several walks from one locus spelling the same k-mer -> the first one to
//     claim (chain head, locus) in a device hash set reports, the others skip.

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

struct ReadIndexSink {
  KmerTable rt;                 // chunk read index
  const uint32_t* next;         // seed chains
  KmerTable pt;                 // path index
  const uint32_t* multi;
  uint32_t has_index;
  unsigned long long* dedup;    // hash set of (chain head << 32 | locus)
  uint64_t dedup_mask;
  Hit* hits;
  uint64_t hits_cap;
  unsigned long long* hit_count;
  unsigned long long* walk_count;
  unsigned long long* err;
  uint32_t walks;

  __device__ bool skip(uint32_t) const { return false; }

  __device__ bool claim(uint64_t key)
  {
    uint64_t p = mix64(key) & dedup_mask;
    for (int i = 0; i < 128; ++i) {
      const unsigned long long old = atomicCAS(dedup + p, EMPTY8, (unsigned long long)key);
      if (old == EMPTY8) return true;
      if (old == key) return false;
      p = (p + 1) & dedup_mask;
    }
    atomicOr(err, 4ull);  // set too full: host grows it and retries
    return true;
  }

  __device__ void complete(uint64_t kmer, uint32_t origin)
  {
    ++walks;
    uint32_t head;
    bool m;
    const bool found = rt.fmt == 8 ? table_find<8>(rt, kmer, head, m) : table_find<16>(rt, kmer, head, m);
    if (!found) return;
    if (has_index && index_contains(pt, multi, kmer, origin)) return;
    if (!claim(((uint64_t)head << 32) | origin)) return;
    uint32_t n = 0;
    for (uint32_t s = head; s != NIL32; s = __ldg(next + s)) ++n;
    uint64_t slot = atomicAdd(hit_count, (unsigned long long)n);
    for (uint32_t s = head; s != NIL32; s = __ldg(next + s), ++slot)
      if (slot < hits_cap) hits[slot] = Hit{ s, origin };
  }

  __device__ void finish()
  {
    unsigned long long w = walks;
#pragma unroll
    for (int d = 16; d; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (lane_id() == 0 && w) atomicAdd(walk_count, w);
  }
};

__global__ void __launch_bounds__(WALK_WARPS * 32)
seeds_off_paths_kernel(GraphView g, uint32_t k, uint64_t n_loci, const uint32_t* loci_node, const uint32_t* loci_off,
                       ReadIndexSink sink, unsigned long long* work, WalkItem* spill, uint32_t spill_items)
{
  __shared__ WalkItem smem[WALK_WARPS * WALK_SMEM_ITEMS];
  LociSource src{ g, loci_node, loci_off };
  sink.walks = 0;
  walk_all(g, k, n_loci, work, smem, spill, spill_items, sink.err, src, sink);
}

// ============================================================= resolve ==
//
// Compact hit (seed, global position) -> the reference's output record
// (seed.hpp:32-46 as written by src/psikt.cpp:172-181): node_id, node_offset,
// read_id, read_offset, 4 x u64.  read_id / read_offset replace
// Records::position_to_id/offset (sequence.hpp:1201-1213,1277-1289); the node
// lookup replaces position_to_id/offset(PathIndex) (pathindex.hpp:378-416).
__global__ void __launch_bounds__(256)
resolve_hits_kernel(GraphView g, const uint64_t* __restrict__ node_id, const Hit* __restrict__ hits, uint64_t n_hits,
                    const uint32_t* __restrict__ seed_read, const uint32_t* __restrict__ seed_first,
                    uint32_t d, uint64_t first_read_id, uint64_t* __restrict__ records)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hits) return;
  const Hit h = hits[i];
  const uint32_t r = __ldg(seed_read + h.seed);
  const uint64_t read_off = (uint64_t)(h.seed - __ldg(seed_first + r)) * d;
  const uint32_t v = node_of_pos(g, h.gpos);
  const uint64_t node_off = h.gpos - __ldg(&g.rec[v].seq_start);
  ulonglong2* out = reinterpret_cast<ulonglong2*>(records + 4 * i);
  out[0] = make_ulonglong2(__ldg(node_id + v), node_off);
  out[1] = make_ulonglong2(first_read_id + r, read_off);
}

// ================================================================ host ==

static uint64_t next_pow2(uint64_t x)
{
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

void engine_seeds(Ctx& c, unsigned flags)
{
  if (!c.has_chunk) throw StateError("seeds_all: no read chunk submitted");
  PSI_CUDA(cudaSetDevice(c.device));
  const bool do_on = (flags & PSI_B200_ON_PATHS) && c.sh->has_index;
  const bool do_off = (flags & PSI_B200_OFF_PATHS) && c.sh->n_loci > 0;
  const GraphView g = make_graph_view(c);
  unsigned long long* dc = c.dev_counters.p;
  c.records_valid = false;
  c.ev_state[T_ON] = c.ev_state[T_OFF] = c.ev_state[T_RESOLVE] = c.ev_state[T_SORT] = c.ev_state[T_D2H] = 0;

  if (c.hits.cap == 0) c.hits.ensure(std::max<uint64_t>(2 * c.n_seeds_cap, 1u << 20));
  if (c.dedup.cap == 0) c.dedup.ensure(1u << 20);
  uint64_t n_on = 0, n_total = 0, n_walks = 0, n_sectors = 0, n_seeds = 0;

  for (int attempt = 0;; ++attempt) {
    if (attempt > 12) throw OverflowError("seeds_all: device buffers keep overflowing");
    PSI_CUDA(cudaMemsetAsync(dc + DC_HITS, 0, 5 * sizeof(unsigned long long), c.stream));  // HITS..WORK
    const uint64_t hits_cap = c.hits.cap;

    PhaseTimer t_on(c, T_ON);
    if (do_on) {
      constexpr int ITEMS = 4;
      const unsigned grid = grid_for(c.n_seeds_cap, 256, ITEMS);
      if (c.sh->index.view.fmt == 8)
        seeds_on_paths_kernel<8, ITEMS, false><<<grid, 256, 0, c.stream>>>(c.sh->index.view, c.sh->multi.p, c.seed_kmer.p, c.seed_valid.p,
                                                                         dc + DC_SEEDS, c.hits.p, hits_cap, dc + DC_HITS, dc + DC_SECTORS);
      else
        seeds_on_paths_kernel<16, ITEMS, false><<<grid, 256, 0, c.stream>>>(c.sh->index.view, c.sh->multi.p, c.seed_kmer.p, c.seed_valid.p,
                                                                          dc + DC_SEEDS, c.hits.p, hits_cap, dc + DC_HITS, dc + DC_SECTORS);
      ++c.counters.launches;
    }
    t_on.stop();
    PSI_CUDA(cudaMemcpyAsync(c.h_pinned + 0, dc + DC_HITS, sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));

    if (do_off) {
      engine_index_chunk(c);
      PhaseTimer t_off(c, T_OFF);
      const uint64_t dedup_slots = next_pow2(c.dedup.cap) == c.dedup.cap ? c.dedup.cap : next_pow2(c.dedup.cap) / 2;
      PSI_CUDA(cudaMemsetAsync(c.dedup.p, 0xff, dedup_slots * sizeof(unsigned long long), c.stream));
      const unsigned grid = (unsigned)c.sm_count * 8;
      c.walk_spill.ensure((size_t)grid * WALK_WARPS * c.spill_items * sizeof(WalkItem));
      ReadIndexSink sink;
      sink.rt = c.read_index.view;
      sink.rt.stash_nonempty = 1;  // not known without a sync; probing an empty stash costs one load, and only for full lines
      sink.next = c.seed_next.p;
      sink.pt = c.sh->index.view;
      sink.multi = c.sh->multi.p;
      sink.has_index = c.sh->has_index ? 1u : 0u;
      sink.dedup = c.dedup.p;
      sink.dedup_mask = dedup_slots - 1;
      sink.hits = c.hits.p;
      sink.hits_cap = hits_cap;
      sink.hit_count = dc + DC_HITS;
      sink.walk_count = dc + DC_WALKS;
      sink.err = dc + DC_ERR;
      sink.walks = 0;
      seeds_off_paths_kernel<<<grid, WALK_WARPS * 32, 0, c.stream>>>(g, c.k, c.sh->n_loci, c.sh->loci_node.p, c.sh->loci_off.p, sink,
                                                                     dc + DC_WORK, (WalkItem*)c.walk_spill.p, c.spill_items);
      ++c.counters.launches;
      t_off.stop();
    }
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(c.h_pinned + 1, dc, DC_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    n_on = c.h_pinned[0];
    n_total = c.h_pinned[1 + DC_HITS];
    n_walks = c.h_pinned[1 + DC_WALKS];
    n_sectors = c.h_pinned[1 + DC_SECTORS];
    n_seeds = c.h_pinned[1 + DC_SEEDS];
    const uint64_t err = c.h_pinned[1 + DC_ERR];
    bool retry = false;
    if (n_total > hits_cap) { c.hits.ensure(n_total, 1.25); retry = true; }
    if (err & 1ull) {
      if (c.spill_items >= (1u << 20)) throw OverflowError("seeds_off_paths: walk frontier exceeds 2^20 states per warp");
      c.spill_items *= 4;
      retry = true;
    }
    if (err & 2ull) throw OverflowError("read index: hash stash exhausted");
    if (err & 4ull) { c.dedup.ensure(c.dedup.cap * 4); retry = true; }
    if (!retry) break;
  }

  c.n_hits = n_total;
  c.counters.n_seeds = n_seeds;
  c.counters.n_hits_on = n_on;
  c.counters.n_hits_off = n_total - n_on;
  c.counters.n_hits = n_total;
  c.counters.n_walks = n_walks;
  c.counters.n_on_probe_sectors = n_sectors;

  if ((flags & PSI_B200_SORTED) && n_total > 1) {
    PhaseTimer t_sort(c, T_SORT);
    // canonical order: (seed index, global position) ascending == (read_id, read_offset, node rank, node_offset)
    DevBuf<unsigned long long> tmp_keys;
    tmp_keys.ensure(n_total);
    size_t tmp = 0;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(c.hits.p);
    // Hit{seed, gpos} is little-endian (seed low, gpos high): sort by gpos bits first, then seed bits
    cub::DoubleBuffer<unsigned long long> db(keys, tmp_keys.p);
    PSI_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, db, (int64_t)n_total, 0, 64, c.stream));
    c.scan_tmp.ensure(tmp);
    // 64-bit value = gpos << 32 | seed; canonical order needs seed major: swap halves via two stable passes
    PSI_CUDA(cub::DeviceRadixSort::SortKeys(c.scan_tmp.p, tmp, db, (int64_t)n_total, 32, 64, c.stream));
    PSI_CUDA(cub::DeviceRadixSort::SortKeys(c.scan_tmp.p, tmp, db, (int64_t)n_total, 0, 32, c.stream));
    if (db.Current() != keys)
      PSI_CUDA(cudaMemcpyAsync(keys, db.Current(), n_total * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream));
    t_sort.stop();
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    c.counters.launches += 8;
  }

  if (!(flags & PSI_B200_NO_RESOLVE)) {
    PhaseTimer t_res(c, T_RESOLVE);
    c.records.ensure(4 * std::max<uint64_t>(n_total, 1), 1.25);
    if (n_total) {
      resolve_hits_kernel<<<grid_for(n_total, 256), 256, 0, c.stream>>>(g, c.sh->node_id.p, c.hits.p, n_total, c.seed_read.p,
                                                                      c.seed_first.p, c.distance, c.first_read_id, c.records.p);
      ++c.counters.launches;
    }
    t_res.stop();
    PSI_CUDA(cudaGetLastError());
    c.records_valid = true;
  }
}

void engine_fetch(Ctx& c, uint64_t* hits, uint64_t cap)
{
  if (!c.records_valid) throw StateError("fetch: no resolved seed records (call seeds_all without NO_RESOLVE first)");
  PSI_CUDA(cudaSetDevice(c.device));
  const uint64_t n = c.n_hits < cap ? c.n_hits : cap;
  PhaseTimer t(c, T_D2H);
  if (n) PSI_CUDA(cudaMemcpyAsync(hits, c.records.p, n * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
  t.stop();
  PSI_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace psi_b200
