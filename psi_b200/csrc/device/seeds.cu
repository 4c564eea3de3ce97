// seeds.cu -- the query kernels: seeds_on_paths (K3), seeds_off_paths (K5) and
// the resolve step that turns compact hits into the reference's seed records.
#include "engine.hpp"
#include "walker.cuh"
#include "records.cuh"

#include <algorithm>
#include <atomic>

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
// steps per located occurrence); here each seed is ONE probe of ONE 128-byte
// bucket line (the DRAM access unit, profiles/r01c_gather_peak.md):
//   packed k-mer (coalesced 8 B) -> home line -> the whole line is copied into shared
//   memory asynchronously (cp.async, 8 lanes x 16 B) -> 16 tag compares -> the seed's
//   result: seed_hit[s] = locus, seed_kind[s] = 1 (on an indexed path) / 2 (only
//   reachable by an off-path walk) / 0 (none).
// No atomics, no compaction here: the results are per seed, in seed order; the
// compaction happens in compact_resolve_kernel.  The rare seeds that one line
// cannot settle (locus lists; home line full and key displaced) are queued for
// seeds_slow_kernel.  When the off-path walks are materialised in the index
// (offpath_mode 2) the same probe also answers seeds_off_paths.
// Because the index holds DISTINCT (k-mer, locus) pairs and a seed index is
// unique, the hits are already a set.

// One thread per seed; a CTA handles 256 consecutive seeds:
//   1. coalesced loads of k-mer + validity, hash -> home line and tag (registers);
//   2. the warp copies its 32 home lines into shared memory with cp.async (LDGSTS): 8 lanes x 16 bytes
//      per line, 8 instructions per lane, the line index of each seed comes by shuffle from its owner.
//      No data registers are tied up while the lines are in flight, so ~6 CTAs x 256 lines = 1500
//      lines per SM are outstanding -- what the random-line rate needs (profiles/r01j_gather_peak.md);
//   3. every thread scans ITS seed's line in shared memory (8 x LDS.128, chunk order rotated by the lane
//      so that a warp's loads do not collide on banks), compares the 16 tags and stores the seed's result:
//      seed_hit[s], seed_kind[s] -- fully coalesced stores, no atomics, no ballots.
template <int FMT>
__global__ void __launch_bounds__(256)
seeds_on_paths_kernel(KmerTable t, const uint64_t* __restrict__ seed_kmer, const uint8_t* __restrict__ seed_valid,
                      const unsigned long long* __restrict__ n_seeds_p, uint32_t mode,
                      uint64_t* __restrict__ seed_hit, uint8_t* __restrict__ seed_kind,
                      uint32_t* __restrict__ slow_queue, unsigned long long* __restrict__ slow_count)
{
  __shared__ __align__(128) unsigned char s_lines[256 * 128];
  const uint32_t n_seeds = (uint32_t)*n_seeds_p;
  const uint32_t block_base = blockIdx.x * 256u;
  if (block_base >= n_seeds) return;   // whole CTA out of range
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t s = block_base + threadIdx.x;
  const bool ok = s < n_seeds && __ldg(seed_valid + s) != 0;
  const uint64_t kmer = ok ? __ldg(seed_kmer + s) : 0;
  const Home h = home_of<FMT>(t, kmer);
  const uint32_t my_line = (uint32_t)h.line;           // line_bits <= 32 (checked when the table is allocated)

  unsigned char* warp_lines = s_lines + ((size_t)warp << 12);   // 32 lines of 128 bytes
  const uint32_t sub = lane & 7u;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t owner = it * 4u + (lane >> 3);
    const uint32_t line = __shfl_sync(0xffffffffu, my_line, owner);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(warp_lines + (owner << 7) + (sub << 4));
    const char* src = (const char*)t.slots + ((uint64_t)line << 7) + (sub << 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const uint4* ln = reinterpret_cast<const uint4*>(warp_lines + (lane << 7));
  bool hit = false, empty = false;
  uint32_t pl = 0, ph = 0;                             // the matching slot's payload word and high word
  if (FMT == 8) {
    const uint32_t want = (uint32_t)h.tag;             // tag | displacement 0 (<= 30 bits)
    const uint32_t tsh = 2u + t.pay_hi;                // high word: [tag | disp][2 flag bits][pay_hi payload bits]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 w = ln[(j + lane) & 7u];             // two slots: (w.x, w.y) and (w.z, w.w), high word second
      empty |= (w.y == 0xffffffffu) | (w.w == 0xffffffffu);   // no valid entry has both flag bits set
      if ((w.y >> tsh) == want) { hit = true; pl = w.x; ph = w.y; }
      if ((w.w >> tsh) == want) { hit = true; pl = w.z; ph = w.w; }
    }
  }
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 w = ln[(j + lane) & 7u];             // one slot: key (w.x, w.y), payload w.z, flags + payload high bits w.w
      empty |= w.w == NIL32;
      if (w.w != NIL32 && (((uint64_t)w.y << 32) | w.x) == kmer) { hit = true; pl = w.z; ph = w.w; }
    }
  }
  const uint32_t fl = FMT == 8 ? (ph >> t.pay_hi) & 3u : ph & 3u;
  const uint64_t payload = FMT == 8 ? ((uint64_t)(ph & ((1u << t.pay_hi) - 1u)) << 32) | pl : slot16_payload(pl, ph);
  if (s >= n_seeds) return;
  uint8_t kind = 0;
  if (ok) {
    if (hit) {
      kind = kind_of(fl, mode);
      if (fl & FLAG_MULTI) { kind = 3; slow_queue[atomicAdd(slow_count, 1ull)] = s; }
      seed_hit[s] = payload;
    }
    else if (!empty) {
      // the line is full and does not hold the key: it may sit in a following line (~1 % of the lines are full)
      kind = 3;
      slow_queue[atomicAdd(slow_count, 1ull)] = s;
    }
  }
  seed_kind[s] = kind;
}

// The seeds the probe kernel could not settle from one line: full search (following lines, stash) and
// locus lists, whose entries go to the overflow hit list.  One thread per queued seed.
__global__ void __launch_bounds__(256)
seeds_slow_kernel(KmerTable t, const uint32_t* __restrict__ multi, const uint64_t* __restrict__ seed_kmer,
                  const uint32_t* __restrict__ slow_queue, const unsigned long long* __restrict__ slow_count, uint32_t mode,
                  uint64_t* __restrict__ seed_hit, uint8_t* __restrict__ seed_kind,
                  Hit* __restrict__ ovf, uint8_t* __restrict__ ovf_kind, uint64_t ovf_cap, unsigned long long* __restrict__ ovf_count)
{
  const uint64_t n = *slow_count;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = slow_queue[q];
    Found f;
    uint8_t kind = 0;
    if (table_find_any(t, __ldg(seed_kmer + s), f)) {
      if (!(f.flags & FLAG_MULTI)) {
        kind = kind_of(f.flags, mode);
        seed_hit[s] = f.payload;
      }
      else {
        const uint32_t n_on = __ldg(multi + f.payload), n_all = __ldg(multi + f.payload + 1);
        const uint32_t from = (mode & PSI_B200_ON_PATHS) ? 0u : n_on;
        const uint32_t to = (mode & PSI_B200_OFF_PATHS) ? n_all : n_on;
        if (to > from) {
          uint64_t slot = atomicAdd(ovf_count, (unsigned long long)(to - from));
          for (uint32_t j = from; j < to; ++j, ++slot)
            if (slot < ovf_cap) { ovf[slot] = Hit{ s, __ldg(multi + f.payload + 2 + j) }; ovf_kind[slot] = j < n_on ? 1 : 2; }
        }
      }
    }
    seed_kind[s] = kind;
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

struct ReadIndexSink {
  GraphView g;
  const uint32_t* pfx_bits;     // bitmap of the first `pfx` bases of every read seed (2 pfx bits of index)
  uint32_t pfx;                 // 0: no prefix filter
  KmerTable rt;                 // chunk read index
  const uint32_t* next;         // seed chains
  KmerTable pt;                 // path index
  const uint32_t* multi;
  uint32_t has_index;
  unsigned long long* dedup;    // hash set of (chain head << 32 | locus)
  uint64_t dedup_mask;
  Hit* hits;                    // overflow hit list (shared with seeds_slow_kernel)
  uint8_t* hit_kind;
  uint64_t hits_cap;
  unsigned long long* hit_count;
  unsigned long long* walk_count;
  unsigned long long* err;
  uint32_t walks;

  __device__ bool skip(uint32_t) const { return false; }

  // The reference's traverser stops a walk as soon as its prefix is no prefix of a read seed (it descends the seeds'
  // suffix tree base by base, traverser_bfs.hpp:124-131).  Same pruning, one bit test: when a walk's depth crosses
  // `pfx`, its first pfx bases must be the prefix of some seed of the chunk.
  __device__ bool prune(uint64_t kmer, uint32_t d0, uint32_t d1) const
  {
    if (d0 >= pfx || d1 < pfx) return false;
    const uint32_t i = (uint32_t)(kmer & low_mask64(2u * pfx));
    return !((__ldg(pfx_bits + (i >> 5)) >> (i & 31u)) & 1u);
  }

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
    Found f;
    if (!table_find_any(rt, kmer, f)) return;
    const uint32_t head = (uint32_t)f.payload;
    if (has_index && index_contains(g, pt, multi, kmer, origin)) return;
    if (!claim(((uint64_t)head << 32) | origin)) return;
    uint32_t n = 0;
    for (uint32_t s = head; s != NIL32; s = __ldg(next + s)) ++n;
    uint64_t slot = atomicAdd(hit_count, (unsigned long long)n);
    for (uint32_t s = head; s != NIL32; s = __ldg(next + s), ++slot)
      if (slot < hits_cap) { hits[slot] = Hit{ s, origin }; hit_kind[slot] = 2; }
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
// Per-seed results + overflow list -> the dense output: the reference's records
// (seed.hpp:32-46 as written by src/psikt.cpp:172-181): node_id, node_offset,
// read_id, read_offset, 4 x u64.  read_id / read_offset replace
// Records::position_to_id/offset (sequence.hpp:1201-1213,1277-1289); the node
// half comes out of the locus code the probe found (no memory access for by-id
// codes), or -- overflow entries: locus lists, walker hits -- from the global
// position through the rank structure (pathindex.hpp:378-416).
// A CTA takes 512 consecutive items (seeds first, then overflow entries), counts
// its hits with ballots, reserves their output range with ONE atomic and writes
// them in item order, so the output of a chunk is ordered by (read, offset) up
// to the CTA granularity and the stores of a warp are contiguous.
// RECORDS: 0 = (locus code, seed) pairs only, 1 = 4 x u64 records, 2 = 4 x u32 records
template <int RECORDS, int ITEMS, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
compact_resolve_kernel(GraphView g,
                       const uint64_t* __restrict__ seed_hit, const uint8_t* __restrict__ seed_kind,
                       const unsigned long long* __restrict__ n_seeds_p,
                       const Hit* __restrict__ ovf, const uint8_t* __restrict__ ovf_kind,
                       const unsigned long long* __restrict__ n_ovf_p, uint64_t ovf_cap,
                       const uint32_t* __restrict__ seed_read, const uint32_t* __restrict__ seed_first,
                       uint32_t d, uint64_t first_read_id,
                       uint64_t* __restrict__ records, uint8_t* __restrict__ rec_kind,
                       uint64_t* __restrict__ out_code, uint32_t* __restrict__ out_seed, uint64_t cap,
                       unsigned long long* __restrict__ total, unsigned long long* __restrict__ total_on)
{
  __shared__ uint32_t s_cnt[8 * ITEMS];
  __shared__ uint32_t s_on[8];
  __shared__ unsigned long long s_base;
  const uint64_t n_seeds = *n_seeds_p;
  uint64_t n_ovf = *n_ovf_p;
  if (n_ovf > ovf_cap) n_ovf = ovf_cap;
  const uint64_t n_items = n_seeds + n_ovf;
  const uint64_t base = (uint64_t)blockIdx.x * (256u * ITEMS);
  if (base >= n_items) return;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;

  // all independent loads first: kind and locus of every item (the locus is loaded whether or not it is a hit)
  uint32_t seed[ITEMS];
  uint64_t code[ITEMS];     // per-seed items: the locus code; overflow items: the global position
  uint8_t kind[ITEMS];
  bool is_ovf[ITEMS];
#pragma unroll
  for (int h = 0; h < ITEMS; ++h) {
    const uint64_t i = base + h * 256u + threadIdx.x;
    kind[h] = 0; seed[h] = 0; code[h] = 0; is_ovf[h] = false;
    if (i < n_seeds) {
      kind[h] = __ldg(seed_kind + i);
      code[h] = __ldg(seed_hit + i);
      seed[h] = (uint32_t)i;
    }
    else if (i < n_items) {
      const Hit x = ovf[i - n_seeds];
      kind[h] = ovf_kind[i - n_seeds];
      seed[h] = x.seed;
      code[h] = x.gpos;
      is_ovf[h] = true;
    }
  }
#pragma unroll
  for (int h = 0; h < ITEMS; ++h) if (kind[h] > 2) kind[h] = 0;

  // count, and let one thread reserve the CTA's output range (two atomics per CTA: same-address atomics serialise
  // in L2, so their number, not their latency, is what matters); the round trip overlaps the gathers below
  uint32_t m[ITEMS];
  uint32_t on = 0;
#pragma unroll
  for (int h = 0; h < ITEMS; ++h) {
    m[h] = __ballot_sync(0xffffffffu, kind[h] != 0);
    on += __popc(__ballot_sync(0xffffffffu, kind[h] == 1));
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < ITEMS; ++h) s_cnt[8 * h + warp] = __popc(m[h]);
    s_on[warp] = on;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0, run_on = 0;
#pragma unroll
    for (int w = 0; w < 8 * ITEMS; ++w) run += s_cnt[w];
#pragma unroll
    for (int w = 0; w < 8; ++w) run_on += s_on[w];
    s_base = run ? atomicAdd(total, (unsigned long long)run) : 0ull;
    if (run_on) atomicAdd(total_on, (unsigned long long)run_on);
  }
  // resolution of all items: ITEMS independent chains per thread
  Resolved r[ITEMS];
#pragma unroll
  for (int h = 0; h < ITEMS; ++h) {
    if (!kind[h]) continue;
    if (RECORDS) {
      resolve_read(seed[h], seed_read, seed_first, d, first_read_id, r[h]);
      if (is_ovf[h]) resolve_node(g, (uint32_t)code[h], r[h].node_id, r[h].node_off);
      else decode_code(g, code[h], r[h].node_id, r[h].node_off);
    }
    else if (is_ovf[h]) code[h] = code_of_gpos(g, (uint32_t)code[h]);
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t run = 0;
#pragma unroll
  for (int h = 0; h < ITEMS; ++h) {
    uint32_t pre = run;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t c = s_cnt[8 * h + w];
      if (w < (int)warp) pre += c;
      run += c;
    }
    const uint64_t out = s_base + pre + __popc(m[h] & lt);
    if (!kind[h] || out >= cap) continue;
    if (RECORDS) {
      if (RECORDS == 2) st_record32(records + 2 * out, r[h]);
      else st_record(records + 4 * out, r[h]);
      rec_kind[out] = kind[h];
    }
    else { out_code[out] = code[h]; out_seed[out] = seed[h]; }
  }
}

// sorted (locus code, seed) pairs -> records (only used with PSI_B200_SORTED)
template <bool COMPACT>
__global__ void __launch_bounds__(256)
resolve_hits_kernel(GraphView g, const uint64_t* __restrict__ codes, const uint32_t* __restrict__ seeds, uint64_t n_hits,
                    const uint32_t* __restrict__ seed_read, const uint32_t* __restrict__ seed_first,
                    uint32_t d, uint64_t first_read_id, uint64_t* __restrict__ records)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hits) return;
  Resolved r;
  resolve_read(seeds[i], seed_read, seed_first, d, first_read_id, r);
  decode_code(g, codes[i], r.node_id, r.node_off);
  if (COMPACT) { st_record32(records + 2 * i, r); return; }
  ulonglong2* o = reinterpret_cast<ulonglong2*>(records + 4 * i);
  o[0] = make_ulonglong2(r.node_id, r.node_off);
  o[1] = make_ulonglong2(r.read_id, r.read_off);
}

// ================================================================ host ==

static uint64_t next_pow2(uint64_t x)
{
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

static void engine_seeds_impl(Ctx& c, unsigned flags, bool async);

void engine_seeds(Ctx& c, unsigned flags) { engine_seeds_impl(c, flags, false); }

// Queues the step and returns when the fused kernel serves it; the other routes (walk mode, SORTED, NO_RESOLVE)
// run to completion here, and psi_b200_wait then only reports.
void engine_seeds_async(Ctx& c, unsigned flags) { engine_seeds_impl(c, flags, true); }

static void engine_seeds_impl(Ctx& c, unsigned flags, bool async)
{
  if (!c.has_chunk) throw StateError("seeds_all: no read chunk submitted");
  if (c.pending) throw StateError("seeds_all: a step is in flight on this context (call psi_b200_wait first)");
  PSI_CUDA(cudaSetDevice(c.device));
  Shared& sh = *c.sh;
  // index mode: the off-path walks are entries of the table, one probe answers both questions.
  const bool index_mode = sh.offpath_indexed;
  const bool want_on = (flags & PSI_B200_ON_PATHS) && sh.has_index;
  const bool want_off = (flags & PSI_B200_OFF_PATHS) && sh.n_loci > 0;
  const bool do_probe = sh.has_table && (want_on || (want_off && index_mode && sh.n_off_pairs > 0));
  const bool do_walk = want_off && !index_mode;
  const unsigned probe_mode = (want_on ? PSI_B200_ON_PATHS : 0u) | (want_off && index_mode ? PSI_B200_OFF_PATHS : 0u);
  const bool sorted = (flags & PSI_B200_SORTED) != 0;
  const bool resolve = !(flags & PSI_B200_NO_RESOLVE);
  const bool compact = (flags & PSI_B200_COMPACT) != 0;
  const bool dense5 = (flags & PSI_B200_DENSE5) != 0;
  const bool dense = (flags & PSI_B200_DENSE) != 0 || dense5;
  if (dense && (sorted || !resolve || compact)) throw ArgError("seeds_all: PSI_B200_DENSE excludes SORTED, NO_RESOLVE and COMPACT");
  if (dense5 && !dense5_available(sh))
    throw ArgError("seeds_all: PSI_B200_DENSE5 needs an index and (node id << offset bits | offset) below 2^39 - 1 (psi_b200_dense5_layout)");
  if ((compact || dense) && resolve) {
    if (sh.max_node_id >= 0xffffffffull) throw ArgError("seeds_all: PSI_B200_COMPACT / PSI_B200_DENSE need node ids below 2^32 - 1");
    if (c.first_read_id + c.n_reads > 0x100000000ull) throw ArgError("seeds_all: PSI_B200_COMPACT / PSI_B200_DENSE need read ids below 2^32");
  }
  c.records_valid = false;
  c.kinds_valid = false;
  c.records_dense = false;
  c.counters.fused = 0;
  // the usual case -- the index answers every requested phase, records wanted in emission order -- is ONE kernel
  const bool probe_only = !do_walk && !sorted && resolve;
  if ((c.opt_fused || dense) && do_probe && probe_only) {
    if (async) engine_seeds_fused_async(c, probe_mode, compact ? 1 : dense5 ? 3 : dense ? 2 : 0);
    else engine_seeds_fused(c, probe_mode, compact ? 1 : dense5 ? 3 : dense ? 2 : 0);
    return;
  }
  if (dense) {
    // dense results come from the fused kernel only; an index that cannot answer the requested phases hits nothing
    if (do_walk) throw ArgError("seeds_all: PSI_B200_DENSE needs the off-path walks materialised in the index (offpath_mode 0 or 2)");
    if (!sh.has_table) throw StateError("seeds_all: PSI_B200_DENSE needs an index (set_paths / set_loci first)");
    if (async) engine_seeds_fused_async(c, 0, dense5 ? 3 : 2);
    else engine_seeds_fused(c, 0, dense5 ? 3 : 2);
    return;
  }
  engine_seed_chunk(c);
  const GraphView g = make_graph_view(c);
  unsigned long long* dc = c.dev_counters.p;
  c.ev_state[T_ON] = c.ev_state[T_PROBE] = c.ev_state[T_OFF] = c.ev_state[T_RESOLVE] = c.ev_state[T_SORT] = c.ev_state[T_D2H] = 0;

  c.seed_hit.ensure(c.n_seeds_cap, 1.25);
  c.seed_kind.ensure(c.n_seeds_cap + 16, 1.25);
  if (c.hits.cap == 0) { c.hits.ensure(1u << 20); c.hit_kind.ensure(c.hits.cap); }
  uint64_t out_cap_want = std::max<uint64_t>(c.n_seeds_cap + c.n_seeds_cap / 4, 1u << 20);
  uint64_t n_total = 0, n_on = 0, n_walks = 0, n_seeds = 0, n_slow = 0;
  bool redo_search = true;

  for (int attempt = 0;; ++attempt) {
    if (attempt > 16) throw OverflowError("seeds_all: device buffers keep overflowing");
    if (redo_search) {
      PSI_CUDA(cudaMemsetAsync(dc + DC_HITS, 0, 5 * sizeof(unsigned long long), c.stream));  // HITS..WORK
      PSI_CUDA(cudaMemsetAsync(dc + DC_OVF, 0, 2 * sizeof(unsigned long long), c.stream));   // OVF, SLOW
      PhaseTimer t_on(c, T_ON);
      if (do_probe) {
        const unsigned grid = grid_for(c.n_seeds_cap, 256);
        c.slow_queue.ensure(c.n_seeds_cap, 1.25);
        // 6 CTAs x 32 KB of line buffers per SM need the large shared-memory split (an attribute per device)
        static std::atomic<uint64_t> carveout_done{ 0 };
        const uint64_t bit = 1ull << (c.device & 63);
        if (c.device >= 64 || !(carveout_done.load(std::memory_order_acquire) & bit)) {
          PSI_CUDA(cudaFuncSetAttribute(seeds_on_paths_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          PSI_CUDA(cudaFuncSetAttribute(seeds_on_paths_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          carveout_done.fetch_or(bit, std::memory_order_release);
        }
        PhaseTimer t_probe(c, T_PROBE);
        if (sh.index.view.fmt == 8)
          seeds_on_paths_kernel<8><<<grid, 256, 0, c.stream>>>(sh.index.view, c.seed_kmer.p, c.seed_valid.p, dc + DC_SEEDS,
                                                              probe_mode, c.seed_hit.p, c.seed_kind.p, c.slow_queue.p, dc + DC_SLOW);
        else
          seeds_on_paths_kernel<16><<<grid, 256, 0, c.stream>>>(sh.index.view, c.seed_kmer.p, c.seed_valid.p, dc + DC_SEEDS,
                                                               probe_mode, c.seed_hit.p, c.seed_kind.p, c.slow_queue.p, dc + DC_SLOW);
        t_probe.stop();
        seeds_slow_kernel<<<(unsigned)c.sm_count * 2, 256, 0, c.stream>>>(sh.index.view, sh.multi.p, c.seed_kmer.p, c.slow_queue.p,
                                                                         dc + DC_SLOW, probe_mode, c.seed_hit.p, c.seed_kind.p,
                                                                         c.hits.p, c.hit_kind.p, c.hits.cap, dc + DC_OVF);
        c.counters.launches += 2;
      }
      else {
        PSI_CUDA(cudaMemsetAsync(c.seed_kind.p, 0, c.n_seeds_cap, c.stream));
      }
      t_on.stop();

      if (do_walk) {
        engine_index_chunk(c);
        PhaseTimer t_off(c, T_OFF);
        if (c.dedup.cap == 0) c.dedup.ensure(1u << 20);
        const uint64_t dedup_slots = next_pow2(c.dedup.cap) == c.dedup.cap ? c.dedup.cap : next_pow2(c.dedup.cap) / 2;
        PSI_CUDA(cudaMemsetAsync(c.dedup.p, 0xff, dedup_slots * sizeof(unsigned long long), c.stream));
        const unsigned grid = (unsigned)c.sm_count * 8;
        c.walk_spill.ensure((size_t)grid * WALK_WARPS * c.spill_items * sizeof(WalkItem));
        ReadIndexSink sink;
        sink.g = g;
        sink.pfx_bits = c.filter_bits.p;
        sink.pfx = c.filter_pfx;
        sink.rt = c.read_index.view;
        sink.rt.stash_nonempty = 1;  // not known without a sync; probing an empty stash costs one load, and only for full lines
        sink.next = c.seed_next.p;
        sink.pt = sh.index.view;
        sink.multi = sh.multi.p;
        sink.has_index = sh.has_table ? 1u : 0u;
        sink.dedup = c.dedup.p;
        sink.dedup_mask = dedup_slots - 1;
        sink.hits = c.hits.p;
        sink.hit_kind = c.hit_kind.p;
        sink.hits_cap = c.hits.cap;
        sink.hit_count = dc + DC_OVF;
        sink.walk_count = dc + DC_WALKS;
        sink.err = dc + DC_ERR;
        sink.walks = 0;
        seeds_off_paths_kernel<<<grid, WALK_WARPS * 32, 0, c.stream>>>(g, c.k, sh.n_loci, sh.loci_node.p, sh.loci_off.p, sink,
                                                                       dc + DC_WORK, (WalkItem*)c.walk_spill.p, c.spill_items);
        ++c.counters.launches;
        t_off.stop();
      }
      redo_search = false;
    }
    else {
      PSI_CUDA(cudaMemsetAsync(dc + DC_HITS, 0, sizeof(unsigned long long), c.stream));
      PSI_CUDA(cudaMemsetAsync(dc + DC_HITS_ON, 0, sizeof(unsigned long long), c.stream));
    }

    // compaction (+ resolution unless the hits are to be sorted first)
    PhaseTimer t_res(c, T_RESOLVE);
    uint64_t out_cap = 0;
    {
      const unsigned items = (unsigned)c.opt_resolve_items;
      const unsigned grid = grid_for(c.n_seeds_cap + c.hits.cap, 256, items);
#define PSI_RESOLVE_ARGS(REC, KINDS, CODES, SEEDS) \
  g, c.seed_hit.p, c.seed_kind.p, dc + DC_SEEDS, c.hits.p, c.hit_kind.p, dc + DC_OVF, c.hits.cap, c.seed_read.p, \
  c.seed_first.p, c.distance, c.first_read_id, REC, KINDS, CODES, SEEDS, out_cap, dc + DC_HITS, dc + DC_HITS_ON
      if (sorted || !resolve) {
        c.sorted_code.ensure(out_cap_want);
        c.sorted_seed.ensure(out_cap_want);
        out_cap = std::min<uint64_t>(c.sorted_code.cap, c.sorted_seed.cap);
        if (items == 4) compact_resolve_kernel<0, 4, 4><<<grid, 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(nullptr, nullptr, c.sorted_code.p, c.sorted_seed.p));
        else compact_resolve_kernel<0, 2, 5><<<grid, 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(nullptr, nullptr, c.sorted_code.p, c.sorted_seed.p));
      }
      else {
        c.records.ensure(4 * out_cap_want);
        c.rec_kind.ensure(out_cap_want);
        out_cap = std::min<uint64_t>(c.records.cap / 4, c.rec_kind.cap);
        if (compact) compact_resolve_kernel<2, 2, 6><<<grid_for(c.n_seeds_cap + c.hits.cap, 256, 2), 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(c.records.p, c.rec_kind.p, nullptr, nullptr));
        else if (items == 4) compact_resolve_kernel<1, 4, 4><<<grid, 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(c.records.p, c.rec_kind.p, nullptr, nullptr));
        else if (c.opt_resolve_ctas >= 6) compact_resolve_kernel<1, 2, 6><<<grid, 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(c.records.p, c.rec_kind.p, nullptr, nullptr));
        else compact_resolve_kernel<1, 2, 5><<<grid, 256, 0, c.stream>>>(PSI_RESOLVE_ARGS(c.records.p, c.rec_kind.p, nullptr, nullptr));
      }
#undef PSI_RESOLVE_ARGS
      ++c.counters.launches;
    }
    t_res.stop();
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(c.h_pinned, dc, DC_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    n_total = c.h_pinned[DC_HITS];
    n_on = c.h_pinned[DC_HITS_ON];
    n_walks = c.h_pinned[DC_WALKS];
    n_seeds = c.h_pinned[DC_SEEDS];
    n_slow = c.h_pinned[DC_SLOW];
    const uint64_t n_ovf = c.h_pinned[DC_OVF];
    const uint64_t err = c.h_pinned[DC_ERR];
    bool retry = false;
    if (n_ovf > c.hits.cap) { c.hits.ensure(n_ovf, 1.25); c.hit_kind.ensure(c.hits.cap); retry = redo_search = true; }
    if (err & 1ull) {
      if (c.spill_items >= (1u << 20)) throw OverflowError("seeds_off_paths: walk frontier exceeds 2^20 states per warp");
      c.spill_items *= 4;
      retry = redo_search = true;
    }
    if (err & 2ull) throw OverflowError("read index: hash stash exhausted");
    if (err & 4ull) { c.dedup.ensure(c.dedup.cap * 4); retry = redo_search = true; }
    if (!retry && n_total > out_cap) { out_cap_want = n_total + n_total / 8; retry = true; }   // only the compaction is redone
    if (!retry) break;
  }

  c.n_hits = n_total;
  c.counters.n_seeds = n_seeds;
  c.counters.n_hits_on = n_on;
  c.counters.n_hits_off = n_total - n_on;
  c.counters.n_hits = n_total;
  c.counters.n_walks = n_walks;
  c.counters.n_on_probe_sectors = n_slow;   // seeds that needed more than their home line
  c.counters.offpath_mode = index_mode ? 2u : 1u;

  const uint64_t* codes = c.sorted_code.p;
  const uint32_t* seeds = c.sorted_seed.p;
  if (sorted && n_total > 1) {
    PhaseTimer t_sort(c, T_SORT);
    // canonical order: (seed index, locus code) ascending == (read_id, read_offset, node id or rank, node_offset):
    // LSD with two stable passes, by code, then by seed
    DevBuf<uint64_t> code_b;
    DevBuf<uint32_t> seed_b;
    code_b.ensure(n_total); seed_b.ensure(n_total);
    size_t tmp1 = 0, tmp2 = 0;
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp1, c.sorted_code.p, code_b.p, c.sorted_seed.p, seed_b.p, (int64_t)n_total, 0, 64, c.stream));
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, seed_b.p, c.sorted_seed.p, code_b.p, c.sorted_code.p, (int64_t)n_total, 0, 32, c.stream));
    c.scan_tmp.ensure(std::max(tmp1, tmp2));
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(c.scan_tmp.p, tmp1, c.sorted_code.p, code_b.p, c.sorted_seed.p, seed_b.p, (int64_t)n_total, 0, 64, c.stream));
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(c.scan_tmp.p, tmp2, seed_b.p, c.sorted_seed.p, code_b.p, c.sorted_code.p, (int64_t)n_total, 0, 32, c.stream));
    t_sort.stop();
    PSI_CUDA(cudaStreamSynchronize(c.stream));   // the temporaries die here
    c.counters.launches += 12;
  }
  if (sorted && resolve) {
    c.records.ensure(4 * std::max<uint64_t>(n_total, 1), 1.25);
    if (n_total) {
      if (compact)
        resolve_hits_kernel<true><<<grid_for(n_total, 256), 256, 0, c.stream>>>(g, codes, seeds, n_total, c.seed_read.p,
                                                                                c.seed_first.p, c.distance, c.first_read_id, c.records.p);
      else
        resolve_hits_kernel<false><<<grid_for(n_total, 256), 256, 0, c.stream>>>(g, codes, seeds, n_total, c.seed_read.p,
                                                                                 c.seed_first.p, c.distance, c.first_read_id, c.records.p);
      ++c.counters.launches;
    }
    PSI_CUDA(cudaGetLastError());
  }
  c.records_valid = resolve;
  c.records_compact = resolve && compact;
  c.kinds_valid = resolve && !sorted;
}

void engine_fetch(Ctx& c, void* hits, uint64_t cap, bool compact)
{
  if (c.pending) throw StateError("fetch: a step is in flight on this context (call psi_b200_wait first)");
  if (!c.records_valid || c.records_dense) throw StateError("fetch: no resolved seed records (call seeds_all without NO_RESOLVE / DENSE first)");
  if (compact != c.records_compact)
    throw StateError(compact ? "fetch32: the last seeds_all was not run with PSI_B200_COMPACT"
                             : "fetch: the last seeds_all was run with PSI_B200_COMPACT (use psi_b200_fetch32)");
  PSI_CUDA(cudaSetDevice(c.device));
  const uint64_t n = c.n_hits < cap ? c.n_hits : cap;
  PhaseTimer t(c, T_D2H);
  if (n) PSI_CUDA(cudaMemcpyAsync(hits, c.records.p, n * (compact ? 16u : 32u), cudaMemcpyDeviceToHost, c.stream));
  t.stop();
  ctx_wait(c);
}

void engine_fetch_kinds(Ctx& c, uint8_t* kinds, uint64_t cap)
{
  if (c.pending) throw StateError("fetch_kinds: a step is in flight on this context (call psi_b200_wait first)");
  if (!c.kinds_valid) throw StateError("fetch_kinds: no unsorted resolved records");
  PSI_CUDA(cudaSetDevice(c.device));
  const uint64_t n = c.n_hits < cap ? c.n_hits : cap;
  if (n) PSI_CUDA(cudaMemcpyAsync(kinds, c.rec_kind.p, n, cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace psi_b200
