// index_build.cu -- GPU construction of the path index and of the starting loci.
//
// Path index: stands behind PathIndex::create_index + psi::FMIndex (reference
// include/psi/pathindex.hpp:235-243, fmindex.hpp:257-271), which build
// csa_wt<wt_huff<>,32,64> over rev(path_0)$rev(path_1)$...  The only queries
// the seed-finding path ever makes against it are "all occurrences of a k-mer"
// with fixed k (index_iter.hpp:808-852), each mapped to the (node, offset) of
// the occurrence's first base (pathindex.hpp:378-416).  The device index
// therefore stores exactly that relation: the distinct (k-mer, locus) pairs of
// all k-windows of all paths, in a bucketised hash keyed by the k-mer.
//
// Starting loci: stands behind SeedFinder::add_uncovered_loci
// (seed_finder.hpp:1481-1541).  A locus is kept iff some k-walk starting there
// spells a k-mer w with (w, locus) absent from the path index -- then, and only
// then, seeds_on_paths cannot report every hit at this locus.
#include "engine.hpp"
#include "walker.cuh"

#include <algorithm>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace psi_b200 {

using namespace dev;

// ------------------------------------------------------ path k-windows --

struct PathsView {
  const uint64_t* path_ptr;    // n_paths + 1
  const uint32_t* nodes;       // entries
  const uint32_t* head_off;    // may be null
  const uint32_t* tail_trim;   // may be null
  uint32_t n_paths;
};

__device__ __forceinline__ uint32_t path_of_entry(const PathsView& p, uint64_t e)
{
  uint32_t lo = 0, hi = p.n_paths;  // path_ptr[lo] <= e < path_ptr[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(p.path_ptr + mid) <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// One warp per path entry (a node visit); lanes stride over the offsets inside
// the node.  For every offset the k-window that starts there is gathered along
// the PATH's successors.  count_only: just count valid windows.
// MODE 0: emit (k-mer, locus) pairs; 1: count the valid windows; 2: count the windows PER K-MER in a counting table
// (the k-mer's number of occurrences in the path text, what the reference compares with its gocc threshold,
// index_iter.hpp:842-848)
// A slice (sliced builds of very large graphs): only the windows whose FIRST slice_bits / 2 bases spell slice_id are
// taken; a window whose first bases lie inside its node is rejected before anything else is read.
struct WindowSlice {
  uint32_t bits;   // 0: every window
  uint32_t id;
  uint64_t cap;    // MODE 0: room in out_kmer / out_gpos (the count goes on beyond it; nothing is written)
};

template <int MODE>
__global__ void __launch_bounds__(256)
path_windows_kernel(GraphView g, PathsView p, uint32_t k, uint64_t e_begin, uint64_t e_end,
                    uint64_t* __restrict__ out_kmer, uint32_t* __restrict__ out_gpos,
                    unsigned long long* __restrict__ out_count, KmerTable counts, unsigned long long* __restrict__ err_flag,
                    WindowSlice slice)
{
  constexpr bool COUNT_ONLY = MODE == 1;
  const uint32_t lane = lane_id();
  const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long local_count = 0;
  for (uint64_t e = e_begin + warp0; e < e_end; e += n_warps) {
    const uint32_t pi = path_of_entry(p, e);
    const uint64_t pbeg = __ldg(p.path_ptr + pi), pend = __ldg(p.path_ptr + pi + 1);
    const uint32_t head = p.head_off ? __ldg(p.head_off + pi) : 0;
    const uint32_t tail = p.tail_trim ? __ldg(p.tail_trim + pi) : 0;
    const uint32_t node = __ldg(p.nodes + e);
    const NodeRec r = g.rec[node];
    const uint32_t from = (e == pbeg) ? head : 0;
    uint32_t to = r.seq_len;
    if (e + 1 == pend) to = tail < to ? to - tail : 0;
    for (uint32_t o0 = from; o0 < to; o0 += 32) {   // warp-uniform bounds
      const uint32_t o = o0 + lane;
      bool ok = o < to;
      uint64_t kmer = 0;
      if (ok && slice.bits && to - o >= (slice.bits >> 1))
        ok = (uint32_t)extract_bases(g.seq2, (uint64_t)r.seq_start + o, slice.bits >> 1) == slice.id;
      if (ok) {
        uint32_t depth = 0, off = o;
        uint64_t ce = e;
        NodeRec cr = r;
        uint32_t cto = to;
        while (true) {
          const uint32_t avail = cto > off ? cto - off : 0;
          const uint32_t want = k - depth;
          const uint32_t take = avail < want ? avail : want;
          if (take) {
            const uint64_t pos = (uint64_t)cr.seq_start + off;
            if (g.has_n && extract_nmask(g.nmask, pos, take)) { ok = false; break; }
            kmer |= extract_bases(g.seq2, pos, take) << (2u * depth);
            depth += take;
          }
          if (depth == k) break;
          if (++ce == pend) { ok = false; break; }   // window runs off the path
          cr = g.rec[__ldg(p.nodes + ce)];
          off = 0;
          cto = cr.seq_len;
          if (ce + 1 == pend) cto = tail < cto ? cto - tail : 0;
        }
        if (ok && slice.bits && (uint32_t)(kmer & ((1ull << slice.bits) - 1ull)) != slice.id) ok = false;
      }
      if (COUNT_ONLY) local_count += ok;
      else if (MODE == 2) {
        if (ok && !(counts.fmt == 8 ? table_count<8>(counts, kmer) : table_count<16>(counts, kmer))) atomicOr(err_flag, 2ull);
      }
      else {
        const uint64_t slot = warp_reserve(out_count, ok ? 1u : 0u);
        if (ok && slot < slice.cap) { out_kmer[slot] = kmer; out_gpos[slot] = r.seq_start + o; }
      }
    }
  }
  if (COUNT_ONLY) {
#pragma unroll
    for (int d = 16; d; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
    if (lane == 0 && local_count) atomicAdd(out_count, local_count);
  }
}

// ------------------------------------------- unique pairs and k-mer runs --

// flag[i] = 1 iff pair i differs from pair i-1 (pairs sorted by (kmer, gpos)).
__global__ void __launch_bounds__(256)
flag_unique_pairs_kernel(const uint64_t* __restrict__ kmer, const uint32_t* __restrict__ gpos, uint64_t n,
                         uint32_t* __restrict__ flag)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || kmer[i] != kmer[i - 1] || gpos[i] != gpos[i - 1]) ? 1u : 0u;
}

// Scatter unique pairs: dst index = exclusive scan of flag.
__global__ void __launch_bounds__(256)
compact_pairs_kernel(const uint64_t* __restrict__ kmer, const uint32_t* __restrict__ gpos,
                     const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan, uint64_t n,
                     uint64_t* __restrict__ ukmer, uint32_t* __restrict__ ugpos)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  ukmer[scan[i]] = kmer[i];
  ugpos[scan[i]] = gpos[i];
}

// Concatenate the on-path pairs and the off-path pairs: value = gpos | (off << 32).
__global__ void __launch_bounds__(256)
concat_pairs_kernel(const uint64_t* __restrict__ on_kmer, const uint32_t* __restrict__ on_gpos, uint64_t n_on,
                    const uint64_t* __restrict__ off_kmer, const uint32_t* __restrict__ off_gpos, uint64_t n_off,
                    uint64_t* __restrict__ key, uint64_t* __restrict__ val)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_on) { key[i] = on_kmer[i]; val[i] = on_gpos[i]; }
  else if (i < n_on + n_off) { key[i] = off_kmer[i - n_on]; val[i] = (uint64_t)off_gpos[i - n_on] | (1ull << 32); }
}

// head[i] = 1 iff pair i starts a new k-mer run.
__global__ void __launch_bounds__(256)
flag_run_heads_kernel(const uint64_t* __restrict__ ukmer, uint64_t n, uint32_t* __restrict__ head)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || ukmer[i] != ukmer[i - 1]) ? 1u : 0u;
}

// run_start[run id] = i for every run head (run id = exclusive scan of head).
__global__ void __launch_bounds__(256)
scatter_run_starts_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ scan, uint64_t n,
                          uint32_t n_runs, uint32_t* __restrict__ run_start)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) run_start[n_runs] = (uint32_t)n;
  if (i >= n || !head[i]) return;
  run_start[scan[i]] = (uint32_t)i;
}

// words each run needs in the multi array: 0 for single-locus k-mers, else 2 + count
__global__ void __launch_bounds__(256)
run_multi_words_kernel(const uint32_t* __restrict__ run_start, uint32_t n_runs, uint32_t* __restrict__ words)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint32_t len = run_start[r + 1] - run_start[r];
  words[r] = len > 1 ? len + 2 : 0;
}

// One thread per distinct k-mer: write its locus list (if several) and insert
// it into the table.  Within a run the on-path loci come first (stable sort of
// the concatenation), each group ascending.
template <int FMT>
__global__ void __launch_bounds__(256)
insert_runs_kernel(KmerTable t, GraphView g, const uint64_t* __restrict__ key, const uint64_t* __restrict__ val,
                   const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ multi_off,
                   uint32_t n_runs, uint32_t* __restrict__ multi, unsigned long long* __restrict__ err_flag, uint32_t multi_base)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint32_t b = run_start[r], e = run_start[r + 1];
  const uint64_t kmer = key[b];
  uint64_t payload;
  uint32_t flags;
  if (e - b == 1) {
    // single locus: the entry carries the locus code, so that a probe that finds it needs nothing else
    const uint64_t v = val[b];
    payload = code_of_gpos(g, (uint32_t)v);
    flags = (v >> 32) ? FLAG_OFF : 0u;
  }
  else {
    const uint32_t m = multi_base + multi_off[r];
    uint32_t n_on = 0;
    for (uint32_t i = b; i < e; ++i) {
      const uint64_t v = val[i];
      multi[m + 2 + (i - b)] = (uint32_t)v;
      n_on += (v >> 32) ? 0u : 1u;
    }
    multi[m] = n_on;
    multi[m + 1] = e - b;
    payload = m;
    flags = FLAG_MULTI;
  }
  uint32_t prev;
  if (!table_insert<FMT>(t, kmer, payload, flags, false, prev)) atomicOr(err_flag, 2ull);
}

// ------------------------------------------------ gocc threshold (-r) --
//
// The reference skips, in seeds_on_paths, every k-mer that occurs more than gocc_threshold times in the path text
// (kmer_exact_matches, index_iter.hpp:842-848); seeds_off_paths is not affected and still reports every k-walk from a
// starting locus.  With the off-path walks materialised that is a filter on the index: the on-path entries of such a
// k-mer go, except those at a starting locus, which stay as off-path entries.

__global__ void __launch_bounds__(256)
mark_loci_kernel(const NodeRec* __restrict__ rec, const uint32_t* __restrict__ loci_node, const uint32_t* __restrict__ loci_off,
                 uint64_t n_loci, uint32_t n_nodes, uint32_t* __restrict__ bits)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loci) return;
  const uint32_t v = loci_node[i];
  if (v >= n_nodes || loci_off[i] >= rec[v].seq_len) return;
  const uint32_t gpos = rec[v].seq_start + loci_off[i];
  atomicOr(bits + (gpos >> 5), 1u << (gpos & 31u));
}

// keep[i] = 0 for the on-path pairs of over-represented k-mers that are not at a starting locus; those at one become
// off-path entries (bit 32 of the value).
__global__ void __launch_bounds__(256)
gocc_filter_kernel(const uint64_t* __restrict__ key, uint64_t* __restrict__ val, uint64_t n, KmerTable counts, uint32_t threshold,
                   const uint32_t* __restrict__ loci_bits, uint32_t* __restrict__ keep)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k_ = 1;
  const uint64_t v = val[i];
  if (!(v >> 32)) {
    Found f{};
    if (table_find_any(counts, key[i], f) && (uint32_t)f.payload > threshold) {
      const uint32_t gpos = (uint32_t)v;
      if ((loci_bits[gpos >> 5] >> (gpos & 31u)) & 1u) val[i] = v | (1ull << 32);
      else k_ = 0;
    }
  }
  keep[i] = k_;
}

__global__ void __launch_bounds__(256)
compact_kv_kernel(const uint64_t* __restrict__ key, const uint64_t* __restrict__ val, const uint32_t* __restrict__ keep,
                  const uint32_t* __restrict__ scan, uint64_t n, uint64_t* __restrict__ okey, uint64_t* __restrict__ oval)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  okey[scan[i]] = key[i];
  oval[scan[i]] = val[i];
}

// ------------------------------------------------------------ host side --

static void exclusive_scan_u32(Ctx& c, const uint32_t* in, uint32_t* out, uint64_t n)
{
  size_t tmp = 0;
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, c.stream));
  c.scan_tmp.ensure(tmp);
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(c.scan_tmp.p, tmp, in, out, (int64_t)n, c.stream));
  c.counters.launches += 2;
}

// Sort n (kmer, gpos) pairs by (kmer, gpos) and drop duplicates.  On return the unique pairs are
// in (kmer_out, gpos_out)[0, result); the inputs are clobbered.
static uint64_t sort_unique_pairs(Ctx& c, DevBuf<uint64_t>& kmer_a, DevBuf<uint32_t>& gpos_a, uint64_t n,
                                  DevBuf<uint64_t>& kmer_out, DevBuf<uint32_t>& gpos_out)
{
  if (n == 0) return 0;
  DevBuf<uint64_t> kmer_b;
  DevBuf<uint32_t> gpos_b;
  kmer_b.ensure(n); gpos_b.ensure(n);
  // LSD: stable sort by gpos, then stable sort by kmer
  size_t tmp1 = 0, tmp2 = 0;
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp1, gpos_a.p, gpos_b.p, kmer_a.p, kmer_b.p, (int64_t)n, 0, 32, c.stream));
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, kmer_b.p, kmer_a.p, gpos_b.p, gpos_a.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
  DevBuf<char> tmp;
  tmp.ensure(std::max(tmp1, tmp2));
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp1, gpos_a.p, gpos_b.p, kmer_a.p, kmer_b.p, (int64_t)n, 0, 32, c.stream));
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp2, kmer_b.p, kmer_a.p, gpos_b.p, gpos_a.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
  c.counters.launches += 12;
  // sorted pairs are in (kmer_a, gpos_a)
  DevBuf<uint32_t> flag, scan;
  flag.ensure(n + 1); scan.ensure(n + 1);
  flag_unique_pairs_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(kmer_a.p, gpos_a.p, n, flag.p);
  ++c.counters.launches;
  PSI_CUDA(cudaMemsetAsync(flag.p + n, 0, sizeof(uint32_t), c.stream));
  exclusive_scan_u32(c, flag.p, scan.p, n + 1);
  uint32_t n_unique = 0;
  PSI_CUDA(cudaMemcpyAsync(&n_unique, scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  kmer_out.ensure(n_unique); gpos_out.ensure(n_unique);
  compact_pairs_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(kmer_a.p, gpos_a.p, flag.p, scan.p, n, kmer_out.p, gpos_out.p);
  ++c.counters.launches;
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  return n_unique;
}

// ------------------------------------------------------------ sliced build --
//
// The one-shot build above counts pairs, runs and list words in 32 bits.  A graph with more than ~4 G distinct
// (k-mer, locus) pairs (BASELINE configs[4]: 3.1 Gbp) is built in slices of the k-mer space instead: a k-mer belongs to
// the slice its first slice_bits / 2 bases spell, every slice holds its own sorted distinct pairs (Shared::slices) and
// is run-length grouped and inserted into the ONE table by itself, so that every count is per slice.

template <class T>
static void grow_preserve(Ctx& c, DevBuf<T>& b, size_t used, size_t want)
{
  if (want <= b.cap) return;
  T* np = nullptr;
  PSI_CUDA(cudaMalloc((void**)&np, want * sizeof(T)));
  if (used && b.p) PSI_CUDA(cudaMemcpyAsync(np, b.p, used * sizeof(T), cudaMemcpyDeviceToDevice, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  if (b.p) cudaFree(b.p);
  b.p = np;
  b.cap = want;
}

__global__ void __launch_bounds__(256)
flag_slice_kernel(const uint64_t* __restrict__ kmer, uint64_t n, uint64_t mask, uint32_t id, uint32_t* __restrict__ flag)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (kmer[i] & mask) == id ? 1u : 0u;
}

static void empty_table(Ctx& c)
{
  Shared& sh = *c.sh;
  table_alloc(c, sh.index, 1, 2 * c.k, 1024);
  sh.index.view.stash_nonempty = 0;
  sh.multi.ensure(4);
  sh.has_table = true;
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.counters.n_index_entries = c.counters.n_index_kmers = 0;
}

static void build_table_sliced(Ctx& c, const uint64_t* off_kmer, const uint32_t* off_gpos, uint64_t n_off)
{
  Shared& sh = *c.sh;
  if (sh.pairs_released)
    throw StateError("index: the sorted pairs of a sliced build were released when the starting loci were set; call set_paths again");
  if (sh.gocc_threshold) throw ArgError("index: a gocc threshold is not served by a sliced build (build_slices)");
  if (n_off >= 0xfffffff0ull) throw ArgError("index: more than 2^32 off-path pairs");
  const uint32_t n_slices = (uint32_t)sh.slices.size();
  const uint64_t slice_mask = (1ull << sh.slice_bits) - 1ull;
  sh.table_filtered = false;
  sh.has_table = false;
  sh.n_off_pairs = n_off;
  unsigned long long* d_err = c.dev_counters.p + DC_ERR;
  uint64_t runs_bound = n_off;
  for (auto& sl : sh.slices) runs_bound += sl->n_kmers;
  if (runs_bound == 0) { empty_table(c); return; }
  PSI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), c.stream));
  // locus codes as in build_table; list offsets are taken as 32 bits wide (the lists are not known yet)
  {
    auto bits_of = [](uint64_t x) { uint32_t b = 0; while (x) { ++b; x >>= 1; } return b; };
    const uint32_t off_bits = bits_of(sh.max_node_len ? sh.max_node_len - 1 : 0);
    const uint32_t id_bits = bits_of(sh.max_node_id), rank_bits = bits_of(sh.n_nodes ? sh.n_nodes - 1 : 0);
    const uint32_t list_bits = 32;
    const uint32_t cap8 = table_fmt8_payload_bits(runs_bound, 2 * c.k, c.opt_index_slack);
    uint32_t need;
    sh.code_off_bits = off_bits;
    if (std::max(id_bits + off_bits, list_bits) <= cap8) { sh.code_by_rank = false; need = id_bits + off_bits; }
    else if (std::max(rank_bits + off_bits, list_bits) <= cap8) { sh.code_by_rank = true; need = rank_bits + off_bits; }
    else if (id_bits + off_bits <= 62) { sh.code_by_rank = false; need = id_bits + off_bits; }
    else { sh.code_by_rank = true; need = rank_bits + off_bits; }
    if (c.opt_code_by_rank) { sh.code_by_rank = true; need = rank_bits + off_bits; }
    need = std::max(need, list_bits);
    if (need > 62) throw ArgError("index: node labels too long for the locus codes");
    table_alloc(c, sh.index, runs_bound, 2 * c.k, runs_bound / 512 + 1024, c.opt_index_slack, need);
  }
  const GraphView g = make_graph_view(c);   // after the code layout has been chosen
  if (sh.multi.cap < (1u << 20)) sh.multi.ensure(1u << 20);
  uint64_t multi_used = 0, n_total = 0, n_runs_total = 0;
  DevBuf<uint32_t> oflag, oscan, offs_g;
  DevBuf<uint64_t> offs_k;
  if (n_off) { oflag.ensure(n_off + 1); oscan.ensure(n_off + 1); }
  for (uint32_t si = 0; si < n_slices; ++si) {
    const Shared::PairSlice& sl = *sh.slices[si];
    uint32_t n_off_s = 0;
    if (n_off) {   // the off-path pairs of this slice (the input is sorted, so is the selection)
      flag_slice_kernel<<<grid_for(n_off, 256), 256, 0, c.stream>>>(off_kmer, n_off, slice_mask, si, oflag.p);
      PSI_CUDA(cudaMemsetAsync(oflag.p + n_off, 0, sizeof(uint32_t), c.stream));
      exclusive_scan_u32(c, oflag.p, oscan.p, n_off + 1);
      PSI_CUDA(cudaMemcpyAsync(&n_off_s, oscan.p + n_off, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      if (n_off_s) {
        offs_k.ensure(n_off_s); offs_g.ensure(n_off_s);
        compact_pairs_kernel<<<grid_for(n_off, 256), 256, 0, c.stream>>>(off_kmer, off_gpos, oflag.p, oscan.p, n_off, offs_k.p, offs_g.p);
      }
      c.counters.launches += 2;
    }
    const uint64_t n = sl.n + n_off_s;
    if (n == 0) continue;
    if (n >= 0xfffffff0ull) throw ArgError("index: more than 2^32 pairs in one slice (raise build_slices)");
    DevBuf<uint64_t> key_a, val_a, key_b, val_b;
    key_a.ensure(n); val_a.ensure(n);
    concat_pairs_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(sl.kmer.p, sl.gpos.p, sl.n, offs_k.p, offs_g.p, n_off_s, key_a.p, val_a.p);
    ++c.counters.launches;
    uint64_t* key = key_a.p;
    uint64_t* val = val_a.p;
    if (n_off_s && sl.n) {
      key_b.ensure(n); val_b.ensure(n);
      size_t tmp_bytes = 0;
      PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a.p, key_b.p, val_a.p, val_b.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
      DevBuf<char> tmp;
      tmp.ensure(tmp_bytes);
      PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.p, key_b.p, val_a.p, val_b.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
      c.counters.launches += 6;
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      key = key_b.p;
      val = val_b.p;
    }
    DevBuf<uint32_t> flag, scan;
    flag.ensure(n + 1); scan.ensure(n + 1);
    flag_run_heads_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(key, n, flag.p);
    ++c.counters.launches;
    PSI_CUDA(cudaMemsetAsync(flag.p + n, 0, sizeof(uint32_t), c.stream));
    exclusive_scan_u32(c, flag.p, scan.p, n + 1);
    uint32_t n_runs = 0;
    PSI_CUDA(cudaMemcpyAsync(&n_runs, scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    DevBuf<uint32_t> run_start, words, multi_off;
    run_start.ensure((uint64_t)n_runs + 2); words.ensure((uint64_t)n_runs + 2); multi_off.ensure((uint64_t)n_runs + 2);
    scatter_run_starts_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(flag.p, scan.p, n, n_runs, run_start.p);
    run_multi_words_kernel<<<grid_for(n_runs, 256), 256, 0, c.stream>>>(run_start.p, n_runs, words.p);
    c.counters.launches += 2;
    PSI_CUDA(cudaMemsetAsync(words.p + n_runs, 0, sizeof(uint32_t), c.stream));
    exclusive_scan_u32(c, words.p, multi_off.p, (uint64_t)n_runs + 1);
    uint32_t multi_words = 0;
    PSI_CUDA(cudaMemcpyAsync(&multi_words, multi_off.p + n_runs, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (multi_used + multi_words + 4 >= 0xfffffff0ull) throw ArgError("index: locus lists beyond 2^32 words");
    if (multi_used + multi_words + 4 > sh.multi.cap)
      grow_preserve(c, sh.multi, multi_used, std::max<uint64_t>(2 * sh.multi.cap, multi_used + multi_words + 4));
    if (sh.index.view.fmt == 8)
      insert_runs_kernel<8><<<grid_for(n_runs, 256), 256, 0, c.stream>>>(sh.index.view, g, key, val, run_start.p, multi_off.p, n_runs, sh.multi.p, d_err,
                                                                        (uint32_t)multi_used);
    else
      insert_runs_kernel<16><<<grid_for(n_runs, 256), 256, 0, c.stream>>>(sh.index.view, g, key, val, run_start.p, multi_off.p, n_runs, sh.multi.p, d_err,
                                                                         (uint32_t)multi_used);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaStreamSynchronize(c.stream));     // the slice's temporaries die here
    multi_used += multi_words;
    n_total += n;
    n_runs_total += n_runs;
  }
  unsigned long long err = 0;
  uint32_t stash_used = 0;
  PSI_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaMemcpyAsync(&stash_used, sh.index.stash_used.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  if (err & 2ull) throw OverflowError("path index: hash stash exhausted");
  sh.index.view.stash_nonempty = stash_used ? 1u : 0u;
  sh.has_table = true;
  c.counters.n_index_entries = n_total;
  c.counters.n_index_kmers = n_runs_total;
  c.counters.n_offpath_entries = n_off;
  c.counters.index_buckets = sh.index.n_lines;
  c.counters.index_bytes = sh.index.n_lines * 128 + (multi_used + 4) * 4 + ((uint64_t)sh.index.view.stash_mask + 1) * sizeof(Slot16);
  c.counters.index_slot_bytes = sh.index.view.fmt;
  c.counters.index_stash_used = stash_used;
}

// (Re)build the device index from the resident on-path pairs plus `n_off` off-path pairs
// (both sorted by (kmer, gpos) and duplicate free).  final: the starting loci are known and this is the table the
// chunks will probe -- the gocc threshold, if one is set, is applied here (never to the table the loci are derived from).
static void build_table(Ctx& c, const uint64_t* off_kmer, const uint32_t* off_gpos, uint64_t n_off, bool final = false)
{
  Shared& sh = *c.sh;
  if (!sh.slices.empty() || sh.pairs_released) { build_table_sliced(c, off_kmer, off_gpos, n_off); return; }
  const uint64_t n_on = sh.n_on_pairs;
  uint64_t n = n_on + n_off;
  sh.table_filtered = false;
  unsigned long long* d_err = c.dev_counters.p + DC_ERR;
  sh.has_table = false;
  sh.n_off_pairs = n_off;
  if (n >= 0xfffffff0ull) throw ArgError("index: more than 2^32 (k-mer, locus) pairs");
  if (n == 0) {
    table_alloc(c, sh.index, 1, 2 * c.k, 1024);
    sh.index.view.stash_nonempty = 0;
    sh.multi.ensure(4);
    sh.has_table = true;
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    c.counters.n_index_entries = c.counters.n_index_kmers = 0;
    return;
  }
  PSI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), c.stream));
  DevBuf<uint64_t> key_a, val_a, key_b, val_b;
  key_a.ensure(n); val_a.ensure(n);
  concat_pairs_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(sh.on_kmer.p, sh.on_gpos.p, n_on, off_kmer, off_gpos, n_off, key_a.p, val_a.p);
  ++c.counters.launches;
  uint64_t* key = key_a.p;
  uint64_t* val = val_a.p;
  if (n_off && n_on) {  // stable sort by k-mer keeps on-path loci ahead of off-path loci, each group ascending
    key_b.ensure(n); val_b.ensure(n);
    size_t tmp_bytes = 0;
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a.p, key_b.p, val_a.p, val_b.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
    DevBuf<char> tmp;
    tmp.ensure(tmp_bytes);
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.p, key_b.p, val_a.p, val_b.p, (int64_t)n, 0, (int)(2 * c.k), c.stream));
    c.counters.launches += 6;
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    key = key_b.p;
    val = val_b.p;
  }

  DevBuf<uint64_t> key_f, val_f;
  if (final && sh.gocc_threshold && n_on) {
    // occurrences of every k-mer in the path text, counted over the retained paths
    if (!sh.paths_kept) throw StateError("gocc threshold: the paths were indexed before the threshold was set");
    HostTable counts;
    for (int grow = 0;; ++grow) {
      table_alloc(c, counts, n_on, 2 * c.k, 1024, grow);
      PSI_CUDA(cudaMemsetAsync(counts.slots.p, 0xff, counts.n_lines * 128, c.stream));
      PSI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), c.stream));
      const GraphView g0 = make_graph_view(c);
      PathsView pv{ sh.path_ptr.p, sh.path_nodes.p, sh.path_head.p, sh.path_tail.p, (uint32_t)sh.n_paths };
      const uint64_t n_entries = sh.n_path_entries;
      const unsigned wgrid = (unsigned)std::min<uint64_t>((n_entries + 7) / 8 + 1, (uint64_t)c.sm_count * 32);
      path_windows_kernel<2><<<wgrid, 256, 0, c.stream>>>(g0, pv, c.k, 0, n_entries, nullptr, nullptr, nullptr, counts.view, d_err, WindowSlice{ 0, 0, 0 });
      ++c.counters.launches;
      unsigned long long err = 0;
      PSI_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, c.stream));
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      if (!(err & 2ull)) break;
      if (grow >= 3) throw OverflowError("gocc threshold: counting table overflow");
    }
    counts.view.stash_nonempty = 0;
    const uint64_t n_words = (sh.n_bases + 31) >> 5;
    DevBuf<uint32_t> loci_bits, keep, kscan;
    loci_bits.ensure(n_words + 1);
    PSI_CUDA(cudaMemsetAsync(loci_bits.p, 0, (n_words + 1) * sizeof(uint32_t), c.stream));
    if (sh.n_loci)
      mark_loci_kernel<<<grid_for(sh.n_loci, 256), 256, 0, c.stream>>>(sh.node_rec.p, sh.loci_node.p, sh.loci_off.p, sh.n_loci, sh.n_nodes, loci_bits.p);
    keep.ensure(n + 1); kscan.ensure(n + 1);
    gocc_filter_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(key, val, n, counts.view, sh.gocc_threshold, loci_bits.p, keep.p);
    PSI_CUDA(cudaMemsetAsync(keep.p + n, 0, sizeof(uint32_t), c.stream));
    exclusive_scan_u32(c, keep.p, kscan.p, n + 1);
    uint32_t n_kept = 0;
    PSI_CUDA(cudaMemcpyAsync(&n_kept, kscan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    key_f.ensure((uint64_t)n_kept + 1); val_f.ensure((uint64_t)n_kept + 1);
    compact_kv_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(key, val, keep.p, kscan.p, n, key_f.p, val_f.p);
    c.counters.launches += 3;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    c.counters.n_gocc_dropped = n - n_kept;
    key = key_f.p;
    val = val_f.p;
    n = n_kept;
    sh.table_filtered = true;
    if (n == 0) {
      table_alloc(c, sh.index, 1, 2 * c.k, 1024);
      sh.index.view.stash_nonempty = 0;
      sh.multi.ensure(4);
      sh.has_table = true;
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      c.counters.n_index_entries = c.counters.n_index_kmers = 0;
      return;
    }
  }

  // k-mer runs
  DevBuf<uint32_t> flag, scan;
  flag.ensure(n + 1); scan.ensure(n + 1);
  flag_run_heads_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(key, n, flag.p);
  ++c.counters.launches;
  PSI_CUDA(cudaMemsetAsync(flag.p + n, 0, sizeof(uint32_t), c.stream));
  exclusive_scan_u32(c, flag.p, scan.p, n + 1);
  uint32_t n_runs = 0;
  PSI_CUDA(cudaMemcpyAsync(&n_runs, scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  DevBuf<uint32_t> run_start, words, multi_off;
  run_start.ensure((uint64_t)n_runs + 2); words.ensure((uint64_t)n_runs + 2); multi_off.ensure((uint64_t)n_runs + 2);
  scatter_run_starts_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(flag.p, scan.p, n, n_runs, run_start.p);
  run_multi_words_kernel<<<grid_for(n_runs, 256), 256, 0, c.stream>>>(run_start.p, n_runs, words.p);
  c.counters.launches += 2;
  PSI_CUDA(cudaMemsetAsync(words.p + n_runs, 0, sizeof(uint32_t), c.stream));
  exclusive_scan_u32(c, words.p, multi_off.p, (uint64_t)n_runs + 1);
  uint32_t multi_words = 0;
  PSI_CUDA(cudaMemcpyAsync(&multi_words, multi_off.p + n_runs, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  sh.multi.ensure((uint64_t)multi_words + 4);

  // Locus codes: (node id << off_bits) | offset when that fits the payload of 8-byte slots, else (node rank << off_bits)
  // | offset (one gather of node_id per hit), else 16-byte slots (62 payload bits: by id unless the ids are wider still).
  {
    auto bits_of = [](uint64_t x) { uint32_t b = 0; while (x) { ++b; x >>= 1; } return b; };
    const uint32_t off_bits = bits_of(sh.max_node_len ? sh.max_node_len - 1 : 0);
    const uint32_t id_bits = bits_of(sh.max_node_id), rank_bits = bits_of(sh.n_nodes ? sh.n_nodes - 1 : 0);
    const uint32_t list_bits = bits_of((uint64_t)multi_words + 4);
    const uint32_t cap8 = table_fmt8_payload_bits(n_runs, 2 * c.k, c.opt_index_slack);
    uint32_t need;
    sh.code_off_bits = off_bits;
    if (std::max(id_bits + off_bits, list_bits) <= cap8) { sh.code_by_rank = false; need = id_bits + off_bits; }
    else if (std::max(rank_bits + off_bits, list_bits) <= cap8) { sh.code_by_rank = true; need = rank_bits + off_bits; }
    else if (id_bits + off_bits <= 62) { sh.code_by_rank = false; need = id_bits + off_bits; }
    else { sh.code_by_rank = true; need = rank_bits + off_bits; }
    if (c.opt_code_by_rank) { sh.code_by_rank = true; need = rank_bits + off_bits; }
    need = std::max(need, list_bits);
    if (need > 62) throw ArgError("index: node labels too long for the locus codes");
    table_alloc(c, sh.index, n_runs, 2 * c.k, (uint64_t)n_runs / 512 + 1024, c.opt_index_slack, need);
  }
  const GraphView g = make_graph_view(c);   // after the code layout has been chosen
  if (sh.index.view.fmt == 8)
    insert_runs_kernel<8><<<grid_for(n_runs, 256), 256, 0, c.stream>>>(sh.index.view, g, key, val, run_start.p, multi_off.p, n_runs, sh.multi.p, d_err, 0u);
  else
    insert_runs_kernel<16><<<grid_for(n_runs, 256), 256, 0, c.stream>>>(sh.index.view, g, key, val, run_start.p, multi_off.p, n_runs, sh.multi.p, d_err, 0u);
  ++c.counters.launches;
  PSI_CUDA(cudaGetLastError());
  unsigned long long err = 0;
  uint32_t stash_used = 0;
  PSI_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaMemcpyAsync(&stash_used, sh.index.stash_used.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  if (err & 2ull) throw OverflowError("path index: hash stash exhausted");
  sh.index.view.stash_nonempty = stash_used ? 1u : 0u;
  sh.has_table = true;
  c.counters.n_index_entries = n;
  c.counters.n_index_kmers = n_runs;
  c.counters.n_offpath_entries = n_off;
  c.counters.index_buckets = sh.index.n_lines;
  c.counters.index_bytes = sh.index.n_lines * 128 + ((uint64_t)multi_words + 4) * 4 +
                           ((uint64_t)sh.index.view.stash_mask + 1) * sizeof(Slot16);
  c.counters.index_slot_bytes = sh.index.view.fmt;
  c.counters.index_stash_used = stash_used;
}

void engine_build_index(Ctx& c, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                        const uint32_t* head_off, const uint32_t* tail_trim)
{
  if (!c.sh->has_graph) throw StateError("set_paths: no graph");
  if (c.sh.use_count() > 1) throw StateError("set_paths: the index is shared with forked contexts");
  PSI_CUDA(cudaSetDevice(c.device));
  Shared& sh = *c.sh;
  sh.has_index = false;
  sh.has_table = false;
  sh.offpath_indexed = false;
  sh.n_on_pairs = sh.n_off_pairs = 0;
  sh.slices.clear();
  sh.slice_bits = 0;
  sh.pairs_released = false;
  c.counters.n_path_bases = c.counters.n_index_entries = c.counters.n_index_kmers = c.counters.n_offpath_entries = 0;
  c.counters.index_bytes = c.counters.index_buckets = 0;
  c.counters.ms_index_build = 0;
  if (n_paths == 0) return;
  if (!path_ptr || !path_nodes) throw ArgError("set_paths: null arrays");
  if (n_paths >= 0x7fffffffull) throw ArgError("set_paths: too many paths");
  // the C-ABI is public: nothing below may index outside what the caller handed over
  if (path_ptr[0] != 0) throw ArgError("set_paths: path_ptr[0] must be 0");
  for (uint64_t p = 0; p < n_paths; ++p)
    if (path_ptr[p + 1] < path_ptr[p]) throw ArgError("set_paths: path_ptr is not monotone");
  const uint64_t n_entries = path_ptr[n_paths];
  for (uint64_t e = 0; e < n_entries; ++e)
    if (path_nodes[e] >= sh.n_nodes) throw ArgError("set_paths: node rank out of range");

  PhaseTimer timer(c, T_INDEX);
  DevBuf<uint64_t> d_path_ptr;
  DevBuf<uint32_t> d_nodes, d_head, d_tail;
  d_path_ptr.ensure(n_paths + 1);
  d_nodes.ensure(n_entries + 1);
  PSI_CUDA(cudaMemcpyAsync(d_path_ptr.p, path_ptr, (n_paths + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
  if (n_entries) PSI_CUDA(cudaMemcpyAsync(d_nodes.p, path_nodes, n_entries * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  PathsView pv{ d_path_ptr.p, d_nodes.p, nullptr, nullptr, (uint32_t)n_paths };
  if (head_off) {
    d_head.ensure(n_paths);
    PSI_CUDA(cudaMemcpyAsync(d_head.p, head_off, n_paths * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    pv.head_off = d_head.p;
  }
  if (tail_trim) {
    d_tail.ensure(n_paths);
    PSI_CUDA(cudaMemcpyAsync(d_tail.p, tail_trim, n_paths * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    pv.tail_trim = d_tail.p;
  }
  const GraphView g = make_graph_view(c);
  unsigned long long* d_cnt = c.dev_counters.p + DC_AUX;
  PSI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));

  // The paths are indexed in groups: a group's windows are materialised, sorted and made unique (32 B of scratch per
  // window), then merged into the resident set of distinct pairs.  A group holds as many consecutive paths as fit the
  // window budget, so the usual case -- everything fits -- is one group and no merge.
  std::vector<NodeRec> h_rec(sh.n_nodes);
  PSI_CUDA(cudaMemcpyAsync(h_rec.data(), sh.node_rec.p, (size_t)sh.n_nodes * sizeof(NodeRec), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  // (head offsets / tail trims longer than their node are clamped by path_windows_kernel: such a path end is empty)
  uint64_t budget = c.opt_build_group_windows;
  if (budget == 0) {
    size_t free_b = 0, total_b = 0;
    budget = 1ull << 30;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::min<uint64_t>(budget, std::max<uint64_t>(free_b / 48, 1u << 20));
  }
  sh.on_kmer.release(); sh.on_gpos.release();
  sh.n_on_pairs = 0;
  sh.gocc_threshold = c.opt_gocc_threshold;
  sh.paths_kept = false;
  if (sh.gocc_threshold) {
    // the k-mer occurrence counts are taken when the final table is built (the loci are known then): keep the paths
    sh.path_ptr.ensure(n_paths + 1); sh.path_nodes.ensure(n_entries + 1); sh.path_head.ensure(n_paths + 1); sh.path_tail.ensure(n_paths + 1);
    PSI_CUDA(cudaMemcpyAsync(sh.path_ptr.p, d_path_ptr.p, (n_paths + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c.stream));
    if (n_entries) PSI_CUDA(cudaMemcpyAsync(sh.path_nodes.p, d_nodes.p, n_entries * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
    if (head_off) PSI_CUDA(cudaMemcpyAsync(sh.path_head.p, d_head.p, n_paths * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
    else PSI_CUDA(cudaMemsetAsync(sh.path_head.p, 0, n_paths * sizeof(uint32_t), c.stream));
    if (tail_trim) PSI_CUDA(cudaMemcpyAsync(sh.path_tail.p, d_tail.p, n_paths * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
    else PSI_CUDA(cudaMemsetAsync(sh.path_tail.p, 0, n_paths * sizeof(uint32_t), c.stream));
    sh.n_paths = n_paths;
    sh.n_path_entries = n_entries;
    sh.paths_kept = true;
  }
  uint64_t n_windows_total = 0;
  sh.slices.clear();
  sh.slice_bits = 0;
  sh.pairs_released = false;
  c.counters.index_build_slices = 1;
  // one-shot or sliced?  The bases of the paths bound their k-windows, twice the graph's bases is a generous estimate of
  // the distinct (k-mer, locus) pairs among them.
  std::vector<uint64_t> path_bases(n_paths, 0);
  uint64_t bases_total = 0;
  for (uint64_t p = 0; p < n_paths; ++p) {
    uint64_t b = 0;
    for (uint64_t e = path_ptr[p]; e < path_ptr[p + 1]; ++e) b += h_rec[path_nodes[e]].seq_len;
    path_bases[p] = b;
    bases_total += b;
  }
  uint32_t n_slices = 1;
  if (c.opt_build_slices > 0) n_slices = (uint32_t)c.opt_build_slices;
  else {
    const uint64_t est = std::min<uint64_t>(bases_total, 2 * sh.n_bases);
    if (est > (3ull << 30)) { n_slices = 4; while (n_slices < 256 && est / n_slices > (1ull << 29)) n_slices *= 4; }
  }
  if (n_slices > 1) {
    if (sh.gocc_threshold) throw ArgError("set_paths: a gocc threshold is not served by a sliced index build (build_slices)");
    uint32_t sb = 0;
    while ((1u << sb) < n_slices) sb += 2;     // whole bases: n_slices is a power of 4
    if (sb > 2 * c.k) throw ArgError("set_paths: more build slices than k-mers");
    sh.slice_bits = sb;
    c.counters.index_build_slices = n_slices;
    const uint64_t group_bases = std::max<uint64_t>(budget, 1u << 20) * n_slices / 2;
    uint64_t n_on_total = 0;
    for (uint32_t si = 0; si < n_slices; ++si) {
      auto sl = std::make_unique<Shared::PairSlice>();
      uint64_t n_acc = 0;
      for (uint64_t g0 = 0; g0 < n_paths;) {
        uint64_t g1 = g0, bases = 0;
        while (g1 < n_paths) {
          if (g1 > g0 && bases + path_bases[g1] > group_bases) break;
          bases += path_bases[g1];
          ++g1;
        }
        const uint64_t e0 = path_ptr[g0], e1 = path_ptr[g1];
        g0 = g1;
        if (e1 == e0) continue;
        const unsigned wgrid = (unsigned)std::min<uint64_t>((e1 - e0 + 7) / 8 + 1, (uint64_t)c.sm_count * 32);
        // no counting pass: the slice's share of the group is guessed and the group redone if it holds more
        uint64_t cap = bases / n_slices + bases / (4ull * n_slices) + (1u << 16);
        unsigned long long n_pairs = 0;
        DevBuf<uint64_t> kmer_a;
        DevBuf<uint32_t> gpos_a;
        for (;;) {
          if (cap + n_acc >= 0xfffffff0ull) throw ArgError("set_paths: more than 2^32 path windows in one slice of one group (raise build_slices)");
          kmer_a.ensure(cap + n_acc); gpos_a.ensure(cap + n_acc);
          PSI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));
          path_windows_kernel<0><<<wgrid, 256, 0, c.stream>>>(g, pv, c.k, e0, e1, kmer_a.p, gpos_a.p, d_cnt, KmerTable{}, nullptr,
                                                             WindowSlice{ sb, si, cap });
          ++c.counters.launches;
          PSI_CUDA(cudaGetLastError());
          PSI_CUDA(cudaMemcpyAsync(&n_pairs, d_cnt, sizeof(n_pairs), cudaMemcpyDeviceToHost, c.stream));
          PSI_CUDA(cudaStreamSynchronize(c.stream));
          if (n_pairs <= cap) break;
          cap = n_pairs;        // the kernel counted on beyond the capacity: exact now
        }
        n_windows_total += n_pairs;
        if (n_pairs == 0) continue;
        if (n_acc) {
          PSI_CUDA(cudaMemcpyAsync(kmer_a.p + n_pairs, sl->kmer.p, n_acc * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c.stream));
          PSI_CUDA(cudaMemcpyAsync(gpos_a.p + n_pairs, sl->gpos.p, n_acc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
          PSI_CUDA(cudaStreamSynchronize(c.stream));
          sl->kmer.release(); sl->gpos.release();
        }
        n_acc = sort_unique_pairs(c, kmer_a, gpos_a, n_pairs + n_acc, sl->kmer, sl->gpos);
      }
      sl->n = n_acc;
      if (n_acc) {   // distinct k-mers of the slice: sizes the table before the slices are grouped into runs
        DevBuf<uint32_t> flag, scan;
        flag.ensure(n_acc + 1); scan.ensure(n_acc + 1);
        flag_run_heads_kernel<<<grid_for(n_acc, 256), 256, 0, c.stream>>>(sl->kmer.p, n_acc, flag.p);
        ++c.counters.launches;
        PSI_CUDA(cudaMemsetAsync(flag.p + n_acc, 0, sizeof(uint32_t), c.stream));
        exclusive_scan_u32(c, flag.p, scan.p, n_acc + 1);
        uint32_t n_k = 0;
        PSI_CUDA(cudaMemcpyAsync(&n_k, scan.p + n_acc, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
        PSI_CUDA(cudaStreamSynchronize(c.stream));
        sl->n_kmers = n_k;
      }
      n_on_total += n_acc;
      sh.slices.push_back(std::move(sl));
    }
    sh.n_on_pairs = n_on_total;
    d_nodes.release(); d_path_ptr.release(); d_head.release(); d_tail.release();   // room for the table
  }
  else
  for (uint64_t g0 = 0; g0 < n_paths;) {
    uint64_t g1 = g0, bases = 0;
    while (g1 < n_paths) {
      uint64_t b = 0;
      for (uint64_t e = path_ptr[g1]; e < path_ptr[g1 + 1]; ++e) b += h_rec[path_nodes[e]].seq_len;
      if (g1 > g0 && bases + b > budget) break;
      bases += b;
      ++g1;
    }
    const uint64_t e0 = path_ptr[g0], e1 = path_ptr[g1];
    g0 = g1;
    if (e1 == e0) continue;
    // pass 1: count valid windows
    const unsigned wgrid = (unsigned)std::min<uint64_t>((e1 - e0 + 7) / 8 + 1, (uint64_t)c.sm_count * 32);
    PSI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));
    path_windows_kernel<1><<<wgrid, 256, 0, c.stream>>>(g, pv, c.k, e0, e1, nullptr, nullptr, d_cnt, KmerTable{}, nullptr, WindowSlice{ 0, 0, 0 });
    ++c.counters.launches;
    unsigned long long n_pairs = 0;
    PSI_CUDA(cudaMemcpyAsync(&n_pairs, d_cnt, sizeof(n_pairs), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (n_pairs + sh.n_on_pairs >= 0xfffffff0ull) throw ArgError("set_paths: more than 2^32 path windows in one group of paths plus the resident pairs");
    n_windows_total += n_pairs;
    if (n_pairs == 0) continue;
    // pass 2: emit (kmer, gpos), sort, unique; merge with what the earlier groups left
    const uint64_t n_acc = sh.n_on_pairs;
    DevBuf<uint64_t> kmer_a;
    DevBuf<uint32_t> gpos_a;
    kmer_a.ensure(n_pairs + n_acc); gpos_a.ensure(n_pairs + n_acc);
    PSI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));
    path_windows_kernel<0><<<wgrid, 256, 0, c.stream>>>(g, pv, c.k, e0, e1, kmer_a.p, gpos_a.p, d_cnt, KmerTable{}, nullptr, WindowSlice{ 0, 0, n_pairs });
    ++c.counters.launches;
    if (n_acc) {
      PSI_CUDA(cudaMemcpyAsync(kmer_a.p + n_pairs, sh.on_kmer.p, n_acc * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c.stream));
      PSI_CUDA(cudaMemcpyAsync(gpos_a.p + n_pairs, sh.on_gpos.p, n_acc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      sh.on_kmer.release(); sh.on_gpos.release();
    }
    sh.n_on_pairs = sort_unique_pairs(c, kmer_a, gpos_a, n_pairs + n_acc, sh.on_kmer, sh.on_gpos);
  }
  c.counters.n_path_bases = n_windows_total;
  build_table(c, nullptr, nullptr, 0);
  sh.has_index = true;
  timer.stop();
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.counters.ms_index_build = timer.ms();
}

// -------------------------------------------------------- starting loci --

struct AllLociSource {
  GraphView g;
  uint32_t step;
  __device__ bool init(uint64_t idx, WalkItem& it) const
  {
    const uint32_t pos = (uint32_t)idx;
    const uint32_t v = node_of_pos(g, pos);
    const uint32_t off = pos - __ldg(&g.rec[v].seq_start);
    if (step > 1 && off % step) return false;
    it.kmer = 0; it.origin = pos; it.node = v; it.off = off; it.depth = 0;
    return true;
  }
};

struct UncoveredSink {
  GraphView g;
  KmerTable t;
  const uint32_t* multi;
  uint32_t* flags;   // bitmap over global positions
  uint32_t has_index;
  __device__ bool skip(uint32_t origin) const
  {
    return (((volatile const uint32_t*)flags)[origin >> 5] >> (origin & 31u)) & 1u;
  }
  __device__ bool prune(uint64_t, uint32_t, uint32_t) const { return false; }
  __device__ void complete(uint64_t kmer, uint32_t origin)
  {
    if (!has_index || !index_contains(g, t, multi, kmer, origin)) atomicOr(flags + (origin >> 5), 1u << (origin & 31u));
  }
  __device__ void finish() {}
};

__global__ void __launch_bounds__(WALK_WARPS * 32)
find_loci_kernel(GraphView g, uint32_t k, uint32_t step, KmerTable t, const uint32_t* multi, uint32_t has_index,
                 uint32_t* flags, unsigned long long* work, WalkItem* spill, uint32_t spill_items,
                 unsigned long long* err)
{
  __shared__ WalkItem smem[WALK_WARPS * WALK_SMEM_ITEMS];
  AllLociSource src{ g, step };
  UncoveredSink sink{ g, t, multi, flags, has_index };
  walk_all(g, k, g.n_bases, work, smem, spill, spill_items, err, src, sink);
}

__global__ void __launch_bounds__(256)
popc_words_kernel(const uint32_t* __restrict__ flags, uint64_t n_words, uint32_t* __restrict__ cnt)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n_words) cnt[w] = __popc(flags[w]);
}

__global__ void __launch_bounds__(256)
emit_loci_kernel(GraphView g, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ scan,
                 uint64_t n_words, uint32_t* __restrict__ loci_node, uint32_t* __restrict__ loci_off)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t m = flags[w];
  uint32_t dst = scan[w];
  while (m) {
    const uint32_t b = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t pos = (uint32_t)(w << 5) + b;
    const uint32_t v = node_of_pos(g, pos);
    loci_node[dst] = v;
    loci_off[dst] = pos - g.rec[v].seq_start;
    ++dst;
  }
}

// ------------------------------------------- off-path walks -> index --
//
// seeds_off_paths of the reference re-walks the graph from every starting locus
// for every read chunk (seed_finder.hpp:1703-1722).  The set of k-mers those
// walks spell does not depend on the reads, so it is enumerated ONCE here -- same
// walker, same rules -- and every (k-mer, start locus) pair that the path index
// does not already hold is merged into the index as an off-path entry.  A chunk
// is then answered by one probe per seed, on-path and off-path alike.  When the
// walks are too many to store (hypervariable graphs, large k) the per-chunk walk
// of seeds.cu is used instead (offpath_mode 1); both give the same seed set.

struct OffPathSink {
  GraphView g;
  KmerTable t;
  const uint32_t* multi;
  uint32_t has_table;
  uint64_t* out_kmer;           // null: count only
  uint32_t* out_gpos;
  uint64_t cap;
  unsigned long long* count;
  __device__ bool skip(uint32_t) const { return false; }
  __device__ bool prune(uint64_t, uint32_t, uint32_t) const { return false; }
  __device__ void complete(uint64_t kmer, uint32_t origin)
  {
    if (has_table && index_contains(g, t, multi, kmer, origin)) return;
    const unsigned long long slot = atomicAdd(count, 1ull);
    if (out_kmer && slot < cap) { out_kmer[slot] = kmer; out_gpos[slot] = origin; }
  }
  __device__ void finish() {}
};

__global__ void __launch_bounds__(WALK_WARPS * 32)
collect_offpath_kernel(GraphView g, uint32_t k, uint64_t n_loci, const uint32_t* loci_node, const uint32_t* loci_off,
                       OffPathSink sink, unsigned long long* work, WalkItem* spill, uint32_t spill_items,
                       unsigned long long* err)
{
  __shared__ WalkItem smem[WALK_WARPS * WALK_SMEM_ITEMS];
  LociSource src{ g, loci_node, loci_off };
  walk_all(g, k, n_loci, work, smem, spill, spill_items, err, src, sink);
}

// Decide the off-path mode for the current loci and, in index mode, merge the walks into the table.
static void materialise_offpath(Ctx& c)
{
  Shared& sh = *c.sh;
  sh.offpath_indexed = false;
  c.counters.n_offpath_walks = 0;
  c.counters.offpath_mode = 1;
  const bool had_off = sh.n_off_pairs != 0;
  const bool gocc = sh.gocc_threshold != 0 && sh.has_index;
  c.counters.n_gocc_dropped = 0;
  auto walk_mode_refused = [&] {
    if (gocc) throw ArgError("a gocc threshold needs the off-path walks materialised in the index (offpath_mode 0 or 2 and a budget that holds them)");
  };
  auto drop_off = [&](bool final) { if (had_off || !sh.has_table || sh.table_filtered || (final && gocc)) build_table(c, nullptr, nullptr, 0, final); };
  if (sh.n_loci == 0) { drop_off(true); sh.offpath_indexed = true; c.counters.offpath_mode = 2; return; }
  if (c.opt_offpath_mode == 1) { walk_mode_refused(); drop_off(false); return; }
  // the walks are compared with the UNFILTERED path index
  if (!sh.has_table || sh.table_filtered) build_table(c, nullptr, nullptr, 0);

  const GraphView g = make_graph_view(c);
  const unsigned grid = (unsigned)c.sm_count * 8;
  unsigned long long* d_err = c.dev_counters.p + DC_ERR;
  unsigned long long* d_work = c.dev_counters.p + DC_WORK;
  unsigned long long* d_cnt = c.dev_counters.p + DC_AUX;
  DevBuf<uint64_t> kmer_a;
  DevBuf<uint32_t> gpos_a;
  uint64_t n_walks = 0;
  // pass 0 counts the uncovered walks, pass 1 stores them
  for (int pass = 0; pass < 2; ++pass) {
    while (true) {
      PSI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), c.stream));
      PSI_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), c.stream));
      PSI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));
      c.walk_spill.ensure((size_t)grid * WALK_WARPS * c.spill_items * sizeof(WalkItem));
      OffPathSink sink{ g, sh.index.view, sh.multi.p, sh.has_table ? 1u : 0u, pass ? kmer_a.p : nullptr, gpos_a.p, n_walks, d_cnt };
      collect_offpath_kernel<<<grid, WALK_WARPS * 32, 0, c.stream>>>(g, c.k, sh.n_loci, sh.loci_node.p, sh.loci_off.p, sink,
                                                                    d_work, (WalkItem*)c.walk_spill.p, c.spill_items, d_err);
      ++c.counters.launches;
      PSI_CUDA(cudaGetLastError());
      unsigned long long h[2] = { 0, 0 };
      PSI_CUDA(cudaMemcpyAsync(&h[0], d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
      PSI_CUDA(cudaMemcpyAsync(&h[1], d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
      PSI_CUDA(cudaStreamSynchronize(c.stream));
      if (h[0] & 1ull) {
        if (c.spill_items >= (1u << 20)) throw OverflowError("off-path walks: frontier exceeds 2^20 states per warp");
        c.spill_items *= 4;
        continue;
      }
      if (pass == 0) n_walks = h[1];
      else if (h[1] != n_walks) throw StateError("off-path walks: enumeration is not reproducible");
      break;
    }
    if (pass == 0) {
      c.counters.n_offpath_walks = n_walks;
      if (n_walks >= 0xfffffff0ull || (c.opt_offpath_mode == 0 && n_walks > c.opt_offpath_max_pairs)) { walk_mode_refused(); drop_off(false); return; }
      if (n_walks == 0) break;
      kmer_a.ensure(n_walks); gpos_a.ensure(n_walks);
    }
  }
  DevBuf<uint64_t> off_kmer;
  DevBuf<uint32_t> off_gpos;
  const uint64_t n_off = sort_unique_pairs(c, kmer_a, gpos_a, n_walks, off_kmer, off_gpos);
  kmer_a.release(); gpos_a.release();
  if (n_off || had_off || gocc) build_table(c, off_kmer.p, off_gpos.p, n_off, true);
  sh.offpath_indexed = true;
  c.counters.offpath_mode = 2;
}

// A large sliced build gives its sorted pairs back once the table that the chunks will probe stands (12 bytes per pair:
// 60 GB at whole-genome scale); changing the loci afterwards needs set_paths again.
static void release_sliced_pairs(Ctx& c)
{
  Shared& sh = *c.sh;
  if (sh.slices.empty() || sh.n_on_pairs < (1ull << 31)) return;
  sh.slices.clear();
  sh.pairs_released = true;
}

void engine_find_loci(Ctx& c, unsigned step)
{
  if (!c.sh->has_graph) throw StateError("find_loci: no graph");
  if (c.sh.use_count() > 1) throw StateError("find_loci: the index is shared with forked contexts");
  PSI_CUDA(cudaSetDevice(c.device));
  if (step == 0) step = 1;
  PhaseTimer timer(c, T_LOCI);
  const uint64_t n_words = (c.sh->n_bases + 31) >> 5;
  DevBuf<uint32_t> flags, cnt, scan;
  flags.ensure(n_words + 1); cnt.ensure(n_words + 1); scan.ensure(n_words + 1);
  const GraphView g = make_graph_view(c);
  const unsigned grid = (unsigned)c.sm_count * 8;
  unsigned long long* d_err = c.dev_counters.p + DC_ERR;
  unsigned long long* d_work = c.dev_counters.p + DC_WORK;
  while (true) {
    PSI_CUDA(cudaMemsetAsync(flags.p, 0, (n_words + 1) * sizeof(uint32_t), c.stream));
    PSI_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), c.stream));
    PSI_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), c.stream));
    c.walk_spill.ensure((size_t)grid * WALK_WARPS * c.spill_items * sizeof(WalkItem));
    find_loci_kernel<<<grid, WALK_WARPS * 32, 0, c.stream>>>(g, c.k, step, c.sh->index.view, c.sh->multi.p, c.sh->has_table ? 1u : 0u,
                                                            flags.p, d_work, (WalkItem*)c.walk_spill.p, c.spill_items, d_err);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    unsigned long long err = 0;
    PSI_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (!(err & 1ull)) break;
    if (c.spill_items >= (1u << 20)) throw OverflowError("find_loci: walk frontier exceeds 2^20 states per warp");
    c.spill_items *= 4;  // frontier overflow: retry with a larger spill area
  }
  popc_words_kernel<<<grid_for(n_words, 256), 256, 0, c.stream>>>(flags.p, n_words, cnt.p);
  ++c.counters.launches;
  PSI_CUDA(cudaMemsetAsync(cnt.p + n_words, 0, sizeof(uint32_t), c.stream));
  exclusive_scan_u32(c, cnt.p, scan.p, n_words + 1);
  uint32_t n_loci = 0;
  PSI_CUDA(cudaMemcpyAsync(&n_loci, scan.p + n_words, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.sh->loci_node.ensure((uint64_t)n_loci + 1);
  c.sh->loci_off.ensure((uint64_t)n_loci + 1);
  if (n_loci) {
    emit_loci_kernel<<<grid_for(n_words, 256), 256, 0, c.stream>>>(g, flags.p, scan.p, n_words, c.sh->loci_node.p, c.sh->loci_off.p);
    ++c.counters.launches;
  }
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.sh->n_loci = n_loci;
  c.counters.n_loci = n_loci;
  materialise_offpath(c);
  release_sliced_pairs(c);
  timer.stop();
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.counters.ms_find_loci = timer.ms();
}

void engine_set_loci(Ctx& c, uint64_t n, const uint32_t* node, const uint32_t* off)
{
  if (!c.sh->has_graph) throw StateError("set_loci: no graph");
  if (c.sh.use_count() > 1) throw StateError("set_loci: the index is shared with forked contexts");
  if (n && (!node || !off)) throw ArgError("set_loci: null arrays");
  PSI_CUDA(cudaSetDevice(c.device));
  c.sh->loci_node.ensure(n + 1);
  c.sh->loci_off.ensure(n + 1);
  if (n) {
    PSI_CUDA(cudaMemcpyAsync(c.sh->loci_node.p, node, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    PSI_CUDA(cudaMemcpyAsync(c.sh->loci_off.p, off, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  }
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.sh->n_loci = n;
  c.counters.n_loci = n;
  materialise_offpath(c);
  release_sliced_pairs(c);
}

void engine_set_option(Ctx& c, const char* name, long long value)
{
  if (!name) throw ArgError("set_option: null name");
  const std::string n = name;
  if (n == "offpath_mode") {
    if (value < 0 || value > 2) throw ArgError("set_option: offpath_mode is 0 (auto), 1 (walk per chunk) or 2 (materialise)");
    c.opt_offpath_mode = (int)value;
  }
  else if (n == "offpath_max_pairs") {
    if (value < 0) throw ArgError("set_option: offpath_max_pairs must be >= 0");
    c.opt_offpath_max_pairs = (uint64_t)value;
  }
  else if (n == "dindex_mode") {
    if (value < 0 || value > 2) throw ArgError("set_option: dindex_mode is 0 (auto), 1 (queries enumerate) or 2 (rows materialised)");
    c.opt_dindex_mode = (int)value;     // takes effect at the next create_distance_index
  }
  else if (n == "dindex_list_cap") {
    if (value < 64 || value > 1024 || (value & (value - 1))) throw ArgError("set_option: dindex_list_cap is a power of two in [64, 1024]");
    c.opt_dindex_list_cap = (uint32_t)value;
  }
  else if (n == "dindex_max_bytes") {
    if (value < 0) throw ArgError("set_option: dindex_max_bytes must be >= 0 (0 = half of the free device memory)");
    c.opt_dindex_max_bytes = (uint64_t)value;
  }
  else if (n == "build_slices") {
    if (value != 0 && value != 1 && value != 4 && value != 16 && value != 64 && value != 256)
      throw ArgError("set_option: build_slices is 0 (auto), 1, 4, 16, 64 or 256");
    c.opt_build_slices = (int)value;     // takes effect at the next set_paths
  }
  else if (n == "build_group_windows") {
    if (value < 0) throw ArgError("set_option: build_group_windows must be >= 0 (0 = from the free device memory)");
    c.opt_build_group_windows = (uint64_t)value;     // takes effect at the next set_paths
  }
  else if (n == "index_slack") {
    if (value < -1 || value > 3) throw ArgError("set_option: index_slack is -1 (auto), 0, 1, 2 or 3 extra doublings of the bucket count");
    c.opt_index_slack = (int)value;     // takes effect at the next set_paths / find_loci / set_loci
  }
  else if (n == "blocking_sync") {
    c.opt_blocking_sync = value != 0;
  }
  else if (n == "fused") {
    c.opt_fused = value != 0;
  }
  else if (n == "timers") {
    c.opt_timers = value != 0;
  }
  else if (n == "gocc_threshold") {
    if (value < 0 || value > 0xffffffffll) throw ArgError("set_option: gocc_threshold must be in [0, 2^32)");
    c.opt_gocc_threshold = (uint32_t)value;     // takes effect at the next set_paths
  }
  else if (n == "code_by_rank") {
    c.opt_code_by_rank = value != 0;     // takes effect at the next set_paths / find_loci / set_loci
  }
  else if (n == "seeding_mode") {
    if (value < 0 || value > 1) throw ArgError("set_option: seeding_mode is 0 (direct from ASCII) or 1 (2-bit staging)");
    c.opt_seeding_mode = (int)value;
  }
  else if (n == "resolve_ctas") {
    if (value != 5 && value != 6) throw ArgError("set_option: resolve_ctas is 5 or 6");
    c.opt_resolve_ctas = (int)value;
  }
  else if (n == "resolve_items") {
    if (value != 2 && value != 4) throw ArgError("set_option: resolve_items is 2 or 4");
    c.opt_resolve_items = (int)value;
  }
  else throw ArgError("set_option: unknown option '" + n + "'");
}

void engine_get_loci(Ctx& c, uint32_t* node, uint32_t* off, uint64_t cap)
{
  PSI_CUDA(cudaSetDevice(c.device));
  const uint64_t n = c.sh->n_loci < cap ? c.sh->n_loci : cap;
  if (n && node) PSI_CUDA(cudaMemcpyAsync(node, c.sh->loci_node.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  if (n && off) PSI_CUDA(cudaMemcpyAsync(off, c.sh->loci_off.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace psi_b200
