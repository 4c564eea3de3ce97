// mems.cu -- MEM mode: SeedFinder::seeds_on_paths(sequence, callback) -> find_mems
// (reference include/psi/seed_finder.hpp:1459-1479, include/psi/index_iter.hpp:854-906).
//
// The reference walks the FM-index of the (reversed) path text with a greedy left-to-right scan of the read:
//   extend pattern[start : start + plen] one character at a time while it still occurs in the path text; as soon as it
//   is at least minlen (= the seed length) long and occurs at most gocc_threshold times, report ALL its occurrences
//   (node, offset of the occurrence's first base; match_len = plen; gocc = number of occurrences), stop after max_mem
//   hits, and restart one character further on; a character that cannot be appended, or an 'N', restarts the scan
//   behind it.
// The fixed-k hash of the seed path cannot answer "does this prefix of arbitrary length occur", so MEM mode has its own
// device index: the SUFFIX TABLE of the path text -- every text position with the (up to) 32 characters that follow
// it on its path as one 64-bit key (2 bits per character, first character in the top bits, so that lexicographic order
// is numeric order), the number of valid characters (a path end or an 'N' cuts a suffix short) and the graph position --
// sorted by (key, valid length).  A prefix of length p <= 32 is then a contiguous range of the table found by two binary
// searches inside the range of the previous, shorter prefix; a direct-addressed table of the 4^12 twelve-character
// prefixes replaces the first twelve descents by two loads each.  Beyond 32 characters (only reached when a gocc
// threshold keeps a scan going) the range is filtered by reading on along each occurrence's path.
// One thread scans one read; the raw hits (one per text occurrence, like the reference's) are sorted and made unique
// on the device, so the result is the SET of (read, offset, length, node, offset) -- every hit once -- with gocc kept
// as the raw occurrence count.
#include "engine.hpp"
#include "records.cuh"

#include <algorithm>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace psi_b200 {

using namespace dev;

constexpr uint32_t MEM_KEY_CHARS = 32;
constexpr uint32_t MEM_PFX = 12;          // direct-addressed prefixes: 4^12 + 1 offsets (64 MB)

struct MemPaths {
  const uint64_t* path_ptr;
  const uint32_t* nodes;
  const uint32_t* head;
  const uint32_t* tail;
  const uint64_t* entry_start;   // text position of the first character an entry contributes
  uint32_t n_paths;
};

__device__ __forceinline__ uint32_t mem_path_of_entry(const MemPaths& p, uint64_t e)
{
  uint32_t lo = 0, hi = p.n_paths;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(p.path_ptr + mid) <= e) lo = mid; else hi = mid; }
  return lo;
}

// characters an entry (one node visit of a path) contributes to the path text: [from, to) of the node's label
__device__ __forceinline__ void mem_entry_span(const GraphView& g, const MemPaths& p, uint64_t e, uint32_t pi, uint32_t& from, uint32_t& to)
{
  const uint64_t pbeg = __ldg(p.path_ptr + pi), pend = __ldg(p.path_ptr + pi + 1);
  const uint32_t len = g.rec[__ldg(p.nodes + e)].seq_len;
  from = e == pbeg ? min(__ldg(p.head + pi), len) : 0u;
  to = len;
  if (e + 1 == pend) { const uint32_t t = __ldg(p.tail + pi); to = t < to ? to - t : 0u; }
  if (to < from) to = from;
}

__global__ void __launch_bounds__(256)
mem_entry_len_kernel(GraphView g, MemPaths p, uint64_t n_entries, uint32_t* __restrict__ len)
{
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  uint32_t from, to;
  mem_entry_span(g, p, e, mem_path_of_entry(p, e), from, to);
  len[e] = to - from;
}

// One warp per path entry, lanes over its offsets: the key (next <= 32 characters along the PATH), the number of valid
// characters and where the suffix starts.
__global__ void __launch_bounds__(256)
mem_suffix_kernel(GraphView g, MemPaths p, uint64_t n_entries, uint64_t* __restrict__ key, uint8_t* __restrict__ vlen,
                  uint32_t* __restrict__ gpos, uint32_t* __restrict__ ent)
{
  const uint32_t lane = lane_id();
  const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t e = warp0; e < n_entries; e += n_warps) {
    const uint32_t pi = mem_path_of_entry(p, e);
    const uint64_t pend = __ldg(p.path_ptr + pi + 1);
    uint32_t from, to;
    mem_entry_span(g, p, e, pi, from, to);
    const NodeRec r = g.rec[__ldg(p.nodes + e)];
    const uint64_t t0 = __ldg(p.entry_start + e);
    for (uint32_t o = from + lane; o < to; o += 32) {
      uint64_t k = 0;
      uint32_t depth = 0, off = o, cto = to;
      uint64_t ce = e;
      NodeRec cr = r;
      while (depth < MEM_KEY_CHARS) {
        const uint32_t avail = cto > off ? cto - off : 0;
        uint32_t take = min(avail, MEM_KEY_CHARS - depth);
        if (take) {
          const uint64_t pos = (uint64_t)cr.seq_start + off;
          if (g.has_n) {
            const uint32_t nm = extract_nmask(g.nmask, pos, take);
            if (nm) take = __ffs(nm) - 1;              // characters before the first one outside A/C/G/T
          }
          const uint64_t bits = take ? extract_bases(g.seq2, pos, take) : 0;    // base i at bits [2i, 2i + 2)
          for (uint32_t i = 0; i < take; ++i) k |= ((bits >> (2 * i)) & 3ull) << (62 - 2 * (depth + i));
          depth += take;
          if (take < min(avail, MEM_KEY_CHARS - (depth - take))) break;         // stopped at an N
        }
        if (depth == MEM_KEY_CHARS) break;
        if (++ce == pend) break;                                                  // path end
        uint32_t f2, t2;
        mem_entry_span(g, p, ce, pi, f2, t2);
        cr = g.rec[__ldg(p.nodes + ce)];
        off = f2;
        cto = t2;
      }
      const uint64_t t = t0 + (o - from);
      key[t] = k;
      vlen[t] = (uint8_t)depth;
      gpos[t] = r.seq_start + o;
      ent[t] = (uint32_t)e;
    }
  }
}

__global__ void __launch_bounds__(256)
mem_iota_kernel(uint32_t* __restrict__ idx, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (uint32_t)i;
}

template <class T>
__global__ void __launch_bounds__(256)
mem_gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, T* __restrict__ dst)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256)
mem_prefix_table_kernel(const uint64_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ pstart)
{
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_pfx = 1ull << (2 * MEM_PFX);
  if (x > n_pfx) return;
  if (x == n_pfx) { pstart[x] = (uint32_t)n; return; }
  const uint64_t want = x << (64 - 2 * MEM_PFX);
  uint64_t lo = 0, hi = n;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (key[mid] < want) lo = mid + 1; else hi = mid; }
  pstart[x] = (uint32_t)lo;
}

struct MemIndexView {
  const uint64_t* key;
  const uint8_t* vlen;
  const uint32_t* gpos;
  const uint32_t* ent;
  const uint32_t* pstart;
  uint64_t n;
};

// first index in [lo, hi) whose (key, vlen) is >= (k, v)
__device__ __forceinline__ uint64_t mem_lower_bound(const MemIndexView& ix, uint64_t lo, uint64_t hi, uint64_t k, uint32_t v)
{
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    const uint64_t km = __ldg(ix.key + mid);
    if (km < k || (km == k && (uint32_t)__ldg(ix.vlen + mid) < v)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// character j (>= 0) after the start of the occurrence in table slot i, along its path; 4 = none (path end / N)
__device__ uint32_t mem_char_along(const GraphView& g, const MemPaths& p, const MemIndexView& ix, uint64_t i, uint32_t j)
{
  uint64_t e = __ldg(ix.ent + i);
  const uint32_t pi = mem_path_of_entry(p, e);
  const uint64_t pend = __ldg(p.path_ptr + pi + 1);
  uint32_t from, to;
  mem_entry_span(g, p, e, pi, from, to);
  NodeRec r = g.rec[__ldg(p.nodes + e)];
  uint32_t off = __ldg(ix.gpos + i) - r.seq_start;
  while (true) {
    const uint32_t avail = to > off ? to - off : 0;
    if (j < avail) {
      const uint64_t pos = (uint64_t)r.seq_start + off + j;
      if (g.has_n && extract_nmask(g.nmask, pos, 1)) return 4;
      return (uint32_t)extract_bases(g.seq2, pos, 1);
    }
    j -= avail;
    if (++e == pend) return 4;
    mem_entry_span(g, p, e, pi, from, to);
    r = g.rec[__ldg(p.nodes + e)];
    off = from;
  }
}

struct MemChunk {
  const char* bases;
  const uint64_t* words;
  const uint64_t* read_ptr;
  const uint64_t* exc;
  uint64_t n_exc;
  uint64_t n_reads;
  uint32_t read_len;
};

struct alignas(16) MemRaw {
  uint64_t key;      // read (local) << 32 | start << 16 | length
  uint32_t gpos;
  uint32_t gocc;
};

// One thread per read: the scan of find_mems (index_iter.hpp:854-906).  pass 0 counts the raw hits, pass 1 stores them.
__global__ void __launch_bounds__(128)
find_mems_kernel(GraphView g, MemPaths p, MemIndexView ix, MemChunk ch, uint32_t minlen, uint32_t gocc_threshold, uint32_t max_mem,
                 MemRaw* __restrict__ out, uint64_t cap, unsigned long long* __restrict__ n_out)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ch.n_reads) return;
  const uint64_t p0 = ch.read_ptr ? ch.read_ptr[r] : r * ch.read_len;
  const uint32_t len = (uint32_t)((ch.read_ptr ? ch.read_ptr[r + 1] : p0 + ch.read_len) - p0);
  // this read's slice of the exception list (2-bit chunks)
  uint64_t e0 = 0, e1 = 0;
  if (ch.words && ch.n_exc) {
    uint64_t lo = 0, hi = ch.n_exc;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(ch.exc + mid) < p0) lo = mid + 1; else hi = mid; }
    e0 = lo;
    hi = ch.n_exc;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(ch.exc + mid) < p0 + len) lo = mid + 1; else hi = mid; }
    e1 = lo;
  }
  auto read_char = [&](uint32_t i) -> uint32_t {       // 0..3, or 4 for anything outside A/C/G/T
    const uint64_t pos = p0 + i;
    if (ch.words) {
      for (uint64_t e = e0; e < e1; ++e) if (__ldg(ch.exc + e) == pos) return 4u;
      return (uint32_t)((__ldg(ch.words + (pos >> 5)) >> (2 * (pos & 31))) & 3ull);
    }
    return base_code((unsigned char)ch.bases[pos]);
  };
  if (gocc_threshold == 0) gocc_threshold = 0xffffffffu;
  if (max_mem == 0) max_mem = 0xffffffffu;

  uint32_t start = 0, plen = 0;
  bool has_hit = false;
  uint64_t nof = 0;
  uint64_t q = 0, lo = 0, hi = ix.n;      // range of pattern[start : start + min(plen, 32)]
  // number of occurrences of the current pattern: the range, filtered by the characters beyond the 32nd
  auto matches_beyond = [&](uint64_t i) -> bool {
    for (uint32_t j = MEM_KEY_CHARS; j < plen; ++j) if (mem_char_along(g, p, ix, i, j) != read_char(start + j)) return false;
    return true;
  };
  auto count_occ = [&]() -> uint64_t {
    if (plen <= MEM_KEY_CHARS) return hi - lo;
    uint64_t c = 0;
    for (uint64_t i = lo; i < hi; ++i) c += matches_beyond(i) ? 1 : 0;
    return c;
  };
  while (start + plen < len) {
    if (plen == 0 && !has_hit) {
      // A fresh pattern: nothing is reported below minlen characters, and a pattern occurs only if all its prefixes do,
      // so the first J = min(MEM_PFX, minlen) characters are appended in ONE step when pattern[start : start + J]
      // occurs -- one bucket of the prefix table instead of J rounds of searches in quarter-of-the-table brackets.  If
      // it does not occur (or holds a character outside A/C/G/T) the character-by-character scan below finds where.
      const uint32_t J = minlen < MEM_PFX ? minlen : MEM_PFX;
      if (J >= 2 && start + J <= len) {
        uint64_t qn = 0;
        bool ok = true;
        for (uint32_t j = 0; j < J && ok; ++j) {
          const uint32_t c = read_char(start + j);
          ok = c < 4;
          qn |= (uint64_t)c << (62 - 2 * j);
        }
        if (ok) {
          const uint64_t x0 = qn >> (64 - 2 * MEM_PFX);
          const uint64_t blo = __ldg(ix.pstart + x0), bhi = __ldg(ix.pstart + x0 + (1ull << (2 * (MEM_PFX - J))));
          const uint64_t nlo = mem_lower_bound(ix, blo, bhi, qn, J);
          const uint64_t step = 1ull << (64 - 2 * J);
          const uint64_t nhi = (qn + step < qn) ? bhi : mem_lower_bound(ix, nlo, bhi, qn + step, 0);
          if (nlo < nhi) { q = qn; lo = nlo; hi = nhi; plen = J; continue; }
        }
      }
    }
    if (plen >= minlen) {
      const uint64_t cnt = count_occ();
      if (cnt <= gocc_threshold) {
        has_hit = true;
        for (uint64_t i = lo; i < hi; ++i) {
          if (plen > MEM_KEY_CHARS && !matches_beyond(i)) continue;
          const unsigned long long slot = atomicAdd(n_out, 1ull);
          if (slot < cap) out[slot] = MemRaw{ (r << 32) | ((uint64_t)start << 16) | plen, __ldg(ix.gpos + i), (uint32_t)min(cnt, (uint64_t)0xffffffffu) };
        }
        nof += cnt;
        if (nof >= max_mem) break;
      }
    }
    bool down = false;
    if (!has_hit) {
      const uint32_t c = read_char(start + plen);
      if (c < 4) {
        if (plen < MEM_KEY_CHARS) {
          const uint32_t pl = plen + 1;
          const uint64_t qn = q | ((uint64_t)c << (62 - 2 * plen));
          uint64_t blo = lo, bhi = hi;
          if (pl <= MEM_PFX) {                         // the bracket comes from the direct-addressed prefix table
            const uint64_t x0 = qn >> (64 - 2 * MEM_PFX);
            blo = __ldg(ix.pstart + x0);
            bhi = __ldg(ix.pstart + x0 + (1ull << (2 * (MEM_PFX - pl))));
          }
          const uint64_t nlo = mem_lower_bound(ix, blo, bhi, qn, pl);
          const uint64_t step = pl < 32 ? 1ull << (64 - 2 * pl) : 1ull;
          const uint64_t nhi = (qn + step < qn) ? bhi : mem_lower_bound(ix, nlo, bhi, qn + step, 0);     // wrap: the last prefix
          if (nlo < nhi) { down = true; q = qn; lo = nlo; hi = nhi; }
        }
        else {
          // beyond the key: some occurrence of the current pattern must go on with c
          for (uint64_t i = lo; i < hi && !down; ++i) down = matches_beyond(i) && mem_char_along(g, p, ix, i, plen) == c;
        }
      }
    }
    if (!down) {          // has_hit, an N, or the character cannot be appended: restart behind it
      start = start + plen + 1;
      plen = 0;
      has_hit = false;
      q = 0; lo = 0; hi = ix.n;
      continue;
    }
    ++plen;
  }
}

__global__ void __launch_bounds__(256)
mem_split_kernel(const MemRaw* __restrict__ raw, uint64_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ gpos)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = raw[i].key; gpos[i] = raw[i].gpos; }
}

__global__ void __launch_bounds__(256)
mem_flag_unique_kernel(const MemRaw* __restrict__ raw, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ flag)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const MemRaw a = raw[idx[i]];
  bool first = i == 0;
  if (!first) { const MemRaw b = raw[idx[i - 1]]; first = a.key != b.key || a.gpos != b.gpos; }
  flag[i] = first ? 1u : 0u;
}

// psi_b200_mem_hit records: node_id, node_off, read_id, read_off, match_len, gocc (6 x u64)
__global__ void __launch_bounds__(256)
mem_records_kernel(GraphView g, const MemRaw* __restrict__ raw, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ flag,
                   const uint32_t* __restrict__ scan, uint64_t n, uint64_t first_read_id, uint64_t* __restrict__ records)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const MemRaw a = raw[idx[i]];
  uint64_t id, off;
  resolve_node(g, a.gpos, id, off);
  uint64_t* o = records + 6 * (uint64_t)scan[i];
  o[0] = id;
  o[1] = off;
  o[2] = first_read_id + (a.key >> 32);
  o[3] = (a.key >> 16) & 0xffffu;
  o[4] = a.key & 0xffffu;
  o[5] = a.gocc;
}

// ---------------------------------------------------------------- host --

static void mem_scan_u32(Ctx& c, const uint32_t* in, uint32_t* out, uint64_t n)
{
  size_t tmp = 0;
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, c.stream));
  c.scan_tmp.ensure(tmp);
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(c.scan_tmp.p, tmp, in, out, (int64_t)n, c.stream));
}

// PathIndex::create_index for the MEM queries (pathindex.hpp:235-243): the suffix table of the path text.
void engine_build_mem_index(Ctx& c, uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* path_nodes,
                            const uint32_t* head_off, const uint32_t* tail_trim)
{
  if (!c.sh->has_graph) throw StateError("build_mem_index: no graph");
  if (c.sh.use_count() > 1) throw StateError("build_mem_index: the index is shared with forked contexts");
  if (n_paths && (!path_ptr || !path_nodes)) throw ArgError("build_mem_index: null arrays");
  if (n_paths >= 0x7fffffffull) throw ArgError("build_mem_index: too many paths");
  PSI_CUDA(cudaSetDevice(c.device));
  Shared& sh = *c.sh;
  sh.has_mem_index = false;
  sh.mem_n = 0;
  if (n_paths == 0) { sh.has_mem_index = true; sh.mem_paths_n = 0; return; }
  if (path_ptr[0] != 0) throw ArgError("build_mem_index: path_ptr[0] must be 0");
  for (uint64_t p = 0; p < n_paths; ++p) if (path_ptr[p + 1] < path_ptr[p]) throw ArgError("build_mem_index: path_ptr is not monotone");
  const uint64_t n_entries = path_ptr[n_paths];
  if (n_entries >= 0xfffffff0ull) throw ArgError("build_mem_index: too many path entries");
  for (uint64_t e = 0; e < n_entries; ++e) if (path_nodes[e] >= sh.n_nodes) throw ArgError("build_mem_index: node rank out of range");
  nvtxRangePushA("index-paths (MEM suffix table)");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop_range;

  sh.mem_path_ptr.ensure(n_paths + 1); sh.mem_nodes.ensure(n_entries + 1); sh.mem_head.ensure(n_paths + 1); sh.mem_tail.ensure(n_paths + 1);
  sh.mem_entry_start.ensure(n_entries + 2);
  PSI_CUDA(cudaMemcpyAsync(sh.mem_path_ptr.p, path_ptr, (n_paths + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
  if (n_entries) PSI_CUDA(cudaMemcpyAsync(sh.mem_nodes.p, path_nodes, n_entries * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  if (head_off) PSI_CUDA(cudaMemcpyAsync(sh.mem_head.p, head_off, n_paths * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  else PSI_CUDA(cudaMemsetAsync(sh.mem_head.p, 0, n_paths * sizeof(uint32_t), c.stream));
  if (tail_trim) PSI_CUDA(cudaMemcpyAsync(sh.mem_tail.p, tail_trim, n_paths * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  else PSI_CUDA(cudaMemsetAsync(sh.mem_tail.p, 0, n_paths * sizeof(uint32_t), c.stream));
  sh.mem_paths_n = (uint32_t)n_paths;
  const GraphView g = make_graph_view(c);
  MemPaths mp{ sh.mem_path_ptr.p, sh.mem_nodes.p, sh.mem_head.p, sh.mem_tail.p, nullptr, (uint32_t)n_paths };

  // text position of every entry: exclusive scan of the entries' lengths (64-bit: the text may exceed 2^32 characters
  // only beyond what the 32-bit table offsets allow, which is checked below)
  DevBuf<uint32_t> len;
  len.ensure(n_entries + 1);
  mem_entry_len_kernel<<<grid_for(n_entries, 256), 256, 0, c.stream>>>(g, mp, n_entries, len.p);
  std::vector<uint32_t> h_len(n_entries);
  PSI_CUDA(cudaMemcpyAsync(h_len.data(), len.p, n_entries * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<uint64_t> h_start(n_entries + 1);
  uint64_t T = 0;
  for (uint64_t e = 0; e < n_entries; ++e) { h_start[e] = T; T += h_len[e]; }
  h_start[n_entries] = T;
  if (T >= 0xfffffff0ull) throw ArgError("build_mem_index: path text of more than 2^32 characters");
  PSI_CUDA(cudaMemcpyAsync(sh.mem_entry_start.p, h_start.data(), (n_entries + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
  mp.entry_start = sh.mem_entry_start.p;
  sh.mem_n = T;
  const uint64_t n_pfx = (1ull << (2 * MEM_PFX)) + 1;
  sh.mem_pstart.ensure(n_pfx + 1);
  if (T == 0) {
    PSI_CUDA(cudaMemsetAsync(sh.mem_pstart.p, 0, n_pfx * sizeof(uint32_t), c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    sh.has_mem_index = true;
    return;
  }

  // suffixes in text order, then the permutation that sorts them by (key, valid length): LSD, two stable passes
  DevBuf<uint64_t> key;
  DevBuf<uint8_t> vlen;
  DevBuf<uint32_t> gpos, ent, idx_a, idx_b;
  key.ensure(T); vlen.ensure(T); gpos.ensure(T); ent.ensure(T); idx_a.ensure(T); idx_b.ensure(T);
  const unsigned wgrid = (unsigned)std::min<uint64_t>((n_entries + 7) / 8 + 1, (uint64_t)c.sm_count * 32);
  mem_suffix_kernel<<<wgrid, 256, 0, c.stream>>>(g, mp, n_entries, key.p, vlen.p, gpos.p, ent.p);
  mem_iota_kernel<<<grid_for(T, 256), 256, 0, c.stream>>>(idx_a.p, T);
  {
    DevBuf<uint8_t> vlen_s;
    DevBuf<uint64_t> key_g, key_s;
    vlen_s.ensure(T); key_g.ensure(T); key_s.ensure(T);
    size_t t1 = 0, t2 = 0;
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, vlen.p, vlen_s.p, idx_a.p, idx_b.p, (int64_t)T, 0, 8, c.stream));
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, key_g.p, key_s.p, idx_b.p, idx_a.p, (int64_t)T, 0, 64, c.stream));
    DevBuf<char> tmp;
    tmp.ensure(std::max(t1, t2));
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t1, vlen.p, vlen_s.p, idx_a.p, idx_b.p, (int64_t)T, 0, 8, c.stream));
    mem_gather_kernel<uint64_t><<<grid_for(T, 256), 256, 0, c.stream>>>(key.p, idx_b.p, T, key_g.p);
    PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t2, key_g.p, key_s.p, idx_b.p, idx_a.p, (int64_t)T, 0, 64, c.stream));
    // idx_a: table slot -> text position; key_s is the sorted key column
    sh.mem_key.ensure(T); sh.mem_vlen.ensure(T); sh.mem_gpos.ensure(T); sh.mem_ent.ensure(T);
    PSI_CUDA(cudaMemcpyAsync(sh.mem_key.p, key_s.p, T * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c.stream));
    mem_gather_kernel<uint8_t><<<grid_for(T, 256), 256, 0, c.stream>>>(vlen.p, idx_a.p, T, sh.mem_vlen.p);
    mem_gather_kernel<uint32_t><<<grid_for(T, 256), 256, 0, c.stream>>>(gpos.p, idx_a.p, T, sh.mem_gpos.p);
    mem_gather_kernel<uint32_t><<<grid_for(T, 256), 256, 0, c.stream>>>(ent.p, idx_a.p, T, sh.mem_ent.p);
    PSI_CUDA(cudaStreamSynchronize(c.stream));     // the temporaries die here
  }
  mem_prefix_table_kernel<<<grid_for(n_pfx, 256), 256, 0, c.stream>>>(sh.mem_key.p, T, sh.mem_pstart.p);
  c.counters.launches += 16;
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  sh.has_mem_index = true;
}

// find_mems over every read of the submitted chunk; the result is the set of hits as 6 x u64 records.
void engine_find_mems(Ctx& c, unsigned max_mem)
{
  if (!c.has_chunk) throw StateError("find_mems: no read chunk submitted");
  if (c.pending) throw StateError("find_mems: a step is in flight on this context (call psi_b200_wait first)");
  Shared& sh = *c.sh;
  if (!sh.has_mem_index) throw StateError("find_mems: no MEM index (psi_b200_build_mem_index first)");
  PSI_CUDA(cudaSetDevice(c.device));
  c.mem_valid = false;
  c.n_mems = 0;
  if (c.n_reads == 0 || sh.mem_n == 0) { c.mem_valid = true; return; }
  nvtxRangePushA("query-paths (find_mems)");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop_range;
  const GraphView g = make_graph_view(c);
  MemPaths mp{ sh.mem_path_ptr.p, sh.mem_nodes.p, sh.mem_head.p, sh.mem_tail.p, sh.mem_entry_start.p, sh.mem_paths_n };
  MemIndexView ix{ sh.mem_key.p, sh.mem_vlen.p, sh.mem_gpos.p, sh.mem_ent.p, sh.mem_pstart.p, sh.mem_n };
  MemChunk ch{ c.chunk_packed ? nullptr : c.d_bases, c.chunk_packed ? c.d_words : nullptr, c.d_read_ptr, c.d_exc,
               c.chunk_packed ? c.n_exc : 0, c.n_reads, c.read_len };
  unsigned long long* d_n = c.dev_counters.p + DC_AUX;
  if (c.mem_raw.cap == 0) c.mem_raw.ensure(std::max<uint64_t>(c.n_seeds_cap, 1u << 16));
  unsigned long long n_raw = 0;
  PhaseTimer t_on(c, T_ON);
  for (int attempt = 0; attempt < 3; ++attempt) {
    PSI_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), c.stream));
    find_mems_kernel<<<grid_for(c.n_reads, 128), 128, 0, c.stream>>>(g, mp, ix, ch, c.k, sh.gocc_threshold ? sh.gocc_threshold : c.opt_gocc_threshold,
                                                                      max_mem, reinterpret_cast<MemRaw*>(c.mem_raw.p), c.mem_raw.cap, d_n);
    ++c.counters.launches;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(&n_raw, d_n, sizeof(n_raw), cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaStreamSynchronize(c.stream));
    if (n_raw <= c.mem_raw.cap) break;
    c.mem_raw.ensure(n_raw, 1.1);
  }
  t_on.stop();
  if (n_raw > c.mem_raw.cap) throw OverflowError("find_mems: hit buffer keeps overflowing");
  if (n_raw >= 0xfffffff0ull) throw OverflowError("find_mems: more than 2^32 raw hits in one chunk (use smaller chunks)");
  if (n_raw == 0) { c.mem_valid = true; return; }
  // the set of hits: sort by (read | offset | length, locus), drop repeats (one locus reached through several paths)
  const uint64_t n = n_raw;
  const MemRaw* raw = reinterpret_cast<const MemRaw*>(c.mem_raw.p);
  DevBuf<uint64_t> key_a, key_b;
  DevBuf<uint32_t> gp_a, gp_b, idx_a, idx_b, flag, scan;
  key_a.ensure(n); key_b.ensure(n); gp_a.ensure(n); gp_b.ensure(n); idx_a.ensure(n); idx_b.ensure(n); flag.ensure(n + 1); scan.ensure(n + 1);
  mem_split_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(raw, n, key_a.p, gp_a.p);
  mem_iota_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(idx_a.p, n);
  size_t t1 = 0, t2 = 0;
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, gp_a.p, gp_b.p, idx_a.p, idx_b.p, (int64_t)n, 0, 32, c.stream));
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, key_b.p, key_a.p, idx_b.p, idx_a.p, (int64_t)n, 0, 64, c.stream));
  DevBuf<char> tmp;
  tmp.ensure(std::max(t1, t2));
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t1, gp_a.p, gp_b.p, idx_a.p, idx_b.p, (int64_t)n, 0, 32, c.stream));
  mem_gather_kernel<uint64_t><<<grid_for(n, 256), 256, 0, c.stream>>>(key_a.p, idx_b.p, n, key_b.p);
  PSI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t2, key_b.p, key_a.p, idx_b.p, idx_a.p, (int64_t)n, 0, 64, c.stream));
  mem_flag_unique_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(raw, idx_a.p, n, flag.p);
  PSI_CUDA(cudaMemsetAsync(flag.p + n, 0, sizeof(uint32_t), c.stream));
  mem_scan_u32(c, flag.p, scan.p, n + 1);
  uint32_t n_unique = 0;
  PSI_CUDA(cudaMemcpyAsync(&n_unique, scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.mem_records.ensure(6 * (uint64_t)n_unique + 6);
  mem_records_kernel<<<grid_for(n, 256), 256, 0, c.stream>>>(g, raw, idx_a.p, flag.p, scan.p, n, c.first_read_id, c.mem_records.p);
  c.counters.launches += 12;
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaStreamSynchronize(c.stream));
  c.n_mems = n_unique;
  c.n_mems_raw = n_raw;
  c.mem_valid = true;
}

void engine_fetch_mems(Ctx& c, uint64_t* hits, uint64_t cap)
{
  if (!c.mem_valid) throw StateError("fetch_mems: no MEM results (psi_b200_find_mems first)");
  PSI_CUDA(cudaSetDevice(c.device));
  const uint64_t n = std::min<uint64_t>(c.n_mems, cap);
  if (n) PSI_CUDA(cudaMemcpyAsync(hits, c.mem_records.p, n * 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
  PSI_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace psi_b200
