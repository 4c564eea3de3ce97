// fused.cu -- the whole per-chunk hot path in ONE kernel (index mode):
//   ASCII reads -> seeds (K1) -> one bucket-line probe per seed (K3) -> seed records (resolve).
//
// Stands behind the loop body of find_seeds() (reference src/psikt.cpp:190-208):
//   get_seeds (seed_finder.hpp:1099-1109, sequence.hpp:1688-1745)
//   -> seeds_all = seeds_on_paths + seeds_off_paths (seed_finder.hpp:1426-1457,1703-1743)
//   -> write_callback (src/psikt.cpp:172-181),
// when the k-walks from the starting loci are materialised in the index (offpath_mode 2), so that one probe
// answers both phases.  The separate kernels (chunk.cu, seeds.cu) stay for walk mode, PSI_B200_SORTED and
// PSI_B200_NO_RESOLVE; both routes produce the same set and the parity suite runs every case through both.
//
// Why fuse: per step of 1 M x 100 bp reads the three-kernel route writes and re-reads 13 B (k-mer, validity,
// read index) + 5 B (locus, kind) per seed between its kernels and pays three launch tails; the fused kernel
// touches only what the problem requires: the chunk's bytes once, one 128-byte index line per seed, one record
// per hit.  The packing arithmetic and the record resolution run in the issue slots the line waits leave idle.
//
// Shape: a CTA owns FUSED_READS consecutive reads (their start offsets and the prefix sums of their seed counts
// live in shared memory) and walks its seeds in seed order in batches of 256, one seed per thread:
//   1. the characters of the batch were requested during the previous batch (cp.async into a per-thread staging
//      slot: no register waits for them); pack -> k-mer + validity, hash -> home line + tag;
//   2. each warp copies the 32 home lines of its seeds into shared memory with cp.async (8 lanes x 16 B per
//      line; invalid / inactive seeds issue a zero-fill copy that reads nothing);
//   3. the characters of the NEXT batch are requested;
//   4. wait for the lines only, every thread scans its own line (16 tag compares);
//   5. the warp appends its hits (locus, seed index) to a CTA-local list -- no CTA-wide barrier inside a batch;
//   6. every FUSED_FLUSH batches and after the last one the list becomes records: ONE atomic reserves the CTA's
//      output range, every thread has the position -> node gathers of up to FUSED_FLUSH hits in flight together.
// The ~0.3 % of the seeds one line cannot settle (locus lists, displaced keys) are queued with their k-mer and
// (read, offset); seeds_slow_fused_kernel resolves them and appends their records.
#include "engine.hpp"
#include "records.cuh"
#include "seeding.cuh"

#include <algorithm>

namespace psi_b200 {

using namespace dev;

constexpr int FUSED_READS = 256;
// Bucket lines sit in shared memory at a stride of 144 bytes: lane l then reads 16-byte chunk j of its own line at
// bank 4 (l + j) mod 32 -- conflict-free within every quarter warp, no per-chunk address arithmetic.
constexpr uint32_t FUSED_LINE_STRIDE = 144;
// Hits are collected in a CTA-local list and turned into records every FUSED_FLUSH batches (and after the last one):
// one output reservation and one burst of position -> node gathers per flush instead of per batch, and no CTA-wide
// barrier inside a batch.  A seed settled by the one-line probe has at most one hit, so the list cannot overflow.
constexpr int FUSED_FLUSH = 5;
constexpr int FUSED_LIST = FUSED_FLUSH * 256;
constexpr int FUSED_WIDTH = 3;      // hits per thread whose gathers are in flight together during a flush
// The characters of a seed are staged in shared memory as K4 + 1 aligned 32-bit words at an odd word stride per thread.
template <int K4> struct FusedCfg {
  static constexpr uint32_t WORDS = (K4 + 1) | 1;
  static constexpr size_t SMEM = (size_t)256 * FUSED_LINE_STRIDE + (size_t)256 * WORDS * 4;
};

template <int FMT, int K4, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
seeds_fused_kernel(KmerTable t, GraphView g, const uint64_t* __restrict__ node_id,
                   const char* __restrict__ bases, const uint64_t* __restrict__ read_ptr, uint64_t n_reads, uint32_t R,
                   uint32_t k, uint32_t d, uint32_t mode, uint64_t first_read_id, uint32_t compact,
                   uint64_t* __restrict__ records, uint8_t* __restrict__ rec_kind, uint64_t cap,
                   SlowItem* __restrict__ slow_queue, uint64_t slow_cap, unsigned long long* __restrict__ dc)
{
  constexpr uint32_t STRIDE = FUSED_LINE_STRIDE;
  constexpr uint32_t WORDS = FusedCfg<K4>::WORDS;
  extern __shared__ __align__(128) unsigned char s_dyn[];       // 256 bucket lines, then 256 x WORDS staged characters
  unsigned char* s_lines = s_dyn;
  uint32_t* s_ascii = reinterpret_cast<uint32_t*>(s_dyn + 256 * STRIDE);
  __shared__ uint32_t s_first[FUSED_READS + 1];
  __shared__ uint64_t s_ptr[FUSED_READS];
  __shared__ uint32_t s_warp[8];
  // the CTA's hits since the last flush: locus, and seed index relative to the window start (bit 15: off-path entry)
  __shared__ uint32_t s_hit_gpos[FUSED_LIST];
  __shared__ uint16_t s_hit_idx[FUSED_LIST];
  __shared__ uint32_t s_count, s_on;
  __shared__ unsigned long long s_base;

  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint64_t r_base = (uint64_t)blockIdx.x * R;     // R <= FUSED_READS reads per CTA (host: enough CTAs to fill the GPU)

  // ---- the CTA's reads: start offsets and CTA-local prefix sums of the seed counts (sequence.hpp:1712) ----
  {
    const uint64_t r = r_base + threadIdx.x;
    uint64_t p0 = 0;
    uint32_t mine = 0;
    if (threadIdx.x < R && r < n_reads) {
      p0 = read_ptr[r];
      const uint64_t len = read_ptr[r + 1] - p0;
      mine = len >= k ? (uint32_t)((len - k) / d) + 1 : 0;      // reads shorter than k have no seeds
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
    s_first[threadIdx.x] = before + incl - mine;
    s_ptr[threadIdx.x] = p0;
    if (threadIdx.x == FUSED_READS - 1) { s_first[FUSED_READS] = before + incl; s_count = 0; s_on = 0; }   // threads >= R hold the total
    __syncthreads();
  }
  const uint32_t n_cta_seeds = s_first[FUSED_READS];
  if (n_cta_seeds == 0) return;
  // reads of one length (the usual case) have the same number of seeds: the read of a seed is then a division,
  // done as a multiplication by floor(2^32 / per_read) plus one correction step
  const uint32_t per_read = s_first[1];
  bool uniform, one_each;
  {
    const uint32_t r_in_cta = (uint32_t)min((uint64_t)R, n_reads - r_base);
    const uint32_t mine = threadIdx.x < r_in_cta ? s_first[threadIdx.x + 1] - s_first[threadIdx.x] : per_read;
    uniform = __syncthreads_and(mine == per_read) && per_read > 1u;
    one_each = !uniform && __syncthreads_and(mine == 1u) && per_read == 1u;
  }
  const uint32_t magic = uniform ? (uint32_t)(0x100000000ull / per_read) : 0u;
  const uint32_t tail = k - 4u * (K4 - 1);                           // characters in the last group, 1..4
  const uint32_t tail_mask = tail >= 4u ? 0xffffffffu : (1u << (8u * tail)) - 1u;

  unsigned char* warp_lines = s_lines + (size_t)warp * (32 * STRIDE);
  uint32_t* my_words = s_ascii + threadIdx.x * WORDS;
  const uint32_t my_words_sa = (uint32_t)__cvta_generic_to_shared(my_words);
  const uint32_t sub = lane & 7u;

  // CTA-local seed index -> the read (CTA-local) it belongs to
  auto read_of = [&](uint32_t ls) -> uint32_t {
    uint32_t lo = 0;
    if (uniform) {
      lo = __umulhi(ls, magic);
      if ((lo + 1u) * per_read <= ls) ++lo;
    }
    else if (one_each) lo = ls;
    else {
      uint32_t hi = FUSED_READS;
#pragma unroll 1
      while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_first[mid] <= ls) lo = mid; else hi = mid; }
    }
    return lo;
  };

  // Locate this thread's seed of the batch starting at `base` and request its characters: K4 + 1 aligned words go to
  // the thread's staging slot with cp.async, so no register waits for them while the previous batch is processed.
  uint32_t rd = 0, roff = 0, sh = 0;
  auto fetch_batch = [&](uint32_t base) {
    const uint32_t ls_raw = base + threadIdx.x;
    const uint32_t ls = ls_raw < n_cta_seeds ? ls_raw : 0u;     // inactive slots re-read seed 0 (valid memory), emit nothing
    const uint32_t lo = read_of(ls);
    rd = lo;
    roff = (ls - s_first[lo]) * d;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(bases + s_ptr[lo] + roff);
    const char* w = reinterpret_cast<const char*>(addr & ~uintptr_t(3));
    const uint32_t off = (uint32_t)(addr & 3u);
    sh = off * 8u;
    // words 0 .. K4-1 always hold bytes of the k-mer (k > 4 (K4 - 1)); word K4 only when the k-mer spills into it
#pragma unroll
    for (int i = 0; i < K4; ++i)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(my_words_sa + 4u * i), "l"(w + 4 * i) : "memory");
    const uint32_t last = off + k > 4u * K4 ? 4u : 0u;          // 0: zero fill, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(my_words_sa + 4u * K4), "l"(last ? w + 4 * K4 : w), "r"(last) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch_batch(0);

  // Hits of the window -> records: ONE reservation of the CTA's output range, all gathers of a thread in flight together.
  uint32_t window = 0;       // CTA-local index of the first seed of the current window
  auto flush = [&]() {
    __syncthreads();                                   // every warp has appended its hits
    const uint32_t n = s_count;
    if (threadIdx.x == 0) s_base = n ? atomicAdd(dc + DC_HITS, (unsigned long long)n) : 0ull;
    const bool fast = g.rank16 != nullptr;
    uint32_t on = 0;
    bool base_ready = false;
    unsigned long long out0 = 0;
    // FUSED_WIDTH hits per thread and pass: their gathers are in flight together
#pragma unroll
    for (int p0 = 0; p0 < FUSED_FLUSH; p0 += FUSED_WIDTH) {
      if (p0 * 256u >= n) break;                       // CTA-uniform
      uint32_t gp[FUSED_WIDTH], ix[FUSED_WIDTH];
      uint4 rw[FUSED_WIDTH];
#pragma unroll
      for (int i = 0; i < FUSED_WIDTH; ++i) {
        const uint32_t e = (p0 + i) * 256u + threadIdx.x;
        gp[i] = e < n ? s_hit_gpos[e] : 0u;
        ix[i] = e < n ? (uint32_t)s_hit_idx[e] : 0u;
        if (fast && e < n) rw[i] = __ldg(reinterpret_cast<const uint4*>(g.rank16 + (gp[i] >> 6)));
      }
#pragma unroll
      for (int i = 0; i < FUSED_WIDTH; ++i) {
        const uint32_t e = (p0 + i) * 256u + threadIdx.x;
        if (fast && e < n) {
          const uint64_t bits = ((uint64_t)rw[i].y << 32) | rw[i].x;
          const uint32_t v = rw[i].z + (uint32_t)__popcll(bits & (~0ull >> (63u - (gp[i] & 63u)))) - 1u;
          rw[i] = __ldg(reinterpret_cast<const uint4*>(g.node_res + v));
        }
        on += (e < n && !(ix[i] & 0x8000u)) ? 1u : 0u;
      }
      if (!base_ready) {
        __syncthreads();                               // s_base is visible
        out0 = s_base;
        base_ready = true;
      }
#pragma unroll
      for (int i = 0; i < FUSED_WIDTH; ++i) {
        const uint32_t e = (p0 + i) * 256u + threadIdx.x;
        const uint64_t out = out0 + e;
        if (e < n && out < cap) {
          Resolved r;
          const uint32_t ls = window + (ix[i] & 0x7fffu);
          const uint32_t lo = read_of(ls);
          r.read_id = first_read_id + r_base + lo;
          r.read_off = (ls - s_first[lo]) * d;
          if (fast) { r.node_off = gp[i] - rw[i].x; r.node_id = ((uint64_t)rw[i].w << 32) | rw[i].z; }
          else resolve_node(g, node_id, gp[i], r.node_id, r.node_off);
          if (compact) st_record32(records + 2 * out, r);
          else st_record(records + 4 * out, r);
          rec_kind[out] = (ix[i] & 0x8000u) ? 2 : 1;
        }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) on += __shfl_xor_sync(0xffffffffu, on, o);
    if (lane == 0 && on) atomicAdd(&s_on, on);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();                                   // the list is free again
  };

  uint32_t batch = 0;
  for (uint32_t base = 0; base < n_cta_seeds; base += 256u, ++batch) {
    // ---- 1. pack + hash ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");        // this thread's characters have arrived
    AsciiWords<K4> aw;
#pragma unroll
    for (int i = 0; i <= K4; ++i) aw.x[i] = my_words[i];
    aw.sh = sh;
    bool valid;
    const uint64_t kmer = pack_ascii_words<K4>(aw, tail_mask, valid);
    const bool ok = valid && base + threadIdx.x < n_cta_seeds;
    const Home hm = home_of<FMT>(t, kmer);
    const uint32_t my_line = (uint32_t)hm.line;              // line_bits <= 32 (checked when the table is allocated)
    const uint32_t want = (uint32_t)hm.tag;                  // fmt 8: tag | displacement 0 (30 bits)
    const uint32_t cur_rd = rd, cur_off = roff;
    // ---- 2. the warp copies its home lines into shared memory ----
    {
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(warp_lines + (lane >> 3) * STRIDE + (sub << 4));
      const char* src0 = (const char*)t.slots + (sub << 4);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t owner = it * 4u + (lane >> 3);
        const uint32_t line = __shfl_sync(0xffffffffu, my_line, owner);
        const uint32_t bytes = (okmask >> owner) & 1u ? 16u : 0u;      // 0: zero fill, nothing is read
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                     :: "r"(dst0 + it * 4u * STRIDE), "l"(src0 + (bytes ? (uint64_t)line << 7 : 0ull)), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- 3. characters of the next batch ----
    const bool more = base + 256u < n_cta_seeds;
    if (more) {
      fetch_batch(base + 256u);
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // the lines have arrived, the characters may still fly
    }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- 4. scan ----
    uint32_t gpos = 0;
    uint32_t kind = 0;
    {
      const uint4* ln = reinterpret_cast<const uint4*>(warp_lines + lane * STRIDE);
      bool hit = false;
      uint32_t fl = 0;
      if (FMT == 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 w = ln[j];                             // two slots: (w.x, w.y) and (w.z, w.w), high word second
          if ((w.y >> 2) == want) { hit = true; gpos = w.x; fl = w.y & 3u; }
          if ((w.w >> 2) == want) { hit = true; gpos = w.z; fl = w.w & 3u; }
        }
      }
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 w = ln[j];                             // one slot: key (w.x, w.y), payload w.z, flags w.w
          if (w.w != NIL32 && (((uint64_t)w.y << 32) | w.x) == kmer) { hit = true; gpos = w.z; fl = w.w & 3u; }
        }
      }
      if (ok) {
        bool slow = false;
        if (hit) {
          if (fl & FLAG_MULTI) slow = true;
          else kind = kind_of(fl, mode);
        }
        else {
          // a miss is final only when the line has a free slot (else the key may sit in a following line): second
          // look, taken by the seeds that missed only
          bool empty = false;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 w = ln[j];
            if (FMT == 8) empty |= (w.y == 0xffffffffu) | (w.w == 0xffffffffu);   // no valid entry has all of rem/disp/flags set
            else empty |= w.w == NIL32;
          }
          slow = !empty;
        }
        if (slow) {
          const unsigned long long q = atomicAdd(dc + DC_SLOW, 1ull);
          if (q < slow_cap) slow_queue[q] = SlowItem{ kmer, (uint32_t)(r_base + cur_rd), cur_off };
        }
      }
    }
    __syncwarp();            // every lane has read its line: the warp's buffers may be overwritten by the next batch
    // ---- 5. the warp appends its hits to the CTA's list ----
    const uint32_t m = __ballot_sync(0xffffffffu, kind != 0);
    if (m) {
      uint32_t wbase = 0;
      if (lane == 0) wbase = atomicAdd(&s_count, (uint32_t)__popc(m));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (kind) {
        const uint32_t e = wbase + __popc(m & ((1u << lane) - 1u));
        s_hit_gpos[e] = gpos;
        s_hit_idx[e] = (uint16_t)((base + threadIdx.x - window) | (kind == 2 ? 0x8000u : 0u));
      }
    }
    if (batch % FUSED_FLUSH == FUSED_FLUSH - 1 || !more) {
      flush();
      window = base + 256u;
    }
  }
  if (threadIdx.x == 0) {
    atomicAdd(dc + DC_SEEDS, (unsigned long long)n_cta_seeds);
    if (s_on) atomicAdd(dc + DC_HITS_ON, (unsigned long long)s_on);
  }
}

// The queued seeds: full search (following lines, stash) and locus lists; their records are appended to the
// CTAs' output.  One thread per queued seed; a warp reserves the output range of its 32 seeds with ONE atomic
// (same-address atomics serialise in L2: ~15 000 queued seeds per 1 M reads would otherwise queue up there).
__global__ void __launch_bounds__(256)
seeds_slow_fused_kernel(KmerTable t, const uint32_t* __restrict__ multi, GraphView g, const uint64_t* __restrict__ node_id,
                        const SlowItem* __restrict__ slow_queue, uint64_t slow_cap, uint32_t mode, uint64_t first_read_id,
                        uint32_t compact, uint64_t* __restrict__ records, uint8_t* __restrict__ rec_kind, uint64_t cap,
                        unsigned long long* __restrict__ dc)
{
  uint64_t n = dc[DC_SLOW];
  if (n > slow_cap) n = slow_cap;     // the host grows the queue and repeats the step
  const uint32_t lane = lane_id();
  // warp-uniform trip count: every lane of a warp takes part in the reservation
  for (uint64_t q0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; q0 < n; q0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = q0 + lane;
    SlowItem it{};
    Found f{};
    uint32_t from = 0, to = 0, n_on = 0;      // this seed's hits: loci from..to of its list (or the single payload), the first n_on on paths
    bool single = false;
    if (q < n) {
      it = slow_queue[q];
      if (table_find_any(t, it.kmer, f)) {
        if (!(f.flags & FLAG_MULTI)) {
          const uint8_t kind = kind_of(f.flags, mode);
          single = true;
          to = kind ? 1u : 0u;
          n_on = kind == 1 ? 1u : 0u;
        }
        else {
          const uint32_t l_on = __ldg(multi + f.payload), l_all = __ldg(multi + f.payload + 1);
          from = (mode & PSI_B200_ON_PATHS) ? 0u : l_on;
          to = (mode & PSI_B200_OFF_PATHS) ? l_all : l_on;
          if (to < from) to = from;
          n_on = l_on > from ? l_on - from : 0u;
        }
      }
    }
    const uint32_t cnt = to - from;
    uint64_t out = warp_reserve(dc + DC_HITS, cnt);
    uint32_t on_sum = n_on;
#pragma unroll
    for (int o = 16; o; o >>= 1) on_sum += __shfl_xor_sync(0xffffffffu, on_sum, o);
    if (lane == 0 && on_sum) atomicAdd(dc + DC_HITS_ON, (unsigned long long)on_sum);
    Resolved r;
    r.read_id = first_read_id + it.read;
    r.read_off = it.off;
    for (uint32_t j = from; j < to; ++j, ++out) {
      if (out >= cap) break;
      const uint32_t gpos = single ? f.payload : __ldg(multi + f.payload + 2 + j);
      resolve_node(g, node_id, gpos, r.node_id, r.node_off);
      if (compact) st_record32(records + 2 * out, r);
      else st_record(records + 4 * out, r);
      rec_kind[out] = single ? (n_on ? 1 : 2) : (j - from < n_on ? 1 : 2);
    }
  }
}

// ---------------------------------------------------------------- host --

template <int FMT, int K4, int MIN_CTAS>
static void launch_fused(Ctx& c, const GraphView& g, unsigned probe_mode, bool compact, uint64_t out_cap)
{
  Shared& sh = *c.sh;
  auto kern = seeds_fused_kernel<FMT, K4, MIN_CTAS>;
  const size_t smem = FusedCfg<K4>::SMEM;
  static bool attr_set[64] = {};    // per instantiation and device (function attributes are per device)
  const int dv = c.device & 63;
  if (!attr_set[dv]) {
    PSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[dv] = true;
  }
  // Reads per CTA: 256 when that gives at least 4 CTAs per resident slot; fewer when the reads are few or long (d = 1,
  // long reads), but never so few that a CTA has less than 5 batches of seeds (n_seeds_cap is an upper bound).
  const uint64_t slots = (uint64_t)c.sm_count * MIN_CTAS;
  uint32_t R = FUSED_READS;
  if (c.n_reads / FUSED_READS < 4 * slots) {
    const uint64_t seeds_per_read = std::max<uint64_t>(1, c.n_seeds_cap / std::max<uint64_t>(1, c.n_reads));
    const uint64_t r_min = (5 * 256 + seeds_per_read - 1) / seeds_per_read;
    const uint64_t r_fill = (c.n_reads + 4 * slots - 1) / (4 * slots);
    R = (uint32_t)std::min<uint64_t>(FUSED_READS, std::max<uint64_t>(std::max(r_min, r_fill), 1));
  }
  const unsigned grid = (unsigned)std::max<uint64_t>(1, (c.n_reads + R - 1) / R);
  kern<<<grid, 256, smem, c.stream>>>(sh.index.view, g, sh.node_id.p, c.d_bases, c.d_read_ptr, c.n_reads, R, c.k, c.distance,
                                      probe_mode, c.first_read_id, compact ? 1u : 0u, c.records.p, c.rec_kind.p, out_cap,
                                      c.slow_items.p, c.slow_items.cap, c.dev_counters.p);
}

template <int FMT, int MIN_CTAS>
static void launch_fused_k(Ctx& c, const GraphView& g, unsigned probe_mode, bool compact, uint64_t out_cap)
{
  switch ((c.k + 3) / 4) {
    case 1: launch_fused<FMT, 1, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 2: launch_fused<FMT, 2, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 3: launch_fused<FMT, 3, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 4: launch_fused<FMT, 4, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 5: launch_fused<FMT, 5, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 6: launch_fused<FMT, 6, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    case 7: launch_fused<FMT, 7, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
    default: launch_fused<FMT, 8, MIN_CTAS>(c, g, probe_mode, compact, out_cap); break;
  }
}

// seeds_all of the submitted chunk through the fused kernel.  Preconditions (checked by engine_seeds): the index
// answers every requested phase by itself (probe_mode != 0, no per-chunk walk), records are wanted, unsorted.
void engine_seeds_fused(Ctx& c, unsigned probe_mode, bool compact)
{
  Shared& sh = *c.sh;
  const GraphView g = make_graph_view(c);
  unsigned long long* dc = c.dev_counters.p;
  uint64_t out_cap_want = std::max<uint64_t>(c.n_seeds_cap + c.n_seeds_cap / 4, 1u << 20);
  if (c.slow_items.cap == 0) c.slow_items.ensure(std::max<uint64_t>(c.n_seeds_cap / 16, 1u << 16));
  c.ev_state[T_PACK] = c.ev_state[T_READ_INDEX] = c.ev_state[T_RESOLVE] = 0;   // no such phases on this route
  c.ev_state[T_OFF] = c.ev_state[T_SORT] = c.ev_state[T_D2H] = 0;

  for (int attempt = 0;; ++attempt) {
    if (attempt > 16) throw OverflowError("seeds_all: device buffers keep overflowing");
    c.records.ensure(4 * out_cap_want);
    c.rec_kind.ensure(out_cap_want);
    const uint64_t out_cap = std::min<uint64_t>(c.records.cap / 4, c.rec_kind.cap);
    PSI_CUDA(cudaMemsetAsync(dc, 0, DC_COUNT * sizeof(unsigned long long), c.stream));
    PhaseTimer t_on(c, T_ON);
    PhaseTimer t_probe(c, T_PROBE);
    const bool pin = c.l2_window_bytes != 0 && sh.has_rank16;
    if (pin) {   // the position -> node gathers persist in the L2 set-aside, the index lines and the chunk stream through
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.base_ptr = sh.gather_pool.p;
      attr.accessPolicyWindow.num_bytes = c.l2_window_bytes;
      attr.accessPolicyWindow.hitRatio = c.l2_persist_bytes >= c.l2_window_bytes ? 1.0f : (float)c.l2_persist_bytes / (float)c.l2_window_bytes;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      PSI_CUDA(cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    if (sh.index.view.fmt == 8) {
      if (c.opt_fused_ctas == 5) launch_fused_k<8, 5>(c, g, probe_mode, compact, out_cap);
      else if (c.opt_fused_ctas == 3) launch_fused_k<8, 3>(c, g, probe_mode, compact, out_cap);
      else launch_fused_k<8, 4>(c, g, probe_mode, compact, out_cap);
    }
    else {
      if (c.opt_fused_ctas == 5) launch_fused_k<16, 5>(c, g, probe_mode, compact, out_cap);
      else if (c.opt_fused_ctas == 3) launch_fused_k<16, 3>(c, g, probe_mode, compact, out_cap);
      else launch_fused_k<16, 4>(c, g, probe_mode, compact, out_cap);
    }
    if (pin) {
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.num_bytes = 0;
      PSI_CUDA(cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    t_probe.stop();
    seeds_slow_fused_kernel<<<(unsigned)c.sm_count, 256, 0, c.stream>>>(sh.index.view, sh.multi.p, g, sh.node_id.p, c.slow_items.p,
                                                                           c.slow_items.cap, probe_mode, c.first_read_id,
                                                                           compact ? 1u : 0u, c.records.p, c.rec_kind.p, out_cap, dc);
    t_on.stop();
    c.counters.launches += 2;
    PSI_CUDA(cudaGetLastError());
    PSI_CUDA(cudaMemcpyAsync(c.h_pinned, dc, DC_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    ctx_wait(c);
    const uint64_t n_total = c.h_pinned[DC_HITS], n_slow = c.h_pinned[DC_SLOW];
    bool retry = false;
    if (n_slow > c.slow_items.cap) { c.slow_items.ensure(n_slow, 1.25); retry = true; }
    if (n_total > out_cap) { out_cap_want = n_total + n_total / 8; retry = true; }
    if (retry) continue;
    c.n_hits = n_total;
    c.counters.n_seeds = c.h_pinned[DC_SEEDS];
    c.counters.n_hits_on = c.h_pinned[DC_HITS_ON];
    c.counters.n_hits_off = n_total - c.h_pinned[DC_HITS_ON];
    c.counters.n_hits = n_total;
    c.counters.n_walks = 0;
    c.counters.n_on_probe_sectors = n_slow;
    c.counters.offpath_mode = 2u;
    c.counters.fused = 1u;
    break;
  }
  c.records_valid = true;
  c.records_compact = compact;
  c.kinds_valid = true;
}

}  // namespace psi_b200
