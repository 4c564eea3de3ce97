// fused.cu -- the whole per-chunk hot path in ONE kernel (index mode):
//   reads (ASCII or 2-bit words) -> seeds (K1) -> one bucket-line probe per seed (K3) -> results.
//
// Stands behind the loop body of find_seeds() (reference src/psikt.cpp:190-208):
//   get_seeds (seed_finder.hpp:1099-1109, sequence.hpp:1688-1745)
//   -> seeds_all = seeds_on_paths + seeds_off_paths (seed_finder.hpp:1426-1457,1703-1743)
//   -> write_callback (src/psikt.cpp:172-181),
// when the k-walks from the starting loci are materialised in the index (offpath_mode 2), so that one probe
// answers both phases.  The separate kernels (chunk.cu, seeds.cu) stay for walk mode, PSI_B200_SORTED and
// PSI_B200_NO_RESOLVE; both routes produce the same set and the parity suite runs every case through both.
//
// What a step touches is what the problem requires and nothing else: the chunk's bytes once, one 128-byte index
// line per seed, one result per seed (dense) or per hit (records).  A single-locus index entry carries its LOCUS
// CODE -- (node id, offset in the node) packed into the slot's payload (common.cuh) -- so a hit needs no further
// memory access: the position -> node gathers of the first version of this kernel are gone.
//
// Shape: a CTA owns R <= 256 consecutive reads (their start offsets and the prefix sums of their seed counts
// live in shared memory) and walks its seeds in seed order in batches of 256, one seed per thread:
//   1. the seed's characters were requested during the previous batch (ASCII: cp.async into a per-thread staging
//      slot; 2-bit words: two 64-bit loads into registers); -> k-mer + validity, hash -> home line + tag;
//   2. each warp copies the 32 home lines of its seeds into shared memory with cp.async (8 lanes x 16 B per
//      line; invalid / inactive seeds issue a zero-fill copy that reads nothing);
//   3. the characters of the NEXT batch are requested;
//   4. wait for the lines only, every thread scans its own line (16 tag compares);
//   5. results
//      DENSE:   two coalesced stores per seed into two planes: node id (u32, NIL = no hit) and node offset with the
//               off-path flag in its top bit (u16 when no node is longer than 32 768 bases, else u32).
//               No atomics, no lists, no CTA-wide barrier in the loop.  Read id and offset are implied by the
//               seed's position in the planes (seed order of the chunk).
//      RECORDS: the warp appends its hits (locus code, seed index) to a CTA-local list; every FUSED_FLUSH batches
//               the list becomes 4 x u64 / 4 x u32 records behind ONE atomic reservation of the CTA's output range.
// The ~0.3 % of the seeds one line cannot settle (locus lists, displaced keys) are queued with their k-mer and
// (read, offset, seed index); seeds_slow_fused_kernel resolves them.
#include "engine.hpp"
#include "records.cuh"
#include "seeding.cuh"

#include <algorithm>
#include <atomic>

namespace psi_b200 {

using namespace dev;

constexpr int FUSED_READS = 256;
// Bucket lines sit in shared memory at a stride of 144 bytes: lane l then reads 16-byte chunk j of its own line at
// bank 4 (l + j) mod 32 -- conflict-free within every quarter warp, no per-chunk address arithmetic.
constexpr uint32_t FUSED_LINE_STRIDE = 144;
// RECORDS: hits are collected in a CTA-local list and turned into records every FUSED_FLUSH batches (and after the
// last one): one output reservation per flush instead of per batch, and no CTA-wide barrier inside a batch.  A seed
// settled by the one-line probe has at most one hit, so the list cannot overflow.
constexpr int FUSED_FLUSH = 3;
constexpr int FUSED_LIST = FUSED_FLUSH * 256;
// ASCII input: the characters of a seed are staged in shared memory as K4 + 1 aligned 32-bit words at an odd word
// stride per thread.  2-bit input needs no staging (two 64-bit words per seed travel in registers).
template <int K4, bool PACKED> struct FusedCfg {
  static constexpr uint32_t WORDS = PACKED ? 0u : (uint32_t)((K4 + 1) | 1);
  static constexpr size_t SMEM = (size_t)256 * FUSED_LINE_STRIDE + (size_t)256 * WORDS * 4;
};

struct FusedChunk {
  const char* bases;            // ASCII chunk (null for 2-bit chunks)
  const uint64_t* words;        // 2-bit chunk: 32 bases per word, first base in the low bits, reads back to back
  const uint64_t* read_ptr;     // n_reads + 1 base offsets; null: every read has read_len bases
  const uint64_t* exc;          // 2-bit chunk: sorted base positions of the characters outside A/C/G/T
  uint64_t n_exc;
  uint64_t n_reads;
  uint64_t first_read_id;
  const uint32_t* cta_first;    // DENSE, ragged reads: first seed of every CTA; null: seeds_per_read each
  uint32_t seeds_per_read;
  uint32_t read_len;
  uint32_t R;                   // reads per CTA
  uint32_t k, d;
};

struct FusedOut {
  uint64_t* records;            // RECORDS: 4 x u64 (or 4 x u32 when compact) per hit
  uint8_t* rec_kind;
  uint64_t cap;
  uint32_t compact;
  uint32_t* dense_id;           // DENSE: plane of node ids, one per seed (NIL32 = no hit)
  void* dense_off;              //        plane of node offsets | off-path flag in the top bit: u32, u16 or (packed) u8
  uint32_t off_mode;            // 0: u32 offsets, 1: u16 offsets, 2: DENSE5 -- the 39-bit entry (id << off_bits | offset) split
                                // into its low 32 bits (dense_id plane) and a byte holding the rest + the off-path flag
  SlowItem* slow_queue;
  uint64_t slow_cap;
  unsigned long long* dc;
};

// one seed's slot of the dense planes; kind 0 = no hit, 1 = on an indexed path, 2 = off the indexed paths
__device__ __forceinline__ void st_dense(const FusedOut& out, const GraphView& g, uint32_t s, uint64_t id, uint32_t off, uint32_t kind)
{
  if (out.off_mode == 2) {
    uint32_t lo = NIL32, hi = 0xffu;
    if (kind) {
      const uint64_t e = (id << g.code_off_bits) | off;
      lo = (uint32_t)e;
      hi = (uint32_t)(e >> 32) | (kind == 2 ? 0x80u : 0u);
    }
    out.dense_id[s] = lo;
    static_cast<uint8_t*>(out.dense_off)[s] = (uint8_t)hi;
  }
  else {
    out.dense_id[s] = kind ? (uint32_t)id : NIL32;
    const uint32_t v = kind ? off : 0u;
    if (out.off_mode == 1) static_cast<uint16_t*>(out.dense_off)[s] = (uint16_t)(v | (kind == 2 ? 0x8000u : 0u));
    else static_cast<uint32_t*>(out.dense_off)[s] = v | (kind == 2 ? 0x80000000u : 0u);
  }
}

template <int FMT, int K4, int MIN_CTAS, bool DENSE, bool PACKED>
__global__ void __launch_bounds__(256, MIN_CTAS)
seeds_fused_kernel(KmerTable t, GraphView g, FusedChunk ch, uint32_t mode, FusedOut out)
{
  constexpr uint32_t STRIDE = FUSED_LINE_STRIDE;
  constexpr uint32_t WORDS = FusedCfg<K4, PACKED>::WORDS;
  constexpr int LIST = DENSE ? 1 : FUSED_LIST;
  extern __shared__ __align__(128) unsigned char s_dyn[];       // 256 bucket lines, then (ASCII) 256 x WORDS staged characters
  unsigned char* s_lines = s_dyn;
  __shared__ uint32_t s_first[FUSED_READS + 1];
  __shared__ uint64_t s_ptr[FUSED_READS];
  __shared__ uint32_t s_warp[8];
  // RECORDS: the CTA's hits since the last flush: locus code, and seed index relative to the window start (bit 15: off-path entry)
  __shared__ uint64_t s_hit_code[LIST];
  __shared__ uint16_t s_hit_idx[LIST];
  __shared__ uint32_t s_count, s_on, s_hits;
  __shared__ unsigned long long s_base;
  __shared__ uint64_t s_exc[2];

  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t R = ch.R, k = ch.k, d = ch.d;
  const uint64_t n_reads = ch.n_reads;
  const uint64_t r_base = (uint64_t)blockIdx.x * R;     // R <= FUSED_READS reads per CTA (host: enough CTAs to fill the GPU)

  // ---- the CTA's reads: start offsets and CTA-local prefix sums of the seed counts (sequence.hpp:1712) ----
  {
    const uint64_t r = r_base + threadIdx.x;
    uint64_t p0 = 0;
    uint32_t mine = 0;
    if (threadIdx.x < R && r < n_reads) {
      uint64_t len;
      if (ch.read_ptr) { p0 = ch.read_ptr[r]; len = ch.read_ptr[r + 1] - p0; }
      else { p0 = r * ch.read_len; len = ch.read_len; }
      mine = len >= k ? (uint32_t)((len - k) / d) + 1 : 0;      // reads shorter than k have no seeds
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
    if (lane == 31u) s_warp[warp] = incl;
    if (PACKED && threadIdx.x == 0) {
      // the CTA's slice of the exception list: positions in [start of its first read, end of its last read)
      uint64_t e0 = 0, e1 = 0;
      if (ch.n_exc) {
        const uint64_t r_end = min(r_base + R, n_reads);
        const uint64_t p_end = ch.read_ptr ? ch.read_ptr[r_end] : r_end * ch.read_len;
        uint64_t lo = 0, hi = ch.n_exc;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(ch.exc + mid) < p0) lo = mid + 1; else hi = mid; }
        e0 = lo;
        hi = ch.n_exc;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(ch.exc + mid) < p_end) lo = mid + 1; else hi = mid; }
        e1 = lo;
      }
      s_exc[0] = e0;
      s_exc[1] = e1;
    }
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
    s_first[threadIdx.x] = before + incl - mine;
    s_ptr[threadIdx.x] = p0;
    if (threadIdx.x == FUSED_READS - 1) { s_first[FUSED_READS] = before + incl; s_count = 0; s_on = 0; s_hits = 0; }   // threads >= R hold the total
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
  const uint32_t tail = k - 4u * (K4 - 1);                           // ASCII: characters in the last group, 1..4
  const uint32_t tail_mask = tail >= 4u ? 0xffffffffu : (1u << (8u * tail)) - 1u;
  const uint64_t kmask = low_mask64(2u * k);
  const uint64_t exc0 = PACKED ? s_exc[0] : 0, exc1 = PACKED ? s_exc[1] : 0;
  // DENSE: index of the CTA's first seed within the chunk
  const uint32_t seed0 = ch.cta_first ? __ldg(ch.cta_first + blockIdx.x) : (uint32_t)r_base * ch.seeds_per_read;

  unsigned char* warp_lines = s_lines + (size_t)warp * (32 * STRIDE);
  uint32_t* my_words = reinterpret_cast<uint32_t*>(s_dyn + 256 * STRIDE) + threadIdx.x * WORDS;
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

  // Locate this thread's seed of the batch starting at `base` and request its characters.
  //   ASCII: K4 + 1 aligned words go to the thread's staging slot with cp.async, so no register waits for them while
  //          the previous batch is processed;
  //   2-bit: the two 64-bit words holding the k-mer are loaded into registers (consumed one batch later).
  uint32_t rd = 0, roff = 0, sh = 0;
  uint64_t w0 = 0, w1 = 0, pos = 0;
  auto fetch_batch = [&](uint32_t base) {
    const uint32_t ls_raw = base + threadIdx.x;
    const uint32_t ls = ls_raw < n_cta_seeds ? ls_raw : 0u;     // inactive slots re-read seed 0 (valid memory), emit nothing
    const uint32_t lo = read_of(ls);
    rd = lo;
    roff = (ls - s_first[lo]) * d;
    if (PACKED) {
      pos = s_ptr[lo] + roff;
      const uint64_t* w = ch.words + (pos >> 5);
      sh = (uint32_t)(pos & 31u) * 2u;
      w0 = __ldg(w);
      w1 = __ldg(w + 1);         // the word buffer ends with one spare word
    }
    else {
      const uintptr_t addr = reinterpret_cast<uintptr_t>(ch.bases + s_ptr[lo] + roff);
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
    }
  };
  fetch_batch(0);

  // RECORDS: hits of the window -> records behind ONE reservation of the CTA's output range.
  uint32_t window = 0;       // CTA-local index of the first seed of the current window
  auto flush = [&]() {
    __syncthreads();                                   // every warp has appended its hits
    const uint32_t n = s_count;
    if (threadIdx.x == 0) s_base = n ? atomicAdd(out.dc + DC_HITS, (unsigned long long)n) : 0ull;
    uint32_t on = 0;
    uint64_t id[FUSED_FLUSH], noff[FUSED_FLUSH];
    uint32_t ix[FUSED_FLUSH];
#pragma unroll
    for (int p = 0; p < FUSED_FLUSH; ++p) {
      const uint32_t e = p * 256u + threadIdx.x;
      ix[p] = 0; id[p] = 0; noff[p] = 0;
      if (e < n) {
        ix[p] = (uint32_t)s_hit_idx[e];
        decode_code(g, s_hit_code[e], id[p], noff[p]);   // by-rank codes: one gather each, all in flight together
        on += (ix[p] & 0x8000u) ? 0u : 1u;
      }
    }
    __syncthreads();                                   // s_base is visible; the list has been read
    const unsigned long long out0 = s_base;
#pragma unroll
    for (int p = 0; p < FUSED_FLUSH; ++p) {
      const uint32_t e = p * 256u + threadIdx.x;
      const uint64_t o = out0 + e;
      if (e < n && o < out.cap) {
        Resolved r;
        const uint32_t ls = window + (ix[p] & 0x7fffu);
        const uint32_t lo = read_of(ls);
        r.read_id = ch.first_read_id + r_base + lo;
        r.read_off = (ls - s_first[lo]) * d;
        r.node_id = id[p];
        r.node_off = noff[p];
        if (out.compact) st_record32(out.records + 2 * o, r);
        else st_record(out.records + 4 * o, r);
        out.rec_kind[o] = (ix[p] & 0x8000u) ? 2 : 1;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) on += __shfl_xor_sync(0xffffffffu, on, o);
    if (lane == 0 && on) atomicAdd(&s_on, on);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();                                   // the list is free again
  };

  uint32_t n_hit = 0, n_on = 0;     // DENSE: this thread's hits
  uint32_t batch = 0;
  for (uint32_t base = 0; base < n_cta_seeds; base += 256u, ++batch) {
    // ---- 1. k-mer + hash ----
    uint64_t kmer;
    bool valid;
    if constexpr (PACKED) {
      kmer = ((w0 >> sh) | (sh ? w1 << (64u - sh) : 0ull)) & kmask;
      valid = true;
      if (exc0 < exc1) {          // CTA-uniform and rare: some read of this CTA holds a character outside A/C/G/T
        uint64_t lo = exc0, hi = exc1;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(ch.exc + mid) < pos) lo = mid + 1; else hi = mid; }
        valid = !(lo < exc1 && __ldg(ch.exc + lo) < pos + k);
      }
    }
    else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");        // this thread's characters have arrived
      AsciiWords<K4> aw;
#pragma unroll
      for (int i = 0; i <= K4; ++i) aw.x[i] = my_words[i];
      aw.sh = sh;
      kmer = pack_ascii_words<K4>(aw, tail_mask, valid);
    }
    const bool active = base + threadIdx.x < n_cta_seeds;
    const bool ok = valid && active;
    const Home hm = home_of<FMT>(t, kmer);
    const uint32_t my_line = (uint32_t)hm.line;              // line_bits <= 32 (checked when the table is allocated)
    const uint32_t want = (uint32_t)hm.tag;                  // fmt 8: tag | displacement 0 (<= 30 bits)
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
    if (more) fetch_batch(base + 256u);
    if (!PACKED && more) asm volatile("cp.async.wait_group 1;" ::: "memory");   // the lines have arrived, the characters may still fly
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- 4. scan ----
    uint32_t pl = 0, ph = 0;       // the matching slot's payload word and its high word (tag | flags | payload high bits)
    bool hit = false;
    const uint4* ln = reinterpret_cast<const uint4*>(warp_lines + lane * STRIDE);
    if (FMT == 8) {
      const uint32_t tsh = 2u + t.pay_hi;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 w = ln[j];                             // two slots: (w.x, w.y) and (w.z, w.w), high word second
        if ((w.y >> tsh) == want) { hit = true; pl = w.x; ph = w.y; }
        if ((w.w >> tsh) == want) { hit = true; pl = w.z; ph = w.w; }
      }
    }
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 w = ln[j];                             // one slot: key (w.x, w.y), payload w.z, flags + payload high bits w.w
        if (w.w != NIL32 && (((uint64_t)w.y << 32) | w.x) == kmer) { hit = true; pl = w.z; ph = w.w; }
      }
    }
    const uint32_t fl = FMT == 8 ? (ph >> t.pay_hi) & 3u : ph & 3u;
    uint32_t kind = 0;
    bool slow = false;
    if (ok) {
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
          if (FMT == 8) empty |= (w.y == 0xffffffffu) | (w.w == 0xffffffffu);   // no valid entry has both flag bits set
          else empty |= w.w == NIL32;
        }
        slow = !empty;
      }
      if (slow) {
        const unsigned long long q = atomicAdd(out.dc + DC_SLOW, 1ull);
        if (q < out.slow_cap) out.slow_queue[q] = SlowItem{ kmer, (uint32_t)(r_base + cur_rd), cur_off, seed0 + base + threadIdx.x, 0u };
      }
    }
    __syncwarp();            // every lane has read its line: the warp's buffers may be overwritten by the next batch
    const uint64_t code = FMT == 8 ? ((uint64_t)(ph & ((1u << t.pay_hi) - 1u)) << 32) | pl : slot16_payload(pl, ph);
    // ---- 5. results ----
    if (DENSE) {
      if (active) {
        uint64_t id = 0, noff = 0;                         // a queued seed shows "no hit" until the slow kernel has run
        if (kind) {
          decode_code(g, code, id, noff);
          ++n_hit;
          n_on += kind == 1 ? 1u : 0u;
        }
        st_dense(out, g, seed0 + base + threadIdx.x, id, (uint32_t)noff, kind);
      }
    }
    else {
      const uint32_t m = __ballot_sync(0xffffffffu, kind != 0);
      if (m) {
        uint32_t wbase = 0;
        if (lane == 0) wbase = atomicAdd(&s_count, (uint32_t)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (kind) {
          const uint32_t e = wbase + __popc(m & ((1u << lane) - 1u));
          s_hit_code[e] = code;
          s_hit_idx[e] = (uint16_t)((base + threadIdx.x - window) | (kind == 2 ? 0x8000u : 0u));
        }
      }
      if (batch % FUSED_FLUSH == FUSED_FLUSH - 1 || !more) {
        flush();
        window = base + 256u;
      }
    }
  }
  if (DENSE) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { n_hit += __shfl_xor_sync(0xffffffffu, n_hit, o); n_on += __shfl_xor_sync(0xffffffffu, n_on, o); }
    if (lane == 0 && n_hit) { atomicAdd(&s_hits, n_hit); atomicAdd(&s_on, n_on); }
    __syncthreads();
    if (threadIdx.x == 0 && s_hits) atomicAdd(out.dc + DC_HITS, (unsigned long long)s_hits);
  }
  if (threadIdx.x == 0) {
    atomicAdd(out.dc + DC_SEEDS, (unsigned long long)n_cta_seeds);
    if (s_on) atomicAdd(out.dc + DC_HITS_ON, (unsigned long long)s_on);
  }
}

// The queued seeds: full search (following lines, stash) and locus lists.  One thread per queued seed; a warp
// reserves the output range of its 32 seeds with ONE atomic (same-address atomics serialise in L2: ~15 000 queued
// seeds per 1 M reads would otherwise queue up there).
//   RECORDS: the records are appended to the CTAs' output;
//   DENSE:   the seed's first hit fills its slot of the dense planes, further hits (locus lists) become 4 x u32
//            records {node_id, node_off, read_id, read_off | off-path << 31} in the extra list.
template <bool DENSE>
__global__ void __launch_bounds__(256)
seeds_slow_fused_kernel(KmerTable t, const uint32_t* __restrict__ multi, GraphView g,
                        uint32_t mode, uint64_t first_read_id, FusedOut out, uint32_t* __restrict__ extra, uint64_t extra_cap)
{
  unsigned long long* dc = out.dc;
  uint64_t n = dc[DC_SLOW];
  if (n > out.slow_cap) n = out.slow_cap;     // the host grows the queue and repeats the step
  const uint32_t lane = lane_id();
  // warp-uniform trip count: every lane of a warp takes part in the reservation
  for (uint64_t q0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; q0 < n; q0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = q0 + lane;
    SlowItem it{};
    Found f{};
    uint32_t from = 0, to = 0, n_on = 0;      // this seed's hits: loci from..to of its list (or the single payload), the first n_on on paths
    bool single = false;
    if (q < n) {
      it = out.slow_queue[q];
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
    // DENSE: the first hit lives in the dense array, the extra list takes the rest
    uint64_t o = warp_reserve(DENSE ? dc + DC_EXTRA : dc + DC_HITS, DENSE ? (cnt ? cnt - 1u : 0u) : cnt);
    uint32_t on_sum = n_on, cnt_sum = cnt;
#pragma unroll
    for (int s = 16; s; s >>= 1) { on_sum += __shfl_xor_sync(0xffffffffu, on_sum, s); cnt_sum += __shfl_xor_sync(0xffffffffu, cnt_sum, s); }
    if (lane == 0 && on_sum) atomicAdd(dc + DC_HITS_ON, (unsigned long long)on_sum);
    if (DENSE && lane == 0 && cnt_sum) atomicAdd(dc + DC_HITS, (unsigned long long)cnt_sum);
    Resolved r;
    r.read_id = first_read_id + it.read;
    r.read_off = it.off;
    for (uint32_t j = from; j < to; ++j) {
      if (single) decode_code(g, f.payload, r.node_id, r.node_off);
      else resolve_node(g, __ldg(multi + f.payload + 2 + j), r.node_id, r.node_off);
      const uint32_t kind = single ? (n_on ? 1u : 2u) : (j - from < n_on ? 1u : 2u);
      if (DENSE) {
        if (j == from) st_dense(out, g, it.seed, r.node_id, (uint32_t)r.node_off, kind);
        else {
          if (o < extra_cap)
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(extra + 4 * o), "r"((uint32_t)r.node_id), "r"((uint32_t)r.node_off),
                         "r"((uint32_t)r.read_id), "r"((uint32_t)r.read_off | (kind == 2 ? 0x80000000u : 0u)) : "memory");
          ++o;
        }
      }
      else {
        if (o < out.cap) {
          if (out.compact) st_record32(out.records + 2 * o, r);
          else st_record(out.records + 4 * o, r);
          out.rec_kind[o] = (uint8_t)kind;
        }
        ++o;
      }
    }
  }
}

// DENSE with reads of several lengths: seeds per CTA of the fused kernel (R reads each), scanned into cta_first.
__global__ void __launch_bounds__(256)
count_cta_seeds_kernel(const uint64_t* __restrict__ read_ptr, uint64_t n_reads, uint32_t R, uint32_t k, uint32_t d,
                       uint32_t* __restrict__ cta_count)
{
  __shared__ uint32_t s_warp[8];
  const uint64_t r = (uint64_t)blockIdx.x * R + threadIdx.x;
  uint32_t c = 0;
  if (threadIdx.x < R && r < n_reads) {
    const uint64_t len = read_ptr[r + 1] - read_ptr[r];
    c = len >= k ? (uint32_t)((len - k) / d) + 1 : 0;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_warp[w];
    cta_count[blockIdx.x] = tot;
  }
}

// ---------------------------------------------------------------- host --

void launch_scan_cta_counts(Ctx& c, const uint32_t* cta_count, uint32_t n_ctas, uint32_t* cta_first, unsigned long long* n_seeds_out);

template <int FMT, int K4, bool DENSE, bool PACKED>
static void launch_fused(Ctx& c, const GraphView& g, const FusedChunk& ch, unsigned probe_mode, const FusedOut& out, unsigned grid)
{
  constexpr int MIN_CTAS = 4;
  auto kern = seeds_fused_kernel<FMT, K4, MIN_CTAS, DENSE, PACKED>;
  const size_t smem = FusedCfg<K4, PACKED>::SMEM;
  // function attributes are per device: one bit per device ordinal and instantiation, set after the (idempotent)
  // attribute calls have succeeded, so concurrent first launches from several host threads are harmless
  static std::atomic<uint64_t> attr_done{ 0 };
  const uint64_t bit = 1ull << (c.device & 63);
  if (c.device >= 64 || !(attr_done.load(std::memory_order_acquire) & bit)) {
    PSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  kern<<<grid, 256, smem, c.stream>>>(c.sh->index.view, g, ch, probe_mode, out);
}

template <int FMT, bool DENSE>
static void launch_fused_k(Ctx& c, const GraphView& g, const FusedChunk& ch, unsigned probe_mode, const FusedOut& out, unsigned grid)
{
  if (ch.words) { launch_fused<FMT, 1, DENSE, true>(c, g, ch, probe_mode, out, grid); return; }
  switch ((c.k + 3) / 4) {
    case 1: launch_fused<FMT, 1, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 2: launch_fused<FMT, 2, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 3: launch_fused<FMT, 3, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 4: launch_fused<FMT, 4, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 5: launch_fused<FMT, 5, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 6: launch_fused<FMT, 6, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    case 7: launch_fused<FMT, 7, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
    default: launch_fused<FMT, 8, DENSE, false>(c, g, ch, probe_mode, out, grid); break;
  }
}

// Queue one step on the context's stream: counters cleared, fused kernel, slow-queue kernel, counters copied to pinned
// host memory.  Returns without waiting.  out_kind: 0 = 4 x u64 records, 1 = 4 x u32 records, 2 = dense.
static void fused_enqueue(Ctx& c)
{
  Shared& sh = *c.sh;
  const GraphView g = make_graph_view(c);
  unsigned long long* dc = c.dev_counters.p;
  const int out_kind = c.pending_out_kind;
  const unsigned probe_mode = c.pending_probe_mode;
  const bool dense = out_kind >= 2;

  // Reads per CTA: 256 when that gives at least 4 CTAs per resident slot; fewer when the reads are few or long (d = 1,
  // long reads), but never so few that a CTA has less than 5 batches of seeds (n_seeds_cap is an upper bound).
  const uint64_t slots = (uint64_t)c.sm_count * 4;
  uint32_t R = FUSED_READS;
  if (c.n_reads / FUSED_READS < 4 * slots) {
    const uint64_t seeds_per_read = std::max<uint64_t>(1, c.n_seeds_cap / std::max<uint64_t>(1, c.n_reads));
    const uint64_t r_min = (5 * 256 + seeds_per_read - 1) / seeds_per_read;
    const uint64_t r_fill = (c.n_reads + 4 * slots - 1) / (4 * slots);
    R = (uint32_t)std::min<uint64_t>(FUSED_READS, std::max<uint64_t>(std::max(r_min, r_fill), 1));
  }
  const unsigned grid = (unsigned)std::max<uint64_t>(1, (c.n_reads + R - 1) / R);

  FusedChunk ch{};
  ch.bases = c.chunk_packed ? nullptr : c.d_bases;
  ch.words = c.chunk_packed ? c.d_words : nullptr;
  ch.read_ptr = c.d_read_ptr;
  ch.exc = c.d_exc;
  ch.n_exc = c.chunk_packed ? c.n_exc : 0;
  ch.n_reads = c.n_reads;
  ch.first_read_id = c.first_read_id;
  ch.read_len = c.read_len;
  ch.R = R;
  ch.k = c.k;
  ch.d = c.distance;
  ch.cta_first = nullptr;
  ch.seeds_per_read = 0;
  nvtxRangePushA("seeds-on-paths + seeds-off-path (fused)");
  struct Pop { ~Pop() { nvtxRangePop(); } } pop_range;
  PSI_CUDA(cudaMemsetAsync(dc, 0, DC_COUNT * sizeof(unsigned long long), c.stream));
  if (dense) {
    if (c.read_len && !c.d_read_ptr) ch.seeds_per_read = c.read_len >= c.k ? (c.read_len - c.k) / c.distance + 1 : 0;
    else {
      // reads of several lengths: where each CTA's seeds start in the dense array
      c.cta_first.ensure(2 * (size_t)grid + 2, 1.25);
      uint32_t* cta_count = c.cta_first.p + grid + 1;
      count_cta_seeds_kernel<<<grid, 256, 0, c.stream>>>(c.d_read_ptr, c.n_reads, R, c.k, c.distance, cta_count);
      launch_scan_cta_counts(c, cta_count, grid, c.cta_first.p, dc + DC_AUX2);
      c.counters.launches += 2;
      ch.cta_first = c.cta_first.p;
    }
  }

  FusedOut out{};
  out.records = c.records.p;
  out.rec_kind = c.rec_kind.p;
  out.cap = dense ? 0 : std::min<uint64_t>(c.records.cap / 4, c.rec_kind.cap);
  out.compact = out_kind == 1 ? 1u : 0u;
  out.dense_id = reinterpret_cast<uint32_t*>(c.records.p);
  out.dense_off = reinterpret_cast<char*>(c.records.p) + c.dense_off_plane;
  out.off_mode = c.dense_off_bytes == 1 ? 2u : c.dense_off_bytes == 2 ? 1u : 0u;
  out.slow_queue = c.slow_items.p;
  out.slow_cap = c.slow_items.cap;
  out.dc = dc;

  if (c.opt_timers) { PSI_CUDA(cudaEventRecord(c.ev[2 * T_ON], c.stream)); PSI_CUDA(cudaEventRecord(c.ev[2 * T_PROBE], c.stream)); }
  if (sh.index.view.fmt == 8) {
    if (dense) launch_fused_k<8, true>(c, g, ch, probe_mode, out, grid);
    else launch_fused_k<8, false>(c, g, ch, probe_mode, out, grid);
  }
  else {
    if (dense) launch_fused_k<16, true>(c, g, ch, probe_mode, out, grid);
    else launch_fused_k<16, false>(c, g, ch, probe_mode, out, grid);
  }
  if (c.opt_timers) PSI_CUDA(cudaEventRecord(c.ev[2 * T_PROBE + 1], c.stream));
  if (dense)
    seeds_slow_fused_kernel<true><<<(unsigned)c.sm_count, 256, 0, c.stream>>>(sh.index.view, sh.multi.p, g, probe_mode, c.first_read_id, out,
                                                                                c.extra.p, c.extra.cap / 4);
  else
    seeds_slow_fused_kernel<false><<<(unsigned)c.sm_count, 256, 0, c.stream>>>(sh.index.view, sh.multi.p, g, probe_mode, c.first_read_id, out,
                                                                                 nullptr, 0);
  if (c.opt_timers) PSI_CUDA(cudaEventRecord(c.ev[2 * T_ON + 1], c.stream));
  c.ev_state[T_ON] = c.ev_state[T_PROBE] = c.opt_timers ? 2 : 0;
  c.counters.launches += 2;
  PSI_CUDA(cudaGetLastError());
  PSI_CUDA(cudaMemcpyAsync(c.h_pinned, dc, DC_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
}

// the dense results -> host memory, queued behind the step (no wait)
static void dense_copy_enqueue(Ctx& c)
{
  if (c.pending_dense_dst && c.n_dense_seeds) {
    // the two planes, made adjacent on the host: ids[n_seeds], then offsets[n_seeds]
    const uint64_t n = c.n_dense_seeds;
    char* dst = static_cast<char*>(c.pending_dense_dst);
    PSI_CUDA(cudaMemcpyAsync(dst, c.records.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaMemcpyAsync(dst + n * 4, reinterpret_cast<const char*>(c.records.p) + c.dense_off_plane, n * c.dense_off_bytes,
                             cudaMemcpyDeviceToHost, c.stream));
  }
  c.pending_extra_copied = 0;
  if (c.pending_extra_dst && c.pending_extra_cap) {
    // the length of the extra list is not known yet: copy a fixed share now, the rest (rarely any) at wait time
    const uint64_t n = std::min<uint64_t>(std::min<uint64_t>(c.pending_extra_cap, c.extra.cap / 4), c.n_seeds_cap / 256 + 256);
    PSI_CUDA(cudaMemcpyAsync(c.pending_extra_dst, c.extra.p, n * 16, cudaMemcpyDeviceToHost, c.stream));
    c.pending_extra_copied = n;
  }
}

// seeds_all of the submitted chunk through the fused kernel, queued without waiting (engine_wait completes it).
// Preconditions (checked by engine_seeds): the index answers every requested phase by itself, records are wanted, unsorted.
void engine_seeds_fused_async(Ctx& c, unsigned probe_mode, int out_kind)
{
  const bool dense = out_kind >= 2;
  if (c.slow_items.cap == 0) c.slow_items.ensure(std::max<uint64_t>(c.n_seeds_cap / 16, 1u << 16));
  c.ev_state[T_PACK] = c.ev_state[T_READ_INDEX] = c.ev_state[T_RESOLVE] = 0;   // no such phases on this route
  c.ev_state[T_OFF] = c.ev_state[T_SORT] = c.ev_state[T_D2H] = 0;
  c.n_dense_seeds = 0;
  if (dense) {
    // two planes in one buffer: 4 bytes of node id per seed, then 2 or 4 bytes of node offset per seed
    c.dense_off_bytes = out_kind == 3 ? 1u : c.sh->max_node_len <= 32768u ? 2u : 4u;
    c.dense_off_plane = ((c.n_seeds_cap + 1) * 4 + 255) & ~(uint64_t)255;
    c.records.ensure((c.dense_off_plane + (c.n_seeds_cap + 1) * c.dense_off_bytes + 7) / 8);
    if (c.extra.cap == 0) c.extra.ensure(4 * std::max<uint64_t>(c.n_seeds_cap / 16, 1u << 16));
    // known now only when all reads have one length; otherwise the step's seed counter tells (engine_wait)
    if (c.read_len && !c.d_read_ptr)
      c.n_dense_seeds = c.read_len >= c.k ? c.n_reads * ((c.read_len - c.k) / c.distance + 1) : 0;
  }
  else {
    const uint64_t want = std::max<uint64_t>(c.pending_out_cap_want, std::max<uint64_t>(c.n_seeds_cap + c.n_seeds_cap / 4, 1u << 20));
    c.records.ensure(4 * want);
    c.rec_kind.ensure(want);
  }
  c.pending_out_kind = out_kind;
  c.pending_probe_mode = probe_mode;
  c.pending_dense_dst = c.pending_extra_dst = nullptr;
  c.pending_dense_cap = c.pending_extra_cap = c.pending_extra_copied = 0;
  c.pending_attempts = 0;
  fused_enqueue(c);
  c.pending = true;
}

void engine_fetch_dense_async(Ctx& c, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra)
{
  if (!c.pending || c.pending_out_kind < 2) throw StateError("fetch_dense_async: no PSI_B200_DENSE step in flight on this context");
  PSI_CUDA(cudaSetDevice(c.device));
  c.pending_dense_dst = dense;
  c.pending_dense_cap = cap_seeds;
  c.pending_extra_dst = extra;
  c.pending_extra_cap = cap_extra;
  if (c.n_dense_seeds == 0 && c.n_reads) {
    // reads of several lengths: the number of seeds comes back with the step's counters
    ctx_wait(c);
    c.n_dense_seeds = c.h_pinned[DC_SEEDS];
  }
  if (c.n_dense_seeds > cap_seeds) throw ArgError("fetch_dense: the dense buffer is smaller than the chunk's seed count");
  if (c.opt_timers) PSI_CUDA(cudaEventRecord(c.ev[2 * T_D2H], c.stream));
  dense_copy_enqueue(c);
  if (c.opt_timers) { PSI_CUDA(cudaEventRecord(c.ev[2 * T_D2H + 1], c.stream)); c.ev_state[T_D2H] = 2; }
}

// Completes the step in flight: waits, grows what overflowed and repeats the step if it must (rare), finalises counters.
void engine_wait(Ctx& c)
{
  if (!c.pending) return;
  PSI_CUDA(cudaSetDevice(c.device));
  const bool dense = c.pending_out_kind >= 2;
  while (true) {
    ctx_wait(c);
    const uint64_t n_total = c.h_pinned[DC_HITS], n_slow = c.h_pinned[DC_SLOW], n_extra = c.h_pinned[DC_EXTRA];
    bool retry = false;
    if (n_slow > c.slow_items.cap) { c.slow_items.ensure(n_slow, 1.25); retry = true; }
    if (dense) {
      if (n_extra > c.extra.cap / 4) { c.extra.ensure(4 * n_extra, 1.25); retry = true; }
    }
    else {
      const uint64_t out_cap = std::min<uint64_t>(c.records.cap / 4, c.rec_kind.cap);
      if (n_total > out_cap) {
        c.pending_out_cap_want = n_total + n_total / 8;
        c.records.ensure(4 * c.pending_out_cap_want);
        c.rec_kind.ensure(c.pending_out_cap_want);
        retry = true;
      }
    }
    if (retry) {
      if (++c.pending_attempts > 16) { c.pending = false; throw OverflowError("seeds_all: device buffers keep overflowing"); }
      fused_enqueue(c);
      if (dense && (c.pending_dense_dst || c.pending_extra_dst)) dense_copy_enqueue(c);
      continue;
    }
    c.n_hits = n_total;
    c.counters.n_seeds = c.h_pinned[DC_SEEDS];
    c.counters.n_hits_on = c.h_pinned[DC_HITS_ON];
    c.counters.n_hits_off = n_total - c.h_pinned[DC_HITS_ON];
    c.counters.n_hits = n_total;
    c.counters.n_walks = 0;
    c.counters.n_on_probe_sectors = n_slow;
    c.counters.offpath_mode = 2u;
    c.counters.fused = 1u;
    c.n_extra = dense ? n_extra : 0;
    if (dense) {
      c.n_dense_seeds = c.h_pinned[DC_SEEDS];
      // what the speculative copy of the extra list did not cover
      if (c.pending_extra_dst && n_extra > c.pending_extra_copied) {
        const uint64_t upto = std::min<uint64_t>(n_extra, c.pending_extra_cap);
        if (upto > c.pending_extra_copied) {
          PSI_CUDA(cudaMemcpyAsync(static_cast<uint32_t*>(c.pending_extra_dst) + 4 * c.pending_extra_copied, c.extra.p + 4 * c.pending_extra_copied,
                                   (upto - c.pending_extra_copied) * 16, cudaMemcpyDeviceToHost, c.stream));
          ctx_wait(c);
        }
      }
    }
    break;
  }
  if (c.opt_timers) {
    c.counters.ms_probe_sum += PhaseTimer::timer_ms(c, T_PROBE);
    c.counters.ms_on_sum += PhaseTimer::timer_ms(c, T_ON);
    ++c.counters.timed_steps;
  }
  c.pending = false;
  c.records_valid = true;
  c.records_compact = c.pending_out_kind == 1;
  c.records_dense = dense;
  c.kinds_valid = !dense;
}

void engine_seeds_fused(Ctx& c, unsigned probe_mode, int out_kind)
{
  engine_seeds_fused_async(c, probe_mode, out_kind);
  engine_wait(c);
}

// Synchronous copy of the dense results of the last completed step.
void engine_fetch_dense(Ctx& c, void* dense, uint64_t cap_seeds, uint32_t* extra, uint64_t cap_extra)
{
  if (c.pending) throw StateError("fetch_dense: a step is in flight on this context (call psi_b200_wait first)");
  if (!c.records_valid || !c.records_dense) throw StateError("fetch_dense: the last seeds_all was not run with PSI_B200_DENSE");
  PSI_CUDA(cudaSetDevice(c.device));
  if (cap_seeds && c.n_dense_seeds > cap_seeds) throw ArgError("fetch_dense: the dense buffer is smaller than the chunk's seed count");
  PhaseTimer t(c, T_D2H);
  if (dense && cap_seeds && c.n_dense_seeds) {
    const uint64_t n = c.n_dense_seeds;
    PSI_CUDA(cudaMemcpyAsync(dense, c.records.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    PSI_CUDA(cudaMemcpyAsync(static_cast<char*>(dense) + n * 4, reinterpret_cast<const char*>(c.records.p) + c.dense_off_plane,
                             n * c.dense_off_bytes, cudaMemcpyDeviceToHost, c.stream));
  }
  const uint64_t ne = std::min<uint64_t>(c.n_extra, cap_extra);
  if (extra && ne) PSI_CUDA(cudaMemcpyAsync(extra, c.extra.p, ne * 16, cudaMemcpyDeviceToHost, c.stream));
  t.stop();
  ctx_wait(c);
}

}  // namespace psi_b200
