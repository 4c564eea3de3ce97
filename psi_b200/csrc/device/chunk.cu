// chunk.cu -- read chunk upload, seeding and 2-bit k-mer packing (K1) and the
// per-chunk read index (K4).
//
// Stands behind SeedFinder::get_seeds -> seeding() (reference
// include/psi/seed_finder.hpp:1099-1109, sequence.hpp:1688-1745) and
// SeedFinder::index_reads (seed_finder.hpp:1089-1097).  The reference copies
// every seed into a StringSet and builds (lazily) a suffix tree over them; here
// a seed is one 64-bit word and the "index" is a hash from the packed k-mer to
// the chain of seeds that spell it.
//   seeds of read r: offsets 0, d, 2d, ... while off + k <= len  (sequence.hpp:1712)
//   seed -> (read id, offset): SeedMap (sequence.hpp:1148-1220) is replaced by
//   seed_read[] plus the exclusive scan seed_first[] (offset = (s - first) * d).
#include "engine.hpp"
#include "seeding.cuh"

#include <cub/device/device_scan.cuh>

namespace psi_b200 {

using namespace dev;

// K1a: the chunk's bases, ASCII -> 2 bits per base + a 1-bit "not A/C/G/T" mask, whatever the read
// boundaries are (positions stay global byte offsets).  One thread per 32 bases: two 16-byte loads,
// eight 4-byte groups converted with byte-parallel arithmetic:
//   x = (c >> 1) & 3 ; x ^= x >> 1        maps A,C,G,T (either case) to 0,1,2,3
//   PRMT("ACGT", codes) == upper(c)       validates all four characters with one byte permute
__device__ __forceinline__ uint32_t pack4(uint32_t w4, uint32_t& bad)
{
  uint32_t x = (w4 >> 1) & 0x03030303u;
  x ^= (x >> 1) & 0x01010101u;
  const uint32_t y = (x | (x >> 4)) & 0x00ff00ffu;
  const uint32_t sel = (y | (y >> 8)) & 0xffffu;            // one code per nibble
  const uint32_t expected = __byte_perm(0x54474341u, 0u, sel);
  bad = expected ^ (w4 & 0xdfdfdfdfu);                      // non-zero bytes are not A/C/G/T
  const uint32_t p = (sel | (sel >> 2)) & 0x0f0fu;
  return (p | (p >> 4)) & 0xffu;                            // four codes, first base in the low bits
}

template <bool VEC>
__global__ void __launch_bounds__(256)
pack_reads_kernel(const char* __restrict__ bases, uint64_t n_bases, uint64_t* __restrict__ seq2, uint32_t* __restrict__ nmask)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b0 = w << 5;
  if (b0 >= n_bases) return;
  uint64_t bits = 0;
  uint32_t mask = 0;
  if (VEC && b0 + 32 <= n_bases) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(bases + b0));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(bases + b0 + 16));
    const uint32_t ws[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint32_t bad;
      bits |= (uint64_t)pack4(ws[g], bad) << (8 * g);
      if (bad) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if ((bad >> (8 * q)) & 0xffu) mask |= 1u << (4 * g + q);
      }
    }
  }
  else {
    const uint32_t cnt = (uint32_t)min((uint64_t)32, n_bases - b0);
    for (uint32_t i = 0; i < cnt; ++i) {
      const uint32_t c = base_code((unsigned char)bases[b0 + i]);
      if (c > 3) mask |= 1u << i;
      else bits |= (uint64_t)c << (2 * i);
    }
  }
  seq2[w] = bits;
  nmask[w] = mask;
}

// K1b: seeding.  A CTA owns READS_PER_CTA consecutive reads.
//   pass 1 (count_seeds_kernel):   seeds per CTA -> cta_count[]
//   pass 2 (scan_cta_counts_kernel, one CTA): exclusive scan -> cta_first[], total -> n_seeds
//   pass 3 (extract_seeds_kernel): recomputes the per-read counts, scans them inside the CTA
//          (-> seed_first[], the replacement of SeedMap's rank/select, sequence.hpp:1148-1220) and then
//          walks the CTA's seeds IN SEED ORDER -- thread t takes seeds t, t + 256, ... and finds the read by a
//          binary search of the CTA-local prefix sums in shared memory -- so that the stores of seed_kmer /
//          seed_valid / seed_read are contiguous; the k-mers are cut out of the 2-bit array.
constexpr int READS_PER_CTA = 1024;   // 256 threads x 4 reads

__device__ __forceinline__ uint32_t seed_count(const uint64_t* __restrict__ read_ptr, uint64_t r, uint64_t n_reads, uint32_t k, uint32_t d)
{
  if (r >= n_reads) return 0;
  const uint64_t len = read_ptr[r + 1] - read_ptr[r];
  return len >= k ? (uint32_t)((len - k) / d) + 1 : 0;     // reads shorter than k have no seeds
}

// sum over the CTA; result valid in every thread
__device__ __forceinline__ uint32_t cta_sum(uint32_t x, uint32_t* s_warp)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t t = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += s_warp[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256)
count_seeds_kernel(const uint64_t* __restrict__ read_ptr, uint64_t n_reads, uint32_t k, uint32_t d, uint32_t* __restrict__ cta_count)
{
  __shared__ uint32_t s_warp[8];
  const uint64_t r0 = (uint64_t)blockIdx.x * READS_PER_CTA + threadIdx.x * 4u;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) c += seed_count(read_ptr, r0 + i, n_reads, k, d);
  c = cta_sum(c, s_warp);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
scan_cta_counts_kernel(const uint32_t* __restrict__ cta_count, uint32_t n_ctas, uint32_t* __restrict__ cta_first,
                       unsigned long long* __restrict__ n_seeds_out)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_ctas; base += 1024u) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t x = i < n_ctas ? cta_count[i] : 0;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31u) >= (uint32_t)o) incl += y; }
    if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = s_carry;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += s_warp[w];
    if (i < n_ctas) cta_first[i] = before + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023u) s_carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { cta_first[n_ctas] = s_carry; *n_seeds_out = s_carry; }
}

__global__ void __launch_bounds__(256)
extract_seeds_kernel(const uint64_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, const uint64_t* __restrict__ read_ptr,
                     const uint32_t* __restrict__ cta_first, uint64_t n_reads, uint32_t k, uint32_t d,
                     uint32_t* __restrict__ seed_first, uint64_t* __restrict__ seed_kmer, uint8_t* __restrict__ seed_valid,
                     uint32_t* __restrict__ seed_read)
{
  __shared__ uint32_t s_first[READS_PER_CTA + 1];   // CTA-local exclusive prefix sums of the seed counts
  __shared__ uint32_t s_warp[8];
  const uint64_t r_base = (uint64_t)blockIdx.x * READS_PER_CTA;
  const uint64_t r0 = r_base + threadIdx.x * 4u;
  uint32_t c[4];
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i] = seed_count(read_ptr, r0 + i, n_reads, k, d); mine += c[i]; }
  // exclusive scan of `mine` over the CTA
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
  if (lane == 31u) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
  uint32_t run = before + incl - mine;
  const uint32_t first0 = cta_first[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s_first[threadIdx.x * 4u + i] = run;
    if (r0 + i <= n_reads) seed_first[r0 + i] = first0 + run;   // r == n_reads gets the total
    run += c[i];
  }
  if (threadIdx.x == 255u) s_first[READS_PER_CTA] = run;
  __syncthreads();
  const uint32_t n_cta_seeds = s_first[READS_PER_CTA];
  // reads of one length (the usual case) have the same number of seeds: the read of a seed is then a division
  const uint32_t per_read = s_first[1];
  const bool uniform = __syncthreads_and(c[0] == per_read && c[1] == per_read && c[2] == per_read && c[3] == per_read) && per_read != 0;
  for (uint32_t ls = threadIdx.x; ls < n_cta_seeds; ls += 256u) {
    // read = last index with s_first[index] <= ls
    uint32_t lo = 0, hi = READS_PER_CTA;
    if (uniform) lo = ls / per_read;
    else {
#pragma unroll 1
      while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_first[mid] <= ls) lo = mid; else hi = mid; }
    }
    const uint64_t r = r_base + lo;
    const uint64_t pos = __ldg(read_ptr + r) + (uint64_t)(ls - s_first[lo]) * d;
    const uint32_t s = first0 + ls;
    seed_kmer[s] = extract_bases(seq2, pos, k);
    seed_valid[s] = extract_nmask(nmask, pos, k) ? 0 : 1;
    seed_read[s] = (uint32_t)r;
  }
}

// K1 direct: seeding straight from the ASCII chunk (no 2-bit staging of the reads).  A CTA owns DIRECT_READS
// consecutive reads, one per thread: their start offsets and the CTA-local prefix sums of their seed counts go to
// shared memory, then the CTA walks its seeds IN SEED ORDER, DIRECT_UNROLL seeds per thread and round, so that
// every thread has DIRECT_UNROLL independent groups of loads in flight and all stores are contiguous.
// Per seed: k bytes in (each chunk byte is read from DRAM once when d >= k), 8 + 1 + 4 bytes out.
constexpr int DIRECT_READS = 256;
constexpr int DIRECT_UNROLL = 4;

__global__ void __launch_bounds__(256)
count_seeds_direct_kernel(const uint64_t* __restrict__ read_ptr, uint64_t n_reads, uint32_t k, uint32_t d, uint32_t* __restrict__ cta_count)
{
  __shared__ uint32_t s_warp[8];
  const uint32_t c = cta_sum(seed_count(read_ptr, (uint64_t)blockIdx.x * DIRECT_READS + threadIdx.x, n_reads, k, d), s_warp);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = c;
}

template <int K4>
__global__ void __launch_bounds__(256)
seed_reads_kernel(const char* __restrict__ bases, const uint64_t* __restrict__ read_ptr, const uint32_t* __restrict__ cta_first,
                  uint64_t n_reads, uint32_t k, uint32_t d, uint32_t* __restrict__ seed_first, uint64_t* __restrict__ seed_kmer,
                  uint8_t* __restrict__ seed_valid, uint32_t* __restrict__ seed_read)
{
  __shared__ uint32_t s_first[DIRECT_READS + 1];
  __shared__ uint64_t s_ptr[DIRECT_READS];
  __shared__ uint32_t s_warp[8];
  const uint64_t r_base = (uint64_t)blockIdx.x * DIRECT_READS;
  const uint64_t r = r_base + threadIdx.x;
  uint64_t p0 = 0;
  uint32_t mine = 0;
  if (r < n_reads) {
    p0 = read_ptr[r];
    const uint64_t len = read_ptr[r + 1] - p0;
    mine = len >= k ? (uint32_t)((len - k) / d) + 1 : 0;      // reads shorter than k have no seeds
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
  if (lane == 31u) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
  const uint32_t excl = before + incl - mine;
  const uint32_t first0 = cta_first[blockIdx.x];
  s_first[threadIdx.x] = excl;
  s_ptr[threadIdx.x] = p0;
  if (r <= n_reads) seed_first[r] = first0 + excl;              // r == n_reads gets the total
  if (threadIdx.x == DIRECT_READS - 1) s_first[DIRECT_READS] = excl + mine;
  __syncthreads();
  const uint32_t n_cta_seeds = s_first[DIRECT_READS];
  // reads of one length (the usual case) have the same number of seeds: the read of a seed is then a division,
  // done as a multiplication by floor(2^32 / per_read) plus one correction step
  const uint32_t per_read = s_first[1];
  const bool uniform = __syncthreads_and(mine == per_read) && per_read > 1u;
  const bool one_each = !uniform && __syncthreads_and(mine == 1u);   // per_read == 1: the seed index is the read index
  const uint32_t magic = uniform ? (uint32_t)(0x100000000ull / per_read) : 0u;
  const uint32_t tail = k - 4u * (K4 - 1);                           // characters in the last group, 1..4
  const uint32_t tail_mask = tail >= 4u ? 0xffffffffu : (1u << (8u * tail)) - 1u;
  for (uint32_t base = 0; base < n_cta_seeds; base += 256u * DIRECT_UNROLL) {
    AsciiWords<K4> aw[DIRECT_UNROLL];
    uint32_t rd[DIRECT_UNROLL];
#pragma unroll
    for (int u = 0; u < DIRECT_UNROLL; ++u) {
      const uint32_t ls_raw = base + u * 256u + threadIdx.x;
      const uint32_t ls = ls_raw < n_cta_seeds ? ls_raw : 0u;   // inactive slots re-read seed 0 (valid memory), store nothing
      uint32_t lo = 0;
      if (uniform) {
        lo = __umulhi(ls, magic);
        if ((lo + 1u) * per_read <= ls) ++lo;
      }
      else if (one_each) lo = ls;
      else {
        uint32_t hi = DIRECT_READS;
#pragma unroll 1
        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (s_first[mid] <= ls) lo = mid; else hi = mid; }
      }
      rd[u] = lo;
      load_ascii_words<K4>(bases + s_ptr[lo] + (uint64_t)(ls - s_first[lo]) * d, k, aw[u]);
    }
#pragma unroll
    for (int u = 0; u < DIRECT_UNROLL; ++u) {
      bool valid;
      const uint64_t kmer = pack_ascii_words<K4>(aw[u], tail_mask, valid);
      const uint32_t ls = base + u * 256u + threadIdx.x;
      if (ls < n_cta_seeds) {
        const uint32_t s = first0 + ls;
        seed_kmer[s] = kmer;
        seed_valid[s] = valid ? 1 : 0;
        seed_read[s] = (uint32_t)(r_base + rd[u]);
      }
    }
  }
}

// 2-bit chunks on the separate-kernel route: the exception list becomes the 1-bit "not A/C/G/T" mask the seed
// extraction kernel reads, and equal-length chunks without offsets get them written out.
__global__ void __launch_bounds__(256)
mark_exceptions_kernel(const uint64_t* __restrict__ exc, uint64_t n_exc, uint32_t* __restrict__ nmask)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_exc) return;
  const uint64_t p = exc[i];
  atomicOr(nmask + (p >> 5), 1u << (p & 31u));
}

__global__ void __launch_bounds__(256)
uniform_read_ptr_kernel(uint64_t n_reads, uint32_t read_len, uint64_t* __restrict__ read_ptr)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_reads) read_ptr[r] = r * read_len;
}

void launch_scan_cta_counts(Ctx& c, const uint32_t* cta_count, uint32_t n_ctas, uint32_t* cta_first, unsigned long long* n_seeds_out)
{
  scan_cta_counts_kernel<<<1, 1024, 0, c.stream>>>(cta_count, n_ctas, cta_first, n_seeds_out);
}

template <int K4>
static void launch_seed_reads(Ctx& c, unsigned n_ctas, uint64_t n_reads, unsigned distance)
{
  seed_reads_kernel<K4><<<n_ctas, 256, 0, c.stream>>>(c.d_bases, c.d_read_ptr, c.cta_first.p, n_reads, c.k, distance,
                                                       c.seed_first.p, c.seed_kmer.p, c.seed_valid.p, c.seed_read.p);
}

// K4: one thread per valid seed; k-mer -> chain of seeds.
template <int FMT>
__global__ void __launch_bounds__(256)
build_read_index_kernel(KmerTable t, const uint64_t* __restrict__ seed_kmer, const uint8_t* __restrict__ seed_valid,
                        const unsigned long long* __restrict__ n_seeds_p, uint32_t* __restrict__ seed_next,
                        unsigned long long* __restrict__ err_flag, uint32_t* __restrict__ pfx_bits, uint32_t pfx)
{
  const uint32_t n_seeds = (uint32_t)*n_seeds_p;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seeds) return;
  if (!seed_valid[s]) { seed_next[s] = NIL32; return; }
  {
    // the prefix filter the walker prunes with (seeds.cu, ReadIndexSink::prune)
    const uint32_t i = (uint32_t)(seed_kmer[s] & low_mask64(2u * pfx));
    atomicOr(pfx_bits + (i >> 5), 1u << (i & 31u));
  }
  uint32_t prev = NIL32;
  if (!table_insert<FMT>(t, seed_kmer[s], s, 0, true, prev)) { atomicOr(err_flag, 2ull); prev = NIL32; }
  seed_next[s] = prev;
}

void engine_submit_chunk(Ctx& c, uint64_t n_reads, const uint64_t* read_ptr, const char* bases,
                         uint64_t n_bases, uint64_t first_read_id, unsigned distance, bool on_device)
{
  if (!c.sh->has_graph) throw StateError("submit_chunk: no graph");
  if (c.pending) throw StateError("submit_chunk: a step is in flight on this context (call psi_b200_wait first)");
  if (n_reads && (!read_ptr || !bases)) throw ArgError("submit_chunk: null arrays");
  if (n_reads >= 0x7fffffffull) throw ArgError("submit_chunk: more than 2^31 reads in one chunk");
  PSI_CUDA(cudaSetDevice(c.device));
  c.chunk_packed = false;
  c.read_len = 0;
  c.n_exc = 0;
  if (distance == 0) distance = c.k;  // src/psikt.cpp:469
  c.has_chunk = false;
  c.chunk_indexed = false;
  c.records_valid = false;
  c.n_hits = 0;
  if (!on_device) n_bases = n_reads ? read_ptr[n_reads] : 0;
  // seeds <= n_bases / d + n_reads ; no host pass over the read lengths
  const uint64_t seeds_cap = n_bases / distance + n_reads + 1;
  if (seeds_cap >= 0xfffffff0ull) throw ArgError("submit_chunk: too many seeds in one chunk (use smaller chunks)");

  c.ev_state[T_READ_INDEX] = 0;
  PhaseTimer t_h2d(c, T_H2D);
  if (on_device) {
    c.d_bases = bases;
    c.d_read_ptr = read_ptr;
  }
  else {
    c.bases.ensure(n_bases + 64, 1.25);
    c.read_ptr.ensure(n_reads + 1, 1.25);
    if (n_bases) PSI_CUDA(cudaMemcpyAsync(c.bases.p, bases, n_bases, cudaMemcpyHostToDevice, c.stream));
    PSI_CUDA(cudaMemcpyAsync(c.read_ptr.p, read_ptr, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
    c.d_bases = c.bases.p;
    c.d_read_ptr = c.read_ptr.p;
  }
  t_h2d.stop();

  c.n_reads = n_reads;
  c.n_read_bases = n_bases;
  c.first_read_id = first_read_id;
  c.distance = distance;
  c.n_seeds_cap = seeds_cap;
  c.ev_state[T_PACK] = 0;
  c.chunk_seeded = false;    // seeding is deferred: the fused route (fused.cu) seeds inside its own kernel
  c.has_chunk = true;
  c.counters.n_reads = n_reads;
}

// get_seeds for a chunk that arrives as 2-bit words (psi_b200_packed_chunk).
void engine_submit_chunk_packed(Ctx& c, const psi_b200_packed_chunk& ch, unsigned distance, bool on_device)
{
  if (!c.sh->has_graph) throw StateError("submit_chunk: no graph");
  if (c.pending) throw StateError("submit_chunk: a step is in flight on this context (call psi_b200_wait first)");
  const uint64_t n_reads = ch.n_reads, n_bases = ch.n_bases;
  if (n_reads && !ch.words) throw ArgError("submit_chunk_packed: null words");
  if (n_reads && !ch.read_ptr && ch.read_len == 0) throw ArgError("submit_chunk_packed: neither read offsets nor a read length");
  if (ch.n_exc && !ch.exc) throw ArgError("submit_chunk_packed: null exception list");
  if (n_reads >= 0x7fffffffull) throw ArgError("submit_chunk: more than 2^31 reads in one chunk");
  if (!ch.read_ptr && ch.read_len && n_bases != n_reads * ch.read_len) throw ArgError("submit_chunk_packed: n_bases != n_reads * read_len");
  if (!on_device && ch.read_ptr && n_reads && ch.read_ptr[n_reads] != n_bases) throw ArgError("submit_chunk_packed: read_ptr[n_reads] != n_bases");
  PSI_CUDA(cudaSetDevice(c.device));
  if (distance == 0) distance = c.k;  // src/psikt.cpp:469
  c.has_chunk = false;
  c.chunk_indexed = false;
  c.records_valid = false;
  c.n_hits = 0;
  const uint64_t seeds_cap = n_bases / distance + n_reads + 1;
  if (seeds_cap >= 0xfffffff0ull) throw ArgError("submit_chunk: too many seeds in one chunk (use smaller chunks)");
  const uint64_t n_words = n_bases / 32 + 2;

  c.ev_state[T_READ_INDEX] = 0;
  PhaseTimer t_h2d(c, T_H2D);
  if (on_device) {
    c.d_words = ch.words;
    c.d_read_ptr = ch.read_ptr;
    c.d_exc = ch.exc;
  }
  else {
    c.words.ensure(n_words + 2, 1.25);
    if (n_reads) PSI_CUDA(cudaMemcpyAsync(c.words.p, ch.words, n_words * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
    c.d_words = c.words.p;
    c.d_read_ptr = nullptr;
    if (ch.read_ptr) {
      c.read_ptr.ensure(n_reads + 1, 1.25);
      PSI_CUDA(cudaMemcpyAsync(c.read_ptr.p, ch.read_ptr, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
      c.d_read_ptr = c.read_ptr.p;
    }
    c.d_exc = nullptr;
    if (ch.n_exc) {
      c.exc.ensure(ch.n_exc, 1.25);
      PSI_CUDA(cudaMemcpyAsync(c.exc.p, ch.exc, ch.n_exc * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
      c.d_exc = c.exc.p;
    }
  }
  t_h2d.stop();

  c.chunk_packed = true;
  c.read_len = ch.read_ptr ? 0u : ch.read_len;   // offsets given: the kernels use them (equal lengths are then detected per CTA)
  c.n_exc = ch.n_exc;
  c.d_bases = nullptr;
  c.n_reads = n_reads;
  c.n_read_bases = n_bases;
  c.first_read_id = ch.first_read_id;
  c.distance = distance;
  c.n_seeds_cap = seeds_cap;
  c.ev_state[T_PACK] = 0;
  c.chunk_seeded = false;
  c.has_chunk = true;
  c.counters.n_reads = n_reads;
}

// K1 of the separate-kernel route: seed_first / seed_kmer / seed_valid / seed_read of the submitted chunk.
void engine_seed_chunk(Ctx& c)
{
  if (c.chunk_seeded) return;
  const uint64_t n_reads = c.n_reads, n_bases = c.n_read_bases, seeds_cap = c.n_seeds_cap;
  const unsigned distance = c.distance;
  PhaseTimer t_pack(c, T_PACK);
  c.seed_first.ensure(n_reads + 2, 1.25);
  c.seed_read.ensure(seeds_cap, 1.25);
  c.seed_kmer.ensure(seeds_cap, 1.25);
  c.seed_valid.ensure(seeds_cap + 16, 1.25);

  if (c.chunk_packed) {
    // the chunk IS the 2-bit array the staged seeding reads; only the N mask and (equal-length chunks) the offsets are made
    const uint64_t n_words = (n_bases + 31) >> 5;
    c.reads_n.ensure(n_words + 2, 1.25);
    PSI_CUDA(cudaMemsetAsync(c.reads_n.p, 0, (n_words + 2) * sizeof(uint32_t), c.stream));
    if (c.n_exc) {
      mark_exceptions_kernel<<<grid_for(c.n_exc, 256), 256, 0, c.stream>>>(c.d_exc, c.n_exc, c.reads_n.p);
      ++c.counters.launches;
    }
    if (!c.d_read_ptr) {
      c.read_ptr.ensure(n_reads + 1, 1.25);
      uniform_read_ptr_kernel<<<grid_for(n_reads + 1, 256), 256, 0, c.stream>>>(n_reads, c.read_len, c.read_ptr.p);
      ++c.counters.launches;
      c.d_read_ptr = c.read_ptr.p;
      c.read_len = 0;
    }
    const unsigned n_ctas = (unsigned)((n_reads + 1 + READS_PER_CTA - 1) / READS_PER_CTA);   // + 1: seed_first[n_reads] = total
    c.cta_first.ensure(2 * (size_t)n_ctas + 2, 1.25);
    uint32_t* cta_count = c.cta_first.p + n_ctas + 1;
    count_seeds_kernel<<<n_ctas, 256, 0, c.stream>>>(c.d_read_ptr, n_reads, c.k, distance, cta_count);
    scan_cta_counts_kernel<<<1, 1024, 0, c.stream>>>(cta_count, n_ctas, c.cta_first.p, c.dev_counters.p + DC_SEEDS);
    extract_seeds_kernel<<<n_ctas, 256, 0, c.stream>>>(c.d_words, c.reads_n.p, c.d_read_ptr, c.cta_first.p, n_reads, c.k, distance,
                                                        c.seed_first.p, c.seed_kmer.p, c.seed_valid.p, c.seed_read.p);
    c.counters.launches += 3;
  }
  else if (c.opt_seeding_mode == 0) {
    // direct: ASCII -> seeds (3 launches)
    const unsigned n_ctas = (unsigned)((n_reads + 1 + DIRECT_READS - 1) / DIRECT_READS);   // + 1: seed_first[n_reads] = total
    c.cta_first.ensure(2 * (size_t)n_ctas + 2, 1.25);
    uint32_t* cta_count = c.cta_first.p + n_ctas + 1;
    count_seeds_direct_kernel<<<n_ctas, 256, 0, c.stream>>>(c.d_read_ptr, n_reads, c.k, distance, cta_count);
    scan_cta_counts_kernel<<<1, 1024, 0, c.stream>>>(cta_count, n_ctas, c.cta_first.p, c.dev_counters.p + DC_SEEDS);
    switch ((c.k + 3) / 4) {
      case 1: launch_seed_reads<1>(c, n_ctas, n_reads, distance); break;
      case 2: launch_seed_reads<2>(c, n_ctas, n_reads, distance); break;
      case 3: launch_seed_reads<3>(c, n_ctas, n_reads, distance); break;
      case 4: launch_seed_reads<4>(c, n_ctas, n_reads, distance); break;
      case 5: launch_seed_reads<5>(c, n_ctas, n_reads, distance); break;
      case 6: launch_seed_reads<6>(c, n_ctas, n_reads, distance); break;
      case 7: launch_seed_reads<7>(c, n_ctas, n_reads, distance); break;
      default: launch_seed_reads<8>(c, n_ctas, n_reads, distance); break;
    }
    c.counters.launches += 3;
  }
  else {
    // staged: ASCII -> 2-bit reads -> seeds (4 launches); kept for A/B measurements
    const uint64_t n_words = (n_bases + 31) >> 5;
    c.reads2.ensure(n_words + 2, 1.25);
    c.reads_n.ensure(n_words + 2, 1.25);
    if (n_words) {
      // the word after the last one is read by extract_bases of the final seed
      PSI_CUDA(cudaMemsetAsync(c.reads2.p + n_words, 0, 2 * sizeof(uint64_t), c.stream));
      PSI_CUDA(cudaMemsetAsync(c.reads_n.p + n_words, 0, 2 * sizeof(uint32_t), c.stream));
      if ((reinterpret_cast<uintptr_t>(c.d_bases) & 15u) == 0)
        pack_reads_kernel<true><<<grid_for(n_words, 256), 256, 0, c.stream>>>(c.d_bases, n_bases, c.reads2.p, c.reads_n.p);
      else
        pack_reads_kernel<false><<<grid_for(n_words, 256), 256, 0, c.stream>>>(c.d_bases, n_bases, c.reads2.p, c.reads_n.p);
    }
    const unsigned n_ctas = (unsigned)((n_reads + 1 + READS_PER_CTA - 1) / READS_PER_CTA);   // + 1: seed_first[n_reads] = total
    c.cta_first.ensure(2 * (size_t)n_ctas + 2, 1.25);
    uint32_t* cta_count = c.cta_first.p + n_ctas + 1;
    count_seeds_kernel<<<n_ctas, 256, 0, c.stream>>>(c.d_read_ptr, n_reads, c.k, distance, cta_count);
    scan_cta_counts_kernel<<<1, 1024, 0, c.stream>>>(cta_count, n_ctas, c.cta_first.p, c.dev_counters.p + DC_SEEDS);
    extract_seeds_kernel<<<n_ctas, 256, 0, c.stream>>>(c.reads2.p, c.reads_n.p, c.d_read_ptr, c.cta_first.p, n_reads, c.k, distance,
                                                        c.seed_first.p, c.seed_kmer.p, c.seed_valid.p, c.seed_read.p);
    c.counters.launches += 4;
  }
  t_pack.stop();
  PSI_CUDA(cudaGetLastError());
  c.chunk_seeded = true;
}

// Build the read index of the current chunk (lazily, only when seeds_off_paths
// has loci to walk).
void engine_index_chunk(Ctx& c)
{
  if (c.chunk_indexed) return;
  PhaseTimer t(c, T_READ_INDEX);
  c.seed_next.ensure(c.n_seeds_cap, 1.25);
  unsigned long long* d_err = c.dev_counters.p + DC_ERR;
  table_alloc(c, c.read_index, c.n_seeds_cap, 2 * c.k, c.n_seeds_cap / 256 + 1024);
  c.filter_pfx = c.k < FILTER_PFX ? c.k : FILTER_PFX;
  const size_t filter_words = ((size_t)1 << (2 * c.filter_pfx)) / 32 + 1;
  c.filter_bits.ensure(filter_words);
  PSI_CUDA(cudaMemsetAsync(c.filter_bits.p, 0, filter_words * sizeof(uint32_t), c.stream));
  if (c.read_index.view.fmt == 8)
    build_read_index_kernel<8><<<grid_for(c.n_seeds_cap, 256), 256, 0, c.stream>>>(
        c.read_index.view, c.seed_kmer.p, c.seed_valid.p, c.dev_counters.p + DC_SEEDS, c.seed_next.p, d_err,
        c.filter_bits.p, c.filter_pfx);
  else
    build_read_index_kernel<16><<<grid_for(c.n_seeds_cap, 256), 256, 0, c.stream>>>(
        c.read_index.view, c.seed_kmer.p, c.seed_valid.p, c.dev_counters.p + DC_SEEDS, c.seed_next.p, d_err,
        c.filter_bits.p, c.filter_pfx);
  ++c.counters.launches;
  t.stop();
  PSI_CUDA(cudaGetLastError());
  c.chunk_indexed = true;
}

}  // namespace psi_b200
