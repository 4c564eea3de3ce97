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

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

namespace psi_b200 {

using namespace dev;

// seeds of read r: offsets 0, d, 2d, ... while off + k <= len; reads shorter than k have none
struct SeedCountOp {
  const uint64_t* read_ptr;
  uint64_t n_reads;
  uint32_t k, d;
  __host__ __device__ uint32_t operator()(uint64_t r) const
  {
    if (r >= n_reads) return 0;
    const uint64_t len = read_ptr[r + 1] - read_ptr[r];
    return len >= k ? (uint32_t)((len - k) / d) + 1 : 0;
  }
};

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

// K1b: one thread per read: its seeds (offsets 0, d, 2d, ...) cut out of the 2-bit array.
__global__ void __launch_bounds__(256)
extract_seeds_kernel(const uint64_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, const uint64_t* __restrict__ read_ptr,
                     const uint32_t* __restrict__ seed_first, uint64_t n_reads, uint32_t k, uint32_t d,
                     uint64_t* __restrict__ seed_kmer, uint8_t* __restrict__ seed_valid, uint32_t* __restrict__ seed_read,
                     unsigned long long* __restrict__ n_seeds_out)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) *n_seeds_out = seed_first[n_reads];
  if (r >= n_reads) return;
  const uint32_t first = seed_first[r];
  const uint32_t cnt = seed_first[r + 1] - first;
  uint64_t pos = read_ptr[r];
  for (uint32_t j = 0; j < cnt; ++j, pos += d) {
    seed_kmer[first + j] = extract_bases(seq2, pos, k);
    seed_valid[first + j] = extract_nmask(nmask, pos, k) ? 0 : 1;
    seed_read[first + j] = (uint32_t)r;
  }
}

// K4: one thread per valid seed; k-mer -> chain of seeds.
template <int FMT>
__global__ void __launch_bounds__(256)
build_read_index_kernel(KmerTable t, const uint64_t* __restrict__ seed_kmer, const uint8_t* __restrict__ seed_valid,
                        const unsigned long long* __restrict__ n_seeds_p, uint32_t* __restrict__ seed_next,
                        unsigned long long* __restrict__ err_flag)
{
  const uint32_t n_seeds = (uint32_t)*n_seeds_p;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seeds) return;
  if (!seed_valid[s]) { seed_next[s] = NIL32; return; }
  uint32_t prev = NIL32;
  if (!table_insert<FMT>(t, seed_kmer[s], s, 0, true, prev)) { atomicOr(err_flag, 2ull); prev = NIL32; }
  seed_next[s] = prev;
}

void engine_submit_chunk(Ctx& c, uint64_t n_reads, const uint64_t* read_ptr, const char* bases,
                         uint64_t n_bases, uint64_t first_read_id, unsigned distance, bool on_device)
{
  if (!c.sh->has_graph) throw StateError("submit_chunk: no graph");
  if (n_reads && (!read_ptr || !bases)) throw ArgError("submit_chunk: null arrays");
  if (n_reads >= 0x7fffffffull) throw ArgError("submit_chunk: more than 2^31 reads in one chunk");
  PSI_CUDA(cudaSetDevice(c.device));
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

  PhaseTimer t_pack(c, T_PACK);
  c.seed_first.ensure(n_reads + 2, 1.25);
  c.seed_read.ensure(seeds_cap, 1.25);
  c.seed_kmer.ensure(seeds_cap, 1.25);
  c.seed_valid.ensure(seeds_cap + 16, 1.25);
  c.n_reads = n_reads;
  c.n_read_bases = n_bases;
  c.first_read_id = first_read_id;
  c.distance = distance;
  c.n_seeds_cap = seeds_cap;

  {
    SeedCountOp op{ c.d_read_ptr, n_reads, c.k, distance };
    auto counts = thrust::make_transform_iterator(thrust::counting_iterator<uint64_t>(0), op);
    size_t tmp = 0;
    PSI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, c.seed_first.p, (int64_t)(n_reads + 1), c.stream));
    c.scan_tmp.ensure(tmp);
    PSI_CUDA(cub::DeviceScan::ExclusiveSum(c.scan_tmp.p, tmp, counts, c.seed_first.p, (int64_t)(n_reads + 1), c.stream));
  }
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
  extract_seeds_kernel<<<grid_for(n_reads + 1, 256), 256, 0, c.stream>>>(c.reads2.p, c.reads_n.p, c.d_read_ptr, c.seed_first.p, n_reads,
                                                                         c.k, distance, c.seed_kmer.p, c.seed_valid.p,
                                                                         c.seed_read.p, c.dev_counters.p + DC_SEEDS);
  c.counters.launches += 4;
  t_pack.stop();
  PSI_CUDA(cudaGetLastError());
  c.has_chunk = true;
  c.counters.n_reads = n_reads;
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
  if (c.read_index.view.fmt == 8)
    build_read_index_kernel<8><<<grid_for(c.n_seeds_cap, 256), 256, 0, c.stream>>>(
        c.read_index.view, c.seed_kmer.p, c.seed_valid.p, c.dev_counters.p + DC_SEEDS, c.seed_next.p, d_err);
  else
    build_read_index_kernel<16><<<grid_for(c.n_seeds_cap, 256), 256, 0, c.stream>>>(
        c.read_index.view, c.seed_kmer.p, c.seed_valid.p, c.dev_counters.p + DC_SEEDS, c.seed_next.p, d_err);
  ++c.counters.launches;
  t.stop();
  PSI_CUDA(cudaGetLastError());
  c.chunk_indexed = true;
}

}  // namespace psi_b200
