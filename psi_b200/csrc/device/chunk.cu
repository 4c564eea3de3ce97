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

namespace psi_b200 {

using namespace dev;

// seeds per read -> cnt[r]; cnt[n_reads] = 0
__global__ void __launch_bounds__(256)
count_seeds_kernel(const uint64_t* __restrict__ read_ptr, uint64_t n_reads, uint32_t k, uint32_t d,
                   uint32_t* __restrict__ cnt)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_reads) return;
  uint32_t c = 0;
  if (r < n_reads) {
    const uint64_t len = read_ptr[r + 1] - read_ptr[r];
    if (len >= k) c = (uint32_t)((len - k) / d) + 1;   // reads shorter than k: no seeds
  }
  cnt[r] = c;
}

// seed_read[s] = r for the seeds of read r
__global__ void __launch_bounds__(256)
fill_seed_read_kernel(const uint32_t* __restrict__ seed_first, uint64_t n_reads, uint32_t* __restrict__ seed_read)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t b = seed_first[r], e = seed_first[r + 1];
  for (uint32_t s = b; s < e; ++s) seed_read[s] = (uint32_t)r;
}

// K1: one thread per seed; ASCII -> packed k-mer + validity bit.
__global__ void __launch_bounds__(256)
pack_seeds_kernel(const char* __restrict__ bases, const uint64_t* __restrict__ read_ptr,
                  const uint32_t* __restrict__ seed_first, const uint32_t* __restrict__ seed_read,
                  uint64_t n_reads, uint32_t k, uint32_t d,
                  uint64_t* __restrict__ seed_kmer, uint32_t* __restrict__ seed_valid,
                  unsigned long long* __restrict__ n_seeds_out)
{
  const uint32_t n_seeds = seed_first[n_reads];
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) *n_seeds_out = n_seeds;
  // whole warps stay in the ballot below
  const bool in_range = s < n_seeds;
  uint64_t kmer = 0;
  bool valid = false;
  if (in_range) {
    const uint32_t r = seed_read[s];
    const uint64_t p = read_ptr[r] + (uint64_t)(s - seed_first[r]) * d;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(bases) + p;
    valid = true;
    for (uint32_t i = 0; i < k; ++i) {
      const uint32_t c = base_code(__ldg(b + i));
      if (c > 3) valid = false;
      kmer |= (uint64_t)(c & 3u) << (2u * i);
    }
    seed_kmer[s] = kmer;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31u) == 0) seed_valid[s >> 5] = m;
}

// K4: one thread per valid seed; k-mer -> chain of seeds.
template <int FMT>
__global__ void __launch_bounds__(256)
build_read_index_kernel(KmerTable t, const uint64_t* __restrict__ seed_kmer, const uint32_t* __restrict__ seed_valid,
                        const unsigned long long* __restrict__ n_seeds_p, uint32_t* __restrict__ seed_next,
                        unsigned long long* __restrict__ err_flag)
{
  const uint32_t n_seeds = (uint32_t)*n_seeds_p;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seeds) return;
  if (!((seed_valid[s >> 5] >> (s & 31u)) & 1u)) { seed_next[s] = NIL32; return; }
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
  c.seed_valid.ensure((seeds_cap >> 5) + 16, 1.25);
  c.n_reads = n_reads;
  c.n_read_bases = n_bases;
  c.first_read_id = first_read_id;
  c.distance = distance;
  c.n_seeds_cap = seeds_cap;

  uint32_t* cnt = c.seed_first.p;  // scanned in place
  count_seeds_kernel<<<grid_for(n_reads + 1, 256), 256, 0, c.stream>>>(c.d_read_ptr, n_reads, c.k, distance, cnt);
  size_t tmp = 0;
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, cnt, (int64_t)(n_reads + 1), c.stream));
  c.scan_tmp.ensure(tmp);
  PSI_CUDA(cub::DeviceScan::ExclusiveSum(c.scan_tmp.p, tmp, cnt, cnt, (int64_t)(n_reads + 1), c.stream));
  if (n_reads) fill_seed_read_kernel<<<grid_for(n_reads, 256), 256, 0, c.stream>>>(c.seed_first.p, n_reads, c.seed_read.p);
  pack_seeds_kernel<<<grid_for(seeds_cap, 256), 256, 0, c.stream>>>(c.d_bases, c.d_read_ptr, c.seed_first.p, c.seed_read.p,
                                                                    n_reads, c.k, distance, c.seed_kmer.p, c.seed_valid.p,
                                                                    c.dev_counters.p + DC_SEEDS);
  c.counters.launches += 5;
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
  // per-chunk table: inflate by at most 64 MiB to keep the compact slot format
  table_alloc(c, c.read_index, c.n_seeds_cap, 2 * c.k, 64ull << 20, c.n_seeds_cap / 256 + 1024);
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
