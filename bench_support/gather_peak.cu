// gather_peak.cu -- measurement tool (not product code): what can a B200 deliver for
// RANDOM 32-byte sector reads out of a table much larger than L2?  This is the
// access pattern of the seeds_on_paths probe (one bucket = one sector per query
// seed), so its result is the practical ceiling beside the streaming-copy peak
// in MEASURED_PEAKS.json.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_peak gather_peak.cu
//   ./gather_peak [table MiB = 2048] [probes = 5000000] [reps = 10]
//
// Prints one JSON line per configuration:
//   per-thread sector loads (1/2/4 sectors per probe), lane-cooperative line loads,
//   16-byte loads, L2-prefetch-then-load and TMA bulk (cp.async.bulk) gathers.
// cudaLimitMaxL2FetchGranularity (32/64/128) was measured to make no difference
// (profiles/r01b_gather_peak.jsonl), so it is not swept here.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

__device__ __forceinline__ void ld_sector(const void* p, uint64_t (&v)[4])
{
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

// each thread: ITEMS probes, SECTORS adjacent sectors per probe, all loads issued before the first use
template <int ITEMS, int SECTORS>
__global__ void __launch_bounds__(256)
gather_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
              unsigned long long* __restrict__ sink)
{
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
  uint64_t v[ITEMS][SECTORS][4];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    uint64_t sec = mix64(key) & n_sectors_mask;
    sec &= ~(uint64_t)(SECTORS - 1);
#pragma unroll
    for (int j = 0; j < SECTORS; ++j) ld_sector(table + ((sec + j) << 5), v[i][j]);
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
#pragma unroll
    for (int j = 0; j < SECTORS; ++j) acc ^= v[i][j][0] ^ v[i][j][1] ^ v[i][j][2] ^ v[i][j][3];
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}


// COOP lanes cooperate on one probe: lane j of the group loads sector j of the probe's line
// (one warp instruction touches 32/COOP lines with COOP sectors each).
template <int ITEMS, int COOP>
__global__ void __launch_bounds__(256)
gather_coop_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                   unsigned long long* __restrict__ sink)
{
  const uint32_t sub = threadIdx.x % COOP;
  const uint32_t base = (blockIdx.x * 256u + threadIdx.x) / COOP;      // probe group of this thread
  const uint32_t stride = gridDim.x * 256u / COOP;
  uint64_t v[ITEMS][4];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * stride;
    const uint64_t key = s < n ? keys[s] : 0;
    uint64_t sec = (mix64(key) & n_sectors_mask) & ~(uint64_t)(COOP - 1);
    ld_sector(table + ((sec + sub) << 5), v[i]);
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) acc ^= v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3];
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// 16 bytes per probe (LDG.128): is the ceiling per request or per byte?
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather16_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                unsigned long long* __restrict__ sink)
{
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
  uint64_t a[ITEMS], b[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    const uint64_t sec = mix64(key) & n_sectors_mask;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a[i]), "=l"(b[i]) : "l"(table + (sec << 5)));
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) acc ^= a[i] ^ b[i];
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// L2 prefetch of every probe's sector first, demand loads afterwards (software pipelining through L2)
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather_prefetch_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                       unsigned long long* __restrict__ sink)
{
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
  uint64_t sec[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    sec[i] = mix64(key) & n_sectors_mask;
    asm volatile("prefetch.global.L2 [%0];" :: "l"(table + (sec[i] << 5)));
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    uint64_t v[4];
    ld_sector(table + (sec[i] << 5), v);
    acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// TMA path: every lane issues ITEMS 32-byte cp.async.bulk global->shared copies that complete on one
// mbarrier per CTA; the data never passes through the LSU/L1 miss path.
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather_bulk_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                   unsigned long long* __restrict__ sink)
{
  __shared__ alignas(128) uint64_t buf[ITEMS][256][4];
  __shared__ alignas(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(ITEMS * 256 * 32) : "memory");
  __syncthreads();
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    const uint64_t sec = mix64(key) & n_sectors_mask;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&buf[i][threadIdx.x][0]);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                 :: "r"(dst), "l"(table + (sec << 5)), "r"(bar_a) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_a) : "memory");
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) acc ^= buf[i][threadIdx.x][0] ^ buf[i][threadIdx.x][1] ^ buf[i][threadIdx.x][2] ^ buf[i][threadIdx.x][3];
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}


// TMA bulk copies of a whole 128-byte line per probe into shared memory: ITEMS lines per thread in flight,
// no registers hold data.  Dynamic shared memory: ITEMS * 256 * 128 bytes.
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather_bulk128_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                      unsigned long long* __restrict__ sink)
{
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ alignas(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(ITEMS * 256 * 128) : "memory");
  __syncthreads();
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    const uint64_t line = (mix64(key) & n_sectors_mask) >> 2;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dyn + ((size_t)(i * 256 + threadIdx.x) << 7));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                 :: "r"(dst), "l"(table + (line << 7)), "r"(bar_a) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_a) : "memory");
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint4* ln = reinterpret_cast<const uint4*>(dyn + ((size_t)(i * 256 + threadIdx.x) << 7));
#pragma unroll
    for (int j = 0; j < 8; ++j) { const uint4 w = ln[(j + threadIdx.x) & 7]; acc ^= w.x ^ w.y ^ w.z ^ w.w; }
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// cp.async (LDGSTS) 16 bytes per lane, 8 lanes per 128-byte line, ITEMS instructions per thread in flight.
// Dynamic shared memory: ITEMS * 256 * 16 bytes.
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather_ldgsts_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                     unsigned long long* __restrict__ sink)
{
  extern __shared__ __align__(128) unsigned char dyn[];
  const uint32_t sub = threadIdx.x & 7u;
  const uint32_t grp = threadIdx.x >> 3;                       // 32 line groups per CTA
  const uint32_t base = blockIdx.x * (32u * ITEMS) + grp;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 32u;
    const uint64_t key = s < n ? keys[s] : 0;
    const uint64_t line = (mix64(key) & n_sectors_mask) >> 2;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dyn + ((size_t)(i * 32 + grp) << 7) + (sub << 4));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(table + (line << 7) + (sub << 4)) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint4 w = *reinterpret_cast<const uint4*>(dyn + ((size_t)(i * 32 + grp) << 7) + (sub << 4));
    acc ^= w.x ^ w.y ^ w.z ^ w.w;
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// one 32-byte sector per probe, one thread per probe, REGS-light: how far does occupancy carry the line rate?
template <int ITEMS, int MINB>
__global__ void __launch_bounds__(256, MINB)
gather_occ_kernel(const char* __restrict__ table, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
                  unsigned long long* __restrict__ sink)
{
  const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
  uint64_t v[ITEMS][4];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t s = base + i * 256u;
    const uint64_t key = s < n ? keys[s] : 0;
    ld_sector(table + ((mix64(key) & n_sectors_mask) << 5), v[i]);
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) acc ^= v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3];
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// sm_100 TMA gather: cp.async.bulk.tensor.2d.tile::gather4 copies FOUR rows of a [n_lines, 128 B] tensor map into shared
// memory with one instruction (one column coordinate, four row coordinates), completion on an mbarrier.  Lanes 0..7 of
// each warp issue one gather4 for four probes each, ITEMS rounds; dynamic shared memory ITEMS * 256 * 128 bytes + 8.
#include <cuda.h>
template <int ITEMS>
__global__ void __launch_bounds__(256)
gather4_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n_sectors_mask, const uint64_t* __restrict__ keys, uint32_t n,
               unsigned long long* __restrict__ sink)
{
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"((uint32_t)(ITEMS * 256 * 128)) : "memory");
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t p0 = blockIdx.x * (256u * ITEMS) + i * 256u + warp * 32u + lane * 4u;
      int32_t row[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t key = p0 + j < n ? keys[p0 + j] : 0;
        row[j] = (int32_t)((mix64(key) & n_sectors_mask) >> 2);
      }
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dyn + ((size_t)(i * 256 + warp * 32 + lane * 4) << 7));
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                   :: "r"(dst), "l"(&tmap), "r"(0), "r"(row[0]), "r"(row[1]), "r"(row[2]), "r"(row[3]), "r"(bar_a) : "memory");
    }
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar_a) : "memory");
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint4* ln = reinterpret_cast<const uint4*>(dyn + ((size_t)(i * 256 + threadIdx.x) << 7));
#pragma unroll
    for (int j = 0; j < 8; ++j) { const uint4 w = ln[(j + threadIdx.x) & 7]; acc ^= w.x ^ w.y ^ w.z ^ w.w; }
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
static bool make_line_tensor_map(CUtensorMap* out, void* table, uint64_t n_lines, unsigned box_rows)
{
  typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
  const cuuint64_t dims[2] = { 128, n_lines };
  const cuuint64_t strides[1] = { 128 };
  const cuuint32_t box[2] = { 128, box_rows };
  const cuuint32_t estr[2] = { 1, 1 };
  const CUresult r = ((encode_t)fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) std::fprintf(stderr, "cuTensorMapEncodeTiled(box rows %u): error %d\n", box_rows, (int)r);
  return r == CUDA_SUCCESS;
}

template <class F>
static float time_launch(F launch, int reps, char* flush, size_t flush_bytes)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemsetAsync(flush, r, flush_bytes));
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    if (r >= 2) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}

__global__ void fill_kernel(uint64_t* p, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = mix64(i);
}

__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int ITEMS, int SECTORS>
static float run(const char* table, uint64_t mask, const uint64_t* keys, uint32_t n, unsigned long long* sink, int reps,
                 char* flush, size_t flush_bytes)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned grid = (n + 256 * ITEMS - 1) / (256 * ITEMS);
  std::vector<float> ms;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemsetAsync(flush, r, flush_bytes));  // evict the table's lines from L2
    CK(cudaEventRecord(e0));
    gather_kernel<ITEMS, SECTORS><<<grid, 256>>>(table, mask, keys, n, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    if (r >= 2) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}

int main(int argc, char** argv)
{
  const uint64_t table_mib = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 2048;
  const uint32_t n = argc > 2 ? (uint32_t)std::strtoull(argv[2], nullptr, 10) : 5000000u;
  const int reps = argc > 3 ? std::atoi(argv[3]) : 10;
  uint64_t bytes = 1; while (bytes < (table_mib << 20)) bytes <<= 1;
  char* table; uint64_t* keys; unsigned long long* sink; char* flush;
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&table, bytes)); CK(cudaMalloc(&keys, (size_t)n * 8)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&flush, flush_bytes));
  fill_kernel<<<1184, 256>>>((uint64_t*)table, bytes / 8);
  fill_kernel<<<1184, 256>>>(keys, n);
  CK(cudaMemset(sink, 0, 8));
  CK(cudaDeviceSynchronize());
  const uint64_t mask = (bytes >> 5) - 1;
  const uint64_t table_bytes = bytes;

  // streaming copy for reference (read + write bytes)
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      CK(cudaEventRecord(e0));
      copy_kernel<<<148 * 16, 512>>>((const uint4*)table, (uint4*)(table + bytes / 2), bytes / 32);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float t; CK(cudaEventElapsedTime(&t, e0, e1)); best = std::min(best, t);
    }
    std::printf("{\"test\": \"stream_copy\", \"bytes\": %llu, \"ms\": %.4f, \"GBps\": %.1f}\n", (unsigned long long)bytes, best, bytes / (best * 1e-3) / 1e9);
  }

  size_t g = 0;
  CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity));
  const char* only = argc > 4 ? argv[4] : "";
  if (std::string(only) == "gather4") {
    // ./gather_peak MiB probes reps gather4 [box rows = 4]: the TMA gather4 variants alone (a faulting variant must
    // not take the other measurements down, so this mode runs in a process of its own)
    const unsigned box_rows = argc > 5 ? (unsigned)std::atoi(argv[5]) : 4u;
    CUtensorMap tmap;
    if (!make_line_tensor_map(&tmap, table, bytes >> 7, box_rows)) { std::printf("{\"test\": \"random_gather\", \"variant\": \"tma_gather4\", \"error\": \"tensor map\"}\n"); return 0; }
    auto rep4 = [&](const char* name, float ms) {
      std::printf("{\"test\": \"random_gather\", \"variant\": \"%s_box%u\", \"table_bytes\": %llu, \"probes\": %u, \"bytes_per_probe\": 128, "
                  "\"ms\": %.4f, \"Gprobes_per_s\": %.2f, \"useful_GBps\": %.1f}\n", name, box_rows, (unsigned long long)table_bytes, n, ms,
                  n / (ms * 1e-3) / 1e9, (double)n * 128 / (ms * 1e-3) / 1e9);
    };
    CK(cudaFuncSetAttribute(gather4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 * 256 * 128));
    CK(cudaFuncSetAttribute(gather4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 128));
    CK(cudaFuncSetAttribute(gather4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 256 * 128));
    rep4("tma_gather4_128B_items1", time_launch([&] { gather4_kernel<1><<<(n + 255) / 256, 256, 1 * 256 * 128>>>(tmap, mask, keys, n, sink); }, reps, flush, flush_bytes));
    rep4("tma_gather4_128B_items2", time_launch([&] { gather4_kernel<2><<<(n + 511) / 512, 256, 2 * 256 * 128>>>(tmap, mask, keys, n, sink); }, reps, flush, flush_bytes));
    rep4("tma_gather4_128B_items4", time_launch([&] { gather4_kernel<4><<<(n + 1023) / 1024, 256, 4 * 256 * 128>>>(tmap, mask, keys, n, sink); }, reps, flush, flush_bytes));
    // the LDGSTS variant the product kernel uses, in the same process for a like-for-like number
    CK(cudaFuncSetAttribute(gather_ldgsts_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 256 * 16));
    rep4("ldgsts_128B_items8", time_launch([&] { gather_ldgsts_kernel<8><<<(n + 255) / 256, 256, 8 * 256 * 16>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes));
    return 0;
  }
  auto report = [&](const char* name, float ms, int bytes, double probes) {
    std::printf("{\"test\": \"random_gather\", \"variant\": \"%s\", \"table_bytes\": %llu, \"probes\": %.0f, \"bytes_per_probe\": %d, "
                "\"ms\": %.4f, \"Gprobes_per_s\": %.2f, \"useful_GBps\": %.1f, \"useful_plus_key_GBps\": %.1f}\n",
                name, (unsigned long long)table_bytes, probes, bytes, ms,
                probes / (ms * 1e-3) / 1e9, probes * bytes / (ms * 1e-3) / 1e9, probes * (bytes + 8) / (ms * 1e-3) / 1e9);
  };
  (void)g;
  report("1x32B_items4", run<4, 1>(table, mask, keys, n, sink, reps, flush, flush_bytes), 32, n);
  report("1x32B_items2", run<2, 1>(table, mask, keys, n, sink, reps, flush, flush_bytes), 32, n);
  report("1x32B_items1", run<1, 1>(table, mask, keys, n, sink, reps, flush, flush_bytes), 32, n);
  report("2x32B_items4", run<4, 2>(table, mask, keys, n, sink, reps, flush, flush_bytes), 64, n);
  report("4x32B_items4", run<4, 4>(table, mask, keys, n, sink, reps, flush, flush_bytes), 128, n);
  {
    const unsigned grid4 = ((uint64_t)n * 4 + 256 * 4 - 1) / (256 * 4);
    report("coop4_128B_items4", time_launch([&] { gather_coop_kernel<4, 4><<<grid4, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
    const unsigned grid2 = ((uint64_t)n * 2 + 256 * 4 - 1) / (256 * 4);
    report("coop2_64B_items4", time_launch([&] { gather_coop_kernel<4, 2><<<grid2, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 64, n);
    const unsigned grid1 = (n + 256 * 4 - 1) / (256 * 4);
    report("16B_items4", time_launch([&] { gather16_kernel<4><<<grid1, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 16, n);
    report("prefetchL2_32B_items4", time_launch([&] { gather_prefetch_kernel<4><<<grid1, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
    const unsigned grid8 = (n + 256 * 8 - 1) / (256 * 8);
    report("prefetchL2_32B_items8", time_launch([&] { gather_prefetch_kernel<8><<<grid8, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
    report("tma_bulk_32B_items4", time_launch([&] { gather_bulk_kernel<4><<<grid1, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
    const unsigned grid2b = (n + 256 * 2 - 1) / (256 * 2);
    report("tma_bulk_32B_items2", time_launch([&] { gather_bulk_kernel<2><<<grid2b, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
  }
    {
      // in-flight capacity experiments
      CK(cudaFuncSetAttribute(gather_bulk128_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 * 256 * 128));
      CK(cudaFuncSetAttribute(gather_bulk128_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 128));
      CK(cudaFuncSetAttribute(gather_bulk128_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 256 * 128));
      report("tma_bulk_128B_items1", time_launch([&] { gather_bulk128_kernel<1><<<(n + 255) / 256, 256, 1 * 256 * 128>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      report("tma_bulk_128B_items2", time_launch([&] { gather_bulk128_kernel<2><<<(n + 511) / 512, 256, 2 * 256 * 128>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      report("tma_bulk_128B_items4", time_launch([&] { gather_bulk128_kernel<4><<<(n + 1023) / 1024, 256, 4 * 256 * 128>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      CK(cudaFuncSetAttribute(gather_ldgsts_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 256 * 16));
      CK(cudaFuncSetAttribute(gather_ldgsts_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 256 * 16));
      report("ldgsts_128B_items4", time_launch([&] { gather_ldgsts_kernel<4><<<(n + 127) / 128, 256, 4 * 256 * 16>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      report("ldgsts_128B_items8", time_launch([&] { gather_ldgsts_kernel<8><<<(n + 255) / 256, 256, 8 * 256 * 16>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      report("ldgsts_128B_items16", time_launch([&] { gather_ldgsts_kernel<16><<<(n + 511) / 512, 256, 16 * 256 * 16>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 128, n);
      report("sector_items1_occ8", time_launch([&] { gather_occ_kernel<1, 8><<<(n + 255) / 256, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
      report("sector_items2_occ8", time_launch([&] { gather_occ_kernel<2, 8><<<(n + 511) / 512, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
      report("sector_items4_occ4", time_launch([&] { gather_occ_kernel<4, 4><<<(n + 1023) / 1024, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
      report("sector_items4_occ2", time_launch([&] { gather_occ_kernel<4, 2><<<(n + 1023) / 1024, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
      report("sector_items8_occ2", time_launch([&] { gather_occ_kernel<8, 2><<<(n + 2047) / 2048, 256>>>(table, mask, keys, n, sink); }, reps, flush, flush_bytes), 32, n);
    }
  return 0;
}
