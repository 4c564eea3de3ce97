// records.cuh -- from a compact hit (seed, global graph position) to the reference's seed record
// (reference include/psi/seed.hpp:32-46 as written by src/psikt.cpp:172-181): node_id, node_offset,
// read_id, read_offset.  Shared by the resolve kernels (seeds.cu) and the fused one-pass kernel (fused.cu).
#ifndef PSI_B200_DEVICE_RECORDS_CUH
#define PSI_B200_DEVICE_RECORDS_CUH

#include "walker.cuh"

namespace psi_b200 {

using namespace dev;

// which API call reports an index entry: 1 = seeds_on_paths, 2 = seeds_off_paths, 0 = not asked for
__device__ __forceinline__ uint8_t kind_of(uint32_t flags, uint32_t mode)
{
  return (flags & FLAG_OFF) ? ((mode & PSI_B200_OFF_PATHS) ? 2 : 0) : ((mode & PSI_B200_ON_PATHS) ? 1 : 0);
}

struct Resolved {
  uint64_t node_id, node_off, read_id, read_off;
};

// global graph position -> (node id, offset in the node); replaces position_to_id/offset(PathIndex)
// (reference include/psi/pathindex.hpp:378-416).  Only locus LISTS and walker hits carry positions; single-locus
// index entries carry a locus code that decode_code (walker.cuh) turns into the same two fields without a gather.
__device__ __forceinline__ void resolve_node(const GraphView& g, uint32_t gpos, uint64_t& id, uint64_t& off)
{
  if (g.rank16) {
    // two dependent 16-byte gathers: start bits + prefix count of the 64 positions, then the node's record
    const uint4 rw = __ldg(reinterpret_cast<const uint4*>(g.rank16 + (gpos >> 6)));
    const uint64_t bits = ((uint64_t)rw.y << 32) | rw.x;
    const uint32_t v = rw.z + (uint32_t)__popcll(bits & (~0ull >> (63u - (gpos & 63u)))) - 1u;
    const uint4 nr = __ldg(reinterpret_cast<const uint4*>(g.node_res + v));
    off = gpos - nr.x;
    id = ((uint64_t)nr.w << 32) | nr.z;
  }
  else {
    const uint32_t v = node_of_pos(g, gpos);
    off = gpos - __ldg(&g.rec[v].seq_start);
    id = __ldg(g.node_id + v);
  }
}

// seed index -> (read id, offset of the seed in the read); replaces Records::position_to_id/offset
// (reference include/psi/sequence.hpp:1201-1213,1277-1289)
__device__ __forceinline__ void resolve_read(uint32_t seed, const uint32_t* __restrict__ seed_read, const uint32_t* __restrict__ seed_first,
                                             uint32_t d, uint64_t first_read_id, Resolved& o)
{
  const uint32_t r = __ldg(seed_read + seed);
  o.read_id = first_read_id + r;
  o.read_off = (uint64_t)(seed - __ldg(seed_first + r)) * d;
}

// one 32-byte record per store instruction (STG.256): every 32-byte sector of the output is written exactly once
__device__ __forceinline__ void st_record(uint64_t* dst, const Resolved& r)
{
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "l"(r.node_id), "l"(r.node_off), "l"(r.read_id), "l"(r.read_off) : "memory");
}

// compact form of the same record: 4 x u32 in the same field order (PSI_B200_COMPACT; the host checked that every node id
// and read id of the chunk fits 32 bits) -- half the bytes to write here and to move over PCIe
__device__ __forceinline__ void st_record32(uint64_t* dst, const Resolved& r)
{
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((uint32_t)r.node_id), "r"((uint32_t)r.node_off),
               "r"((uint32_t)r.read_id), "r"((uint32_t)r.read_off) : "memory");
}

}  // namespace psi_b200
#endif
