// seeding.cuh -- ASCII bases -> packed k-mer, shared by the seeding kernel (chunk.cu) and the fused
// one-pass kernel (fused.cu).  Stands behind seeding() (reference include/psi/sequence.hpp:1688-1745): a seed
// is the k characters at read offset 0, d, 2d, ...; here it becomes one 64-bit word (2 bits per base, first
// base in the low bits) plus a validity flag (any character outside A/C/G/T, either case, makes it invalid:
// 'N' never matches).
#ifndef PSI_B200_DEVICE_SEEDING_CUH
#define PSI_B200_DEVICE_SEEDING_CUH

#include "common.cuh"

namespace psi_b200 {
namespace dev {

// K1 (direct): the k bases at byte address p (any alignment) -> packed k-mer, straight from the ASCII chunk.
// K4 = ceil(k / 4) groups of four characters.  The k bytes lie in K4 or K4 + 1 aligned 32-bit words; a funnel
// shift re-aligns them, then per group (all byte-parallel, ~15 instructions per 4 bases):
//   x = (c >> 1) & 3 ; x ^= x >> 1            A,C,G,T (either case) -> 0,1,2,3 in every byte
//   code = (x * 0x01041040) >> 24             gathers the four 2-bit codes into one byte (the partial
//                                             products occupy disjoint bit fields, so nothing carries)
//   PRMT("ACGT", nibbles(code)) == upper(c)   validates the four characters with one byte permute
// Neighbouring seeds read neighbouring words, so the loads of a warp coalesce in L1 and every chunk byte comes
// from DRAM once.  Loading and packing are separate so that a thread can have the words of several seeds in flight.
template <int K4>
struct AsciiWords {
  uint32_t x[K4 + 1];
  uint32_t sh;
};

template <int K4>
__device__ __forceinline__ void load_ascii_words(const char* p, uint32_t k, AsciiWords<K4>& a)
{
  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~uintptr_t(3));
  const uint32_t off = (uint32_t)(addr & 3u);
  a.sh = off * 8u;
  // words 0 .. K4-1 always hold bytes of the k-mer (k > 4 (K4 - 1)); word K4 only when the k-mer spills into it
#pragma unroll
  for (int i = 0; i < K4; ++i) a.x[i] = __ldg(w + i);
  a.x[K4] = off + k > 4u * K4 ? __ldg(w + K4) : 0u;
}

// tail_mask: byte mask of the characters of the LAST group that belong to the k-mer (all ones when k % 4 == 0)
template <int K4>
__device__ __forceinline__ uint64_t pack_ascii_words(const AsciiWords<K4>& a, uint32_t tail_mask, bool& valid)
{
  uint32_t lo = 0, hi = 0, any_bad = 0;
#pragma unroll
  for (int g = 0; g < K4; ++g) {
    uint32_t c = __funnelshift_r(a.x[g], a.x[g + 1], a.sh);
    if (g == K4 - 1) c = (c & tail_mask) | (0x41414141u & ~tail_mask);    // beyond the k-mer: 'A' = code 0, valid
    uint32_t x = (c >> 1) & 0x03030303u;
    x ^= (x >> 1) & 0x01010101u;
    const uint32_t code = (x * 0x01041040u) >> 24;
    const uint32_t t = (code | (code << 4)) & 0x0f0fu;
    const uint32_t sel = (t | (t << 2)) & 0x3333u;                        // one code per nibble
    any_bad |= (c & 0xdfdfdfdfu) ^ __byte_perm(0x54474341u, 0u, sel);     // non-zero bytes are not A/C/G/T
    if (g < 4) lo |= code << (8 * g);
    else hi |= code << (8 * (g - 4));
  }
  valid = any_bad == 0;
  return ((uint64_t)hi << 32) | lo;
}

}  // namespace dev
}  // namespace psi_b200
#endif
