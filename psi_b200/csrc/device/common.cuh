// common.cuh -- shared device code of libpsi_b200 (sm_100a only).
//
//   * k-mer packing: 2 bits per base, base i of the k-mer at bits [2i, 2i+2)
//   * KmerTable: the bucketised hash used for BOTH device indexes
//       - the path index       (k-mer -> locus / locus list; probed by seeds_on_paths)
//       - the chunk read index (k-mer -> chain of read seeds; probed by seeds_off_paths)
//     A bucket is one 32-byte DRAM sector; 4 buckets form a 128-byte line and a
//     key may only live in its home line (home bucket first, then the other
//     three), so a probe costs one sector in the common case and never leaves
//     one L2 line.  Keys that find their line full go to a small stash.
#ifndef PSI_B200_DEVICE_COMMON_CUH
#define PSI_B200_DEVICE_COMMON_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace psi_b200 {
namespace dev {

constexpr uint32_t NIL32 = 0xffffffffu;
constexpr uint64_t EMPTY8 = ~0ull;

// --------------------------------------------------------------- bases --

// A/C/G/T (either case) -> 0..3, anything else -> 4.
__host__ __device__ __forceinline__ uint32_t base_code(unsigned char c)
{
  c &= 0xdf;  // upper case
  // A=0x41 C=0x43 G=0x47 T=0x54
  if (c == 'A') return 0;
  if (c == 'C') return 1;
  if (c == 'G') return 2;
  if (c == 'T') return 3;
  return 4;
}

__host__ __device__ __forceinline__ uint64_t low_mask64(uint32_t bits)
{
  return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
}

// `count` (1..32) bases starting at base index `pos` of a 2-bit packed array
// (32 bases per 64-bit word), returned with the first base in the low bits.
__device__ __forceinline__ uint64_t extract_bases(const uint64_t* __restrict__ seq2, uint64_t pos, uint32_t count)
{
  const uint64_t w = pos >> 5;
  const uint32_t sh = (uint32_t)(pos & 31u) * 2u;
  uint64_t v = __ldg(seq2 + w) >> sh;
  if (sh != 0 && (64u - sh) < count * 2u) v |= __ldg(seq2 + w + 1) << (64u - sh);
  return v & low_mask64(count * 2u);
}

// `count` (1..32) mask bits starting at bit index `pos` (1 = not A/C/G/T).
__device__ __forceinline__ uint32_t extract_nmask(const uint32_t* __restrict__ nmask, uint64_t pos, uint32_t count)
{
  const uint64_t w = pos >> 5;
  const uint32_t sh = (uint32_t)(pos & 31u);
  uint64_t v = (uint64_t)__ldg(nmask + w);
  if (sh + count > 32u) v |= (uint64_t)__ldg(nmask + w + 1) << 32;
  return (uint32_t)((v >> sh) & low_mask64(count));
}

// -------------------------------------------------------------- hashing --

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Bijection on kb-bit integers (odd multiplications and right xor-shifts are
// invertible modulo 2^kb); the TOP bits of the result select the bucket.
__host__ __device__ __forceinline__ uint64_t mix_kb(uint64_t x, uint32_t kb)
{
  const uint64_t m = low_mask64(kb);
  const uint32_t s = (kb >> 1) ? (kb >> 1) : 1;
  x = (x * 0x9e3779b97f4a7c15ULL) & m;
  x ^= x >> s;
  x = (x * 0xd6e8feb86659fd93ULL) & m;
  x ^= x >> s;
  x = (x * 0xca5a826395121157ULL) & m;
  return x;
}

// ----------------------------------------------------------- KmerTable --

struct alignas(16) Slot16 {
  uint64_t key;
  uint32_t payload;
  uint32_t flags;  // 0 = single, 1 = multi, NIL32 = empty
};

struct KmerTable {
  void*    slots;       // n_lines * 128 bytes
  Slot16*  stash;       // (stash_mask + 1) slots, open addressing on full keys
  uint32_t* stash_used; // device counter of occupied stash slots
  uint32_t line_bits;   // log2(n_lines)
  uint32_t kbits;       // 2k
  uint32_t rem_bits;    // format 8 only: kbits - 2 - line_bits (<= 29)
  uint32_t fmt;         // 8: 4 x 8-byte slots per bucket; 16: 2 x 16-byte slots
  uint32_t stash_mask;
  uint32_t stash_nonempty;  // host-known: 0 lets lookups skip the stash entirely
};

struct Home {
  uint64_t line;
  uint32_t sec;
  uint64_t tag;   // fmt 8: (remainder << 2) | sec ; fmt 16: the k-mer itself
};

__device__ __forceinline__ void ld_sector_nc(const void* p, uint64_t (&v)[4])
{
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

__device__ __forceinline__ bool cas128(Slot16* addr, const Slot16& expect, const Slot16& desired, Slot16& old)
{
  const uint64_t e1 = ((uint64_t)expect.flags << 32) | expect.payload;
  const uint64_t d1 = ((uint64_t)desired.flags << 32) | desired.payload;
  uint64_t o0, o1;
  asm volatile("{\n\t.reg .b128 e, d, o;\n\t"
               "mov.b128 e, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
               "atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n\t"
               "mov.b128 {%0, %1}, o;\n\t}"
               : "=l"(o0), "=l"(o1)
               : "l"(expect.key), "l"(e1), "l"(desired.key), "l"(d1), "l"(addr)
               : "memory");
  old.key = o0;
  old.payload = (uint32_t)o1;
  old.flags = (uint32_t)(o1 >> 32);
  return o0 == expect.key && o1 == e1;
}

__device__ __forceinline__ Slot16 ld_slot16_volatile(const Slot16* p)
{
  Slot16 s;
  uint64_t a, b;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
  s.key = a;
  s.payload = (uint32_t)b;
  s.flags = (uint32_t)(b >> 32);
  return s;
}

template <int FMT>
__device__ __forceinline__ Home home_of(const KmerTable& t, uint64_t kmer)
{
  Home h;
  if (FMT == 8) {
    const uint64_t x = mix_kb(kmer, t.kbits);
    h.line = t.line_bits ? (x >> (t.kbits - t.line_bits)) : 0;
    h.sec = (uint32_t)(x >> t.rem_bits) & 3u;
    h.tag = ((x & low_mask64(t.rem_bits)) << 2) | h.sec;
  }
  else {
    const uint64_t x = mix64(kmer);
    h.line = t.line_bits ? (x >> (64 - t.line_bits)) : 0;
    h.sec = (uint32_t)(x >> 17) & 3u;
    h.tag = kmer;
  }
  return h;
}

// ---- stash (full keys) ----

__device__ __forceinline__ bool stash_find(const KmerTable& t, uint64_t kmer, uint32_t& payload, bool& multi)
{
  if (!t.stash_nonempty) return false;
  uint32_t p = (uint32_t)mix64(kmer ^ 0x5bd1e995u) & t.stash_mask;
  for (uint32_t i = 0; i <= t.stash_mask; ++i) {
    const Slot16 s = ld_slot16_volatile(t.stash + p);
    if (s.flags == NIL32) return false;
    if (s.key == kmer) { payload = s.payload; multi = s.flags == 1; return true; }
    p = (p + 1) & t.stash_mask;
  }
  return false;
}

// Find-or-insert into the stash.  On a fresh insert returns NIL32 in `prev`;
// when the key exists and `chain` is set, the payload is swapped for
// `payload` and the previous payload is returned in `prev`.  Returns false
// when the stash is full.
__device__ __forceinline__ bool stash_insert(const KmerTable& t, uint64_t kmer, uint32_t payload, uint32_t flags,
                                             bool chain, uint32_t& prev)
{
  uint32_t p = (uint32_t)mix64(kmer ^ 0x5bd1e995u) & t.stash_mask;
  const Slot16 empty{ ~0ull, NIL32, NIL32 };
  for (uint32_t i = 0; i <= t.stash_mask; ++i) {
    Slot16 cur = ld_slot16_volatile(t.stash + p);
    while (true) {
      if (cur.flags == NIL32) {
        Slot16 old;
        if (cas128(t.stash + p, empty, Slot16{ kmer, payload, flags }, old)) {
          atomicAdd(t.stash_used, 1u);
          prev = NIL32;
          return true;
        }
        cur = old;
        continue;
      }
      if (cur.key == kmer) {
        if (!chain) { prev = cur.payload; return true; }
        Slot16 old;
        if (cas128(t.stash + p, cur, Slot16{ kmer, payload, cur.flags }, old)) { prev = cur.payload; return true; }
        cur = old;
        continue;
      }
      break;
    }
    p = (p + 1) & t.stash_mask;
  }
  return false;
}

// ---- lookup (read-only tables) ----

template <int FMT>
__device__ __forceinline__ bool table_find_from(const KmerTable& t, const Home& h, uint64_t kmer,
                                                uint32_t& payload, bool& multi, uint32_t& sectors)
{
  const char* line = (const char*)t.slots + h.line * 128u;
#pragma unroll 1
  for (uint32_t i = 0; i < 4; ++i) {
    uint64_t v[4];
    ld_sector_nc(line + (((h.sec + i) & 3u) << 5), v);
    ++sectors;
    bool has_empty = false;
    if (FMT == 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (v[j] == EMPTY8) has_empty = true;
        else if ((v[j] >> 33) == h.tag) { payload = (uint32_t)v[j]; multi = (v[j] >> 32) & 1u; return true; }
      }
    }
    else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t fl = (uint32_t)(v[2 * j + 1] >> 32);
        if (fl == NIL32) has_empty = true;
        else if (v[2 * j] == kmer) { payload = (uint32_t)v[2 * j + 1]; multi = fl == 1; return true; }
      }
    }
    if (has_empty) return false;
  }
  return stash_find(t, kmer, payload, multi);
}

template <int FMT>
__device__ __forceinline__ bool table_find(const KmerTable& t, uint64_t kmer, uint32_t& payload, bool& multi)
{
  uint32_t sectors = 0;
  const Home h = home_of<FMT>(t, kmer);
  return table_find_from<FMT>(t, h, kmer, payload, multi, sectors);
}

// ---- insert (table being built) ----
//
// Find-or-insert `kmer`.  flags: 0 single / 1 multi.  With chain == true an
// existing key gets its payload replaced by `payload` and the old payload is
// returned in `prev` (linked list of read seeds); a fresh key returns NIL32.
// Returns false when line and stash are full (caller raises the overflow flag).
template <int FMT>
__device__ __forceinline__ bool table_insert(const KmerTable& t, uint64_t kmer, uint32_t payload, uint32_t flags,
                                             bool chain, uint32_t& prev)
{
  const Home h = home_of<FMT>(t, kmer);
  char* line = (char*)t.slots + h.line * 128u;
  if (FMT == 8) {
    const uint64_t fresh = (h.tag << 33) | ((uint64_t)(flags & 1u) << 32) | payload;
#pragma unroll 1
    for (uint32_t i = 0; i < 4; ++i) {
      unsigned long long* sec = (unsigned long long*)(line + (((h.sec + i) & 3u) << 5));
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        unsigned long long cur = *(volatile unsigned long long*)(sec + j);
        while (true) {
          if (cur == EMPTY8) {
            const unsigned long long old = atomicCAS(sec + j, EMPTY8, fresh);
            if (old == EMPTY8) { prev = NIL32; return true; }
            cur = old;
            continue;
          }
          if ((cur >> 33) == h.tag) {
            if (!chain) { prev = (uint32_t)cur; return true; }
            const unsigned long long upd = (cur & 0xffffffff00000000ull) | payload;
            const unsigned long long old = atomicCAS(sec + j, cur, upd);
            if (old == cur) { prev = (uint32_t)cur; return true; }
            cur = old;
            continue;
          }
          break;
        }
      }
    }
  }
  else {
    const Slot16 empty{ ~0ull, NIL32, NIL32 };
#pragma unroll 1
    for (uint32_t i = 0; i < 4; ++i) {
      Slot16* sec = (Slot16*)(line + (((h.sec + i) & 3u) << 5));
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        Slot16 cur = ld_slot16_volatile(sec + j);
        while (true) {
          if (cur.flags == NIL32) {
            Slot16 old;
            if (cas128(sec + j, empty, Slot16{ kmer, payload, flags & 1u }, old)) { prev = NIL32; return true; }
            cur = old;
            continue;
          }
          if (cur.key == kmer) {
            if (!chain) { prev = cur.payload; return true; }
            Slot16 old;
            if (cas128(sec + j, cur, Slot16{ kmer, payload, cur.flags }, old)) { prev = cur.payload; return true; }
            cur = old;
            continue;
          }
          break;
        }
      }
    }
  }
  return stash_insert(t, kmer, payload, flags, chain, prev);
}

// membership of (kmer, gpos) in the path index
__device__ __forceinline__ bool index_contains(const KmerTable& t, const uint32_t* __restrict__ multi,
                                               uint64_t kmer, uint32_t gpos)
{
  uint32_t payload;
  bool is_multi;
  const bool found = t.fmt == 8 ? table_find<8>(t, kmer, payload, is_multi) : table_find<16>(t, kmer, payload, is_multi);
  if (!found) return false;
  if (!is_multi) return payload == gpos;
  const uint32_t cnt = __ldg(multi + payload);
  // sorted list: binary search
  uint32_t lo = 0, hi = cnt;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t v = __ldg(multi + payload + 1 + mid);
    if (v == gpos) return true;
    if (v < gpos) lo = mid + 1; else hi = mid;
  }
  return false;
}

// ------------------------------------------------------- warp helpers --

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Reserve `n` consecutive output slots per lane with one atomic per warp.
// All 32 lanes must call.  Returns this lane's first slot.
__device__ __forceinline__ uint64_t warp_reserve(unsigned long long* counter, uint32_t n)
{
  const uint32_t lane = lane_id();
  uint32_t incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += o;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 31 && total) base = atomicAdd(counter, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + (incl - n);
}

}  // namespace dev
}  // namespace psi_b200
#endif
