// common.cuh -- shared device code of libpsi_b200 (sm_100a only).
//
//   * k-mer packing: 2 bits per base, base i of the k-mer at bits [2i, 2i+2)
//   * KmerTable: the bucketised hash used for BOTH device indexes
//       - the path index       (k-mer -> locus / locus list; probed by seeds_on_paths)
//       - the chunk read index (k-mer -> chain of read seeds; probed by seeds_off_paths)
//     A bucket is one 128-byte line = one DRAM access (see the KmerTable section).
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
// invertible modulo 2^kb); the TOP bits of the result select the bucket line.
// Two multiply rounds: every input bit reaches the top bits, and the probe
// kernel stays far from being issue bound (a third round bought nothing
// measurable in bucket balance).
__host__ __device__ __forceinline__ uint64_t mix_kb(uint64_t x, uint32_t kb)
{
  const uint64_t m = low_mask64(kb);
  const uint32_t s = (kb >> 1) ? (kb >> 1) : 1;
  x ^= x >> s;
  x = (x * 0x9e3779b97f4a7c15ULL) & m;
  x ^= x >> s;
  x = (x * 0xd6e8feb86659fd93ULL) & m;
  x ^= x >> s;
  return x;
}

// ----------------------------------------------------------- KmerTable --
//
// Measured on B200 (profiles/r01c_gather_peak.md): a random access costs one
// 128-byte DRAM line whatever part of it is used, ~48 G lines/s.  So a bucket IS
// a 128-byte line: 16 slots of 8 bytes (fmt 8) or 8 slots of 16 bytes (fmt 16).
// A key lives in its home line or, when that line was full at insertion, in one
// of the next MAX_DISP lines (linear probing over lines), else in a small stash.
// Lookups read the home line -- four co-operating lanes load one 32-byte sector
// each in ONE warp instruction (seeds_on_paths), or one thread loads the four
// sectors back to back (walkers) -- and only continue past a line without a free
// slot.
//
// fmt 8 slot (64 bits), valid while R = kbits - line_bits <= 27; P = 27 - R spare bits widen the payload:
//   [63:64-R]     remainder of the bijectively mixed k-mer (the line index holds the rest)
//   next 3 bits   displacement from the home line (0..7)
//   next 2 bits   flags: bit 1 = entry reached only via an off-path walk, bit 0 = payload is a locus list
//   [31+P:0]      payload (32 + P bits): a LOCUS CODE (below), or the offset of the locus list in the multi
//                 array, or (read index) the head of the seed chain
//   An empty slot is all ones; no valid entry has both flag bits set.
// fmt 16 slot: full k-mer (64 bits), payload low word, then flags in bits [1:0] and 30 more payload bits in
//   [31:2] of the last word (NIL32 = empty).
//
// Locus code (path index, single-locus entries): everything a seed record needs from the graph side, so that a hit
// costs NO further memory access: (node id << off_bits) | offset in the node when the ids fit the payload
// (CODE_BY_ID), else (node rank << off_bits) | offset plus one gather of node_id[rank] (CODE_BY_RANK).
// off_bits = bits of the longest node label - 1.  Locus LISTS (multi array) keep 32-bit global positions.

constexpr uint32_t MAX_DISP = 7;
// walk mode: the per-chunk bitmap of seed prefixes has 2 * min(k, FILTER_PFX) bits of index -- 8 MB at most, resident
// in L2 while the graph is walked.  (A second filter on the whole k-mer at depth k was measured too: after the prefix
// pruning too few walks complete for it to pay for the atomics that build it.)
constexpr uint32_t FILTER_PFX = 13;
constexpr uint32_t FLAG_MULTI = 1u;
constexpr uint32_t FLAG_OFF = 2u;

struct alignas(16) Slot16 {
  uint64_t key;
  uint32_t payload;
  uint32_t flags;  // FLAG_* bits; NIL32 = empty
};

struct KmerTable {
  void*    slots;       // n_lines * 128 bytes
  Slot16*  stash;       // (stash_mask + 1) slots, open addressing on full keys
  uint32_t* stash_used; // device counter of occupied stash slots
  uint32_t line_bits;   // log2(n_lines)
  uint32_t kbits;       // 2k
  uint32_t rem_bits;    // fmt 8 only: kbits - line_bits (<= 27)
  uint32_t pay_hi;      // fmt 8 only: 27 - rem_bits payload bits above bit 31
  uint32_t fmt;         // 8 or 16 (bytes per slot)
  uint32_t stash_mask;
  uint32_t stash_nonempty;  // host-known: 0 lets lookups skip the stash entirely
};

// payload bits a slot of this table can hold
__host__ __device__ __forceinline__ uint32_t table_payload_bits(const KmerTable& t)
{
  return t.fmt == 8 ? 32u + t.pay_hi : 62u;
}

struct Home {
  uint64_t line;
  uint64_t tag;   // fmt 8: remainder << 3 (displacement 0); fmt 16: the k-mer itself
};

struct Found {
  uint64_t payload;
  uint32_t flags;
};

// ---- fmt 8 slot words ----
__device__ __forceinline__ uint64_t slot8_make(const KmerTable& t, uint64_t want, uint32_t flags, uint64_t payload)
{
  return (want << (34u + t.pay_hi)) | ((uint64_t)(flags & 3u) << (32u + t.pay_hi)) | payload;
}
__device__ __forceinline__ uint64_t slot8_want(const KmerTable& t, uint64_t slot) { return slot >> (34u + t.pay_hi); }
__device__ __forceinline__ uint32_t slot8_flags(const KmerTable& t, uint64_t slot) { return (uint32_t)(slot >> (32u + t.pay_hi)) & 3u; }
__device__ __forceinline__ uint64_t slot8_payload(const KmerTable& t, uint64_t slot) { return slot & low_mask64(32u + t.pay_hi); }
// ---- fmt 16 slot words ----
__device__ __forceinline__ uint32_t slot16_flagword(uint32_t flags, uint64_t payload) { return (flags & 3u) | ((uint32_t)(payload >> 32) << 2); }
__device__ __forceinline__ uint64_t slot16_payload(uint32_t payload_lo, uint32_t flagword) { return ((uint64_t)(flagword >> 2) << 32) | payload_lo; }

__device__ __forceinline__ void ld_sector_nc(const void* p, uint64_t (&v)[4])
{
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

// coherent variant for tables that are being written by the same kernel
__device__ __forceinline__ void ld_sector_cg(const void* p, uint64_t (&v)[4])
{
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
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
    h.line = t.line_bits ? (x >> t.rem_bits) : 0;        // top line_bits of the kbits-bit value
    h.tag = (x & low_mask64(t.rem_bits)) << 3;
  }
  else {
    const uint64_t x = mix64(kmer);
    h.line = t.line_bits ? (x >> (64 - t.line_bits)) : 0;
    h.tag = kmer;
  }
  return h;
}

// Compare the slots of one 32-byte sector.  `want` = tag | displacement (fmt 8) or the k-mer (fmt 16).
// Returns true on a match; `has_empty` reports a free slot in the sector.
template <int FMT>
__device__ __forceinline__ bool match_sector(const KmerTable& t, const uint64_t (&v)[4], uint64_t want, Found& f, bool& has_empty)
{
  bool hit = false;
  if (FMT == 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (v[j] == EMPTY8) has_empty = true;
      else if (slot8_want(t, v[j]) == want) { f.payload = slot8_payload(t, v[j]); f.flags = slot8_flags(t, v[j]); hit = true; }
    }
  }
  else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t fl = (uint32_t)(v[2 * j + 1] >> 32);
      if (fl == NIL32) has_empty = true;
      else if (v[2 * j] == want) { f.payload = slot16_payload((uint32_t)v[2 * j + 1], fl); f.flags = fl & 3u; hit = true; }
    }
  }
  return hit;
}

// ---- stash (full keys) ----

__device__ __forceinline__ bool stash_find(const KmerTable& t, uint64_t kmer, Found& f)
{
  if (!t.stash_nonempty) return false;
  uint32_t p = (uint32_t)mix64(kmer ^ 0x5bd1e995u) & t.stash_mask;
  for (uint32_t i = 0; i <= t.stash_mask; ++i) {
    const Slot16 s = ld_slot16_volatile(t.stash + p);
    if (s.flags == NIL32) return false;
    if (s.key == kmer) { f.payload = slot16_payload(s.payload, s.flags); f.flags = s.flags & 3u; return true; }
    p = (p + 1) & t.stash_mask;
  }
  return false;
}

// Find-or-insert into the stash.  On a fresh insert returns NIL32 in `prev`;
// when the key exists and `chain` is set, the payload is swapped for
// `payload` and the previous payload is returned in `prev`.  Returns false
// when the stash is full.
__device__ __forceinline__ bool stash_insert(const KmerTable& t, uint64_t kmer, uint64_t payload64, uint32_t flags,
                                             bool chain, uint32_t& prev)
{
  uint32_t p = (uint32_t)mix64(kmer ^ 0x5bd1e995u) & t.stash_mask;
  const Slot16 empty{ ~0ull, NIL32, NIL32 };
  const uint32_t payload = (uint32_t)payload64;
  flags = slot16_flagword(flags, payload64);
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

// ---- lookup by one thread (read-only tables): lines home + d, d = from_disp.. ----

template <int FMT>
__device__ __forceinline__ bool table_find_from(const KmerTable& t, const Home& h, uint64_t kmer, uint32_t from_disp, Found& f)
{
  const uint64_t line_mask = (1ull << t.line_bits) - 1ull;
#pragma unroll 1
  for (uint32_t d = from_disp; d <= MAX_DISP; ++d) {
    const char* line = (const char*)t.slots + (((h.line + d) & line_mask) << 7);
    uint64_t v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld_sector_nc(line + (i << 5), v[i]);
    const uint64_t want = FMT == 8 ? ((h.tag | d) ) : kmer;
    bool has_empty = false, hit = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) hit |= match_sector<FMT>(t, v[i], want, f, has_empty);
    if (hit) return true;
    if (has_empty) return false;
  }
  return stash_find(t, kmer, f);
}

template <int FMT>
__device__ __forceinline__ bool table_find(const KmerTable& t, uint64_t kmer, Found& f)
{
  const Home h = home_of<FMT>(t, kmer);
  return table_find_from<FMT>(t, h, kmer, 0, f);
}

__device__ __forceinline__ bool table_find_any(const KmerTable& t, uint64_t kmer, Found& f)
{
  return t.fmt == 8 ? table_find<8>(t, kmer, f) : table_find<16>(t, kmer, f);
}

// ---- insert (table being built) ----
//
// Find-or-insert `kmer`.  With chain == true an existing key gets its payload
// replaced by `payload` and the old payload is returned in `prev` (linked list
// of read seeds); a fresh key returns NIL32.  Returns false when MAX_DISP + 1
// lines and the stash are full (caller raises the overflow flag).
// `payload` may use table_payload_bits(t) bits; chains (chain == true) keep 32-bit payloads (seed indices).
// where: if not null, receives the address of the slot's low payload word (for counters kept in the payload).
template <int FMT>
__device__ __forceinline__ bool table_insert(const KmerTable& t, uint64_t kmer, uint64_t payload, uint32_t flags,
                                             bool chain, uint32_t& prev, unsigned int** where = nullptr)
{
  const Home h = home_of<FMT>(t, kmer);
  const uint64_t line_mask = (1ull << t.line_bits) - 1ull;
#pragma unroll 1
  for (uint32_t d = 0; d <= MAX_DISP; ++d) {
    char* line = (char*)t.slots + (((h.line + d) & line_mask) << 7);
    if (FMT == 8) {
      const uint64_t want = h.tag | d;
      const uint64_t fresh = slot8_make(t, want, flags, payload);
      unsigned long long* slot = (unsigned long long*)line;
#pragma unroll 1
      for (int j = 0; j < 16; ++j) {
        unsigned long long cur = *(volatile unsigned long long*)(slot + j);
        while (true) {
          if (cur == EMPTY8) {
            const unsigned long long old = atomicCAS(slot + j, EMPTY8, fresh);
            if (old == EMPTY8) { prev = NIL32; if (where) *where = reinterpret_cast<unsigned int*>(slot + j); return true; }
            cur = old;
            continue;
          }
          if (slot8_want(t, cur) == want) {
            if (where) *where = reinterpret_cast<unsigned int*>(slot + j);
            if (!chain) { prev = (uint32_t)cur; return true; }
            const unsigned long long upd = (cur & 0xffffffff00000000ull) | (uint32_t)payload;
            const unsigned long long old = atomicCAS(slot + j, cur, upd);
            if (old == cur) { prev = (uint32_t)cur; return true; }
            cur = old;
            continue;
          }
          break;
        }
      }
    }
    else {
      const Slot16 empty{ ~0ull, NIL32, NIL32 };
      Slot16* slot = (Slot16*)line;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        Slot16 cur = ld_slot16_volatile(slot + j);
        while (true) {
          if (cur.flags == NIL32) {
            Slot16 old;
            if (cas128(slot + j, empty, Slot16{ kmer, (uint32_t)payload, slot16_flagword(flags, payload) }, old)) {
              prev = NIL32;
              if (where) *where = &slot[j].payload;
              return true;
            }
            cur = old;
            continue;
          }
          if (cur.key == kmer) {
            if (where) *where = &slot[j].payload;
            if (!chain) { prev = cur.payload; return true; }
            Slot16 old;
            if (cas128(slot + j, cur, Slot16{ kmer, (uint32_t)payload, cur.flags }, old)) { prev = cur.payload; return true; }
            cur = old;
            continue;
          }
          break;
        }
      }
    }
  }
  if (where) *where = nullptr;      // counters are not kept in the stash: the caller treats this as an overflow
  return where ? false : stash_insert(t, kmer, payload, flags, chain, prev);
}

// Counting table: find-or-insert `kmer` and add 1 to the 32-bit counter in its payload.  False when the key found no
// slot in MAX_DISP + 1 lines (the caller raises the overflow flag; the host retries with a larger table).
template <int FMT>
__device__ __forceinline__ bool table_count(const KmerTable& t, uint64_t kmer)
{
  uint32_t prev;
  unsigned int* where = nullptr;
  if (!table_insert<FMT>(t, kmer, 0, 0, false, prev, &where) || !where) return false;
  atomicAdd(where, 1u);
  return true;
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
