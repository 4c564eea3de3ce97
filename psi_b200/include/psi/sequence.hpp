// psi/sequence.hpp -- read records, chunked loading and seeding parameters.
//
// Stands where the reference uses Records<Dna5QStringSet<>>, readRecords() and
// seeding() (include/psi/sequence.hpp:1130-1294,1608-1624,1688-1745) with
// klibpp::SeqStreamIn as the FASTQ source.  A Records object here is either
//   - a CHUNK: n reads as offsets + their bases as 2-bit words (+ the positions of
//     the characters outside A/C/G/T) in page-locked host memory owned by the
//     stream (readRecords), global ids = rec_offset + i; or
//   - a SEEDS view of a chunk (SeedFinder::get_seeds): the same reads plus
//     (seed length, distance); the k-mers themselves are packed on the GPU, so
//     no seed strings are materialised on the host.  Seed i of read r sits at
//     offset i * distance, for every offset with offset + k <= length
//     (sequence.hpp:1712); position_to_id/offset reproduce the SeedMap arithmetic
//     (sequence.hpp:1148-1220).
#ifndef PSI_B200_PSI_SEQUENCE_HPP
#define PSI_B200_PSI_SEQUENCE_HPP

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/psi_b200.h"

namespace psi {

template <typename TSpec = void>
struct Dna5QStringSet {};
struct InMemory {};
struct DiskBased {};

}  // namespace psi

namespace klibpp {

// FASTQ/FASTA(.gz) input stream (kseq++ SeqStreamIn as used at src/psikt.cpp:258-263).
class SeqStreamIn {
 public:
  explicit SeqStreamIn(const char* path)
  {
    if (psi_b200_reader_open(path, &h_) != PSI_B200_OK) h_ = nullptr;
  }
  SeqStreamIn(const SeqStreamIn&) = delete;
  SeqStreamIn& operator=(const SeqStreamIn&) = delete;
  ~SeqStreamIn() { if (h_) psi_b200_reader_close(h_); }
  explicit operator bool() const { return h_ != nullptr; }
  bool operator!() const { return h_ == nullptr; }
  uint64_t counts() const { return counts_; }
  psi_b200_reader* handle() { return h_; }
  void advance(uint64_t n) { counts_ += n; }

 private:
  psi_b200_reader* h_ = nullptr;
  uint64_t counts_ = 0;
};

}  // namespace klibpp

namespace psi {

template <typename TStringSet = Dna5QStringSet<>>
class Records {
 public:
  typedef uint64_t TPosition;
  typedef uint64_t TSize;
  // chunk part (borrowed from the stream: valid until the call after the next readRecords on it).  The bases are
  // 2-bit words (psi_b200_packed_chunk): 32 per 64-bit word, characters outside A/C/G/T listed in `exc`.
  uint64_t n_reads = 0;
  uint64_t rec_offset = 0;          // sequence.hpp:1616
  uint64_t n_bases = 0;
  uint32_t read_len = 0;            // != 0: every read of the chunk has this length
  const uint64_t* read_ptr = nullptr;
  const uint64_t* words = nullptr;
  const uint64_t* exc = nullptr;
  uint64_t n_exc = 0;
  const uint64_t* name_ptr = nullptr;
  const char* names = nullptr;
  // seeds part
  unsigned seed_len = 0;            // 0 for a plain chunk
  unsigned distance = 0;
  uint64_t serial = 0;              // identifies the submission that packed these seeds
  const void* pipe = nullptr;       // and the pipeline (host thread) it went to

  uint64_t size() const { return n_reads; }
  uint64_t total_length() const { return n_bases; }
  uint64_t read_length(uint64_t i) const { return read_ptr[i + 1] - read_ptr[i]; }
  // the read's characters, upper case; anything that was not A/C/G/T reads as 'N'
  std::string read(uint64_t i) const
  {
    std::string out;
    out.reserve(read_length(i));
    for (uint64_t p = read_ptr[i]; p < read_ptr[i + 1]; ++p) out.push_back("ACGT"[(words[p >> 5] >> (2 * (p & 31))) & 3]);
    const uint64_t* e = std::lower_bound(exc, exc + n_exc, read_ptr[i]);
    for (; e != exc + n_exc && *e < read_ptr[i + 1]; ++e) out[*e - read_ptr[i]] = 'N';
    return out;
  }
  std::string name(uint64_t i) const { return names ? std::string(names + name_ptr[i], names + name_ptr[i + 1]) : std::string(); }
  // global read id of local read i (Records::position_to_id, sequence.hpp:1201-1213)
  TPosition position_to_id(uint64_t i) const { return rec_offset + i; }
  // seeds of read i: offsets 0, d, 2d, ... while offset + k <= length (sequence.hpp:1712)
  uint64_t seeds_of_read(uint64_t i) const
  {
    const uint64_t len = read_length(i);
    return seed_len && len >= seed_len ? (len - seed_len) / distance + 1 : 0;
  }
  void clear() { *this = Records(); }
};

template <typename TStringSet>
inline uint64_t length(const Records<TStringSet>& r) { return r.size(); }
template <typename TStringSet>
inline uint64_t lengthSum(const Records<TStringSet>& r) { return r.total_length(); }

// Loads up to n records (0 = all) into `records`; false at end of input
// (readRecords, sequence.hpp:1608-1624).  The buffers live in the stream's
// page-locked chunk buffers (two sets alternate: a chunk stays valid while the next one is read).
template <typename TStringSet>
inline bool readRecords(Records<TStringSet>& records, klibpp::SeqStreamIn& iss, uint64_t n)
{
  records.clear();
  if (!iss) return false;
  psi_b200_packed_chunk v;
  if (psi_b200_reader_next_packed(iss.handle(), n, &v) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
  if (v.n_reads == 0) return false;
  records.n_reads = v.n_reads;
  records.rec_offset = v.first_read_id;
  records.n_bases = v.n_bases;
  records.read_len = v.read_len;
  records.read_ptr = v.read_ptr;
  records.words = v.words;
  records.exc = v.exc;
  records.n_exc = v.n_exc;
  records.name_ptr = v.name_ptr;
  records.names = v.names;
  iss.advance(v.n_reads);
  return true;
}

}  // namespace psi
#endif
