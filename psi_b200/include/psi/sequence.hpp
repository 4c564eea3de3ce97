// psi/sequence.hpp -- read records, chunked loading and seeding parameters.
//
// Stands where the reference uses Records<Dna5QStringSet<>>, readRecords() and
// seeding() (include/psi/sequence.hpp:1130-1294,1608-1624,1688-1745) with
// klibpp::SeqStreamIn as the FASTQ source.  A Records object here is either
//   - a CHUNK: n reads as offsets + concatenated bases in page-locked host
//     memory owned by the stream (readRecords), global ids = rec_offset + i; or
//   - a SEEDS view of a chunk (SeedFinder::get_seeds): the same reads plus
//     (seed length, distance); the k-mers themselves are packed on the GPU, so
//     no seed strings are materialised on the host.  Seed i of read r sits at
//     offset i * distance, for every offset with offset + k <= length
//     (sequence.hpp:1712); position_to_id/offset reproduce the SeedMap arithmetic
//     (sequence.hpp:1148-1220).
#ifndef PSI_B200_PSI_SEQUENCE_HPP
#define PSI_B200_PSI_SEQUENCE_HPP

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
  // chunk part (borrowed from the stream until the next readRecords on it)
  uint64_t n_reads = 0;
  uint64_t rec_offset = 0;          // sequence.hpp:1616
  const uint64_t* read_ptr = nullptr;
  const char* bases = nullptr;
  const uint64_t* name_ptr = nullptr;
  const char* names = nullptr;
  // seeds part
  unsigned seed_len = 0;            // 0 for a plain chunk
  unsigned distance = 0;
  uint64_t serial = 0;              // identifies the submission that packed these seeds

  uint64_t size() const { return n_reads; }
  uint64_t total_length() const { return n_reads ? read_ptr[n_reads] - read_ptr[0] : 0; }
  uint64_t read_length(uint64_t i) const { return read_ptr[i + 1] - read_ptr[i]; }
  std::string read(uint64_t i) const { return std::string(bases + read_ptr[i], bases + read_ptr[i + 1]); }
  std::string name(uint64_t i) const { return names ? std::string(names + name_ptr[i], names + name_ptr[i + 1]) : std::string(); }
  // global read id of local read i (Records::position_to_id, sequence.hpp:1201-1213)
  TPosition position_to_id(uint64_t i) const { return rec_offset + i; }
  void clear() { *this = Records(); }
};

template <typename TStringSet>
inline uint64_t length(const Records<TStringSet>& r) { return r.size(); }
template <typename TStringSet>
inline uint64_t lengthSum(const Records<TStringSet>& r) { return r.total_length(); }

// Loads up to n records (0 = all) into `records`; false at end of input
// (readRecords, sequence.hpp:1608-1624).  The buffers live in the stream's
// page-locked chunk buffer.
template <typename TStringSet>
inline bool readRecords(Records<TStringSet>& records, klibpp::SeqStreamIn& iss, uint64_t n)
{
  records.clear();
  if (!iss) return false;
  psi_b200_chunk_view v;
  if (psi_b200_reader_next(iss.handle(), n, &v) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
  if (v.n_reads == 0) return false;
  records.n_reads = v.n_reads;
  records.rec_offset = v.first_read_id;
  records.read_ptr = v.read_ptr;
  records.bases = v.bases;
  records.name_ptr = v.name_ptr;
  records.names = v.names;
  iss.advance(v.n_reads);
  return true;
}

}  // namespace psi
#endif
