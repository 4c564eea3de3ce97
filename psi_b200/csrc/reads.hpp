// reads.hpp -- FASTQ/FASTA chunk reader on the host.
//
// Stands behind readRecords(Records&, klibpp::SeqStreamIn&, n) of the reference
// (include/psi/sequence.hpp:1608-1624): up to n records per chunk (0 = all),
// read ids are global ordinals (rec_offset = records consumed before the chunk).
#ifndef PSI_B200_READS_HPP
#define PSI_B200_READS_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace psi_b200 {

// Growable host buffer, page-locked when a CUDA device is present so that chunk
// uploads are asynchronous DMA; plain malloc otherwise (host-only tools/tests).
class HostBuffer {
 public:
  HostBuffer() = default;
  ~HostBuffer();
  HostBuffer(const HostBuffer&) = delete;
  HostBuffer& operator=(const HostBuffer&) = delete;
  void reserve(size_t bytes);
  void append(const void* src, size_t bytes);
  void clear() { size_ = 0; }
  char* data() { return data_; }
  const char* data() const { return data_; }
  size_t size() const { return size_; }
 private:
  char* data_ = nullptr;
  size_t size_ = 0, cap_ = 0;
  bool pinned_ = false;
};

class ChunkReader {
 public:
  explicit ChunkReader(const std::string& path);
  ~ChunkReader();
  // Returns the number of records loaded (0 at end of input).
  uint64_t next(uint64_t max_reads);
  uint64_t first_read_id() const { return first_id_; }
  uint64_t n_reads() const { return read_ptr_.size() - 1; }
  const uint64_t* read_ptr() const { return read_ptr_.data(); }
  const char* bases() const { return bases_.data(); }
  const uint64_t* name_ptr() const { return name_ptr_.data(); }
  const char* names() const { return names_.data(); }
 private:
  bool fill();
  bool getline(std::string& out);
  void* gz_ = nullptr;
  std::vector<char> buf_;
  size_t buf_pos_ = 0, buf_len_ = 0;
  bool eof_ = false;
  std::string pending_;  // header line read ahead (FASTA multi-line)
  bool has_pending_ = false;
  uint64_t consumed_ = 0, first_id_ = 0;
  HostBuffer bases_;
  std::vector<uint64_t> read_ptr_{ 0 };
  std::string names_;
  std::vector<uint64_t> name_ptr_{ 0 };
};

}  // namespace psi_b200
#endif
