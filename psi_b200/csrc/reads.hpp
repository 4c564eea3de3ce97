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
  void resize(size_t bytes) { reserve(bytes); size_ = bytes; }
  char* data() { return data_; }
  const char* data() const { return data_; }
  size_t size() const { return size_; }
 private:
  char* data_ = nullptr;
  size_t size_ = 0, cap_ = 0;
  bool pinned_ = false;
};

// ASCII bases -> 2-bit words (32 bases per 64-bit word, first base in the low bits; A, C, G, T of either case =
// 0..3, anything else 0 and its position appended to `exc`).  words must hold n_bases / 32 + 2 entries; the tail is
// zero filled.  Returns the number of characters outside A/C/G/T (at most exc_cap positions are stored when exc is a
// plain array; the vector form stores all).
uint64_t pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, uint64_t* exc, uint64_t exc_cap);
uint64_t pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, std::vector<uint64_t>& exc);
// the same over several host threads (large chunks; the words and the exception list are identical)
uint64_t pack_bases_parallel(const char* bases, uint64_t n_bases, uint64_t* words, std::vector<uint64_t>& exc, unsigned max_threads = 8);

class ChunkReader {
 public:
  explicit ChunkReader(const std::string& path);
  ~ChunkReader();
  // Returns the number of records loaded (0 at end of input).  packed: the bases are delivered as 2-bit words
  // (words(), exc()) instead of characters (bases()).  Two sets of buffers alternate between calls, so what a call
  // returned stays valid while the next chunk is being parsed (and uploaded asynchronously).
  uint64_t next(uint64_t max_reads, bool packed = false);
  uint64_t first_read_id() const { return first_id_; }
  uint64_t n_reads() const { return set().read_ptr.size() - 1; }
  uint64_t n_bases() const { return set().read_ptr.back(); }
  uint32_t uniform_len() const { return uniform_len_; }    // != 0: every read of the chunk has this length
  const uint64_t* read_ptr() const { return set().read_ptr.data(); }
  const char* bases() const { return set().bases.data(); }
  const uint64_t* words() const { return reinterpret_cast<const uint64_t*>(set().words.data()); }
  const uint64_t* exc() const { return set().exc.data(); }
  uint64_t n_exc() const { return set().exc.size(); }
  const uint64_t* name_ptr() const { return set().name_ptr.data(); }
  const char* names() const { return set().names.data(); }
 private:
  struct Set {
    HostBuffer bases;     // characters (unpacked chunks)
    HostBuffer words;     // 2-bit words (packed chunks)
    std::vector<uint64_t> exc;
    std::vector<uint64_t> read_ptr{ 0 };
    std::string names;
    std::vector<uint64_t> name_ptr{ 0 };
  };
  Set& set() { return sets_[cur_]; }
  const Set& set() const { return sets_[cur_]; }
  // the next line of the input as a view into the read buffer (no copy; valid until the next peek after a consume)
  struct LineView { const char* p; size_t n; };
  bool peek(LineView& lv);
  void consume() { buf_pos_ += line_adv_; line_adv_ = 0; }
  void* gz_ = nullptr;
  std::vector<char> buf_;
  size_t buf_pos_ = 0, buf_len_ = 0, line_adv_ = 0;
  bool eof_ = false;
  uint64_t consumed_ = 0, first_id_ = 0;
  uint32_t uniform_len_ = 0;
  Set sets_[2];
  int cur_ = 1;
  std::vector<char> staging_;   // characters of a chunk that is delivered packed
};

}  // namespace psi_b200
#endif
