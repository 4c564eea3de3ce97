// reads.cpp -- FASTQ/FASTA chunk reader (host).
#include "reads.hpp"

#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include <cuda_runtime_api.h>
#include <zlib.h>

namespace psi_b200 {

// ------------------------------------------------------------ HostBuffer --

HostBuffer::~HostBuffer()
{
  if (!data_) return;
  if (pinned_) cudaFreeHost(data_);
  else std::free(data_);
}

void HostBuffer::reserve(size_t bytes)
{
  if (bytes <= cap_) return;
  size_t ncap = cap_ ? cap_ : (size_t)1 << 20;
  while (ncap < bytes) ncap *= 2;
  char* nd = nullptr;
  bool pinned = false;
  void* p = nullptr;
  if (cudaHostAlloc(&p, ncap, cudaHostAllocDefault) == cudaSuccess) {
    nd = (char*)p;
    pinned = true;
  }
  else {
    (void)cudaGetLastError();  // no device: page-able memory is fine for host-only use
    nd = (char*)std::malloc(ncap);
    if (!nd) throw std::bad_alloc();
  }
  if (size_) std::memcpy(nd, data_, size_);
  if (data_) { if (pinned_) cudaFreeHost(data_); else std::free(data_); }
  data_ = nd;
  cap_ = ncap;
  pinned_ = pinned;
}

void HostBuffer::append(const void* src, size_t bytes)
{
  if (size_ + bytes > cap_) reserve(size_ + bytes);
  std::memcpy(data_ + size_, src, bytes);
  size_ += bytes;
}

// ----------------------------------------------------------- ChunkReader --

ChunkReader::ChunkReader(const std::string& path)
{
  gz_ = gzopen(path.c_str(), "rb");
  if (!gz_) throw std::runtime_error("could not open file '" + path + "'!");
  gzbuffer((gzFile)gz_, 1 << 20);
  buf_.resize(1 << 20);
}

ChunkReader::~ChunkReader()
{
  if (gz_) gzclose((gzFile)gz_);
}

bool ChunkReader::fill()
{
  if (eof_) return false;
  int n = gzread((gzFile)gz_, buf_.data(), (unsigned)buf_.size());
  if (n < 0) throw std::runtime_error("read error in sequence file");
  buf_pos_ = 0;
  buf_len_ = (size_t)n;
  if (n == 0) { eof_ = true; return false; }
  return true;
}

bool ChunkReader::getline(std::string& out)
{
  out.clear();
  bool got = false;
  while (true) {
    if (buf_pos_ == buf_len_ && !fill()) break;
    got = true;
    const char* b = buf_.data() + buf_pos_;
    const char* nl = (const char*)std::memchr(b, '\n', buf_len_ - buf_pos_);
    if (nl) {
      out.append(b, nl - b);
      buf_pos_ += (size_t)(nl - b) + 1;
      break;
    }
    out.append(b, buf_len_ - buf_pos_);
    buf_pos_ = buf_len_;
  }
  if (!out.empty() && out.back() == '\r') out.pop_back();
  return got;
}

uint64_t ChunkReader::next(uint64_t max_reads)
{
  bases_.clear();
  read_ptr_.assign(1, 0);
  names_.clear();
  name_ptr_.assign(1, 0);
  first_id_ = consumed_;  // sequence.hpp:1616
  std::string line, seq, tmp;
  uint64_t n = 0;
  while (max_reads == 0 || n < max_reads) {
    if (has_pending_) { line.swap(pending_); has_pending_ = false; }
    else {
      bool ok;
      do { ok = getline(line); } while (ok && line.empty());
      if (!ok) break;
    }
    if (line[0] != '@' && line[0] != '>') throw std::runtime_error("malformed sequence record");
    const bool fastq = line[0] == '@';
    // name = header up to the first white space (kseq semantics)
    size_t e = 1;
    while (e < line.size() && line[e] != ' ' && line[e] != '\t') ++e;
    names_.append(line, 1, e - 1);
    name_ptr_.push_back(names_.size());
    seq.clear();
    while (getline(tmp)) {
      if (fastq && !tmp.empty() && tmp[0] == '+') break;
      if (!fastq && !tmp.empty() && (tmp[0] == '>' || tmp[0] == '@')) { pending_.swap(tmp); has_pending_ = true; break; }
      seq += tmp;
    }
    if (fastq) {
      size_t q = 0;
      while (q < seq.size() && getline(tmp)) q += tmp.size();
    }
    bases_.append(seq.data(), seq.size());
    read_ptr_.push_back(bases_.size());
    ++n;
  }
  consumed_ += n;
  return n;
}

}  // namespace psi_b200
