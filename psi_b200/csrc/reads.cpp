// reads.cpp -- FASTQ/FASTA chunk reader (host).
#include "reads.hpp"

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <thread>

#ifndef PSI_B200_HOST_ONLY
#include <cuda_runtime_api.h>
#endif
#include <zlib.h>

namespace psi_b200 {

// ------------------------------------------------------------ HostBuffer --

HostBuffer::~HostBuffer()
{
  if (!data_) return;
#ifndef PSI_B200_HOST_ONLY
  if (pinned_) { cudaFreeHost(data_); return; }
#endif
  std::free(data_);
}

void HostBuffer::reserve(size_t bytes)
{
  if (bytes <= cap_) return;
  size_t ncap = cap_ ? cap_ : (size_t)1 << 20;
  while (ncap < bytes) ncap *= 2;
  char* nd = nullptr;
  bool pinned = false;
#ifndef PSI_B200_HOST_ONLY
  void* p = nullptr;
  if (cudaHostAlloc(&p, ncap, cudaHostAllocDefault) == cudaSuccess) {
    nd = (char*)p;
    pinned = true;
  }
  else (void)cudaGetLastError();  // no device: page-able memory is fine for host-only use
#endif
  if (!nd) {
    nd = (char*)std::malloc(ncap);
    if (!nd) throw std::bad_alloc();
  }
  if (size_) std::memcpy(nd, data_, size_);
  if (data_) {
#ifndef PSI_B200_HOST_ONLY
    if (pinned_) cudaFreeHost(data_); else
#endif
    std::free(data_);
  }
  data_ = nd;
  cap_ = ncap;
  pinned_ = pinned;
}

void HostBuffer::append(const void* src, size_t bytes)
{
  if (bytes == 0) return;      // (an empty line before anything is allocated: memcpy must not see a null pointer)
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
  buf_.resize(4u << 20);
}

ChunkReader::~ChunkReader()
{
  if (gz_) gzclose((gzFile)gz_);
}

// The next line without consuming it.  Lines are views into buf_: nothing is copied; when a line straddles the end of
// the buffered data the unread tail moves to the front and the buffer is refilled (and grown for very long lines).
bool ChunkReader::peek(LineView& lv)
{
  for (;;) {
    const char* b = buf_.data() + buf_pos_;
    const size_t avail = buf_len_ - buf_pos_;
    const char* nl = avail ? (const char*)std::memchr(b, '\n', avail) : nullptr;
    if (nl) { lv.p = b; lv.n = (size_t)(nl - b); line_adv_ = lv.n + 1; break; }
    if (eof_) {
      if (!avail) return false;
      lv.p = b; lv.n = avail; line_adv_ = avail;     // last line without a newline
      break;
    }
    if (buf_pos_) { std::memmove(buf_.data(), b, avail); buf_pos_ = 0; buf_len_ = avail; }
    if (buf_len_ == buf_.size()) buf_.resize(buf_.size() * 2);
    const int n = gzread((gzFile)gz_, buf_.data() + buf_len_, (unsigned)std::min<size_t>(buf_.size() - buf_len_, 1u << 30));
    if (n < 0) throw std::runtime_error("read error in sequence file");
    if (n == 0) eof_ = true;
    buf_len_ += (size_t)n;
  }
  if (lv.n && lv.p[lv.n - 1] == '\r') --lv.n;
  return true;
}

// ------------------------------------------------------------ packing --

namespace {

// Eight characters (one per byte of c) -> 16 bits of 2-bit codes; `bad` gets a non-zero byte for every character
// that is not A/C/G/T (either case).  Byte-parallel: x = (c >> 1) & 3; x ^= x >> 1 maps A, C, G, T to 0, 1, 2, 3;
// the expected upper-case character is rebuilt from the code (0x41 + 2 b0 + 6 b1 + 11 b0 b1) and compared.
inline uint64_t pack8(uint64_t c, uint64_t& bad)
{
  const uint64_t ones = 0x0101010101010101ull;
  uint64_t x = (c >> 1) & (3 * ones);
  x ^= (x >> 1) & ones;
  const uint64_t b0 = x & ones, b1 = (x >> 1) & ones, b01 = b0 & b1;
  const uint64_t expect = 0x41 * ones + (b0 << 1) + (b1 << 1) + (b1 << 2) + (b01 << 3) + (b01 << 1) + b01;
  bad = expect ^ (c & (0xdf * ones));
  uint64_t t = (x | (x >> 6)) & 0x000f000f000f000full;
  t = (t | (t >> 12)) & 0x000000ff000000ffull;
  return (t | (t >> 24)) & 0xffffull;
}

template <class Sink>
uint64_t pack_bases_impl(const char* bases, uint64_t n_bases, uint64_t* words, Sink&& on_exception, uint64_t n_words = ~0ull)
{
  if (n_words == ~0ull) n_words = n_bases / 32 + 2;
  uint64_t n_bad = 0, i = 0;
  for (uint64_t w = 0; w < n_words; ++w) {
    uint64_t word = 0;
    for (int g = 0; g < 4 && i < n_bases; ++g) {
      uint64_t c;
      const uint64_t left = n_bases - i;
      if (left >= 8) std::memcpy(&c, bases + i, 8);
      else {
        c = 0x4141414141414141ull;      // pad with 'A' (code 0, valid)
        std::memcpy(&c, bases + i, left);
      }
      uint64_t bad;
      uint64_t code = pack8(c, bad);
      if (bad) {
        for (int b = 0; b < 8; ++b)
          if ((bad >> (8 * b)) & 0xff) {
            on_exception(i + b);
            ++n_bad;
            code &= ~(3ull << (2 * b));   // stored as code 0
          }
      }
      word |= code << (16 * g);
      i += left >= 8 ? 8 : left;
    }
    words[w] = word;
  }
  return n_bad;
}

}  // namespace

uint64_t pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, uint64_t* exc, uint64_t exc_cap)
{
  uint64_t stored = 0;
  return pack_bases_impl(bases, n_bases, words, [&](uint64_t p) { if (exc && stored < exc_cap) exc[stored++] = p; });
}

uint64_t pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, std::vector<uint64_t>& exc)
{
  return pack_bases_impl(bases, n_bases, words, [&](uint64_t p) { exc.push_back(p); });
}

uint64_t pack_bases_parallel(const char* bases, uint64_t n_bases, uint64_t* words, std::vector<uint64_t>& exc, unsigned max_threads)
{
  unsigned hw = std::thread::hardware_concurrency();
  unsigned T = std::min(max_threads, hw ? hw : 1u);
  const uint64_t per = ((n_bases / std::max(T, 1u)) / 32) * 32;      // whole words per thread
  if (T < 2 || per < (1u << 20)) return pack_bases(bases, n_bases, words, exc);
  std::vector<std::vector<uint64_t>> local(T);
  std::vector<uint64_t> bad(T, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      const uint64_t lo = (uint64_t)t * per;
      const bool last = t + 1 == T;
      const uint64_t n = last ? n_bases - lo : per;
      bad[t] = pack_bases_impl(bases + lo, n, words + lo / 32, [&](uint64_t p) { local[t].push_back(lo + p); },
                               last ? n / 32 + 2 : per / 32);
    });
  for (auto& x : th) x.join();
  uint64_t n_bad = 0;
  for (unsigned t = 0; t < T; ++t) { n_bad += bad[t]; exc.insert(exc.end(), local[t].begin(), local[t].end()); }
  return n_bad;
}

uint64_t ChunkReader::next(uint64_t max_reads, bool packed)
{
  cur_ ^= 1;            // the previous chunk's buffers stay untouched: its upload may still be in flight
  Set& s = set();
  s.bases.clear();
  s.words.clear();
  s.exc.clear();
  s.read_ptr.assign(1, 0);
  s.names.clear();
  s.name_ptr.assign(1, 0);
  staging_.clear();
  first_id_ = consumed_;  // sequence.hpp:1616
  uint64_t n = 0, total = 0;
  bool uniform = true;
  uint64_t len0 = 0;
  LineView h, t;
  while (max_reads == 0 || n < max_reads) {
    bool have = false;
    while (peek(h)) {                       // blank lines between records are skipped
      if (h.n) { have = true; break; }
      consume();
    }
    if (!have) break;
    if (h.p[0] != '@' && h.p[0] != '>') throw std::runtime_error("malformed sequence record");
    const bool fastq = h.p[0] == '@';
    // name = header up to the first white space (kseq semantics)
    size_t e = 1;
    while (e < h.n && h.p[e] != ' ' && h.p[e] != '\t') ++e;
    s.names.append(h.p + 1, e - 1);
    s.name_ptr.push_back(s.names.size());
    consume();
    uint64_t seq_len = 0;
    while (peek(t)) {
      if (fastq && t.n && t.p[0] == '+') { consume(); break; }
      if (!fastq && t.n && (t.p[0] == '>' || t.p[0] == '@')) break;       // the next record's header stays unread
      if (packed) staging_.insert(staging_.end(), t.p, t.p + t.n);
      else s.bases.append(t.p, t.n);
      seq_len += t.n;
      consume();
    }
    if (fastq) {
      uint64_t q = 0;
      while (q < seq_len && peek(t)) { q += t.n; consume(); }
    }
    total += seq_len;
    s.read_ptr.push_back(total);
    if (n == 0) len0 = seq_len;
    else if (seq_len != len0) uniform = false;
    ++n;
  }
  uniform_len_ = (n && uniform && len0 && len0 <= 0xffffffffull) ? (uint32_t)len0 : 0u;
  if (packed) {
    const uint64_t n_words = total / 32 + 2;
    s.words.reserve(n_words * sizeof(uint64_t));
    s.words.resize(n_words * sizeof(uint64_t));
    pack_bases_parallel(staging_.data(), total, reinterpret_cast<uint64_t*>(s.words.data()), s.exc);
  }
  consumed_ += n;
  return n;
}

}  // namespace psi_b200
