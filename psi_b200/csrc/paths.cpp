// paths.cpp -- genome-wide path selection (host, index-build time).
#include "paths.hpp"

#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace psi_b200 {

namespace {
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next()
  {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
};
}  // namespace

void pick_paths(const FlatGraph& g, unsigned n, bool patched, unsigned context,
                uint64_t seed, PathSet& out)
{
  // `patched`/`context` only shrink the indexed TEXT in the reference
  // (pathindex.hpp:496-560); the device index stores distinct (k-mer, locus)
  // pairs, for which whole walks are a superset of their patches, so whole
  // walks are always handed over (head_off = tail_trim = 0).
  (void)patched;
  (void)context;
  if (g.paths.empty())
    throw std::runtime_error("No embedded path found in the graph");  // seed_finder.hpp:1145-1147

  out = PathSet();
  const uint64_t nn = g.node_count();
  std::vector<uint32_t> coverage(nn, 0);
  std::vector<uint32_t> stamp(nn, 0);
  uint32_t walk_no = 0;
  SplitMix64 rng(seed);

  for (auto const& region : g.paths) {
    if (region.nodes.empty()) continue;
    const uint32_t start = region.nodes.front();
    for (unsigned i = 0; i < n; ++i) {
      ++walk_no;
      uint32_t cur = start;
      while (true) {
        out.nodes.push_back(cur);
        ++coverage[cur];
        stamp[cur] = walk_no;
        uint64_t eb = g.row_ptr[cur], ee = g.row_ptr[cur + 1];
        if (eb == ee) break;
        // least covered successor not yet on this walk
        uint32_t best_cov = UINT32_MAX, n_best = 0, pick = UINT32_MAX;
        for (uint64_t e = eb; e < ee; ++e) {
          uint32_t to = g.col[e];
          if (stamp[to] == walk_no) continue;  // cyclic graph: never revisit
          uint32_t c = coverage[to];
          if (c < best_cov) { best_cov = c; n_best = 1; pick = to; }
          else if (c == best_cov) {
            ++n_best;
            // reservoir choice among ties; the very first walk keeps the first
            if (i != 0 && rng.next() % n_best == 0) pick = to;
          }
        }
        if (pick == UINT32_MAX) break;
        cur = pick;
      }
      out.path_ptr.push_back(out.nodes.size());
      out.head_off.push_back(0);
      out.tail_trim.push_back(0);
    }
  }
}

// ------------------------------------------- the reference's `<prefix>_paths` --

namespace {

struct Stream {
  std::ifstream in;
  uint64_t file_bytes = 0;
  explicit Stream(const std::string& f) : in(f, std::ifstream::binary | std::ifstream::ate)
  {
    if (!in) throw std::runtime_error("could not open file '" + f + "'!");
    file_bytes = (uint64_t)in.tellg();
    in.seekg(0);
  }
  // nothing read from the file may size an allocation beyond what the file itself still holds
  uint64_t remaining() { const auto p = in.tellg(); return p < 0 ? 0 : file_bytes - (uint64_t)p; }
  uint64_t u64()
  {
    uint64_t v = 0;
    in.read(reinterpret_cast<char*>(&v), sizeof v);
    if (!in) throw std::runtime_error("reference paths file: unexpected end of file");
    return v;
  }
  // sdsl::int_vector<>::load (int_vector.hpp: header = width << 56 | size in bits, then ceil(bits / 64) words)
  void int_vector(std::vector<uint64_t>& words, uint64_t& bits, unsigned& width)
  {
    const uint64_t h = u64();
    bits = h & ((1ull << 56) - 1);
    width = (unsigned)(h >> 56);
    if (bits > (1ull << 40) || ((bits + 63) >> 6) * 8 > remaining())
      throw std::runtime_error("reference paths file: a vector is longer than the file");
    words.resize((bits + 63) >> 6);
    if (!words.empty()) in.read(reinterpret_cast<char*>(words.data()), words.size() * 8);
    if (!in) throw std::runtime_error("reference paths file: unexpected end of file");
  }
};

// bits [pos, pos + len) of an LSB-first bit stream, len <= 64 (sdsl bits::read_int)
uint64_t read_bits(const std::vector<uint64_t>& w, uint64_t pos, unsigned len)
{
  if (len == 0) return 0;
  const uint64_t i = pos >> 6;
  const unsigned sh = (unsigned)(pos & 63);
  uint64_t v = i < w.size() ? w[i] >> sh : 0;
  if (sh && sh + len > 64 && i + 1 < w.size()) v |= w[i + 1] << (64 - sh);
  return len == 64 ? v : v & ((1ull << len) - 1);
}

// one Elias-delta code word at bit position pos (sdsl coder::elias_delta<>::encode, coder_elias_delta.hpp:221-236):
// n zeros and a one, the low n bits of len (n = floor(log2 len)), the low len - 1 bits of x; len = 65 stands for x = 0
uint64_t elias_delta_next(const std::vector<uint64_t>& z, uint64_t z_bits, uint64_t& pos)
{
  unsigned n = 0;
  while (true) {
    if (pos >= z_bits) throw std::runtime_error("reference paths file: truncated delta code");
    if (read_bits(z, pos++, 1)) break;
    if (++n > 7) throw std::runtime_error("reference paths file: malformed delta code");
  }
  uint64_t len = 1;
  if (n) { len = (1ull << n) | read_bits(z, pos, n); pos += n; }
  if (len > 65) throw std::runtime_error("reference paths file: malformed delta code");
  if (len == 65) { pos += 64; return 0; }
  uint64_t x = 1ull << (len - 1);
  if (len > 1) { x |= read_bits(z, pos, (unsigned)(len - 1)); pos += len - 1; }
  return x;
}

// sdsl::enc_vector<coder::elias_delta<>, 128>::load (enc_vector.hpp:236-318,397-403): size, the delta-coded differences,
// and per 128 elements an absolute sample followed by its bit pointer into the deltas
void enc_vector(Stream& s, std::vector<uint64_t>& out)
{
  const uint64_t size = s.u64();
  std::vector<uint64_t> z, sp;
  uint64_t z_bits = 0, sp_bits = 0;
  unsigned zw = 0, spw = 0;
  s.int_vector(z, z_bits, zw);
  s.int_vector(sp, sp_bits, spw);
  // every element costs at least one bit of the delta stream (or a sample)
  if (size > (1ull << 36) || size > z_bits + 64 * (sp_bits / 64 + 1) + 128)
    throw std::runtime_error("reference paths file: implausible path length");
  out.clear();
  out.reserve(size);
  if (size == 0) return;
  if (spw == 0 || spw > 64) throw std::runtime_error("reference paths file: malformed sample vector");
  const uint64_t n_sp = sp_bits / spw;
  auto sample = [&](uint64_t i) {
    if (i >= n_sp) throw std::runtime_error("reference paths file: sample index out of range");
    return read_bits(sp, i * spw, spw);
  };
  uint64_t v = 0, pos = 0;
  for (uint64_t i = 0; i < size; ++i) {
    if (i % 128 == 0) { v = sample(2 * (i / 128)); pos = sample(2 * (i / 128) + 1); }
    else v += elias_delta_next(z, z_bits, pos);      // unsigned wrap-around carries the negative differences
    out.push_back(v);
  }
}

}  // namespace

void load_reference_paths(const FlatGraph& g, const std::string& file, PathSet& out, uint64_t& context)
{
  Stream s(file);
  context = s.u64();
  const uint64_t direction = s.u64();      // 1 = forward text, 0 = reversed (the seed finder's index); either way the paths are the same
  if (direction > 1) throw std::runtime_error("reference paths file: malformed header");
  const uint64_t n_paths = s.u64();
  if (n_paths > (1ull << 32)) throw std::runtime_error("reference paths file: implausible number of paths");
  std::unordered_map<uint64_t, uint32_t> rank_of;
  rank_of.reserve(g.node_count() * 2);
  for (uint64_t r = 0; r < g.node_count(); ++r) rank_of.emplace(g.coord_id[r], (uint32_t)r);
  out = PathSet();
  std::vector<uint64_t> ids, bv;
  for (uint64_t p = 0; p < n_paths; ++p) {
    enc_vector(s, ids);
    const uint64_t left = s.u64(), right = s.u64();
    uint64_t bv_bits = 0;
    unsigned bv_width = 0;
    s.int_vector(bv, bv_bits, bv_width);             // node breaks: implied by the labels
    if (ids.empty()) throw std::runtime_error("reference paths file: empty path");
    for (uint64_t id : ids) {
      auto it = rank_of.find(id);
      if (it == rank_of.end()) throw std::runtime_error("reference paths file: node id " + std::to_string(id) + " is not in the graph");
      out.nodes.push_back(it->second);
    }
    const uint64_t first_len = g.node_length(out.nodes[out.path_ptr.back()]), last_len = g.node_length(out.nodes.back());
    if (left > first_len || right > last_len) throw std::runtime_error("reference paths file: path ends beyond their nodes");
    // left = bases of the first node that belong to the path (its suffix), right = bases of the last node (its prefix)
    // (Path::get_head_offset / get_seqlen_tail, path_base.hpp:241-296)
    out.head_off.push_back(left ? (uint32_t)(first_len - left) : 0u);
    out.tail_trim.push_back(right ? (uint32_t)(last_len - right) : 0u);
    out.path_ptr.push_back(out.nodes.size());
  }
}

}  // namespace psi_b200
