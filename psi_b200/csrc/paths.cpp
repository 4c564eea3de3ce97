// paths.cpp -- genome-wide path selection (host, index-build time).
#include "paths.hpp"

#include <stdexcept>

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

}  // namespace psi_b200
