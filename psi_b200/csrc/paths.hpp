// paths.hpp -- genome-wide path selection on the host (index-build time).
#ifndef PSI_B200_PATHS_HPP
#define PSI_B200_PATHS_HPP

#include <cstdint>
#include <vector>

#include "flat_graph.hpp"

namespace psi_b200 {

struct PathSet {
  std::vector<uint64_t> path_ptr{ 0 };
  std::vector<uint32_t> nodes;
  std::vector<uint32_t> head_off;
  std::vector<uint32_t> tail_trim;
  uint64_t size() const { return head_off.size(); }
};

// Stands behind SeedFinder::pick_paths (reference seed_finder.hpp:1138-1167):
// for every embedded path of the graph, n haplotype-like walks from its first
// node to a sink, each step taking a least-covered successor
// (graph.hpp:216-287), ties broken by a seeded generator.  Throws
// std::runtime_error when the graph embeds no path (seed_finder.hpp:1145-1147).
void pick_paths(const FlatGraph& g, unsigned n, bool patched, unsigned context,
                uint64_t seed, PathSet& out);

}  // namespace psi_b200
#endif
