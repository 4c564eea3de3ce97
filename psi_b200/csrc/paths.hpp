// paths.hpp -- genome-wide path selection on the host (index-build time).
#ifndef PSI_B200_PATHS_HPP
#define PSI_B200_PATHS_HPP

#include <cstdint>
#include <string>
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

// Reads the path set out of the reference's own saved path index: the file `<prefix>_paths` written by
// PathIndex::save_paths_set (reference include/psi/pathindex.hpp:313-332) = context (u64), direction (u64), then
// PathSet::serialize (pathset.hpp:261-273): the number of paths and, per path, Path<graph, Compact>::serialize
// (path_base.hpp:552-560): the node ids in COORDINATE ids as an sdsl enc_vector<coder::elias_delta<>, 128>, `left` and
// `right` (u64: how much of the first / last node belongs to the path, 0 = all of it) and the node-break bit vector.
// What follows the paths in the file (the FM-index over the encoded node ids) is not needed and not read.
// Throws std::runtime_error on a malformed file or a node id the graph does not have.
void load_reference_paths(const FlatGraph& g, const std::string& file, PathSet& out, uint64_t& context);

}  // namespace psi_b200
#endif
