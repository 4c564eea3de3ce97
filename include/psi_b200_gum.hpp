/*
 * psi_b200_gum.hpp -- the adapter a maintainer of the reference adds between gum::SeqGraph / psi::Path and the C-ABI
 * of libpsi_b200 (include/psi_b200.h).  Header-only templates: they compile against the reference's own graph and
 * path types (gum/seqgraph.hpp, psi/path.hpp) and produce exactly the flat arrays psi_b200_set_graph and
 * psi_b200_set_paths take.  tests/test_gum_adapter.py compiles this header against the REAL gum headers and checks
 * the arrays against libpsi_b200's own GFA loader (ranks, ids, out-edge order, labels).
 *
 * Reference interfaces used:
 *   for_each_node(rank, id)          gum/seqgraph_interface.hpp / digraph_succinct.hpp (ranks ascend from 1)
 *   node_sequence(id)                gum/seqgraph_succinct.hpp:194-199
 *   for_each_edges_out(id, f)        gum/digraph_succinct.hpp:595-610  (every link, order of the node's edge list)
 *   id_to_rank(id)                   gum/digraph_succinct.hpp:937-964
 *   Path::get_nodes / get_head_offset / get_sequence_len / get_seqlen_tail    psi/path_base.hpp:232-296
 */
#ifndef PSI_B200_GUM_HPP
#define PSI_B200_GUM_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "psi_b200.h"

namespace psi_b200 {

struct FlatArrays {
  std::vector<uint64_t> seq_start;   // n + 1
  std::string seq;
  std::vector<uint64_t> row_ptr;     // n + 1
  std::vector<uint32_t> col;         // successor ranks (0-based)
  std::vector<uint64_t> node_id;     // the id seed records carry: gum's internal id (src/psikt.cpp:172-181)
  std::vector<uint64_t> coord_id;    // the id of the input file (graph.coordinate_id)
};

/* gum::SeqGraph<Succinct> -> CSR + concatenated labels in rank order. */
template <class TGraph>
inline FlatArrays flatten(const TGraph& g)
{
  FlatArrays a;
  g.for_each_node([&](auto /*rank*/, auto id) {
    a.seq_start.push_back(a.seq.size());
    a.seq += g.node_sequence(id);
    a.row_ptr.push_back(a.col.size());
    g.for_each_edges_out(id, [&](auto to, auto /*linktype*/) {
      a.col.push_back((uint32_t)(g.id_to_rank(to) - 1));
      return true;
    });
    a.node_id.push_back((uint64_t)id);
    a.coord_id.push_back((uint64_t)g.coordinate_id(id));
    return true;
  });
  a.seq_start.push_back(a.seq.size());
  a.row_ptr.push_back(a.col.size());
  return a;
}

/* SeedFinder(graph, k): the borrowed graph goes to the device once. */
template <class TGraph>
inline int set_graph(psi_b200_ctx* ctx, const TGraph& g)
{
  const FlatArrays a = flatten(g);
  return psi_b200_set_graph(ctx, a.node_id.size(), a.seq_start.data(), a.seq.data(), a.row_ptr.data(), a.col.data(), a.node_id.data());
}

struct FlatPaths {
  std::vector<uint64_t> path_ptr{ 0 };
  std::vector<uint32_t> nodes;       // node ranks (0-based)
  std::vector<uint32_t> head_off;    // bases cut from the first node
  std::vector<uint32_t> tail_trim;   // bases cut from the last node
};

/* The picked paths of a PathIndex (pindex.get_paths_set()) -> what psi_b200_set_paths takes. */
template <class TGraph, class TPathSet>
inline FlatPaths flatten_paths(const TGraph& g, const TPathSet& paths)
{
  FlatPaths f;
  for (auto it = paths.begin(); it != paths.end(); ++it) {
    auto const& p = *it;
    uint64_t n = 0, first = 0, last = 0;
    for (auto id : p.get_nodes()) {
      if (n++ == 0) first = (uint64_t)id;
      last = (uint64_t)id;
      f.nodes.push_back((uint32_t)(g.id_to_rank(id) - 1));
    }
    const uint64_t head = n ? (uint64_t)p.get_head_offset() : 0;
    uint64_t tail = 0;
    if (n == 1) tail = (uint64_t)g.node_length(first) - head - (uint64_t)p.get_sequence_len();
    else if (n > 1) tail = (uint64_t)g.node_length(last) - (uint64_t)p.get_seqlen_tail();
    f.head_off.push_back((uint32_t)head);
    f.tail_trim.push_back((uint32_t)tail);
    f.path_ptr.push_back(f.nodes.size());
  }
  return f;
}

template <class TGraph, class TPathSet>
inline int set_paths(psi_b200_ctx* ctx, const TGraph& g, const TPathSet& paths)
{
  const FlatPaths f = flatten_paths(g, paths);
  return psi_b200_set_paths(ctx, f.head_off.size(), f.path_ptr.data(), f.nodes.data(), f.head_off.data(), f.tail_trim.data());
}

}  // namespace psi_b200
#endif
