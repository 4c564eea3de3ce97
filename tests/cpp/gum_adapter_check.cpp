// Compiled against the REAL gum headers of the reference tree (tests/test_gum_adapter.py): loads a GFA with
// gum::util::load(graph, file, sort = true) -- what the reference CLI does (src/psikt.cpp:249-251) -- flattens the gum
// object with include/psi_b200_gum.hpp and compares every array with libpsi_b200's own loader of the same file.
// Usage: gum_adapter_check GRAPH.gfa
#include <cstdio>
#include <cstring>
#include <string>

#include <gum/graph.hpp>
#include <gum/io_utils.hpp>

#include "../../include/psi_b200_gum.hpp"

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  gum::SeqGraph<gum::Succinct> graph;
  gum::util::load(graph, std::string(argv[1]), true);
  const psi_b200::FlatArrays a = psi_b200::flatten(graph);

  psi_b200_graph* h = nullptr;
  if (psi_b200_graph_load_gfa(argv[1], 1, &h) != PSI_B200_OK) { std::fprintf(stderr, "%s\n", psi_b200_global_error()); return 3; }
  psi_b200_graph_view v;
  psi_b200_graph_get_view(h, &v);
  int bad = 0;
  auto expect = [&](bool ok, const char* what) { if (!ok) { std::fprintf(stderr, "MISMATCH: %s\n", what); ++bad; } };
  expect(v.n_nodes == a.node_id.size(), "node count");
  expect(v.n_edges == a.col.size(), "edge count");
  expect(v.n_bases == a.seq.size(), "base count");
  if (!bad) {
    expect(std::memcmp(v.seq_start, a.seq_start.data(), (v.n_nodes + 1) * 8) == 0, "seq_start (label lengths in rank order)");
    expect(std::memcmp(v.seq, a.seq.data(), v.n_bases) == 0, "labels");
    expect(std::memcmp(v.row_ptr, a.row_ptr.data(), (v.n_nodes + 1) * 8) == 0, "row_ptr (out-degrees in rank order)");
    expect(std::memcmp(v.col, a.col.data(), v.n_edges * 4) == 0, "col (out-edge order)");
    expect(std::memcmp(v.internal_id, a.node_id.data(), v.n_nodes * 8) == 0, "internal ids");
    expect(std::memcmp(v.coord_id, a.coord_id.data(), v.n_nodes * 8) == 0, "coordinate ids");
    // embedded paths: same node ranks in the same order
    uint64_t pi = 0;
    graph.for_each_path([&](auto /*rank*/, auto pid) {
      const uint32_t* nodes = nullptr;
      uint64_t n = 0;
      const char* name = nullptr;
      if (psi_b200_graph_path(h, pi, &name, &nodes, &n) != PSI_B200_OK) { ++bad; return false; }
      expect(graph.path_name(pid) == name, "path name");
      uint64_t i = 0;
      bool same = true;
      graph.path(pid).for_each_node([&](auto id, bool /*reversed*/) { same &= i < n && nodes[i] == graph.id_to_rank(id) - 1; ++i; return true; });
      expect(same && i == n, "path nodes");
      ++pi;
      return true;
    });
    expect(pi == v.n_paths, "path count");
  }
  std::printf("{\"nodes\": %llu, \"edges\": %llu, \"bases\": %llu, \"paths\": %llu, \"mismatches\": %d}\n", (unsigned long long)v.n_nodes,
              (unsigned long long)v.n_edges, (unsigned long long)v.n_bases, (unsigned long long)v.n_paths, bad);
  psi_b200_graph_free(h);
  return bad ? 1 : 0;
}
