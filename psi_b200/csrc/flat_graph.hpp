// flat_graph.hpp -- host-side flattened sequence graph of libpsi_b200.
//
// Stands behind gum::SeqGraph<Succinct> as psikt loads it
// (reference src/psikt.cpp:249-251).  Only what the seed-finding path needs is
// kept: rank-ordered nodes, CSR out-adjacency in gum's out-edge order,
// concatenated labels, internal (Succinct) and external (coordinate) ids and
// the embedded paths.
#ifndef PSI_B200_FLAT_GRAPH_HPP
#define PSI_B200_FLAT_GRAPH_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace psi_b200 {

struct EmbeddedPath {
  std::string name;
  std::vector<uint32_t> nodes;  // ranks
};

struct FlatGraph {
  std::vector<uint64_t> seq_start;    // n+1
  std::string seq;                    // concatenated labels
  std::vector<uint64_t> row_ptr;      // n+1
  std::vector<uint32_t> col;          // successor ranks
  std::vector<uint32_t> indeg;        // n
  std::vector<uint64_t> internal_id;  // n
  std::vector<uint64_t> coord_id;     // n
  std::vector<EmbeddedPath> paths;

  uint64_t node_count() const { return coord_id.size(); }
  uint64_t edge_count() const { return col.size(); }
  uint64_t node_length(uint32_t r) const { return seq_start[r + 1] - seq_start[r]; }
};

// Input to the builder: nodes in arbitrary order with external ids, edges as
// (from index, to index) in FILE order, `from_start` marking links that leave
// the start side of their source (GFA '-' orientation).
struct RawGraph {
  std::vector<uint64_t> ids;
  std::vector<std::string> labels;
  struct Edge { uint32_t from, to; bool from_start; };
  std::vector<Edge> edges;
  struct RawPath { std::string name; std::vector<uint32_t> nodes; };
  std::vector<RawPath> paths;  // node = index into ids
};

// Orders the nodes like gum does at load time -- sort by id, then the
// DFS-based topological sort of gum/seqgraph_interface.hpp:252-360 -- and
// assigns gum's Succinct ids (digraph_succinct.hpp:937-964).
void build_flat_graph(RawGraph&& raw, bool sort, FlatGraph& out);

// GFA1 (S/L/P) and GFA2 (S/E/O) reader.  Throws std::runtime_error.
void load_gfa(const std::string& path, bool sort, FlatGraph& out);
void write_gfa1(const FlatGraph& g, const std::string& path);
// vg's protobuf graph files (vg_reader.cpp; no protobuf library involved).  Throws std::runtime_error.
void load_vg(const std::string& path, bool sort, FlatGraph& out);
// by file name like gum::util::load (gum/io_utils.hpp:66-80): *.vg -> load_vg, anything else -> load_gfa
void load_graph_file(const std::string& path, bool sort, FlatGraph& out);

}  // namespace psi_b200
#endif
