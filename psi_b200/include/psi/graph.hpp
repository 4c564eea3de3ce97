// psi/graph.hpp -- the host sequence graph the finder borrows, and Position<>.
//
// Stands where the reference uses gum::SeqGraph<gum::Succinct>
// (src/psikt.cpp:247-251) and psi::Position<> (include/psi/graph.hpp:33-82).
// The accessors carry gum's names and meaning (ranks are 1-based, ids are gum's
// internal Succinct ids, coordinate ids are the ids of the input file; see
// gum/digraph_succinct.hpp:238-242,937-964) so that code written against the
// reference's graph type reads the same; the storage is the flattened host
// graph of libpsi_b200 (psi_b200_graph).
#ifndef PSI_B200_PSI_GRAPH_HPP
#define PSI_B200_PSI_GRAPH_HPP

#include <atomic>
#include <cstdint>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include "../../../include/psi_b200.h"

namespace gum {
struct Succinct {};
struct Dynamic {};
}  // namespace gum

namespace psi {

template <typename TId = int64_t, typename TOffset = uint64_t>
class PositionBase {
 public:
  using id_type = TId;
  using offset_type = TOffset;
  PositionBase(id_type id = 0, offset_type offset = 0) : m_id(id), m_offset(offset) {}
  id_type node_id() const { return m_id; }
  offset_type offset() const { return m_offset; }
  void set_node_id(id_type id) { m_id = id; }
  void set_offset(offset_type o) { m_offset = o; }
  bool operator==(const PositionBase& o) const { return m_id == o.m_id && m_offset == o.m_offset; }

 private:
  TId m_id;
  TOffset m_offset;
};
template <typename TId = int64_t, typename TOffset = uint64_t>
using Position = PositionBase<TId, TOffset>;

// gum::SeqGraph<Succinct> stand-in.  Move-only owner of a psi_b200_graph.
template <typename TSpec = gum::Succinct>
class SeqGraph {
 public:
  typedef TSpec spec_type;
  typedef int64_t id_type;
  typedef uint64_t offset_type;
  typedef uint64_t rank_type;

  SeqGraph() = default;
  SeqGraph(const SeqGraph&) = delete;
  SeqGraph& operator=(const SeqGraph&) = delete;
  SeqGraph(SeqGraph&& o) noexcept { swap(o); }
  SeqGraph& operator=(SeqGraph&& o) noexcept { swap(o); return *this; }
  ~SeqGraph() { psi_b200_graph_free(h_); }

  // takes ownership
  void reset(psi_b200_graph* h)
  {
    psi_b200_graph_free(h_);
    h_ = h;
    by_id_.clear();
    by_id_ready_.store(false);
    if (h_ && psi_b200_graph_get_view(h_, &v_) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
  }
  const psi_b200_graph* handle() const { return h_; }
  const psi_b200_graph_view& view() const { return v_; }

  rank_type get_node_count() const { return h_ ? v_.n_nodes : 0; }
  rank_type get_edge_count() const { return h_ ? v_.n_edges : 0; }
  rank_type get_path_count() const { return h_ ? v_.n_paths : 0; }
  // 1-based rank <-> internal id (gum/digraph_succinct.hpp:937-964)
  id_type rank_to_id(rank_type rank) const { check_rank(rank); return (id_type)v_.internal_id[rank - 1]; }
  rank_type id_to_rank(id_type id) const
  {
    if (!by_id_ready_.load(std::memory_order_acquire)) {     // built once, by whichever thread asks first
      std::lock_guard<std::mutex> lk(by_id_mutex_);
      if (!by_id_ready_.load(std::memory_order_relaxed)) {
        by_id_.clear();
        by_id_.reserve(v_.n_nodes);
        for (uint64_t r = 0; r < v_.n_nodes; ++r) by_id_.emplace(v_.internal_id[r], r + 1);
        by_id_ready_.store(true, std::memory_order_release);
      }
    }
    auto it = by_id_.find((uint64_t)id);
    if (it == by_id_.end()) throw std::runtime_error("node id not found");
    return it->second;
  }
  id_type coordinate_id(id_type id) const { return (id_type)v_.coord_id[id_to_rank(id) - 1]; }
  // external (coordinate) id -> internal id (gum id_by_coordinate); linear scan: a test / tooling accessor
  id_type id_by_coordinate(id_type cid) const
  {
    for (uint64_t r = 0; r < v_.n_nodes; ++r) if (v_.coord_id[r] == (uint64_t)cid) return (id_type)v_.internal_id[r];
    throw std::runtime_error("coordinate id not found");
  }
  offset_type node_length(id_type id) const { rank_type r = id_to_rank(id); return v_.seq_start[r] - v_.seq_start[r - 1]; }
  std::string node_sequence(id_type id) const
  {
    rank_type r = id_to_rank(id);
    return std::string(v_.seq + v_.seq_start[r - 1], v_.seq + v_.seq_start[r]);
  }
  rank_type outdegree(id_type id) const { rank_type r = id_to_rank(id); return v_.row_ptr[r] - v_.row_ptr[r - 1]; }
  bool has_edges_out(id_type id) const { return outdegree(id) != 0; }
  // callback(to_id, linktype) -> bool (false stops), gum/digraph_succinct.hpp:595-610
  bool for_each_edges_out(id_type id, std::function<bool(id_type, int)> cb) const
  {
    rank_type r = id_to_rank(id);
    for (uint64_t e = v_.row_ptr[r - 1]; e < v_.row_ptr[r]; ++e)
      if (!cb((id_type)v_.internal_id[v_.col[e]], 0)) return false;
    return true;
  }
  bool for_each_node(std::function<bool(rank_type, id_type)> cb) const
  {
    for (uint64_t r = 1; r <= v_.n_nodes; ++r)
      if (!cb(r, (id_type)v_.internal_id[r - 1])) return false;
    return true;
  }
  // callback(path rank (1-based), path id) -> bool
  bool for_each_path(std::function<bool(rank_type, id_type)> cb) const
  {
    for (uint64_t p = 1; p <= v_.n_paths; ++p)
      if (!cb(p, (id_type)p)) return false;
    return true;
  }
  std::string path_name(id_type path_id) const
  {
    const char* name = nullptr;
    if (psi_b200_graph_path(h_, (uint64_t)path_id - 1, &name, nullptr, nullptr) != PSI_B200_OK)
      throw std::runtime_error(psi_b200_global_error());
    return name;
  }

 private:
  void check_rank(rank_type rank) const { if (rank == 0 || rank > v_.n_nodes) throw std::runtime_error("rank out of range"); }
  void swap(SeqGraph& o)
  {
    std::swap(h_, o.h_); std::swap(v_, o.v_); by_id_.swap(o.by_id_);
    const bool a = by_id_ready_.load(), b = o.by_id_ready_.load();
    by_id_ready_.store(b); o.by_id_ready_.store(a);
  }
  psi_b200_graph* h_ = nullptr;
  psi_b200_graph_view v_{};
  mutable std::unordered_map<uint64_t, uint64_t> by_id_;
  mutable std::atomic<bool> by_id_ready_{ false };
  mutable std::mutex by_id_mutex_;
};

namespace util {

// gum::util::load(graph, fname, sort): by file name like gum (gum/io_utils.hpp:66-80) -- `*.vg` is read as vg's protobuf
// stream (no protobuf library involved), anything else as GFA1 / GFA2 (gum/gfa_utils.hpp:541-554).
template <typename TSpec>
inline void load(SeqGraph<TSpec>& graph, const std::string& fname, bool sort = true)
{
  psi_b200_graph* h = nullptr;
  if (psi_b200_graph_load(fname.c_str(), sort ? 1 : 0, &h) != PSI_B200_OK)
    throw std::runtime_error(psi_b200_global_error());
  graph.reset(h);
}

// gum::util::ids_in_topological_order (checks coordinate ids along every edge)
template <typename TSpec>
inline bool ids_in_topological_order(const SeqGraph<TSpec>& graph)
{
  const psi_b200_graph_view& v = graph.view();
  for (uint64_t r = 0; r < v.n_nodes; ++r)
    for (uint64_t e = v.row_ptr[r]; e < v.row_ptr[r + 1]; ++e)
      if (v.coord_id[v.col[e]] <= v.coord_id[r]) return false;
  return true;
}

}  // namespace util
}  // namespace psi

namespace gum {
template <typename TSpec = Succinct>
using SeqGraph = ::psi::SeqGraph<TSpec>;
namespace util {
using ::psi::util::ids_in_topological_order;
using ::psi::util::load;
}  // namespace util
}  // namespace gum
#endif
