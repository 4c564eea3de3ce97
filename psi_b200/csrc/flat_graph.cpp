// flat_graph.cpp -- GFA loading and gum-compatible node ordering (host only).
#include "flat_graph.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include <zlib.h>

namespace psi_b200 {

namespace {

// gum::coordinate::Stoid (gum/coordinate.hpp:178-192): leading digits, else
// trailing digits of the segment name.
uint64_t name_to_id(const std::string& name)
{
  if (name.empty()) throw std::runtime_error("empty segment name");
  size_t i = 0;
  while (i < name.size() && std::isdigit((unsigned char)name[i])) ++i;
  if (i > 0) return std::strtoull(name.substr(0, i).c_str(), nullptr, 10);
  size_t j = name.size();
  while (j > 0 && std::isdigit((unsigned char)name[j - 1])) --j;
  if (j == name.size()) throw std::runtime_error("segment name without a numeric id: " + name);
  return std::strtoull(name.substr(j).c_str(), nullptr, 10);
}

std::vector<std::string> split(const std::string& s, char sep)
{
  std::vector<std::string> out;
  size_t b = 0;
  while (true) {
    size_t e = s.find(sep, b);
    if (e == std::string::npos) { out.push_back(s.substr(b)); break; }
    out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
  return out;
}

}  // namespace

void build_flat_graph(RawGraph&& raw, bool sort, FlatGraph& out)
{
  const size_t n = raw.ids.size();
  if (n >= (size_t)UINT32_MAX) throw std::runtime_error("too many nodes");

  // Per-node successor lists in gum's for_each_edges_out order: links leaving
  // the start side first, then links leaving the end side, each in insertion
  // (file) order (gum/digraph_traits_base.hpp:179-187, digraph_dynamic.hpp:619-637).
  std::vector<uint32_t> outdeg(n, 0), indeg(n, 0);
  for (auto const& e : raw.edges) {
    if (e.from >= n || e.to >= n) throw std::runtime_error("edge refers to an unknown node");
    ++outdeg[e.from];
    ++indeg[e.to];
  }
  std::vector<uint64_t> optr(n + 1, 0);
  for (size_t i = 0; i < n; ++i) optr[i + 1] = optr[i] + outdeg[i];
  std::vector<uint32_t> succ(raw.edges.size());
  {
    std::vector<uint64_t> fill(optr.begin(), optr.end() - 1);
    for (int pass = 0; pass < 2; ++pass)
      for (auto const& e : raw.edges)
        if (e.from_start == (pass == 0)) succ[fill[e.from]++] = e.to;
  }

  // order[i] = input index of the node at rank i
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  if (sort) {
    // 1. graph.sort_nodes(): by id (gum/gfa_utils.hpp:552).
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return raw.ids[a] < raw.ids[b]; });
    std::vector<uint32_t> rank_of(n);
    for (size_t r = 0; r < n; ++r) rank_of[order[r]] = (uint32_t)r;

    // 2. topological_sort(graph, force=true): iterative DFS, start nodes (no
    //    in-edge) pushed in rank order, a node's undiscovered successors pushed
    //    in out-edge order, nodes recorded when finished; the new order is the
    //    reverse finishing order (gum/seqgraph_interface.hpp:252-360).
    std::vector<uint8_t> discovered(n, 0), finished(n, 0);
    std::vector<uint32_t> stack, fin;
    fin.reserve(n);
    for (size_t r = 0; r < n; ++r)
      if (indeg[order[r]] == 0) stack.push_back(order[r]);
    size_t next_rank = 0;
    while (true) {
      while (!stack.empty()) {
        uint32_t v = stack.back();
        if (discovered[v]) {
          if (!finished[v]) { finished[v] = 1; fin.push_back(v); }
          stack.pop_back();
          continue;
        }
        discovered[v] = 1;
        for (uint64_t e = optr[v]; e < optr[v + 1]; ++e)
          if (!discovered[succ[e]]) stack.push_back(succ[e]);
      }
      while (next_rank < n && discovered[order[next_rank]]) ++next_rank;
      if (next_rank == n) break;
      stack.push_back(order[next_rank]);
    }
    std::reverse(fin.begin(), fin.end());
    order.swap(fin);
    (void)rank_of;
  }

  std::vector<uint32_t> rank_of(n);
  for (size_t r = 0; r < n; ++r) rank_of[order[r]] = (uint32_t)r;

  out = FlatGraph();
  out.seq_start.assign(n + 1, 0);
  out.row_ptr.assign(n + 1, 0);
  out.coord_id.resize(n);
  out.internal_id.resize(n);
  out.indeg.resize(n);
  uint64_t total = 0;
  for (size_t r = 0; r < n; ++r) total += raw.labels[order[r]].size();
  out.seq.reserve(total);
  out.col.reserve(raw.edges.size());
  // gum Succinct id = 1-based position of the node record in the node array:
  // header of 5 words + 3 words per out-edge and per in-edge
  // (gum/graph_traits_succinct.hpp:27-57, digraph_succinct.hpp:937-964).
  uint64_t next_id = 1;
  for (size_t r = 0; r < n; ++r) {
    uint32_t v = order[r];
    out.seq_start[r] = out.seq.size();
    out.seq += raw.labels[v];
    out.row_ptr[r] = out.col.size();
    for (uint64_t e = optr[v]; e < optr[v + 1]; ++e) out.col.push_back(rank_of[succ[e]]);
    out.coord_id[r] = raw.ids[v];
    out.indeg[r] = indeg[v];
    out.internal_id[r] = next_id;
    next_id += 5 + 3 * (uint64_t)(outdeg[v] + indeg[v]);
  }
  out.seq_start[n] = out.seq.size();
  out.row_ptr[n] = out.col.size();
  for (auto& p : raw.paths) {
    EmbeddedPath ep;
    ep.name = p.name;
    ep.nodes.reserve(p.nodes.size());
    for (uint32_t v : p.nodes) {
      if (v >= n) throw std::runtime_error("path refers to an unknown node");
      ep.nodes.push_back(rank_of[v]);
    }
    out.paths.push_back(std::move(ep));
  }
}

void load_gfa(const std::string& path, bool sort, FlatGraph& out)
{
  // plain or gzip-compressed text (zlib reads both)
  gzFile gz = gzopen(path.c_str(), "rb");
  if (!gz) throw std::runtime_error("could not open file '" + path + "'");
  gzbuffer(gz, 1 << 20);
  std::string content;
  {
    std::vector<char> buf(1 << 20);
    int got;
    while ((got = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) content.append(buf.data(), (size_t)got);
    gzclose(gz);
    if (got < 0) throw std::runtime_error("read error in '" + path + "'");
  }
  std::istringstream in(std::move(content));
  RawGraph raw;
  std::unordered_map<std::string, uint32_t> index_of;
  struct PendingEdge { std::string from, to; bool from_start; };
  std::vector<PendingEdge> pedges;
  struct PendingPath { std::string name; std::vector<std::string> segs; };
  std::vector<PendingPath> ppaths;

  auto strip_orient = [](std::string s, bool& fwd) {
    fwd = true;
    if (!s.empty() && (s.back() == '+' || s.back() == '-')) { fwd = s.back() == '+'; s.pop_back(); }
    return s;
  };

  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    auto f = split(line, '\t');
    const std::string& t = f[0];
    if (t == "S") {
      if (f.size() < 3) throw std::runtime_error("malformed S line");
      // GFA2: S sid slen seq ; GFA1: S name seq
      std::string seqstr;
      bool gfa2 = f.size() >= 4 && !f[2].empty() &&
                  std::all_of(f[2].begin(), f[2].end(), [](unsigned char c) { return std::isdigit(c); });
      seqstr = gfa2 ? f[3] : f[2];
      if (seqstr == "*") seqstr.clear();
      if (index_of.count(f[1])) throw std::runtime_error("duplicate segment " + f[1]);
      index_of[f[1]] = (uint32_t)raw.ids.size();
      raw.ids.push_back(name_to_id(f[1]));
      raw.labels.push_back(seqstr);
    }
    else if (t == "L") {
      if (f.size() < 5) throw std::runtime_error("malformed L line");
      pedges.push_back({ f[1], f[3], f[2] == "-" });
    }
    else if (t == "E") {
      if (f.size() < 4) throw std::runtime_error("malformed E line");
      bool ff, tf;
      std::string from = strip_orient(f[2], ff);
      std::string to = strip_orient(f[3], tf);
      pedges.push_back({ from, to, !ff });
    }
    else if (t == "P") {
      if (f.size() < 3) throw std::runtime_error("malformed P line");
      PendingPath pp;
      pp.name = f[1];
      for (auto& s : split(f[2], ',')) { bool fw; pp.segs.push_back(strip_orient(s, fw)); }
      ppaths.push_back(std::move(pp));
    }
    else if (t == "O") {
      if (f.size() < 3) throw std::runtime_error("malformed O line");
      PendingPath pp;
      pp.name = f[1];
      for (auto& s : split(f[2], ' ')) { if (s.empty()) continue; bool fw; pp.segs.push_back(strip_orient(s, fw)); }
      ppaths.push_back(std::move(pp));
    }
    /* H, C, F, G, U, W ... : not needed on this path */
  }
  for (auto const& e : pedges) {
    auto a = index_of.find(e.from), b = index_of.find(e.to);
    if (a == index_of.end() || b == index_of.end()) throw std::runtime_error("link refers to an unknown segment");
    raw.edges.push_back({ a->second, b->second, e.from_start });
  }
  // gfak keeps paths in a std::map by name (gfakluge.hpp:625) and gum adds them
  // in that order (gum/gfa_utils.hpp:556-558).
  std::sort(ppaths.begin(), ppaths.end(), [](PendingPath const& a, PendingPath const& b) { return a.name < b.name; });
  for (auto& pp : ppaths) {
    RawGraph::RawPath rp;
    rp.name = pp.name;
    for (auto& s : pp.segs) {
      auto it = index_of.find(s);
      if (it == index_of.end()) throw std::runtime_error("path refers to an unknown segment");
      rp.nodes.push_back(it->second);
    }
    raw.paths.push_back(std::move(rp));
  }
  build_flat_graph(std::move(raw), sort, out);
}

void load_graph_file(const std::string& path, bool sort, FlatGraph& out)
{
  if (path.size() > 3 && path.compare(path.size() - 3, 3, ".vg") == 0) load_vg(path, sort, out);
  else load_gfa(path, sort, out);
}

void write_gfa1(const FlatGraph& g, const std::string& path)
{
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("could not open file '" + path + "' for writing");
  std::fputs("H\tVN:Z:1.0\n", f);
  const uint64_t n = g.node_count();
  for (uint64_t r = 0; r < n; ++r) {
    std::fprintf(f, "S\t%llu\t", (unsigned long long)g.coord_id[r]);
    std::fwrite(g.seq.data() + g.seq_start[r], 1, g.seq_start[r + 1] - g.seq_start[r], f);
    std::fputc('\n', f);
  }
  for (uint64_t r = 0; r < n; ++r)
    for (uint64_t e = g.row_ptr[r]; e < g.row_ptr[r + 1]; ++e)
      std::fprintf(f, "L\t%llu\t+\t%llu\t+\t0M\n", (unsigned long long)g.coord_id[r],
                   (unsigned long long)g.coord_id[g.col[e]]);
  for (auto const& p : g.paths) {
    std::fprintf(f, "P\t%s\t", p.name.c_str());
    for (size_t i = 0; i < p.nodes.size(); ++i)
      std::fprintf(f, "%s%llu+", i ? "," : "", (unsigned long long)g.coord_id[p.nodes[i]]);
    std::fputs("\t*\n", f);
  }
  std::fclose(f);
}

}  // namespace psi_b200
