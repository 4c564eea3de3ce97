// vg_reader.cpp -- reads vg's protobuf graph files (.vg) without libprotobuf / libvgio.
//
// Stands behind the reference CLI's graph loading for .vg input -- parse_vg + gum::util::load(graph, path, loader, true)
// (reference src/psikt.cpp:238-250) over its minivgio stream reader (vg/stream.hpp:82-123, vg/vg.proto) and gum's vg
// overloads (gum/io_utils.hpp:77-80, io_utils_vg.hpp, vg_utils.hpp:377-396, chunks merged like merge_vg :436-496):
// the nodes, edges and embedded paths of all Graph chunks of the file, added in file order, then the same id sort +
// topological sort as for GFA (build_flat_graph).  A .vg file is libvgio's stream -- BGZF / gzip (or nothing) around
// groups of `varint count, count x (varint length, message)`, the first message of a group being the type tag "VG" in
// type-tagged files -- of vg.proto's messages:
//     Graph    { repeated Node node = 1; repeated Edge edge = 2; repeated Path path = 3; }
//     Node     { string sequence = 1; string name = 2; int64 id = 3; }
//     Edge     { int64 from = 1; int64 to = 2; bool from_start = 3; bool to_end = 4; int32 overlap = 5; }
//     Path     { string name = 1; repeated Mapping mapping = 2; bool is_circular = 3; int64 length = 4; }
//     Mapping  { Position position = 1; repeated Edit edit = 2; int64 rank = 5; }
//     Position { int64 node_id = 1; int64 offset = 2; bool is_reverse = 4; string name = 5; }
// Only the wire format's varint (0), 64-bit (1), length-delimited (2) and 32-bit (5) types are needed.
#include "flat_graph.hpp"

#include <algorithm>
#include <stdexcept>
#include <unordered_map>

#include <zlib.h>

namespace psi_b200 {

namespace {

struct Span {
  const unsigned char* p;
  const unsigned char* e;
  bool empty() const { return p >= e; }
};

uint64_t varint(Span& s)
{
  uint64_t v = 0;
  for (unsigned shift = 0; shift < 64; shift += 7) {
    if (s.p >= s.e) throw std::runtime_error("truncated vg file (varint)");
    const unsigned char b = *s.p++;
    v |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) return v;
  }
  throw std::runtime_error("malformed vg file (varint longer than 10 bytes)");
}

Span bytes(Span& s)
{
  const uint64_t n = varint(s);
  if (n > (uint64_t)(s.e - s.p)) throw std::runtime_error("truncated vg file (length-delimited field)");
  Span r{ s.p, s.p + n };
  s.p += n;
  return r;
}

// next field of a message: its number and wire type; unknown fields are skipped by the callers through skip()
bool next_field(Span& s, uint32_t& field, uint32_t& wire)
{
  if (s.empty()) return false;
  const uint64_t tag = varint(s);
  field = (uint32_t)(tag >> 3);
  wire = (uint32_t)(tag & 7);
  if (field == 0) throw std::runtime_error("malformed vg file (field number 0)");
  return true;
}

void skip(Span& s, uint32_t wire)
{
  switch (wire) {
    case 0: (void)varint(s); break;
    case 1: if (s.e - s.p < 8) throw std::runtime_error("truncated vg file"); s.p += 8; break;
    case 2: (void)bytes(s); break;
    case 5: if (s.e - s.p < 4) throw std::runtime_error("truncated vg file"); s.p += 4; break;
    default: throw std::runtime_error("malformed vg file (unsupported wire type)");
  }
}

struct Mapping { int64_t rank; uint64_t node_id; };
struct PendingPath { std::string name; std::vector<Mapping> mappings; };

struct Collector {
  RawGraph raw;
  std::unordered_map<uint64_t, uint32_t> index_of;     // node id -> index into raw.ids
  struct PendingEdge { uint64_t from, to; bool from_start; };
  std::vector<PendingEdge> edges;
  std::vector<PendingPath> paths;                       // in order of first appearance (vg_utils.hpp:470-482)

  void node(Span s)
  {
    std::string seq;
    uint64_t id = 0;
    uint32_t f, w;
    while (next_field(s, f, w)) {
      if (f == 1 && w == 2) { Span b = bytes(s); seq.assign((const char*)b.p, (size_t)(b.e - b.p)); }
      else if (f == 3 && w == 0) id = varint(s);
      else skip(s, w);
    }
    auto it = index_of.find(id);
    if (it != index_of.end()) { raw.labels[it->second] = seq; return; }     // add_node(..., force = true) updates
    index_of.emplace(id, (uint32_t)raw.ids.size());
    raw.ids.push_back(id);
    raw.labels.push_back(std::move(seq));
  }

  void edge(Span s)
  {
    PendingEdge e{ 0, 0, false };
    uint32_t f, w;
    while (next_field(s, f, w)) {
      if (f == 1 && w == 0) e.from = varint(s);
      else if (f == 2 && w == 0) e.to = varint(s);
      else if (f == 3 && w == 0) e.from_start = varint(s) != 0;
      else skip(s, w);                                   // to_end, overlap: not needed on this path
    }
    edges.push_back(e);
  }

  void mapping(Span s, PendingPath& pp)
  {
    Mapping m{ 0, 0 };
    uint32_t f, w;
    while (next_field(s, f, w)) {
      if (f == 1 && w == 2) {
        Span pos = bytes(s);
        uint32_t pf, pw;
        while (next_field(pos, pf, pw)) {
          if (pf == 1 && pw == 0) m.node_id = varint(pos);
          else skip(pos, pw);
        }
      }
      else if (f == 5 && w == 0) m.rank = (int64_t)varint(s);
      else skip(s, w);
    }
    pp.mappings.push_back(m);
  }

  void path(Span s)
  {
    // the name may come after the mappings on the wire: collect first, attach then
    PendingPath local;
    uint32_t f, w;
    while (next_field(s, f, w)) {
      if (f == 1 && w == 2) { Span b = bytes(s); local.name.assign((const char*)b.p, (size_t)(b.e - b.p)); }
      else if (f == 2 && w == 2) mapping(bytes(s), local);
      else skip(s, w);
    }
    for (auto& pp : paths)
      if (pp.name == local.name) { pp.mappings.insert(pp.mappings.end(), local.mappings.begin(), local.mappings.end()); return; }
    paths.push_back(std::move(local));
  }

  void graph(Span s)
  {
    uint32_t f, w;
    while (next_field(s, f, w)) {
      if (f == 1 && w == 2) node(bytes(s));
      else if (f == 2 && w == 2) edge(bytes(s));
      else if (f == 3 && w == 2) path(bytes(s));
      else skip(s, w);
    }
  }
};

bool looks_like_tag(Span m)
{
  const size_t n = (size_t)(m.e - m.p);
  if (n == 0 || n > 25) return false;
  for (const unsigned char* c = m.p; c < m.e; ++c)
    if (!((*c >= 'A' && *c <= 'Z') || (*c >= 'a' && *c <= 'z') || (*c >= '0' && *c <= '9') || *c == '_')) return false;
  return true;
}

}  // namespace

void load_vg(const std::string& path, bool sort, FlatGraph& out)
{
  gzFile gz = gzopen(path.c_str(), "rb");       // BGZF is concatenated gzip members; uncompressed files pass through
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
  Span s{ (const unsigned char*)content.data(), (const unsigned char*)content.data() + content.size() };
  // libbdsg's serialised HashGraph shares the .vg extension and starts with its magic number (gum/hg_utils.hpp)
  static const unsigned char HG_MAGIC[4] = { 0x28, 0x4d, 0x4f, 0x38 };
  if (content.size() >= 4 && std::equal(HG_MAGIC, HG_MAGIC + 4, s.p))
    throw std::runtime_error("'" + path + "' is a HashGraph file, not a vg protobuf stream: convert it with `vg convert -v` or to GFA");
  Collector c;
  while (!s.empty()) {
    const uint64_t count = varint(s);
    for (uint64_t i = 0; i < count; ++i) {
      Span m = bytes(s);
      if (i == 0 && looks_like_tag(m)) {
        if (std::string((const char*)m.p, (size_t)(m.e - m.p)) != "VG")
          throw std::runtime_error("'" + path + "' is not a vg graph file (type tag '" + std::string((const char*)m.p, (size_t)(m.e - m.p)) + "')");
        continue;
      }
      c.graph(m);
    }
  }
  for (auto const& e : c.edges) {
    // add_edge(..., force = true) creates missing end points as empty nodes (vg_utils.hpp:199-206)
    for (uint64_t id : { e.from, e.to })
      if (!c.index_of.count(id)) {
        c.index_of.emplace(id, (uint32_t)c.raw.ids.size());
        c.raw.ids.push_back(id);
        c.raw.labels.emplace_back();
      }
    c.raw.edges.push_back({ c.index_of[e.from], c.index_of[e.to], e.from_start });
  }
  for (auto& pp : c.paths) {
    // mappings in rank order (vg_utils.hpp:283-290); files without ranks keep their order
    std::stable_sort(pp.mappings.begin(), pp.mappings.end(), [](Mapping const& a, Mapping const& b) { return a.rank < b.rank; });
    RawGraph::RawPath rp;
    rp.name = pp.name;
    for (auto const& m : pp.mappings) {
      auto it = c.index_of.find(m.node_id);
      if (it == c.index_of.end()) throw std::runtime_error("path '" + pp.name + "' refers to an unknown node");
      rp.nodes.push_back(it->second);
    }
    c.raw.paths.push_back(std::move(rp));
  }
  build_flat_graph(std::move(c.raw), sort, out);
}

}  // namespace psi_b200
