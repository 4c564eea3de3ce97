// bench_support/synth.cpp -- synthetic variation graphs and reads for bench.py
// and the large-size tests.  Bench/test support, not product code
// (the reference's analogue is the `ggsim` tool, out of scope).
//
//   graph : backbone of iid uniform ACGT; variant sites at distinct uniform
//           positions: SNP (one or, with tri_frac, two alternative bases),
//           insertion, deletion (indel length ~ Geometric(0.4) capped at 20);
//           every site is a bubble; node ids 1..N in backbone order; one
//           embedded path along the reference alleles.
//   reads : uniform start base over all graph bases, uniform out-edge at every
//           node end, fixed length, error free; walks that reach a sink early
//           are discarded and redrawn.
// PRNG: splitmix64, so results are identical on every platform.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

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
  uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
  double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct SynthGraph {
  std::vector<uint64_t> seq_start{ 0 };
  std::string seq;
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col;
  std::vector<uint32_t> path;
  std::vector<std::vector<uint32_t>> out;  // temporary adjacency
};

const char ACGT[] = "ACGT";

uint32_t add_node(SynthGraph& g, const char* s, size_t len)
{
  g.seq.append(s, len);
  g.seq_start.push_back(g.seq.size());
  g.out.emplace_back();
  return (uint32_t)(g.out.size() - 1);
}

uint32_t geometric(SplitMix64& r, double p, uint32_t cap)
{
  uint32_t n = 1;
  while (n < cap && r.unit() >= p) ++n;
  return n;
}

}  // namespace

extern "C" {

void* psi_synth_graph_create(uint64_t backbone, uint64_t sites, double p_snp, double p_ins, double tri_frac,
                             uint64_t seed_backbone, uint64_t seed_sites)
{
  SynthGraph* g = new SynthGraph();
  SplitMix64 rb(seed_backbone), rs(seed_sites);
  std::string bb(backbone, 'A');
  for (uint64_t i = 0; i < backbone; ++i) bb[i] = ACGT[rb.next() >> 62];
  // distinct site positions in [1, backbone-1)
  std::vector<uint64_t> pos;
  if (backbone > 2) {
    pos.reserve(sites + sites / 16 + 16);
    while (true) {
      while (pos.size() < sites + sites / 16 + 16) pos.push_back(1 + rs.below(backbone - 2));
      std::sort(pos.begin(), pos.end());
      pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
      if (pos.size() >= sites || pos.size() >= backbone - 2) break;
    }
    if (pos.size() > sites) {
      // drop random surplus positions (keeps the rest uniform)
      for (uint64_t i = pos.size(); i > sites; --i) {
        uint64_t j = rs.below(i);
        pos[j] = pos[i - 1];
        pos.pop_back();
      }
      std::sort(pos.begin(), pos.end());
    }
  }
  g->seq.reserve(backbone + backbone / 16);
  std::vector<uint32_t> tails;
  uint64_t cur = 0;
  auto connect = [&](uint32_t head) { for (uint32_t t : tails) g->out[t].push_back(head); };
  for (size_t si = 0; si < pos.size(); ++si) {
    const uint64_t p = pos[si];
    if (p < cur) continue;  // swallowed by a deletion
    if (p > cur) {
      uint32_t ref = add_node(*g, bb.data() + cur, p - cur);
      connect(ref);
      tails.assign(1, ref);
      g->path.push_back(ref);
      cur = p;
    }
    const double u = rs.unit();
    if (u < p_snp) {
      uint32_t a = add_node(*g, bb.data() + p, 1);
      char alts[3];
      int na = 0;
      for (int b = 0; b < 4; ++b) if (ACGT[b] != bb[p]) alts[na++] = ACGT[b];
      int first = (int)rs.below(3);
      uint32_t b1 = add_node(*g, &alts[first], 1);
      uint32_t b2 = UINT32_MAX;
      if (rs.unit() < tri_frac) {
        int second = (first + 1 + (int)rs.below(2)) % 3;
        b2 = add_node(*g, &alts[second], 1);
      }
      connect(a); connect(b1); if (b2 != UINT32_MAX) connect(b2);
      tails.assign(1, a); tails.push_back(b1); if (b2 != UINT32_MAX) tails.push_back(b2);
      g->path.push_back(a);
      cur = p + 1;
    }
    else if (u < p_snp + p_ins) {
      uint32_t len = geometric(rs, 0.4, 20);
      char buf[32];
      for (uint32_t i = 0; i < len; ++i) buf[i] = ACGT[rs.next() >> 62];
      uint32_t ins = add_node(*g, buf, len);
      connect(ins);
      tails.push_back(ins);
    }
    else {
      uint64_t len = geometric(rs, 0.4, 20);
      uint64_t limit = (si + 1 < pos.size() ? pos[si + 1] : backbone - 1) - p;
      if (len > limit) len = limit;
      if (len == 0) continue;
      uint32_t d = add_node(*g, bb.data() + p, len);
      connect(d);
      tails.push_back(d);
      g->path.push_back(d);
      cur = p + len;
    }
  }
  if (cur < backbone) {
    uint32_t ref = add_node(*g, bb.data() + cur, backbone - cur);
    connect(ref);
    g->path.push_back(ref);
  }
  g->row_ptr.assign(1, 0);
  for (auto& o : g->out) {
    for (uint32_t v : o) g->col.push_back(v);
    g->row_ptr.push_back(g->col.size());
  }
  g->out.clear();
  g->out.shrink_to_fit();
  return g;
}

void psi_synth_graph_sizes(void* h, uint64_t* n_nodes, uint64_t* n_edges, uint64_t* n_bases, uint64_t* path_len)
{
  SynthGraph* g = (SynthGraph*)h;
  *n_nodes = g->seq_start.size() - 1;
  *n_edges = g->col.size();
  *n_bases = g->seq.size();
  *path_len = g->path.size();
}

void psi_synth_graph_fill(void* h, uint64_t* ids, uint64_t* seq_start, char* seq, uint64_t* row_ptr, uint32_t* col,
                          uint32_t* path)
{
  SynthGraph* g = (SynthGraph*)h;
  const uint64_t n = g->seq_start.size() - 1;
  for (uint64_t i = 0; i < n; ++i) ids[i] = i + 1;
  std::memcpy(seq_start, g->seq_start.data(), (n + 1) * sizeof(uint64_t));
  std::memcpy(seq, g->seq.data(), g->seq.size());
  std::memcpy(row_ptr, g->row_ptr.data(), (n + 1) * sizeof(uint64_t));
  if (!g->col.empty()) std::memcpy(col, g->col.data(), g->col.size() * sizeof(uint32_t));
  if (!g->path.empty()) std::memcpy(path, g->path.data(), g->path.size() * sizeof(uint32_t));
}

void psi_synth_graph_free(void* h) { delete (SynthGraph*)h; }

// Random-walk reads of fixed length into bases[n_reads * length].
uint64_t psi_synth_reads(uint64_t n_nodes, const uint64_t* seq_start, const char* seq, const uint64_t* row_ptr,
                         const uint32_t* col, uint64_t n_reads, uint32_t length, uint64_t seed, char* bases)
{
  SplitMix64 r(seed);
  const uint64_t n_bases = seq_start[n_nodes];
  uint64_t done = 0, tries = 0;
  while (done < n_reads && tries < n_reads * 64 + 1024) {
    ++tries;
    const uint64_t pos = r.below(n_bases);
    uint64_t v = (uint64_t)(std::upper_bound(seq_start, seq_start + n_nodes + 1, pos) - seq_start) - 1;
    uint64_t o = pos - seq_start[v];
    char* out = bases + done * length;
    uint32_t got = 0;
    bool ok = true;
    while (got < length) {
      const uint64_t s = seq_start[v] + o, e = seq_start[v + 1];
      const uint32_t take = (uint32_t)std::min<uint64_t>(e - s, length - got);
      std::memcpy(out + got, seq + s, take);
      got += take;
      if (got == length) break;
      const uint64_t b = row_ptr[v], en = row_ptr[v + 1];
      if (b == en) { ok = false; break; }
      v = col[b + r.below(en - b)];
      o = 0;
    }
    if (ok) ++done;
  }
  return done;
}

}  // extern "C"
