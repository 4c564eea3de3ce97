// capi_host.cpp -- the host half of the extern "C" surface (include/psi_b200.h): graphs, path picking, read
// chunking and packing.  No CUDA: together with flat_graph.cpp, paths.cpp and reads.cpp it also builds
// libpsi_b200_host.so, for processes that must not map the CUDA library (bench.py's reference arm).
#include <cstring>
#include <string>

#include "capi_common.hpp"
#include "flat_graph.hpp"
#include "paths.hpp"
#include "reads.hpp"

using namespace psi_b200;

struct psi_b200_graph { FlatGraph g; };
struct psi_b200_pathset { PathSet p; };
struct psi_b200_reader { ChunkReader* r; };

namespace psi_b200 {
std::string& capi_global_error()
{
  thread_local std::string e;
  return e;
}
}  // namespace psi_b200

#define g_error (::psi_b200::capi_global_error())

namespace {
int translate(std::string& err)
{
  try { throw; }
  catch (const std::bad_alloc&) { err = "out of host memory"; return PSI_B200_ERR_NOMEM; }
  catch (const std::exception& e) { err = e.what(); return PSI_B200_ERR_ARG; }
  catch (...) { err = "unknown error"; return PSI_B200_ERR_ARG; }
}
}  // namespace

extern "C" {

const char* psi_b200_global_error(void) { return g_error.c_str(); }

/* ------------------------------------------------------------ graph -- */

int psi_b200_graph_load_gfa(const char* path, int sort, psi_b200_graph** out)
{
  if (!path || !out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  psi_b200_graph* g = nullptr;
  try {
    g = new psi_b200_graph();
    load_gfa(path, sort != 0, g->g);
    *out = g;
    return PSI_B200_OK;
  }
  catch (...) {
    delete g;
    int rc = translate(g_error);
    return rc == PSI_B200_ERR_ARG && g_error.find("could not open") != std::string::npos ? PSI_B200_ERR_IO : rc;
  }
}

static int load_with(void (*loader)(const std::string&, bool, FlatGraph&), const char* path, int sort, psi_b200_graph** out)
{
  if (!path || !out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  psi_b200_graph* g = nullptr;
  try {
    g = new psi_b200_graph();
    loader(path, sort != 0, g->g);
    *out = g;
    return PSI_B200_OK;
  }
  catch (...) {
    delete g;
    int rc = translate(g_error);
    return rc == PSI_B200_ERR_ARG && g_error.find("could not open") != std::string::npos ? PSI_B200_ERR_IO : rc;
  }
}

int psi_b200_graph_load_vg(const char* path, int sort, psi_b200_graph** out) { return load_with(load_vg, path, sort, out); }
int psi_b200_graph_load(const char* path, int sort, psi_b200_graph** out) { return load_with(load_graph_file, path, sort, out); }

int psi_b200_graph_from_arrays(uint64_t n_nodes, const uint64_t* ids, const uint64_t* seq_start, const char* seq,
                               const uint64_t* row_ptr, const uint32_t* col, uint64_t n_paths,
                               const uint64_t* path_ptr, const uint32_t* path_nodes, int sort,
                               psi_b200_graph** out)
{
  if (!out || !ids || !seq_start || !seq || !row_ptr) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  psi_b200_graph* g = nullptr;
  try {
    RawGraph raw;
    raw.ids.assign(ids, ids + n_nodes);
    raw.labels.resize(n_nodes);
    for (uint64_t v = 0; v < n_nodes; ++v) raw.labels[v].assign(seq + seq_start[v], seq + seq_start[v + 1]);
    raw.edges.reserve(row_ptr[n_nodes]);
    for (uint64_t v = 0; v < n_nodes; ++v)
      for (uint64_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e) raw.edges.push_back({ (uint32_t)v, col[e], false });
    for (uint64_t p = 0; p < n_paths; ++p) {
      RawGraph::RawPath rp;
      rp.name = "path" + std::to_string(p);
      rp.nodes.assign(path_nodes + path_ptr[p], path_nodes + path_ptr[p + 1]);
      raw.paths.push_back(std::move(rp));
    }
    g = new psi_b200_graph();
    build_flat_graph(std::move(raw), sort != 0, g->g);
    *out = g;
    return PSI_B200_OK;
  }
  catch (...) { delete g; return translate(g_error); }
}

void psi_b200_graph_free(psi_b200_graph* g) { delete g; }

int psi_b200_graph_get_view(const psi_b200_graph* g, psi_b200_graph_view* v)
{
  if (!g || !v) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  const FlatGraph& f = g->g;
  v->n_nodes = f.node_count();
  v->n_edges = f.edge_count();
  v->n_bases = f.seq.size();
  v->n_paths = f.paths.size();
  v->seq_start = f.seq_start.data();
  v->seq = f.seq.data();
  v->row_ptr = f.row_ptr.data();
  v->col = f.col.data();
  v->internal_id = f.internal_id.data();
  v->coord_id = f.coord_id.data();
  return PSI_B200_OK;
}

int psi_b200_graph_path(const psi_b200_graph* g, uint64_t i, const char** name, const uint32_t** nodes, uint64_t* n_nodes)
{
  if (!g || i >= g->g.paths.size()) { g_error = "path index out of range"; return PSI_B200_ERR_ARG; }
  if (name) *name = g->g.paths[i].name.c_str();
  if (nodes) *nodes = g->g.paths[i].nodes.data();
  if (n_nodes) *n_nodes = g->g.paths[i].nodes.size();
  return PSI_B200_OK;
}

int psi_b200_graph_write_gfa(const psi_b200_graph* g, const char* path)
{
  if (!g || !path) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  try { write_gfa1(g->g, path); return PSI_B200_OK; }
  catch (...) { translate(g_error); return PSI_B200_ERR_IO; }
}

/* ------------------------------------------------------------ paths -- */

int psi_b200_pick_paths(const psi_b200_graph* g, unsigned n, int patched, unsigned context, uint64_t seed,
                        psi_b200_pathset** out)
{
  if (!g || !out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  psi_b200_pathset* p = nullptr;
  try {
    p = new psi_b200_pathset();
    pick_paths(g->g, n, patched != 0, context, seed, p->p);
    *out = p;
    return PSI_B200_OK;
  }
  catch (...) { delete p; return translate(g_error); }
}

int psi_b200_pathset_load_reference(const psi_b200_graph* g, const char* paths_file, psi_b200_pathset** out, uint64_t* context)
{
  if (!g || !paths_file || !out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  psi_b200_pathset* p = nullptr;
  try {
    p = new psi_b200_pathset();
    uint64_t ctx = 0;
    load_reference_paths(g->g, paths_file, p->p, ctx);
    if (context) *context = ctx;
    *out = p;
    return PSI_B200_OK;
  }
  catch (...) { delete p; translate(g_error); return PSI_B200_ERR_IO; }
}

void psi_b200_pathset_free(psi_b200_pathset* p) { delete p; }

int psi_b200_pathset_get_view(const psi_b200_pathset* p, psi_b200_pathset_view* v)
{
  if (!p || !v) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  v->n_paths = p->p.size();
  v->path_ptr = p->p.path_ptr.data();
  v->nodes = p->p.nodes.data();
  v->head_off = p->p.head_off.data();
  v->tail_trim = p->p.tail_trim.data();
  return PSI_B200_OK;
}

/* ------------------------------------------------------------ reads -- */

int psi_b200_reader_open(const char* path, psi_b200_reader** out)
{
  if (!path || !out) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  *out = nullptr;
  try {
    ChunkReader* r = new ChunkReader(path);
    *out = new psi_b200_reader{ r };
    return PSI_B200_OK;
  }
  catch (...) { translate(g_error); return PSI_B200_ERR_IO; }
}

int psi_b200_reader_next(psi_b200_reader* r, uint64_t max_reads, psi_b200_chunk_view* v)
{
  if (!r || !r->r || !v) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  HOST_GUARD({
    r->r->next(max_reads);
    v->n_reads = r->r->n_reads();
    v->first_read_id = r->r->first_read_id();
    v->read_ptr = r->r->read_ptr();
    v->bases = r->r->bases();
    v->name_ptr = r->r->name_ptr();
    v->names = r->r->names();
  })
}

int psi_b200_reader_next_packed(psi_b200_reader* r, uint64_t max_reads, psi_b200_packed_chunk* v)
{
  if (!r || !r->r || !v) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  HOST_GUARD({
    ChunkReader& cr = *r->r;
    cr.next(max_reads, true);
    v->n_reads = cr.n_reads();
    v->first_read_id = cr.first_read_id();
    v->n_bases = cr.n_bases();
    v->read_len = cr.uniform_len();
    v->reserved = 0;
    v->read_ptr = cr.read_ptr();
    v->words = cr.words();
    v->exc = cr.exc();
    v->n_exc = cr.n_exc();
    v->name_ptr = cr.name_ptr();
    v->names = cr.names();
  })
}

int psi_b200_pack_bases(const char* bases, uint64_t n_bases, uint64_t* words, uint64_t* exc, uint64_t exc_cap, uint64_t* n_exc)
{
  if ((n_bases && !bases) || !words) { g_error = "null argument"; return PSI_B200_ERR_ARG; }
  HOST_GUARD({
    const uint64_t n = pack_bases(bases, n_bases, words, exc, exc_cap);
    if (n_exc) *n_exc = n;
  })
}

void psi_b200_reader_close(psi_b200_reader* r)
{
  if (!r) return;
  delete r->r;
  delete r;
}


}  // extern "C"
