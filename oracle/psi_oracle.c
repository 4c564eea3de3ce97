/*
 * oracle/psi_oracle.c -- TEST INFRASTRUCTURE (checker), never the product path.
 * See psi_oracle.h for scope, parity status and the pinning tests.
 *
 * Plain C11, single thread, no dependencies.  k <= 32 (a k-mer is packed into
 * one uint64_t, 2 bits per base, first base in the most significant position).
 */
#include "psi_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- bases -- */

/* A/C/G/T in either case -> 0..3; anything else -> 4 ('N': never matches;
 * traverser_bfs.hpp:124 for the graph side, index_iter.hpp:831-832 for the
 * on-path enumeration over DnaString). */
static inline unsigned base_code(char c)
{
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

static inline uint64_t kmer_mask(unsigned k)
{
  return k >= 32 ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) - 1);
}

/* Pack s[0..k) or return 0 (and leave *out untouched) when a non-ACGT is met. */
static int pack_kmer(const char* s, unsigned k, uint64_t* out)
{
  uint64_t v = 0;
  for (unsigned i = 0; i < k; ++i) {
    unsigned c = base_code(s[i]);
    if (c > 3) return 0;
    v = (v << 2) | c;
  }
  *out = v;
  return 1;
}

/* ------------------------------------------------------- growable tuples -- */

typedef struct {
  uint64_t* data;   /* 4 u64 per tuple */
  uint64_t  n, cap;
} tuple_vec;

static int tv_push(tuple_vec* v, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
  if (v->n == v->cap) {
    uint64_t ncap = v->cap ? v->cap * 2 : 1024;
    uint64_t* nd = (uint64_t*)realloc(v->data, ncap * 4 * sizeof(uint64_t));
    if (!nd) return -1;
    v->data = nd;
    v->cap = ncap;
  }
  uint64_t* t = v->data + 4 * v->n++;
  t[0] = a; t[1] = b; t[2] = c; t[3] = d;
  return 0;
}

static int tuple_cmp(const void* x, const void* y)
{
  const uint64_t* a = (const uint64_t*)x;
  const uint64_t* b = (const uint64_t*)y;
  for (int i = 0; i < 4; ++i) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}

/* Canonical form: sorted, unique. */
static void tv_finish(tuple_vec* v, psi_oracle_result* out)
{
  out->n_raw = v->n;
  if (v->n > 1) {
    qsort(v->data, v->n, 4 * sizeof(uint64_t), tuple_cmp);
    uint64_t w = 1;
    for (uint64_t i = 1; i < v->n; ++i) {
      if (tuple_cmp(v->data + 4 * i, v->data + 4 * (w - 1)) != 0) {
        if (w != i) memcpy(v->data + 4 * w, v->data + 4 * i, 4 * sizeof(uint64_t));
        ++w;
      }
    }
    v->n = w;
  }
  out->tuples = v->data;
  out->n = v->n;
}

/* ---------------------------------------------------------------- seeding -- */

/* sequence.hpp:1688-1718: `for (i = 0; i < len - k + 1; i += step)`; reads
 * shorter than k underflow in the reference -> no seeds here. */
uint64_t psi_oracle_seeding(const psi_oracle_reads* reads, unsigned k, unsigned d,
                            uint64_t* seed_read_id, uint64_t* seed_offset, uint64_t cap)
{
  uint64_t n = 0;
  if (d == 0) d = k;
  for (uint64_t r = 0; r < reads->n_reads; ++r) {
    uint64_t len = reads->read_ptr[r + 1] - reads->read_ptr[r];
    if (len < k) continue;
    for (uint64_t i = 0; i < len - k + 1; i += d) {
      if (n < cap) {
        /* SeedMap, sequence.hpp:1201-1213: id = rec_offset + rank; offset = rank_in_read * d */
        if (seed_read_id) seed_read_id[n] = reads->first_read_id + r;
        if (seed_offset) seed_offset[n] = i;
      }
      ++n;
    }
  }
  return n;
}

/* -------------------------------------------------- read-seed hash (k-mer) -- */

/* Stands in for the reference's suffix tree over the chunk's seeds
 * (seed_finder.hpp:1089-1097): k-mer -> chain of (read, offset). */
typedef struct {
  uint64_t  n_slots;      /* power of two */
  uint64_t* key;
  int64_t*  head;         /* -1 = empty slot */
  uint64_t  n_seeds;
  uint64_t* s_read;       /* global read id */
  uint64_t* s_off;
  int64_t*  s_next;
} seed_hash;

static inline uint64_t mix64(uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

static void sh_free(seed_hash* h)
{
  free(h->key); free(h->head); free(h->s_read); free(h->s_off); free(h->s_next);
  memset(h, 0, sizeof(*h));
}

static int sh_build(seed_hash* h, const psi_oracle_reads* reads, unsigned k, unsigned d)
{
  memset(h, 0, sizeof(*h));
  if (d == 0) d = k;
  uint64_t n = psi_oracle_seeding(reads, k, d, NULL, NULL, 0);
  uint64_t slots = 16;
  while (slots < 2 * n + 1) slots <<= 1;
  h->n_slots = slots;
  h->key = (uint64_t*)malloc(slots * sizeof(uint64_t));
  h->head = (int64_t*)malloc(slots * sizeof(int64_t));
  h->s_read = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
  h->s_off = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
  h->s_next = (int64_t*)malloc((n + 1) * sizeof(int64_t));
  if (!h->key || !h->head || !h->s_read || !h->s_off || !h->s_next) { sh_free(h); return -1; }
  for (uint64_t i = 0; i < slots; ++i) h->head[i] = -1;
  uint64_t s = 0;
  for (uint64_t r = 0; r < reads->n_reads; ++r) {
    uint64_t beg = reads->read_ptr[r];
    uint64_t len = reads->read_ptr[r + 1] - beg;
    if (len < k) continue;
    for (uint64_t i = 0; i < len - k + 1; i += d) {
      uint64_t km;
      h->s_read[s] = reads->first_read_id + r;
      h->s_off[s] = i;
      h->s_next[s] = -1;
      if (pack_kmer(reads->bases + beg + i, k, &km)) {   /* seeds with N never match */
        uint64_t p = mix64(km) & (slots - 1);
        while (h->head[p] >= 0 && h->key[p] != km) p = (p + 1) & (slots - 1);
        h->key[p] = km;
        h->s_next[s] = h->head[p];
        h->head[p] = (int64_t)s;
      }
      ++s;
    }
  }
  h->n_seeds = s;
  return 0;
}

static inline int64_t sh_find(const seed_hash* h, uint64_t km)
{
  uint64_t p = mix64(km) & (h->n_slots - 1);
  while (h->head[p] >= 0) {
    if (h->key[p] == km) return h->head[p];
    p = (p + 1) & (h->n_slots - 1);
  }
  return -1;
}

/* --------------------------------------------- k-mer exact matches (sets) -- */

typedef struct { uint64_t km; } km_rec;

static int km_cmp(const void* a, const void* b)
{
  uint64_t x = ((const km_rec*)a)->km, y = ((const km_rec*)b)->km;
  return x < y ? -1 : (x > y ? 1 : 0);
}

static km_rec* all_kmers(const char* const* set, uint64_t n, unsigned k, uint64_t* count)
{
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t len = strlen(set[i]);
    if (len >= k) total += len - k + 1;
  }
  km_rec* v = (km_rec*)malloc((total + 1) * sizeof(km_rec));
  uint64_t c = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t len = strlen(set[i]);
    for (uint64_t j = 0; j + k <= len; ++j) {
      uint64_t km;
      if (pack_kmer(set[i] + j, k, &km)) v[c++].km = km;
    }
  }
  qsort(v, c, sizeof(km_rec), km_cmp);
  *count = c;
  return v;
}

/* index_iter.hpp:808-852 visits, in lexicographic order, every k-mer over
 * {A,C,G,T} present in both indexes and emits occ1 x occ2 hits
 * (_add_occurrences, index_iter.hpp:728-746).  Two sorted lists walked in
 * step visit the same k-mers in the same order. */
uint64_t psi_oracle_kmer_exact_matches(const char* const* set1, uint64_t n1,
                                       const char* const* set2, uint64_t n2, unsigned k)
{
  if (k == 0 || k > 32) return 0;
  uint64_t c1, c2, hits = 0;
  km_rec* a = all_kmers(set1, n1, k, &c1);
  km_rec* b = all_kmers(set2, n2, k, &c2);
  uint64_t i = 0, j = 0;
  while (i < c1 && j < c2) {
    if (a[i].km < b[j].km) ++i;
    else if (a[i].km > b[j].km) ++j;
    else {
      uint64_t km = a[i].km, x = 0, y = 0;
      while (i < c1 && a[i].km == km) { ++i; ++x; }
      while (j < c2 && b[j].km == km) { ++j; ++y; }
      hits += x * y;
    }
  }
  free(a); free(b);
  return hits;
}

/* sequence.hpp:1639-1667 with continuous=false. */
uint64_t psi_oracle_increment_kmer(char* kmer, uint64_t len, uint64_t pos)
{
  static const char next_of[4] = { 'C', 'G', 'T', 'A' };
  for (uint64_t i = len - 1; i > pos; --i) kmer[i] = 'A';
  while (pos + 1 > 0 && kmer[pos] == 'T') { kmer[pos] = 'A'; --pos; }
  if (pos + 1 > 0) kmer[pos] = next_of[base_code(kmer[pos])];
  return pos;
}

/* ------------------------------------------------------------- on paths -- */

typedef struct {
  const psi_oracle_graph* g;
  const seed_hash* sh;
  unsigned k;
  tuple_vec* out;
  int err;
} emit_ctx;

static void emit_hits(emit_ctx* c, uint64_t km, uint32_t node, uint64_t off)
{
  for (int64_t s = sh_find(c->sh, km); s >= 0; s = c->sh->s_next[s]) {
    if (tv_push(c->out, c->sh->s_read[s], c->sh->s_off[s], c->g->node_id[node], off) != 0) c->err = -1;
  }
}

/* The reference indexes rev(path) texts and maps an occurrence back to the
 * forward (node, offset) of the k-mer's first base (index_iter.hpp:718-723,
 * pathindex.hpp:378-416).  Here the forward path text is scanned directly:
 * every window of k bases that lies inside the (trimmed) path. */
static int on_paths_collect(emit_ctx* c, const psi_oracle_paths* p)
{
  const psi_oracle_graph* g = c->g;
  unsigned k = c->k;
  for (uint64_t pi = 0; pi < p->n_paths; ++pi) {
    uint64_t nb = p->path_ptr[pi], ne = p->path_ptr[pi + 1];
    if (ne == nb) continue;
    /* materialise the path text with, per base, its (node, offset) */
    uint64_t total = 0;
    for (uint64_t j = nb; j < ne; ++j) total += g->seq_start[p->nodes[j] + 1] - g->seq_start[p->nodes[j]];
    char* text = (char*)malloc(total + 1);
    uint32_t* tnode = (uint32_t*)malloc((total + 1) * sizeof(uint32_t));
    uint64_t* toff = (uint64_t*)malloc((total + 1) * sizeof(uint64_t));
    if (!text || !tnode || !toff) { free(text); free(tnode); free(toff); return -1; }
    uint64_t t = 0;
    for (uint64_t j = nb; j < ne; ++j) {
      uint32_t v = p->nodes[j];
      uint64_t len = g->seq_start[v + 1] - g->seq_start[v];
      uint64_t from = (j == nb && p->head_off) ? p->head_off[pi] : 0;
      uint64_t to = len - ((j + 1 == ne && p->tail_trim) ? p->tail_trim[pi] : 0);
      if (to > len) to = 0;  /* trim larger than the node: nothing left */
      for (uint64_t o = from; o < to; ++o) {
        text[t] = g->seq[g->seq_start[v] + o];
        tnode[t] = v;
        toff[t] = o;
        ++t;
      }
    }
    for (uint64_t i = 0; i + k <= t; ++i) {
      uint64_t km;
      if (pack_kmer(text + i, k, &km)) emit_hits(c, km, tnode[i], toff[i]);
    }
    free(text); free(tnode); free(toff);
  }
  return c->err;
}

int psi_oracle_seeds_on_paths(const psi_oracle_graph* g, const psi_oracle_paths* p,
                              const psi_oracle_reads* r, unsigned k, unsigned d,
                              psi_oracle_result* out)
{
  memset(out, 0, sizeof(*out));
  if (k == 0 || k > 32) return -2;
  seed_hash sh;
  if (sh_build(&sh, r, k, d) != 0) return -1;
  tuple_vec tv = { 0, 0, 0 };
  emit_ctx c = { g, &sh, k, &tv, 0 };
  int rc = on_paths_collect(&c, p);
  sh_free(&sh);
  if (rc != 0) { free(tv.data); return rc; }
  tv_finish(&tv, out);
  return 0;
}

/* ------------------------------------------------------------ off paths -- */

/* traverser_bfs.hpp:114-161.  A state walks the label of its current node from
 * cpos.offset, one base per go_down; 'N' kills it (:124); at the node end it
 * moves to the first successor and a copy is pushed for every other one
 * (:146-160, link type ignored, forward strand only); no successor kills it
 * (:140-143); at depth k every read-seed occurrence is reported with the START
 * locus as (node, offset) (:96-110).  The reference additionally drops a state
 * as soon as its prefix is no prefix of any read seed -- a pruning that cannot
 * change which depth-k states match.  Depth-first recursion enumerates the
 * same set of walks as the reference's breadth-first state list. */
typedef void (*walk_cb)(void* user, uint64_t km, uint32_t start_node, uint64_t start_off);

typedef struct {
  const psi_oracle_graph* g;
  unsigned k;
  walk_cb cb;
  void* user;
  uint32_t start_node;
  uint64_t start_off;
} walk_ctx;

static void walk_rec(const walk_ctx* w, uint32_t node, uint64_t off, unsigned depth, uint64_t km)
{
  const psi_oracle_graph* g = w->g;
  uint64_t beg = g->seq_start[node], len = g->seq_start[node + 1] - beg;
  while (off < len && depth < w->k) {
    unsigned c = base_code(g->seq[beg + off]);
    if (c > 3) return;
    km = (km << 2) | c;
    ++off; ++depth;
  }
  if (depth == w->k) { w->cb(w->user, km, w->start_node, w->start_off); return; }
  for (uint64_t e = g->row_ptr[node]; e < g->row_ptr[node + 1]; ++e)
    walk_rec(w, g->col[e], 0, depth, km);
}

static void walks_from(const psi_oracle_graph* g, unsigned k, uint32_t node, uint64_t off,
                       walk_cb cb, void* user)
{
  walk_ctx w = { g, k, cb, user, node, off };
  walk_rec(&w, node, off, 0, 0);
}

static void off_cb(void* user, uint64_t km, uint32_t node, uint64_t off)
{
  emit_hits((emit_ctx*)user, km, node, off);
}

static int off_paths_collect(emit_ctx* c, uint64_t n_loci, const uint32_t* ln, const uint32_t* lo)
{
  for (uint64_t i = 0; i < n_loci; ++i) {
    uint64_t len = c->g->seq_start[ln[i] + 1] - c->g->seq_start[ln[i]];
    if (lo[i] >= len) continue;
    walks_from(c->g, c->k, ln[i], lo[i], off_cb, c);
  }
  return c->err;
}

int psi_oracle_seeds_off_paths(const psi_oracle_graph* g, uint64_t n_loci,
                               const uint32_t* locus_node, const uint32_t* locus_off,
                               const psi_oracle_reads* r, unsigned k, unsigned d,
                               psi_oracle_result* out)
{
  memset(out, 0, sizeof(*out));
  if (k == 0 || k > 32) return -2;
  seed_hash sh;
  if (sh_build(&sh, r, k, d) != 0) return -1;
  tuple_vec tv = { 0, 0, 0 };
  emit_ctx c = { g, &sh, k, &tv, 0 };
  int rc = off_paths_collect(&c, n_loci, locus_node, locus_off);
  sh_free(&sh);
  if (rc != 0) { free(tv.data); return rc; }
  tv_finish(&tv, out);
  return 0;
}

/* seed_finder.hpp:1724-1732 */
int psi_oracle_seeds_all(const psi_oracle_graph* g, const psi_oracle_paths* p,
                         uint64_t n_loci, const uint32_t* locus_node, const uint32_t* locus_off,
                         const psi_oracle_reads* r, unsigned k, unsigned d,
                         psi_oracle_result* out)
{
  memset(out, 0, sizeof(*out));
  if (k == 0 || k > 32) return -2;
  seed_hash sh;
  if (sh_build(&sh, r, k, d) != 0) return -1;
  tuple_vec tv = { 0, 0, 0 };
  emit_ctx c = { g, &sh, k, &tv, 0 };
  int rc = 0;
  if (p && p->n_paths) rc = on_paths_collect(&c, p);
  if (rc == 0) rc = off_paths_collect(&c, n_loci, locus_node, locus_off);
  sh_free(&sh);
  if (rc != 0) { free(tv.data); return rc; }
  tv_finish(&tv, out);
  return 0;
}

int psi_oracle_seeds_closed_form(const psi_oracle_graph* g, const psi_oracle_reads* r,
                                 unsigned k, unsigned d, psi_oracle_result* out)
{
  memset(out, 0, sizeof(*out));
  if (k == 0 || k > 32) return -2;
  seed_hash sh;
  if (sh_build(&sh, r, k, d) != 0) return -1;
  tuple_vec tv = { 0, 0, 0 };
  emit_ctx c = { g, &sh, k, &tv, 0 };
  for (uint64_t v = 0; v < g->n_nodes; ++v) {
    uint64_t len = g->seq_start[v + 1] - g->seq_start[v];
    for (uint64_t o = 0; o < len; ++o) walks_from(g, k, (uint32_t)v, o, off_cb, &c);
  }
  sh_free(&sh);
  if (c.err != 0) { free(tv.data); return c.err; }
  tv_finish(&tv, out);
  return 0;
}

/* -------------------------------------------------------- starting loci -- */

typedef struct { uint64_t km; uint64_t pos; } kp_rec;

static int kp_cmp(const void* a, const void* b)
{
  const kp_rec* x = (const kp_rec*)a; const kp_rec* y = (const kp_rec*)b;
  if (x->km != y->km) return x->km < y->km ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return 0;
}

typedef struct { const kp_rec* v; uint64_t n; const psi_oracle_graph* g; int uncovered; } cover_ctx;

static void cover_cb(void* user, uint64_t km, uint32_t node, uint64_t off)
{
  cover_ctx* c = (cover_ctx*)user;
  if (c->uncovered) return;
  kp_rec key = { km, c->g->seq_start[node] + off };
  if (!bsearch(&key, c->v, c->n, sizeof(kp_rec), kp_cmp)) c->uncovered = 1;
}

int psi_oracle_uncovered_loci(const psi_oracle_graph* g, const psi_oracle_paths* p,
                              unsigned k, unsigned step,
                              uint32_t** locus_node, uint32_t** locus_off, uint64_t* n_loci)
{
  *locus_node = NULL; *locus_off = NULL; *n_loci = 0;
  if (k == 0 || k > 32) return -2;
  if (step == 0) step = 1;
  /* all (k-mer, global position) pairs on the paths */
  uint64_t cap = 1024, n = 0;
  kp_rec* v = (kp_rec*)malloc(cap * sizeof(kp_rec));
  for (uint64_t pi = 0; p && pi < p->n_paths; ++pi) {
    uint64_t nb = p->path_ptr[pi], ne = p->path_ptr[pi + 1];
    uint64_t total = 0;
    for (uint64_t j = nb; j < ne; ++j) total += g->seq_start[p->nodes[j] + 1] - g->seq_start[p->nodes[j]];
    char* text = (char*)malloc(total + 1);
    uint64_t* tpos = (uint64_t*)malloc((total + 1) * sizeof(uint64_t));
    uint64_t t = 0;
    for (uint64_t j = nb; j < ne; ++j) {
      uint32_t nd = p->nodes[j];
      uint64_t len = g->seq_start[nd + 1] - g->seq_start[nd];
      uint64_t from = (j == nb && p->head_off) ? p->head_off[pi] : 0;
      uint64_t to = len - ((j + 1 == ne && p->tail_trim) ? p->tail_trim[pi] : 0);
      if (to > len) to = 0;
      for (uint64_t o = from; o < to; ++o) { text[t] = g->seq[g->seq_start[nd] + o]; tpos[t] = g->seq_start[nd] + o; ++t; }
    }
    for (uint64_t i = 0; i + k <= t; ++i) {
      uint64_t km;
      if (!pack_kmer(text + i, k, &km)) continue;
      if (n == cap) { cap *= 2; v = (kp_rec*)realloc(v, cap * sizeof(kp_rec)); }
      v[n].km = km; v[n].pos = tpos[i]; ++n;
    }
    free(text); free(tpos);
  }
  qsort(v, n, sizeof(kp_rec), kp_cmp);

  uint64_t lc = 1024, ln = 0;
  uint32_t* on = (uint32_t*)malloc(lc * sizeof(uint32_t));
  uint32_t* oo = (uint32_t*)malloc(lc * sizeof(uint32_t));
  for (uint64_t nd = 0; nd < g->n_nodes; ++nd) {
    uint64_t len = g->seq_start[nd + 1] - g->seq_start[nd];
    for (uint64_t o = 0; o < len; o += step) {
      cover_ctx c = { v, n, g, 0 };
      walks_from(g, k, (uint32_t)nd, o, cover_cb, &c);
      if (c.uncovered) {
        if (ln == lc) { lc *= 2; on = (uint32_t*)realloc(on, lc * sizeof(uint32_t)); oo = (uint32_t*)realloc(oo, lc * sizeof(uint32_t)); }
        on[ln] = (uint32_t)nd; oo[ln] = (uint32_t)o; ++ln;
      }
    }
  }
  free(v);
  *locus_node = on; *locus_off = oo; *n_loci = ln;
  return 0;
}

void psi_oracle_free(void* p) { free(p); }
