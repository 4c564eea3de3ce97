/*
 * oracle/psi_oracle.h -- TEST INFRASTRUCTURE (checker), never the product path.
 *
 * A plain-C, single-threaded CPU restatement of the reference's fully-sensitive
 * seed finding (cartoonist/psi).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library.  The product
 * (libpsi_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against
 *   - the reference's own known-answer tests
 *       test/src/test_indexiter.cpp:182-185,230-233,282-285,335-338,394-397
 *       test/src/test_traverser.cpp:81-96
 *       test/src/test_sequence.cpp:1272-1427 (seeding + id/offset mapping)
 *   - golden seed sets produced by the UNMODIFIED reference compiled in this
 *     container (oracle/_ref/psi_ref_driver, see oracle/Makefile and
 *     tests/golden/make_golden.py), committed under tests/golden/.
 *
 * All graph arrays are rank-indexed (0-based ranks).  Sequences are ASCII;
 * A/C/G/T (either case) are bases, every other byte behaves like 'N' and never
 * matches (traverser_bfs.hpp:124, index_iter.hpp:831-832).
 */
#ifndef PSI_ORACLE_H
#define PSI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t        n_nodes;
  const uint64_t* seq_start; /* n_nodes+1 offsets into seq */
  const char*     seq;       /* concatenated node labels */
  const uint64_t* row_ptr;   /* n_nodes+1 offsets into col */
  const uint32_t* col;       /* successor ranks, in the graph's out-edge order */
  const uint64_t* node_id;   /* id reported for each rank */
} psi_oracle_graph;

typedef struct {
  uint64_t        n_reads;
  const uint64_t* read_ptr;  /* n_reads+1 offsets into bases */
  const char*     bases;
  uint64_t        first_read_id; /* Records::rec_offset, sequence.hpp:1616 */
} psi_oracle_reads;

typedef struct {
  uint64_t        n_paths;
  const uint64_t* path_ptr;  /* n_paths+1 offsets into nodes */
  const uint32_t* nodes;     /* node ranks */
  const uint32_t* head_off;  /* bases trimmed from the first node (patches) */
  const uint32_t* tail_trim; /* bases trimmed from the last node */
} psi_oracle_paths;

/* Result: n tuples (read_id, read_offset, node_id, node_offset), 4 x u64 each,
 * sorted ascending and unique.  Free with psi_oracle_free(). */
typedef struct {
  uint64_t* tuples;
  uint64_t  n;
  uint64_t  n_raw;  /* emitted before dedup */
} psi_oracle_result;

/* sequence.hpp:1688-1718 + SeedMap (sequence.hpp:1148-1220): seeds of every
 * read at offsets 0,d,2d,... while off + k <= len.  Returns the number of
 * seeds; fills seed_read_id / seed_offset (either may be NULL) up to cap. */
uint64_t psi_oracle_seeding(const psi_oracle_reads* reads, unsigned k, unsigned d,
                            uint64_t* seed_read_id, uint64_t* seed_offset, uint64_t cap);

/* index_iter.hpp:808-852 on two plain string sets: sum over ACGT-only k-mers w
 * of occ1(w) * occ2(w), all (overlapping) positions of both sets. */
uint64_t psi_oracle_kmer_exact_matches(const char* const* set1, uint64_t n1,
                                       const char* const* set2, uint64_t n2, unsigned k);

/* sequence.hpp:1639-1667: next lexicographic k-mer over ACGT at position pos
 * (continuous=false semantics).  Returns the smallest modified position or
 * UINT64_MAX when the k-mer space is exhausted. */
uint64_t psi_oracle_increment_kmer(char* kmer, uint64_t len, uint64_t pos);

/* seed_finder.hpp:1426-1457 / index_iter.hpp:728-746: every occurrence of
 * every read seed on the indexed paths. */
int psi_oracle_seeds_on_paths(const psi_oracle_graph* g, const psi_oracle_paths* p,
                              const psi_oracle_reads* r, unsigned k, unsigned d,
                              psi_oracle_result* out);

/* seed_finder.hpp:1703-1722 / traverser_bfs.hpp:71-161: all k-walks from the
 * given starting loci against the read seeds. */
int psi_oracle_seeds_off_paths(const psi_oracle_graph* g, uint64_t n_loci,
                               const uint32_t* locus_node, const uint32_t* locus_off,
                               const psi_oracle_reads* r, unsigned k, unsigned d,
                               psi_oracle_result* out);

/* seed_finder.hpp:1724-1732: union of the two. */
int psi_oracle_seeds_all(const psi_oracle_graph* g, const psi_oracle_paths* p,
                         uint64_t n_loci, const uint32_t* locus_node, const uint32_t* locus_off,
                         const psi_oracle_reads* r, unsigned k, unsigned d,
                         psi_oracle_result* out);

/* SURVEY 8a-2 closed form: off-paths with every (node, offset) as a locus. */
int psi_oracle_seeds_closed_form(const psi_oracle_graph* g, const psi_oracle_reads* r,
                                 unsigned k, unsigned d, psi_oracle_result* out);

/* Starting loci by definition (seed_finder.hpp:1481-1541, k-mer level): (v,o),
 * o % step == 0, such that some k-walk from (v,o) spells an ACGT-only k-mer w
 * with (w,(v,o)) not occurring on any path.  Outputs malloc'd arrays. */
int psi_oracle_uncovered_loci(const psi_oracle_graph* g, const psi_oracle_paths* p,
                              unsigned k, unsigned step,
                              uint32_t** locus_node, uint32_t** locus_off, uint64_t* n_loci);

void psi_oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
