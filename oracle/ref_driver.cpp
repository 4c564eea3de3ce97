/**
 *  oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 *  A thin driver around the UNMODIFIED reference headers under /root/reference.
 *  It does what `find_seeds` in the reference CLI does (src/psikt.cpp:83-212):
 *
 *      SeedFinder(graph, k) -> create_path_index(n, patched, context, step)
 *      loop { readRecords; get_seeds; index_reads; seeds_all(write_callback) }
 *
 *  minus spdlog / protobuf / SeqAn ArgumentParser (not buildable in this image),
 *  with GFA as the graph format (gum::util::load(graph, file, sort=true),
 *  src/psikt.cpp:249-251).  It is compiled in place by oracle/Makefile into
 *  oracle/_ref/psi_ref_driver; no reference source is copied into this repo.
 *
 *  Outputs:
 *    --out FILE    canonical seed set: sorted unique tuples
 *                  (read_id, read_offset, coordinate(external) node id, node_offset),
 *                  4 x u64 LE each.
 *    --raw FILE    the reference CLI's byte format (src/psikt.cpp:172-181):
 *                  per hit node_id(internal), node_offset, read_id, read_offset
 *                  as native size_t, in emission order (multiset, path dependent).
 *    --nodes FILE  per rank (1..n): internal id, coordinate id, label length (3 x u64).
 *    --loci FILE   starting loci: internal id, offset (2 x u64).
 *    --paths FILE  the paths the reference picked and indexed (SeedFinder::pick_paths, seed_finder.hpp:1138-1167):
 *                  u64 n_paths, then per path u64 n_nodes, u64 head offset (bases cut from the first node),
 *                  u64 tail trim (bases cut from the last node), u64 text length, the node RANKS (0-based, u64
 *                  each; flattened by include/psi_b200_gum.hpp) and the path's forward text padded to 8 bytes.
 *    --mems FILE   MEM mode (SeedFinder::seeds_on_paths(sequence, callback) -> find_mems, seed_finder.hpp:1459-1479,
 *                  index_iter.hpp:854-906) on every read of --fastq instead of the k-mer path: per hit 6 x u64
 *                  read ordinal, read offset, match length, gocc, node id (coordinate), node offset, in emission order
 *                  (raw: one hit per occurrence in the path text).  -E sets max_mem, -r the gocc threshold.
 *    --dist FILE   paired-end distance verification (SeedFinder::create_distance_index + verify_distance,
 *                  seed_finder.hpp:1193-1317; DiVerG's range sparse boolean matrix powers on Kokkos' Serial backend) for
 *                  the insert sizes -m DMIN -M DMAX: pseudo-random pairs of loci at most DMAX + 8 characters apart in
 *                  rank order, per pair 5 x u64: rank of v (0-based), offset, rank of u, offset, answer (0 / 1).
 *    --save-index P  SeedFinder::serialize_path_index(P, step): the files psikt -I writes (P, P_paths, P_loci_e<step>l<k>).
 *  and one JSON line on stdout with counts and timings.
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <array>
#include <algorithm>
#include <chrono>
#include <functional>
#include <unordered_set>

#include <psi/seed_finder.hpp>
#include <psi/sequence.hpp>
#include <gum/graph.hpp>
#include <gum/io_utils.hpp>
#include <kseq++/seqio.hpp>

#include "../include/psi_b200_gum.hpp"

using namespace psi;

struct Args {
  std::string gfa, fastq, out, raw, nodes, loci, paths, mems, save_index, dist;
  unsigned dmin = 0, dmax = 0;
  unsigned long dist_queries = 20000;
  unsigned max_mem = 0;
  unsigned k = 0, d = 0, n = 0, context = 0, step = 1, gocc = 0;
  unsigned long chunk = 0, first_read = 0, max_reads = 0;
  bool patched = true;
  bool on_only = false, off_only = false;
};

static double now_s()
{
  using namespace std::chrono;
  return duration< double >( steady_clock::now().time_since_epoch() ).count();
}

static void write_u64s( std::FILE* f, const uint64_t* p, size_t n )
{
  if ( std::fwrite( p, sizeof( uint64_t ), n, f ) != n ) { std::perror( "fwrite" ); std::exit( 3 ); }
}

int main( int argc, char** argv )
{
  Args a;
  for ( int i = 1; i < argc; ++i ) {
    std::string s = argv[i];
    auto next = [&]() -> std::string { if ( i + 1 >= argc ) { std::fprintf( stderr, "missing value for %s\n", s.c_str() ); std::exit( 2 ); } return argv[++i]; };
    if ( s == "--gfa" ) a.gfa = next();
    else if ( s == "--fastq" ) a.fastq = next();
    else if ( s == "--out" ) a.out = next();
    else if ( s == "--raw" ) a.raw = next();
    else if ( s == "--nodes" ) a.nodes = next();
    else if ( s == "--loci" ) a.loci = next();
    else if ( s == "--paths" ) a.paths = next();
    else if ( s == "--mems" ) a.mems = next();
    else if ( s == "--dist" ) a.dist = next();
    else if ( s == "-m" ) a.dmin = std::stoul( next() );
    else if ( s == "-M" ) a.dmax = std::stoul( next() );
    else if ( s == "--dist-queries" ) a.dist_queries = std::stoul( next() );
    else if ( s == "--save-index" ) a.save_index = next();   /* SeedFinder::serialize_path_index (psikt -I) */
    else if ( s == "-E" ) a.max_mem = std::stoul( next() );
    else if ( s == "-k" ) a.k = std::stoul( next() );
    else if ( s == "-d" ) a.d = std::stoul( next() );
    else if ( s == "-n" ) a.n = std::stoul( next() );
    else if ( s == "-t" ) a.context = std::stoul( next() );
    else if ( s == "-e" ) a.step = std::stoul( next() );
    else if ( s == "-r" ) a.gocc = std::stoul( next() );   /* seed genome occurrence count threshold (src/psikt.cpp -r) */
    else if ( s == "-c" ) a.chunk = std::stoul( next() );
    else if ( s == "--first-read" ) a.first_read = std::stoul( next() );
    else if ( s == "--max-reads" ) a.max_reads = std::stoul( next() );
    else if ( s == "-P" ) a.patched = false;
    else if ( s == "--on-only" ) a.on_only = true;
    else if ( s == "--off-only" ) a.off_only = true;
    else { std::fprintf( stderr, "unknown option %s\n", s.c_str() ); return 2; }
  }
  if ( a.gfa.empty() || a.k == 0 ) {
    std::fprintf( stderr, "usage: psi_ref_driver --gfa G [--fastq F] -k K [-d D] [-n N] [-P] [-t CTX] [-e STEP] [-c CHUNK]\n"
                          "       [--first-read I] [--max-reads M] [--out F] [--raw F] [--nodes F] [--loci F]\n" );
    return 2;
  }
  if ( a.d == 0 ) a.d = a.k;  /* src/psikt.cpp:469 */

  typedef gum::SeqGraph< gum::Succinct > graph_type;
  typedef Dna5QStringSet<> readsstringset_type;
  typedef SeedFinderTraits< gum::Succinct, readsstringset_type, seqan2::IndexWotd<>, InMemory > traits_type;
  typedef SeedFinder< NoStats, traits_type > finder_type;
  typedef typename finder_type::traverser_type traverser_type;

  double t0 = now_s();
  graph_type graph;
  gum::util::load( graph, a.gfa, true );
  double t_load = now_s() - t0;

  if ( !a.nodes.empty() ) {
    std::FILE* f = std::fopen( a.nodes.c_str(), "wb" );
    graph.for_each_node( [&]( auto rank, auto id ) {
      uint64_t rec[3] = { (uint64_t)id, (uint64_t)graph.coordinate_id( id ), (uint64_t)graph.node_length( id ) };
      (void)rank;
      write_u64s( f, rec, 3 );
      return true;
    } );
    std::fclose( f );
  }

  finder_type finder( graph, a.k, a.gocc, a.max_mem );
  t0 = now_s();
  if ( a.n != 0 ) {
    finder.create_path_index( a.n, a.patched, a.context, a.step );
  }
  double t_index = now_s() - t0;

  if ( !a.dist.empty() ) {
    finder.create_distance_index( a.dmin, a.dmax, PerComponent{} );
    /* global character order of every node start (gum::util::id_to_charorder): rank order */
    std::vector< uint64_t > start, ids, lens;
    uint64_t total = 0;
    graph.for_each_node( [&]( auto rank, auto id ) {
      (void)rank;
      start.push_back( total ); ids.push_back( (uint64_t)id ); lens.push_back( (uint64_t)graph.node_length( id ) );
      total += graph.node_length( id );
      return true;
    } );
    auto locate = [&]( uint64_t pos, uint64_t& r, uint64_t& o ) {
      r = std::upper_bound( start.begin(), start.end(), pos ) - start.begin() - 1;
      o = pos - start[ r ];
    };
    std::FILE* f = std::fopen( a.dist.c_str(), "wb" );
    uint64_t x = 0x9e3779b97f4a7c15ull;
    auto rnd = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for ( unsigned long q = 0; q < a.dist_queries; ++q ) {
      uint64_t pa = rnd() % total;
      uint64_t span = 3ull * a.dmax + 24;            /* rank order interleaves alleles: look well beyond dmax */
      uint64_t pb = pa + rnd() % span;
      if ( q % 7 == 0 && pa >= 5 ) pb = pa - rnd() % 5;   /* some targets behind the source */
      if ( pb >= total ) pb = total - 1;
      uint64_t rv, ov, ru, ou;
      locate( pa, rv, ov );
      locate( pb, ru, ou );
      bool ok = finder.verify_distance( (graph_type::id_type)ids[ rv ], ov, (graph_type::id_type)ids[ ru ], ou );
      uint64_t row[5] = { rv, ov, ru, ou, ok ? 1ull : 0ull };
      write_u64s( f, row, 5 );
    }
    std::fclose( f );
  }

  if ( !a.save_index.empty() && !finder.serialize_path_index( a.save_index, a.step ) ) {
    std::fprintf( stderr, "could not save the path index to %s\n", a.save_index.c_str() );
    return 3;
  }

  if ( !a.loci.empty() ) {
    std::FILE* f = std::fopen( a.loci.c_str(), "wb" );
    for ( auto const& l : finder.get_starting_loci() ) {
      uint64_t rec[2] = { (uint64_t)l.node_id(), (uint64_t)l.offset() };
      write_u64s( f, rec, 2 );
    }
    std::fclose( f );
  }

  if ( !a.paths.empty() ) {
    /* the flat arrays come from the adapter the C-ABI ships for the reference's types (include/psi_b200_gum.hpp):
     * dumping through it pins the adapter's rank / head / tail arithmetic to the reference's own path text */
    std::FILE* f = std::fopen( a.paths.c_str(), "wb" );
    auto const& pset = finder.get_pindex().get_paths_set();
    psi_b200::FlatPaths fp = psi_b200::flatten_paths( graph, pset );
    uint64_t n = fp.head_off.size();
    write_u64s( f, &n, 1 );
    uint64_t pi = 0;
    for ( auto it = pset.begin(); it != pset.end(); ++it, ++pi ) {
      std::string text = sequence( *it, Forward() );
      uint64_t nn = fp.path_ptr[ pi + 1 ] - fp.path_ptr[ pi ];
      uint64_t hdr[4] = { nn, fp.head_off[ pi ], fp.tail_trim[ pi ], (uint64_t)text.size() };
      write_u64s( f, hdr, 4 );
      for ( uint64_t e = fp.path_ptr[ pi ]; e < fp.path_ptr[ pi + 1 ]; ++e ) {
        uint64_t r = fp.nodes[ e ];
        write_u64s( f, &r, 1 );
      }
      text.resize( ( text.size() + 7 ) / 8 * 8, '\0' );
      if ( !text.empty() && std::fwrite( text.data(), 1, text.size(), f ) != text.size() ) { std::perror( "fwrite" ); std::exit( 3 ); }
    }
    std::fclose( f );
  }

  unsigned long long raw_on = 0, raw_off = 0;
  std::vector< std::array< uint64_t, 4 > > hits;
  std::FILE* rawf = a.raw.empty() ? nullptr : std::fopen( a.raw.c_str(), "wb" );
  bool keep = !a.out.empty();
  bool in_off = false;

  std::function< void( typename traverser_type::output_type const& ) > cb =
    [&]( typename traverser_type::output_type const& h ) {
      if ( in_off ) ++raw_off; else ++raw_on;
      if ( rawf ) {
        uint64_t rec[4] = { (uint64_t)h.node_id, (uint64_t)h.node_offset, (uint64_t)h.read_id, (uint64_t)h.read_offset };
        write_u64s( rawf, rec, 4 );
      }
      if ( keep ) {
        hits.push_back( { (uint64_t)h.read_id, (uint64_t)h.read_offset,
                          (uint64_t)graph.coordinate_id( h.node_id ), (uint64_t)h.node_offset } );
      }
    };

  double t_seeding = 0, t_on = 0, t_off = 0;
  unsigned long n_reads = 0, n_seeds = 0, n_chunks = 0;
  if ( !a.mems.empty() && !a.fastq.empty() ) {
    std::FILE* f = std::fopen( a.mems.c_str(), "wb" );
    klibpp::SeqStreamIn iss( a.fastq.c_str() );
    klibpp::KSeq rec;
    uint64_t ordinal = 0;
    t0 = now_s();
    while ( ( a.max_reads == 0 || ordinal < a.max_reads ) && ( iss >> rec ) ) {
      std::string const& sequence = rec.seq;
      finder.seeds_on_paths( sequence, [&]( typename traverser_type::output_type const& h ) {
        uint64_t out6[6] = { ordinal, (uint64_t)h.read_offset, (uint64_t)h.match_len, (uint64_t)h.gocc,
                             (uint64_t)graph.coordinate_id( h.node_id ), (uint64_t)h.node_offset };
        write_u64s( f, out6, 6 );
        ++raw_on;
      } );
      ++ordinal;
    }
    t_on = now_s() - t0;
    n_reads = ordinal;
    std::fclose( f );
  }
  else if ( !a.fastq.empty() ) {
    klibpp::SeqStreamIn iss( a.fastq.c_str() );
    if ( !iss ) { std::fprintf( stderr, "cannot open %s\n", a.fastq.c_str() ); return 2; }
    /* Skip to the first read of this shard; read ids stay global (sequence.hpp:1616). */
    {
      klibpp::KSeq rec;
      for ( unsigned long s = 0; s < a.first_read; ++s ) if ( !( iss >> rec ) ) break;
    }
    auto chunk = finder.create_readrecord();
    auto seeds = finder.create_readrecord();
    auto traverser = finder.create_traverser();
    while ( true ) {
      unsigned long want = a.chunk;
      if ( a.max_reads != 0 ) {
        unsigned long left = a.max_reads - n_reads;
        if ( left == 0 ) break;
        if ( want == 0 || want > left ) want = left;
      }
      if ( !readRecords( chunk, iss, want ) ) break;
      n_reads += length( chunk );
      ++n_chunks;
      t0 = now_s();
      finder.get_seeds( seeds, chunk, a.d );
      auto seeds_index = finder.index_reads( seeds );
      t_seeding += now_s() - t0;
      n_seeds += length( seeds );
      if ( !a.off_only ) {
        t0 = now_s();
        in_off = false;
        finder.seeds_on_paths( seeds, seeds_index, cb );
        t_on += now_s() - t0;
      }
      if ( !a.on_only ) {
        t0 = now_s();
        in_off = true;
        finder.setup_traverser( traverser, seeds, seeds_index );
        finder.seeds_off_paths( traverser, cb );
        t_off += now_s() - t0;
      }
    }
  }
  if ( rawf ) std::fclose( rawf );

  size_t n_unique = 0;
  if ( keep ) {
    std::sort( hits.begin(), hits.end() );
    hits.erase( std::unique( hits.begin(), hits.end() ), hits.end() );
    n_unique = hits.size();
    std::FILE* f = std::fopen( a.out.c_str(), "wb" );
    if ( !f ) { std::perror( "open out" ); return 3; }
    if ( !hits.empty() ) write_u64s( f, &hits[0][0], hits.size() * 4 );
    std::fclose( f );
  }

  std::printf( "{\"impl\": \"reference\", \"nodes\": %lu, \"loci\": %lu, \"loci_nodes\": %lu, \"reads\": %lu, \"query_seeds\": %lu, "
               "\"chunks\": %lu, \"raw_on\": %llu, \"raw_off\": %llu, \"unique\": %lu, "
               "\"t_load\": %.6f, \"t_index\": %.6f, \"t_seeding\": %.6f, \"t_on\": %.6f, \"t_off\": %.6f}\n",
               (unsigned long)graph.get_node_count(), (unsigned long)finder.get_starting_loci().size(),
               (unsigned long)finder.get_nof_uniq_nodes(), n_reads, n_seeds, n_chunks, raw_on, raw_off,
               (unsigned long)n_unique, t_load, t_index, t_seeding, t_on, t_off );
  return 0;
}
