// psi/seed_finder.hpp -- psi::SeedFinder over libpsi_b200 (B200, sm_100a).
//
// Mirror of the reference's API object for fully-sensitive seed finding
// (reference include/psi/seed_finder.hpp:728-1788) for the calls its CLI makes
// (src/psikt.cpp:83-212): same class and method names, argument meaning, error
// behaviour and callback contract; underneath, every method is a call into the
// C-ABI of include/psi_b200.h -- no algorithm lives in this header and there is
// no CPU fallback (construction throws when no sm_100 device is usable).
//
//   reference method (seed_finder.hpp)            C-ABI call
//   SeedFinder(graph, k, ...)          :930-942    psi_b200_create + psi_b200_set_graph
//   pick_paths                         :1138-1167  psi_b200_pick_paths
//   index_paths                        :1169-1176  psi_b200_set_paths
//   add_uncovered_loci                 :1481-1541  psi_b200_find_loci / get_loci
//   load_path_index/serialize_..       :1372-1413  pathset + loci files (own paths format, reference loci format)
//   get_seeds + index_reads            :1089-1109  psi_b200_submit_chunk_packed
//   seeds_on_paths/off_paths/all       :1426-1457,1703-1743  psi_b200_seeds_all_async + psi_b200_fetch_dense_async + psi_b200_wait
//
// Threads: like the reference's finder (seed_finder.hpp:386-399: one const finder, several threads, one chunk
// each) the per-chunk calls may be made from several host threads at once; every thread drives its own pipeline
// (psi_b200_fork: own stream and buffers over the one resident index).  The index-building calls are not
// thread safe, as in the reference.
//
// Differences that are deliberate: callbacks see every hit of the
// seed SET exactly once (the reference repeats a locus once per covering path
// and per walk); Seed<>::gocc is filled in MEM mode only; the distance index and
// approximate matching are outside this build and throw std::runtime_error when
// requested.  A gocc threshold (-r) and a step size (-e) make the reference's seed
// set depend on its (randomly) picked paths and on the loci it derived from them:
// with the same paths and loci (psi_b200_set_paths / set_loci, a shared loci file)
// this build gives the same set (tests/golden/loci, written by the reference).
#ifndef PSI_B200_PSI_SEED_FINDER_HPP
#define PSI_B200_PSI_SEED_FINDER_HPP

#include <array>
#include <atomic>
#include <unistd.h>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/psi_b200.h"
#include "graph.hpp"
#include "seed.hpp"
#include "sequence.hpp"
#include "stats.hpp"

namespace seqan2 {
template <typename T = void> struct IndexWotd {};
template <typename T = void> struct IndexEsa {};
}  // namespace seqan2

namespace psi {

struct BFS {};
struct DFS {};
struct Whole {};
struct PerComponent {};
template <typename, typename> class ExactMatching {};

// Process-wide totals the CLI reports (reference traverser_base.hpp:108-244).
struct TraverserStats {
  static std::atomic<unsigned long long>& seeds_off_paths() { static std::atomic<unsigned long long> v{ 0 }; return v; }
  static std::atomic<unsigned long long>& nof_godowns() { static std::atomic<unsigned long long> v{ 0 }; return v; }
  static unsigned long long get_total_seeds_off_paths() { return seeds_off_paths().load(); }
  // "godown" = one index descent per base in the reference; here one hash probe per completed k-walk
  static unsigned long long get_total_nof_godowns() { return nof_godowns().load(); }
};

template <typename TGraphSpec = gum::Succinct, typename TReadsStringSet = Dna5QStringSet<>,
          typename TReadsIndexSpec = seqan2::IndexWotd<>, typename TPathsStringSetSpec = InMemory,
          typename TStrategy = BFS, template <typename, typename> class TMatchingTraits = ExactMatching>
struct SeedFinderTraits {
  typedef SeqGraph<TGraphSpec> graph_type;
  typedef TReadsStringSet stringset_type;
  typedef TReadsIndexSpec indexspec_type;
  typedef TPathsStringSetSpec pathstrsetspec_type;
  typedef TStrategy strategy_type;
};

template <typename TStatsSpec = NoStats, typename TTraits = SeedFinderTraits<>>
class SeedFinder {
 public:
  typedef TTraits traits_type;
  typedef typename traits_type::graph_type graph_type;
  typedef typename graph_type::id_type id_type;
  typedef typename graph_type::offset_type offset_type;
  typedef typename graph_type::rank_type rank_type;
  typedef Records<typename traits_type::stringset_type> readsrecord_type;
  typedef Seed<> output_type;
  typedef std::function<void(output_type const&)> callback_type;
  // bulk form: n records of 4 x u64 {node_id, node_offset, read_id, read_offset}, the CLI's byte layout
  typedef std::function<void(const uint64_t*, uint64_t)> records_callback_type;

  // The device read index of one submitted chunk (the reference builds a lazy
  // suffix tree here, seed_finder.hpp:1089-1097; the GPU hash is built lazily too).
  struct readsindex_type {
    uint64_t serial = 0;
  };

  // The graph traverser of seeds_off_paths (reference Traverser<...>::Type); the
  // walk itself runs on the device, this object carries the per-thread pipeline.
  class traverser_type {
   public:
    typedef Seed<> output_type;
    typedef TraverserStats stats_type;
    typedef typename traits_type::stringset_type stringset_type;
    typedef readsindex_type index_type;
    traverser_type(const graph_type* g, unsigned len) : graph_ptr(g), seed_len(len) {}
    void set_reads(const readsrecord_type* r) { reads = r; }
    void set_reads_index(readsindex_type* i) { reads_index = i; }
    const graph_type* graph_ptr;
    unsigned seed_len;
    const readsrecord_type* reads = nullptr;
    readsindex_type* reads_index = nullptr;
  };

  // Timers under the reference's names (seed_finder.hpp:427-456): "<finder id><name><thread id>".
  class stats_type {
   public:
    typedef Timer timer_type;
    explicit stats_type(const SeedFinder* f)
    {
      char buf[32];
      std::snprintf(buf, sizeof buf, "%08llx", (unsigned long long)(reinterpret_cast<uintptr_t>(f) & 0xffffffffull));
      id = buf;
    }
    Timer timeit_ts(const std::string& name) const { return Timer(id + name + get_thread_id()); }
    Timer::period_type get_timer(const std::string& name) const { return Timer::get(id + name); }
    Timer::period_type get_timer(const std::string& name, const std::string& tid) const { return Timer::get(id + name + tid); }
    void set_timer(const std::string& name, double seconds) const { Timer::set(id + name + get_thread_id(), seconds); }
    // SIGUSR1 (reference seed_finder.hpp:275-338): a one-line progress report on stderr; async-signal-safe (write only)
    static std::atomic<unsigned long long>& chunks_done() { static std::atomic<unsigned long long> v{ 0 }; return v; }
    static std::atomic<unsigned long long>& reads_done() { static std::atomic<unsigned long long> v{ 0 }; return v; }
    static void signal_handler(int)
    {
      char buf[128];
      int n = std::snprintf(buf, sizeof buf, "psi-b200: %llu chunks, %llu reads, %llu seeds off paths so far\n", chunks_done().load(),
                            reads_done().load(), TraverserStats::seeds_off_paths().load());
      if (n > 0) { ssize_t w = ::write(2, buf, (size_t)n); (void)w; }
    }
    std::string id;
  };

  /* ---- lifecycle ---- */
  SeedFinder(const graph_type& g, unsigned int len, unsigned int gocc_thr = 0, unsigned int mxmem = 0,
             unsigned char mismatches = 0, int device = -1)
      : graph_ptr(&g), seed_len(len), seed_mismatches(mismatches),
        gocc_threshold(gocc_thr != 0 ? gocc_thr : UINT_MAX), max_mem(mxmem != 0 ? mxmem : UINT_MAX),
        stats_ptr(std::make_unique<stats_type>(this))
  {
    if (mismatches != 0) throw std::runtime_error("approximate seed matching is not implemented");
    if (device < 0) {
      const char* e = std::getenv("PSI_B200_DEVICE");
      device = e ? std::atoi(e) : 0;
    }
    if (psi_b200_create(device, len, &ctx) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
    try {
      upload_graph();
      // seeds_on_paths skips k-mers with more occurrences in the path text (index_iter.hpp:842-848); applied to the index
      if (gocc_thr != 0) check(psi_b200_set_option(ctx, "gocc_threshold", gocc_thr));
    }
    catch (...) { psi_b200_destroy(ctx); ctx = nullptr; throw; }
  }
  SeedFinder(const SeedFinder&) = delete;
  SeedFinder& operator=(const SeedFinder&) = delete;
  ~SeedFinder()
  {
    reset_pipes();
    if (pathset) psi_b200_pathset_free(pathset);
    if (ctx) psi_b200_destroy(ctx);
  }

  /* ---- paired-end distance verification (reference seed_finder.hpp:1193-1317) ----
   * The reference materialises DiVerG's boolean matrix of all locus pairs whose distance lies in [dmin, dmax] and looks
   * a pair up; here the device keeps the same relation per node (psi_b200_create_distance_index) and answers queries in
   * bulk.  Whole / PerComponent only choose how the reference assembles its matrix: the relation is the same. */
  template <typename TDIndexMode = PerComponent>
  void create_distance_index(unsigned int dmin, unsigned int dmax, TDIndexMode = {},
                             std::function<void(std::string const&)> info = nullptr,
                             std::function<void(std::string const&)> = nullptr)
  {
    if (dmin == 0 || dmax < dmin) return;   // not constructible (seed_finder.hpp:1198)
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("index-distances");
    if (info) info(std::is_same<TDIndexMode, Whole>::value ? "Constructing distance index for the whole graph..."
                                                            : "Constructing distance index on the device...");
    reset_pipes();
    check(psi_b200_create_distance_index(ctx, dmin, dmax));
    d = std::make_pair(dmin, dmax);
  }

  static std::string get_distance_index_path(std::string prefix, unsigned int dmin, unsigned int dmax)
  {
    return prefix + "_dist_mat_m" + std::to_string(dmin) + "M" + std::to_string(dmax);
  }

  // The device rebuilds its rows from the graph in milliseconds, so the file only records WHICH index was saved (window
  // and graph); `.b200` keeps it apart from the reference's serialised matrix of the same prefix, which is not read.
  bool save_distance_index(std::string prefix) const
  {
    if (d.first == 0) return true;   // empty distance index (seed_finder.hpp:1270)
    std::ofstream ofs(get_distance_index_path(prefix, d.first, d.second) + ".b200", std::ofstream::binary);
    if (!ofs) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("save-dindex");
    const psi_b200_graph_view& gv = graph_ptr->view();
    const uint64_t hdr[6] = { DIST_MAGIC, d.first, d.second, gv.n_nodes, gv.n_bases, graph_checksum() };
    ofs.write(reinterpret_cast<const char*>(hdr), sizeof hdr);
    return (bool)ofs;
  }

  bool open_distance_index(std::string prefix, unsigned int dmin = 0, unsigned int dmax = 0)
  {
    if (dmax == 0) dmax = dmin;
    std::ifstream ifs(get_distance_index_path(prefix, dmin, dmax) + ".b200", std::ifstream::binary);
    if (!ifs) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("load-dindex");
    uint64_t hdr[6] = { 0, 0, 0, 0, 0, 0 };
    ifs.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    const psi_b200_graph_view& gv = graph_ptr->view();
    if (!ifs || hdr[0] != DIST_MAGIC || hdr[1] != dmin || hdr[2] != dmax || hdr[3] != gv.n_nodes || hdr[4] != gv.n_bases ||
        hdr[5] != graph_checksum())
      return false;
    create_distance_index(dmin, dmax);
    return d.first == dmin && d.second == dmax && dmin != 0;
  }

  bool verify_distance(id_type v, offset_type o, id_type u, offset_type p) const
  {
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("query-dindex");
    const uint32_t q[4] = { (uint32_t)(graph_ptr->id_to_rank(v) - 1), clamp32(o), (uint32_t)(graph_ptr->id_to_rank(u) - 1), clamp32(p) };
    uint8_t ok = 0;
    Pipe& pp = pipe();
    pcheck(pp, psi_b200_verify_distance(pp.ctx, 1, q, &ok, 0));
    return ok != 0;
  }

  // Bulk form -- what a paired-end caller uses: n quadruples {v, o, u, p} (node ids as in verify_distance), one answer
  // each, one device pass for all of them.
  std::vector<uint8_t> verify_distances(const std::vector<std::array<uint64_t, 4>>& ends) const
  {
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("query-dindex");
    std::vector<uint32_t> q(4 * ends.size());
    for (size_t i = 0; i < ends.size(); ++i) {
      q[4 * i] = (uint32_t)(graph_ptr->id_to_rank((id_type)ends[i][0]) - 1);
      q[4 * i + 1] = clamp32(ends[i][1]);
      q[4 * i + 2] = (uint32_t)(graph_ptr->id_to_rank((id_type)ends[i][2]) - 1);
      q[4 * i + 3] = clamp32(ends[i][3]);
    }
    std::vector<uint8_t> ok(ends.size());
    Pipe& pp = pipe();
    pcheck(pp, psi_b200_verify_distance(pp.ctx, ends.size(), q.data(), ok.data(), 0));
    return ok;
  }

  /* ---- source compatibility with the reference's lifecycle knobs (seed_finder.hpp:862-874,1066-1081) ----
   * The reference initialises / finalises Kokkos through these (its distance index runs on Kokkos); nothing here needs
   * them, they are kept so that code written against the reference compiles unchanged. */
  static std::atomic_bool& get_kokkos_handling_status() { static std::atomic_bool enabled{ true }; return enabled; }
  static void set_kokkos_handling_status(bool value = true) { get_kokkos_handling_status().store(value); }
  void set_as_finaliser() { finaliser = true; }
  void unset_as_finaliser() { finaliser = false; }
  bool is_finaliser() const { return finaliser; }
  // seeds_on_paths skips k-mers with more occurrences in the path text: takes effect when the paths are indexed next
  void set_gocc_threshold(unsigned int value)
  {
    gocc_threshold = value != 0 ? value : UINT_MAX;
    check(psi_b200_set_option(ctx, "gocc_threshold", value));
  }

  /* ---- accessors ---- */
  const graph_type* get_graph_ptr() const { return graph_ptr; }
  const std::vector<Position<>>& get_starting_loci() const { return starting_loci; }
  unsigned int get_seed_len() const { return seed_len; }
  unsigned char get_seed_mismatches() const { return seed_mismatches; }
  unsigned int get_gocc_threshold() const { return gocc_threshold; }
  const stats_type& get_stats() const { return *stats_ptr; }
  unsigned int get_context() const { return context_; }
  psi_b200_ctx* get_device_context() const { return ctx; }
  psi_b200_counters_t get_counters() const
  {
    psi_b200_counters_t c;
    check(psi_b200_counters(ctx, &c));
    return c;
  }

  /* ---- mutators ---- */
  void set_starting_loci(std::vector<Position<>> loci)
  {
    starting_loci = std::move(loci);
    push_loci();
  }
  void add_start(const Position<>& locus) { starting_loci.push_back(locus); loci_dirty = true; }
  void add_start(id_type node_id, offset_type offset) { add_start(Position<>(node_id, offset)); }

  readsrecord_type create_readrecord() const { return readsrecord_type(); }
  traverser_type create_traverser() const { return traverser_type(graph_ptr, seed_len); }
  void setup_traverser(traverser_type& t, readsrecord_type const& reads, readsindex_type& idx) const
  {
    t.set_reads(&reads);
    t.set_reads_index(&idx);
  }

  /* ---- path index ---- */
  void pick_paths(unsigned int n, bool patched = true, unsigned int context = 0,
                  std::function<void(std::string const&, int)> callback = nullptr,
                  std::function<void(std::string const&)> info = nullptr,
                  std::function<void(std::string const&)> warn = nullptr)
  {
    if (n == 0) return;
    if (graph_ptr->get_path_count() == 0) throw std::runtime_error("no reference path found in the input graph");
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("pick-paths");
    context_ = set_context(context, patched, info, warn);
    if (callback)
      graph_ptr->for_each_path([&](rank_type, id_type pid) {
        for (unsigned i = 0; i < n; ++i) callback(graph_ptr->path_name(pid), (int)i + 1);
        return true;
      });
    if (pathset) { psi_b200_pathset_free(pathset); pathset = nullptr; }
    uint64_t seed = 0x9e3779b97f4a7c15ull;
    if (const char* e = std::getenv("PSI_B200_PATH_SEED")) seed = std::strtoull(e, nullptr, 10);
    if (psi_b200_pick_paths(graph_ptr->handle(), n, patched ? 1 : 0, context_, seed, &pathset) != PSI_B200_OK)
      throw std::runtime_error(psi_b200_global_error());
  }

  void index_paths()
  {
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("index-paths");
    if (!pathset) return;
    psi_b200_pathset_view v;
    if (psi_b200_pathset_get_view(pathset, &v) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
    reset_pipes();     // the index is rebuilt: pipelines forked from the old one go
    check(psi_b200_set_paths(ctx, v.n_paths, v.path_ptr, v.nodes, v.head_off, v.tail_trim));
    remember_paths(v.n_paths, v.path_ptr, v.nodes, v.head_off, v.tail_trim);
    has_index = v.n_paths != 0;
  }

  void add_uncovered_loci(unsigned int step = 1)
  {
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("find-uncovered");
    uint64_t n = 0;
    reset_pipes();
    check(psi_b200_find_loci(ctx, step, &n));
    pull_loci(n);
  }

  template <typename TDIndexMode = PerComponent>
  void create_path_index(unsigned int n, bool patched = true, unsigned int context = 0, unsigned int step_size = 1,
                         unsigned int dmin = 0, unsigned int dmax = 0, TDIndexMode = {},
                         std::function<void(std::string const&)> info = nullptr,
                         std::function<void(std::string const&)> warn = nullptr)
  {
    std::function<void(std::string const&, int)> progress = nullptr;
    if (info) progress = [&info](std::string const& name, int i) {
      info("Selecting path " + std::to_string(i) + " of region " + name + "...");
    };
    pick_paths(n, patched, context, progress, info, warn);
    if (info) info("Indexing the selected paths...");
    index_paths();
    if (info) info("Detecting uncovered loci...");
    if (step_size > 1 && warn)
      warn("Step size > 1: the starting loci are sampled on this build's own lattice (node offsets divisible by the step), "
           "not on the reference's per-walk lattice; share a loci file (-I) for identical seed sets.");
    add_uncovered_loci(step_size);
    if (info) info("Constructing distance index for pair distance queries...");
    create_distance_index(dmin, dmax);
  }

  // files: <prefix>_paths.b200 (picked paths; the device index is rebuilt from them at load) and
  // <prefix>_loci_e<step>l<k> in the reference's own byte format (seed_finder.hpp:884-892,1659-1679).
  // The paths file names the graph it belongs to (node count, base count, a checksum of the labels); a file written
  // for another graph, or a damaged one, is refused (load returns false and the caller rebuilds, as the reference
  // does for a missing piece, seed_finder.hpp:1396-1413).
  // the paths alone (reference serialize_path_index_only / load_path_index_only, seed_finder.hpp:1357-1393)
  bool serialize_path_index_only(std::string const& fpath)
  {
    if (fpath.empty() || !pathset) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("save-pindex");
    psi_b200_pathset_view v;
    if (psi_b200_pathset_get_view(pathset, &v) != PSI_B200_OK) return false;
    std::ofstream ofs(fpath + "_paths.b200", std::ofstream::binary);
    if (!ofs) return false;
    const psi_b200_graph_view& gv = graph_ptr->view();
    const uint64_t hdr[8] = { PATHS_MAGIC, context_, v.n_paths, v.path_ptr[v.n_paths], gv.n_nodes, gv.n_bases, graph_checksum(), 0 };
    ofs.write(reinterpret_cast<const char*>(hdr), sizeof hdr);
    ofs.write(reinterpret_cast<const char*>(v.path_ptr), (v.n_paths + 1) * sizeof(uint64_t));
    // node ranks are stable for a given graph file (the load order is deterministic)
    ofs.write(reinterpret_cast<const char*>(v.nodes), v.path_ptr[v.n_paths] * sizeof(uint32_t));
    ofs.write(reinterpret_cast<const char*>(v.head_off), v.n_paths * sizeof(uint32_t));
    ofs.write(reinterpret_cast<const char*>(v.tail_trim), v.n_paths * sizeof(uint32_t));
    return (bool)ofs;
  }

  bool serialize_path_index(std::string const& fpath, unsigned int step_size = 1)
  {
    // like the reference (seed_finder.hpp:1372-1378): paths, starting loci, distance index
    return serialize_path_index_only(fpath) && save_starts(fpath, seed_len, step_size) && save_distance_index(fpath);
  }

  // `<prefix>_paths.b200` of this build or, failing that, `<prefix>_paths` as the REFERENCE's psikt -I wrote it (coordinate
  // node ids as an sdsl enc_vector, left / right trims; pathindex.hpp:313-332, path_base.hpp:552-560): the device index is
  // rebuilt from the paths, the reference's serialised FM-index (`<prefix>`) is not needed.
  bool load_path_index_only(std::string const& fpath, unsigned int context = 0)
  {
    if (fpath.empty()) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("load-pindex");
    std::ifstream ifs(fpath + "_paths.b200", std::ifstream::binary | std::ifstream::ate);
    if (!ifs) return load_reference_paths(fpath, context);
    const uint64_t file_bytes = (uint64_t)ifs.tellg();
    ifs.seekg(0);
    uint64_t hdr[8];
    ifs.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    if (!ifs || hdr[0] != PATHS_MAGIC) return false;
    const psi_b200_graph_view& gv = graph_ptr->view();
    const uint64_t n_paths = hdr[2], n_entries = hdr[3];
    // the file must be exactly as long as its header says, and belong to this graph
    if (n_paths > file_bytes || n_entries > file_bytes) return false;
    if (file_bytes != sizeof hdr + (n_paths + 1) * 8 + n_entries * 4 + n_paths * 8) return false;
    if (hdr[4] != gv.n_nodes || hdr[5] != gv.n_bases || hdr[6] != graph_checksum()) return false;
    std::vector<uint64_t> path_ptr(n_paths + 1);
    std::vector<uint32_t> nodes(n_entries), head(n_paths), tail(n_paths);
    ifs.read(reinterpret_cast<char*>(path_ptr.data()), path_ptr.size() * sizeof(uint64_t));
    ifs.read(reinterpret_cast<char*>(nodes.data()), nodes.size() * sizeof(uint32_t));
    ifs.read(reinterpret_cast<char*>(head.data()), head.size() * sizeof(uint32_t));
    ifs.read(reinterpret_cast<char*>(tail.data()), tail.size() * sizeof(uint32_t));
    if (!ifs || path_ptr.front() != 0 || path_ptr.back() != nodes.size()) return false;
    for (uint64_t p = 0; p < n_paths; ++p) if (path_ptr[p + 1] < path_ptr[p]) return false;
    for (uint32_t r : nodes) if (r >= gv.n_nodes) return false;
    for (uint64_t p = 0; p < n_paths; ++p) {
      if (path_ptr[p + 1] == path_ptr[p]) { if (head[p] || tail[p]) return false; continue; }
      const uint32_t first = nodes[path_ptr[p]], last = nodes[path_ptr[p + 1] - 1];
      if (head[p] > gv.seq_start[first + 1] - gv.seq_start[first]) return false;
      if (tail[p] > gv.seq_start[last + 1] - gv.seq_start[last]) return false;
    }
    context_ = context ? context : (unsigned)hdr[1];
    reset_pipes();
    check(psi_b200_set_paths(ctx, n_paths, path_ptr.data(), nodes.data(), head.data(), tail.data()));
    remember_paths(n_paths, path_ptr.data(), nodes.data(), head.data(), tail.data());
    has_index = n_paths != 0;
    return true;
  }

  bool load_path_index(std::string const& fpath, unsigned int context = 0, unsigned int step_size = 1,
                       unsigned int dmin = 0, unsigned int dmax = 0)
  {
    if (!load_path_index_only(fpath, context)) return false;      // seed_finder.hpp:1396-1413
    if (!open_starts(fpath, seed_len, step_size)) {
      add_uncovered_loci(step_size);
      save_starts(fpath, seed_len, step_size);
    }
    if (!open_distance_index(fpath, dmin, dmax)) {
      create_distance_index(dmin, dmax);
      save_distance_index(fpath);
    }
    return true;
  }

  bool load_reference_paths(std::string const& fpath, unsigned int context)
  {
    psi_b200_pathset* ps = nullptr;
    uint64_t file_context = 0;
    if (psi_b200_pathset_load_reference(graph_ptr->handle(), (fpath + "_paths").c_str(), &ps, &file_context) != PSI_B200_OK) return false;
    // the reference refuses an index saved with another context (pathindex.hpp:289-291)
    if (context != 0 && file_context != context) { psi_b200_pathset_free(ps); return false; }
    if (pathset) psi_b200_pathset_free(pathset);
    pathset = ps;
    context_ = (unsigned)file_context;
    index_paths();
    return true;
  }

  static std::string get_sloci_filepath(const std::string& prefix, unsigned int seed_len, unsigned int step_size)
  {
    return prefix + "_loci_e" + std::to_string(step_size) + "l" + std::to_string(seed_len);
  }

  bool open_starts(const std::string& prefix, unsigned int len, unsigned int step_size)
  {
    std::ifstream ifs(get_sloci_filepath(prefix, len, step_size), std::ifstream::binary | std::ifstream::ate);
    if (!ifs) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("load-starts");
    const uint64_t file_bytes = (uint64_t)ifs.tellg();
    ifs.seekg(0);
    uint64_t n = 0;
    ifs.read(reinterpret_cast<char*>(&n), sizeof n);
    if (!ifs || file_bytes < 8 || n != (file_bytes - 8) / 16 || (file_bytes - 8) % 16) return false;   // the count must be the file's
    std::unordered_map<int64_t, int64_t> by_coord;   // external id -> internal id
    const psi_b200_graph_view& v = graph_ptr->view();
    for (uint64_t r = 0; r < v.n_nodes; ++r) by_coord.emplace((int64_t)v.coord_id[r], (int64_t)v.internal_id[r]);
    std::vector<Position<>> loci;
    loci.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
      int64_t id; uint64_t off;
      ifs.read(reinterpret_cast<char*>(&id), sizeof id);
      ifs.read(reinterpret_cast<char*>(&off), sizeof off);
      if (!ifs) return false;
      auto it = by_coord.find(id);
      if (it == by_coord.end()) return false;
      if (off >= graph_ptr->node_length(it->second)) return false;      // not a position of this graph
      loci.emplace_back(it->second, off);
    }
    set_starting_loci(std::move(loci));
    return true;
  }

  bool save_starts(const std::string& prefix, unsigned int len, unsigned int step_size)
  {
    std::ofstream ofs(get_sloci_filepath(prefix, len, step_size), std::ofstream::binary);
    if (!ofs) return false;
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("save-starts");
    const uint64_t n = starting_loci.size();
    ofs.write(reinterpret_cast<const char*>(&n), sizeof n);
    for (const auto& l : starting_loci) {   // external (coordinate) ids on disk, as the reference writes them
      const int64_t id = graph_ptr->coordinate_id(l.node_id());
      const uint64_t off = l.offset();
      ofs.write(reinterpret_cast<const char*>(&id), sizeof id);
      ofs.write(reinterpret_cast<const char*>(&off), sizeof off);
    }
    return (bool)ofs;
  }

  std::size_t get_nof_uniq_nodes()
  {
    std::unordered_set<id_type> set;
    for (const auto& l : starting_loci) set.insert(l.node_id());
    return set.size();
  }

  /* ---- per chunk ---- */
  // seeding(): seeds at offsets 0, d, 2d, ... of every read (sequence.hpp:1688-1745).  Uploads the chunk's 2-bit
  // words to this thread's pipeline (asynchronously: `reads` must stay valid until the seeds have been found);
  // the k-mers are cut on the device; `seeds` becomes a view of `reads`.
  template <typename T>
  void get_seeds(readsrecord_type& seeds, readsrecord_type const& reads, T distance) const
  {
    auto timer = stats_ptr->timeit_ts("seeding");
    Pipe& p = pipe();
    if (p.pending) throw std::runtime_error("a chunk is still in flight on this thread (seeds_all_wait first)");
    seeds = reads;
    seeds.seed_len = seed_len;
    seeds.distance = distance ? (unsigned)distance : seed_len;
    seeds.serial = ++p.serial;
    seeds.pipe = &p;
    psi_b200_packed_chunk ch{};
    ch.n_reads = reads.n_reads;
    ch.first_read_id = reads.rec_offset;
    ch.n_bases = reads.n_bases;
    ch.read_len = reads.read_len;
    ch.read_ptr = reads.read_len ? nullptr : reads.read_ptr;
    ch.words = reads.words;
    ch.exc = reads.exc;
    ch.n_exc = reads.n_exc;
    pcheck(p, psi_b200_submit_chunk_packed(p.ctx, &ch, seeds.distance, 0));
    timer.stop();
  }

  readsindex_type index_reads(readsrecord_type const& seeds) const
  {
    [[maybe_unused]] auto timer = stats_ptr->timeit_ts("index-reads");
    return readsindex_type{ seeds.serial };
  }

  void seeds_on_paths(readsrecord_type const& seeds, readsindex_type& reads_index, callback_type callback) const
  {
    if (context_ != 0 && context_ < seed_len) throw std::runtime_error("seed length should not be larger than context size");
    if (!has_index) return;
    run(seeds, reads_index, PSI_B200_ON_PATHS, "seeds-on-paths", callback, nullptr);
  }

  // MEM mode (seed_finder.hpp:1459-1479 -> find_mems, index_iter.hpp:854-906): the greedy scan of one sequence
  // against the text of the indexed paths; hits carry match_len and gocc (occurrences in the path text), read_id is
  // not set by the reference either.  The device index for it (suffix table of the path text) is built on first use.
  template <typename TString>
  void seeds_on_paths(TString const& sequence, callback_type callback) const
  {
    if (!has_index) return;
    ensure_mem_index();
    auto timer = stats_ptr->timeit_ts("query-paths");
    Pipe& p = pipe();
    if (p.pending) throw std::runtime_error("a chunk is still in flight on this thread (seeds_all_wait first)");
    const std::string text(sequence.begin(), sequence.end());
    const uint64_t rp[2] = { 0, text.size() };
    ++p.serial;          // whatever get_seeds submitted to this pipeline is gone
    pcheck(p, psi_b200_submit_chunk(p.ctx, 1, rp, text.data(), 0, 0));
    uint64_t n = 0;
    pcheck(p, psi_b200_find_mems(p.ctx, max_mem == UINT_MAX ? 0u : max_mem, &n));
    std::vector<uint64_t> rec(6 * n);
    if (n) pcheck(p, psi_b200_fetch_mems(p.ctx, rec.data(), n, &n));
    Seed<> hit;
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t* x = rec.data() + 6 * i;
      hit.node_id = x[0]; hit.node_offset = x[1]; hit.read_id = 0; hit.read_offset = x[3]; hit.match_len = x[4]; hit.gocc = x[5];
      if (callback) callback(hit);
    }
  }

  void seeds_off_paths(traverser_type& traverser, callback_type callback) const
  {
    if (!traverser.reads || !traverser.reads_index) throw std::runtime_error("traverser is not set up (setup_traverser)");
    sync_loci();
    if (starting_loci.empty()) return;
    run(*traverser.reads, *traverser.reads_index, PSI_B200_OFF_PATHS, "seeds-off-path", callback, nullptr);
  }

  void seeds_all(readsrecord_type const& seeds, readsindex_type& reads_index, traverser_type& traverser,
                 callback_type callback) const
  {
    if (context_ != 0 && context_ < seed_len) throw std::runtime_error("seed length should not be larger than context size");
    setup_traverser(traverser, seeds, reads_index);
    sync_loci();
    run(seeds, reads_index, PSI_B200_ALL, nullptr, callback, nullptr);
  }

  // callback1 receives the hits found on the indexed paths, callback2 the rest (seed_finder.hpp:1734-1743).
  void seeds_all(readsrecord_type const& seeds, readsindex_type& reads_index, traverser_type& traverser,
                 callback_type callback1, callback_type callback2) const
  {
    if (context_ != 0 && context_ < seed_len) throw std::runtime_error("seed length should not be larger than context size");
    setup_traverser(traverser, seeds, reads_index);
    sync_loci();
    run(seeds, reads_index, PSI_B200_ALL, nullptr, callback1, callback2 ? callback2 : [](output_type const&) {});
  }

  // The same in two halves (not in the reference): seeds_all_begin queues the chunk's kernels and the copy of
  // their results and returns; the caller does host work (parse the next chunk); seeds_all_wait / _wait_records
  // blocks for the results and delivers them.  One chunk in flight per host thread.
  void seeds_all_begin(readsrecord_type const& seeds, readsindex_type& reads_index) const
  {
    if (context_ != 0 && context_ < seed_len) throw std::runtime_error("seed length should not be larger than context size");
    sync_loci();
    begin(pipe(), seeds, reads_index, PSI_B200_ALL);
  }
  void seeds_all_wait(callback_type callback1, callback_type callback2 = nullptr) const { deliver(pipe(), callback1, callback2, nullptr); }
  // bulk form: n records of 4 x u64 {node_id, node_offset, read_id, read_offset}, the CLI's byte layout
  void seeds_all_wait_records(records_callback_type cb) const { deliver(pipe(), nullptr, nullptr, cb); }

  // Bulk variant used by the CLI: one call per chunk with all records in the output byte layout.
  void seeds_all_records(readsrecord_type const& seeds, readsindex_type& reads_index, records_callback_type cb) const
  {
    seeds_all_begin(seeds, reads_index);
    seeds_all_wait_records(cb);
  }

 private:
  static constexpr uint64_t PATHS_MAGIC = 0x3230736874617042ull;  // "Bpaths02"
  static constexpr uint64_t DIST_MAGIC = 0x3130747369644242ull;   // "BBdist01"
  static uint32_t clamp32(uint64_t x) { return x > 0xffffffffull ? 0xffffffffu : (uint32_t)x; }

  // One pipeline per host thread: a fork of the finder's context (own stream and device buffers over the shared
  // resident index) plus the pinned host buffers its results arrive in.
  struct Pipe {
    psi_b200_ctx* ctx = nullptr;
    uint64_t serial = 0;
    void* dense = nullptr;         // ids[n_seeds] u32, then offsets[n_seeds] u16/u32
    uint64_t dense_cap = 0;        // seeds
    uint32_t* extra = nullptr;     // 4 x u32 per further hit of a multi-locus seed
    uint64_t extra_cap = 0;
    uint64_t* records = nullptr;   // per-hit records (routes without dense results) / the CLI's 4 x u64 staging
    uint64_t rec_cap = 0;
    // chunk in flight
    bool pending = false, dense_mode = false, compact = false, dense5 = false;
    unsigned d5_off_bits = 0;
    unsigned flags = 0;
    readsrecord_type seeds;
    ~Pipe()
    {
      if (dense) psi_b200_host_free(dense);
      if (extra) psi_b200_host_free(extra);
      if (records) psi_b200_host_free(records);
      if (ctx) psi_b200_destroy(ctx);
    }
  };

  Pipe& pipe() const
  {
    std::lock_guard<std::mutex> lk(pipes_mutex);
    auto& slot = pipes[std::this_thread::get_id()];
    if (!slot) {
      auto p = std::make_unique<Pipe>();
      if (psi_b200_fork(ctx, &p->ctx) != PSI_B200_OK) throw std::runtime_error(psi_b200_last_error(ctx));
      slot = std::move(p);
    }
    return *slot;
  }

  // Pipelines share the index: before it is rebuilt they go (index building is single-threaded by contract).
  void reset_pipes() const
  {
    std::lock_guard<std::mutex> lk(pipes_mutex);
    pipes.clear();
  }

  void check(int rc) const
  {
    if (rc != PSI_B200_OK) throw std::runtime_error(psi_b200_last_error(ctx));
  }
  static void pcheck(const Pipe& p, int rc)
  {
    if (rc != PSI_B200_OK) throw std::runtime_error(psi_b200_last_error(p.ctx));
  }

  void upload_graph()
  {
    const psi_b200_graph_view& v = graph_ptr->view();
    if (graph_ptr->get_node_count() == 0) throw std::runtime_error("empty graph");
    check(psi_b200_set_graph(ctx, v.n_nodes, v.seq_start, v.seq, v.row_ptr, v.col, v.internal_id));
    max_node_id = 0;
    for (uint64_t i = 0; i < v.n_nodes; ++i) if (v.internal_id[i] > max_node_id) max_node_id = v.internal_id[i];
    unsigned ob = 4;
    check(psi_b200_dense_layout(ctx, &ob));
    dense_off_bytes = ob;
  }

  void remember_paths(uint64_t n_paths, const uint64_t* path_ptr, const uint32_t* nodes, const uint32_t* head, const uint32_t* tail)
  {
    std::lock_guard<std::mutex> lk(mem_mutex);
    mem_built = false;
    kept_ptr.assign(path_ptr, path_ptr + n_paths + 1);
    kept_nodes.assign(nodes, nodes + path_ptr[n_paths]);
    kept_head.assign(head, head + n_paths);
    kept_tail.assign(tail, tail + n_paths);
  }

  // the MEM index is built when the first MEM query comes (before any other thread has a pipeline of its own:
  // building is single-threaded by contract, like the rest of the index construction)
  void ensure_mem_index() const
  {
    std::lock_guard<std::mutex> lk(mem_mutex);
    if (mem_built) return;
    reset_pipes();
    check(psi_b200_build_mem_index(ctx, kept_head.size(), kept_ptr.data(), kept_nodes.data(), kept_head.data(), kept_tail.data()));
    mem_built = true;
  }

  // FNV-1a over the node labels and their boundaries: names the graph a saved path set belongs to
  uint64_t graph_checksum() const
  {
    if (checksum_valid) return checksum;
    const psi_b200_graph_view& v = graph_ptr->view();
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&h](const unsigned char* p, size_t n) { for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; } };
    mix(reinterpret_cast<const unsigned char*>(v.seq), v.n_bases);
    mix(reinterpret_cast<const unsigned char*>(v.seq_start), (v.n_nodes + 1) * sizeof(uint64_t));
    checksum = h;
    checksum_valid = true;
    return h;
  }

  unsigned int set_context(unsigned int context, bool patched, std::function<void(std::string const&)> = nullptr,
                           std::function<void(std::string const&)> warn = nullptr)
  {
    if (!patched) context = 0;
    if (patched && context == 0) {
      if (warn) warn("The context size cannot be zero for patching. Assuming the seed length as the context size...");
      context = seed_len;
    }
    return context;
  }

  void pull_loci(uint64_t n)
  {
    std::vector<uint32_t> node(n), off(n);
    uint64_t got = 0;
    check(psi_b200_get_loci(ctx, node.data(), off.data(), n, &got));
    const psi_b200_graph_view& v = graph_ptr->view();
    starting_loci.clear();
    starting_loci.reserve(n);
    for (uint64_t i = 0; i < n; ++i) starting_loci.emplace_back((int64_t)v.internal_id[node[i]], off[i]);
    loci_dirty = false;
  }

  void push_loci() const
  {
    std::vector<uint32_t> node(starting_loci.size()), off(starting_loci.size());
    for (size_t i = 0; i < starting_loci.size(); ++i) {
      node[i] = (uint32_t)(graph_ptr->id_to_rank(starting_loci[i].node_id()) - 1);
      off[i] = (uint32_t)starting_loci[i].offset();
    }
    reset_pipes();
    check(psi_b200_set_loci(ctx, node.size(), node.data(), off.data()));
    loci_dirty = false;
  }
  void sync_loci() const { if (loci_dirty) push_loci(); }

  static void grow(void*& p, uint64_t& cap, uint64_t want, uint64_t unit)
  {
    if (want <= cap) return;
    if (p) psi_b200_host_free(p);
    p = nullptr;
    cap = 0;
    const uint64_t n = want + want / 4 + 1024;
    if (psi_b200_host_alloc(&p, n * unit) != PSI_B200_OK) throw std::runtime_error(psi_b200_global_error());
    cap = n;
  }

  // Queue the chunk's step and the copy of its results on the pipeline.
  void begin(Pipe& p, readsrecord_type const& seeds, readsindex_type const& idx, unsigned flags) const
  {
    if (p.pending) throw std::runtime_error("a chunk is still in flight on this thread (seeds_all_wait first)");
    if (seeds.serial == 0 || seeds.pipe != &p || seeds.serial != p.serial || idx.serial != seeds.serial)
      throw std::runtime_error("seeds do not belong to the chunk last given to get_seeds()");
    const bool ids32 = max_node_id < 0xffffffffull && seeds.rec_offset + seeds.n_reads <= 0x100000000ull;
    p.flags = flags;
    p.seeds = seeds;
    p.dense_mode = false;
    p.compact = ids32;
    // dense per-seed results whenever the index serves them (walk mode and 64-bit ids keep per-hit records)
    if (ids32 && dense_state != 2) {
      // 5 bytes per seed when the graph's loci fit 39 bits (PSI_B200_DENSE5), else u32 id + u16 / u32 offset
      unsigned d5_bits = 0;
      int d5 = 0;
      if (psi_b200_dense5_layout(p.ctx, &d5_bits, &d5) != PSI_B200_OK) d5 = 0;
      const int rc = psi_b200_seeds_all_async(p.ctx, flags | (d5 ? PSI_B200_DENSE5 : PSI_B200_DENSE));
      if (rc == PSI_B200_OK) {
        dense_state = 1;
        p.dense_mode = true;
        p.dense5 = d5 != 0;
        p.d5_off_bits = d5_bits;
        uint64_t n_seeds = 0;
        if (seeds.read_len) n_seeds = seeds.n_reads * (seeds.read_len >= seed_len ? (seeds.read_len - seed_len) / seeds.distance + 1 : 0);
        else for (uint64_t r = 0; r < seeds.n_reads; ++r) n_seeds += seeds.seeds_of_read(r);
        grow(p.dense, p.dense_cap, n_seeds, 8);
        void* e = p.extra;
        grow(e, p.extra_cap, n_seeds / 16 + 4096, 16);
        p.extra = static_cast<uint32_t*>(e);
        pcheck(p, psi_b200_fetch_dense_async(p.ctx, p.dense, p.dense_cap, p.extra, p.extra_cap));
        p.pending = true;
        return;
      }
      if (rc != PSI_B200_ERR_ARG && rc != PSI_B200_ERR_STATE) pcheck(p, rc);
      if (dense_state == 0 && (flags & PSI_B200_ALL) == PSI_B200_ALL) dense_state = 2;   // this index never serves them
    }
    pcheck(p, psi_b200_seeds_all_async(p.ctx, ids32 ? flags | PSI_B200_COMPACT : flags));
    p.pending = true;
  }

  // Wait for the chunk in flight and hand its hits to the callbacks (cb2 != null: off-path hits go there) or, as
  // 4 x u64 records, to rcb.
  void deliver(Pipe& p, const callback_type& cb1, const callback_type& cb2, const records_callback_type& rcb) const
  {
    if (!p.pending) throw std::runtime_error("no chunk in flight on this thread (seeds_all_begin first)");
    p.pending = false;
    uint64_t n_hits = 0;
    pcheck(p, psi_b200_wait(p.ctx, &n_hits));
    account(p);
    const readsrecord_type& seeds = p.seeds;
    Seed<> hit;
    hit.match_len = seed_len;
    hit.gocc = 0;
    uint64_t* rec = nullptr;
    uint64_t n_rec = 0;
    if (rcb) {
      void* r = p.records;
      grow(r, p.rec_cap, n_hits, 32);
      p.records = static_cast<uint64_t*>(r);
      rec = p.records;
    }
    auto emit = [&](uint64_t node, uint64_t noff, uint64_t read, uint64_t roff, bool off_path) {
      if (rec) {
        uint64_t* o = rec + 4 * n_rec++;
        o[0] = node; o[1] = noff; o[2] = read; o[3] = roff;
        return;
      }
      hit.node_id = node; hit.node_offset = noff; hit.read_id = read; hit.read_offset = roff;
      const callback_type& cb = (off_path && cb2) ? cb2 : cb1;
      if (cb) cb(hit);
    };
    if (p.dense_mode) {
      uint64_t n_seeds = 0, n_extra = 0;
      pcheck(p, psi_b200_dense_counts(p.ctx, &n_seeds, &n_extra));
      if (n_extra > p.extra_cap) {       // rare: more multi-locus hits than the buffer was sized for
        void* e = p.extra;
        grow(e, p.extra_cap, n_extra, 16);
        p.extra = static_cast<uint32_t*>(e);
        pcheck(p, psi_b200_fetch_dense(p.ctx, p.dense, p.dense_cap, p.extra, p.extra_cap, &n_seeds, &n_extra));
      }
      const uint32_t* ids = static_cast<const uint32_t*>(p.dense);
      const uint16_t* off16 = reinterpret_cast<const uint16_t*>(ids + n_seeds);
      const uint32_t* off32 = ids + n_seeds;
      const uint8_t* hi8 = reinterpret_cast<const uint8_t*>(ids + n_seeds);
      const uint64_t d5_mask = (1ull << p.d5_off_bits) - 1ull;
      uint64_t s = 0;
      for (uint64_t r = 0; r < seeds.n_reads; ++r) {
        const uint64_t cnt = seeds.read_len ? (seeds.read_len >= seed_len ? (seeds.read_len - seed_len) / seeds.distance + 1 : 0)
                                            : seeds.seeds_of_read(r);
        for (uint64_t j = 0; j < cnt; ++j, ++s) {
          if (p.dense5) {
            const uint32_t hi = hi8[s];
            if (ids[s] == 0xffffffffu && hi == 0xffu) continue;
            const uint64_t e = (uint64_t)ids[s] | ((uint64_t)(hi & 0x7fu) << 32);
            emit(e >> p.d5_off_bits, e & d5_mask, seeds.rec_offset + r, j * seeds.distance, (hi >> 7) != 0);
            continue;
          }
          if (ids[s] == 0xffffffffu) continue;
          if (dense_off_bytes == 2) emit(ids[s], off16[s] & 0x7fffu, seeds.rec_offset + r, j * seeds.distance, (off16[s] >> 15) != 0);
          else emit(ids[s], off32[s] & 0x7fffffffu, seeds.rec_offset + r, j * seeds.distance, (off32[s] >> 31) != 0);
        }
      }
      if (s != n_seeds) throw std::runtime_error("dense results do not match the chunk's seed count");
      for (uint64_t i = 0; i < n_extra; ++i) {
        const uint32_t* x = p.extra + 4 * i;
        emit(x[0], x[1], x[2], x[3] & 0x7fffffffu, (x[3] >> 31) != 0);
      }
    }
    else if (n_hits) {
      // per-hit records (walk mode / wide ids); the per-record kind routes the hit when there are two callbacks
      void* raw = p.dense;
      grow(raw, p.dense_cap, n_hits * 4, 8);     // room for n_hits x 32 bytes
      p.dense = raw;
      uint64_t got = 0;
      if (p.compact) pcheck(p, psi_b200_fetch32(p.ctx, static_cast<uint32_t*>(raw), n_hits, &got));
      else pcheck(p, psi_b200_fetch(p.ctx, static_cast<uint64_t*>(raw), n_hits, &got));
      std::vector<uint8_t> kinds;
      if (cb2) {
        kinds.resize(n_hits);
        pcheck(p, psi_b200_fetch_kinds(p.ctx, kinds.data(), n_hits, &got));
      }
      for (uint64_t i = 0; i < n_hits; ++i) {
        const bool off_path = !kinds.empty() && kinds[i] == 2;
        if (p.compact) { const uint32_t* x = static_cast<const uint32_t*>(raw) + 4 * i; emit(x[0], x[1], x[2], x[3], off_path); }
        else { const uint64_t* x = static_cast<const uint64_t*>(raw) + 4 * i; emit(x[0], x[1], x[2], x[3], off_path); }
      }
    }
    stats_type::chunks_done() += 1;
    stats_type::reads_done() += seeds.n_reads;
    if (rcb) rcb(rec, n_rec);
  }

  void account(const Pipe& p) const
  {
    psi_b200_counters_t c;
    pcheck(p, psi_b200_counters(p.ctx, &c));
    TraverserStats::seeds_off_paths() += c.n_hits_off;
    TraverserStats::nof_godowns() += c.n_walks;
    stats_ptr->set_timer("seeds-on-paths", (c.ms_on) * 1e-3);
    stats_ptr->set_timer("seeds-off-paths", (c.ms_read_index + c.ms_off) * 1e-3);
    stats_ptr->set_timer("seeds-off-path", (c.ms_read_index + c.ms_off) * 1e-3);
    stats_ptr->set_timer("query-dindex", 0.0);
  }

  void run(readsrecord_type const& seeds, readsindex_type const& idx, unsigned flags, const char* timer_name,
           const callback_type& cb1, const callback_type& cb2) const
  {
    auto timer = timer_name ? std::make_unique<Timer>(stats_ptr->timeit_ts(timer_name)) : nullptr;
    Pipe& p = pipe();
    begin(p, seeds, idx, flags);
    deliver(p, cb1, cb2, nullptr);
  }

  const graph_type* graph_ptr;
  std::vector<Position<>> starting_loci;
  unsigned int seed_len;
  unsigned char seed_mismatches;
  unsigned int gocc_threshold;
  unsigned int max_mem;
  unsigned int context_ = 0;
  std::pair<unsigned int, unsigned int> d{ 0, 0 };   // distance window of the index (reference: SeedFinder::d)
  std::unique_ptr<stats_type> stats_ptr;
  psi_b200_ctx* ctx = nullptr;              // builds and owns the resident graph / index / loci
  psi_b200_pathset* pathset = nullptr;
  bool has_index = false;
  bool finaliser = true;
  mutable bool loci_dirty = false;
  mutable std::mutex pipes_mutex;
  mutable std::map<std::thread::id, std::unique_ptr<Pipe>> pipes;
  mutable std::atomic<int> dense_state{ 0 };
  mutable std::mutex mem_mutex;
  mutable bool mem_built = false;
  std::vector<uint64_t> kept_ptr{ 0 };
  std::vector<uint32_t> kept_nodes, kept_head, kept_tail;             // 0 unknown, 1 the index serves dense results, 2 it does not (walk mode)
  uint64_t max_node_id = 0;
  unsigned dense_off_bytes = 4;
  mutable uint64_t checksum = 0;
  mutable bool checksum_valid = false;
};

}  // namespace psi
#endif
