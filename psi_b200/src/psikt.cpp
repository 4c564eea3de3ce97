// psikt.cpp -- the seed-finding CLI on top of psi::SeedFinder (B200 build).
//
// Drop-in for the reference CLI (src/psikt.cpp) on this path: same options
// (options.hpp), same log lines (they are what the reference's tooling scrapes,
// script/parse2csv_psikt_config.yaml:41-62), same output bytes: per seed hit
// four native-endian size_t  node_id, node_offset, read_id, read_offset, no
// header (src/psikt.cpp:172-181).  The control flow follows find_seeds()
// (src/psikt.cpp:83-212): load or create the path index, then per read chunk
// get_seeds -> index_reads -> seeds_all.  What differs, and is logged as such:
//   - the output is the seed SET (every hit once); the reference's raw stream
//     repeats hits per covering path / walk and varies from run to run;
//   - graphs are read from GFA (1 or 2) and from vg's protobuf files (.vg), chosen by file name like gum::util::load;
//   - all per-chunk times are device times of the CUDA kernels.
#include <csignal>
#include <future>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#include "../include/psi/seed_finder.hpp"
#include "logger.hpp"
#include "options.hpp"

using namespace psi;
using klibpp::SeqStreamIn;

template <typename TSeedFinder>
static void report(TSeedFinder& finder, unsigned long long covered_reads, unsigned long long found)
{
  auto log = get_logger("main");
  log->info("Total number of starting loci: {}", finder.get_starting_loci().size());
  log->info("Total number of seeds found: {}", found);
  log->info("-> of which found off paths: {}", TSeedFinder::traverser_type::stats_type::get_total_seeds_off_paths());
  log->info("Total number of reads covered: {}", covered_reads);
  log->info("Total number of 'godown' operations: {}", TSeedFinder::traverser_type::stats_type::get_total_nof_godowns());
  log->info("All Timers");
  log->info("----------");
  for (const auto& timer : Timer::get_timers()) log->info("{}: {}", timer.first, timer.second.str());
}

template <class TGraph, typename TReadsIndexSpec>
static void find_seeds(TGraph& graph, SeqStreamIn& reads_iss, std::FILE* output_file, Options const& params, TReadsIndexSpec const)
{
  typedef Dna5QStringSet<> readsstringset_type;
  typedef SeedFinderTraits<typename TGraph::spec_type, readsstringset_type, TReadsIndexSpec> finder_traits_type;
  typedef SeedFinder<NoStats, finder_traits_type> finder_type;
  typedef typename finder_type::stats_type stats_type;
  typedef typename stats_type::timer_type timer_type;

  auto log = get_logger("main");
  auto tid = get_thread_id();
  std::signal(SIGUSR1, finder_type::stats_type::signal_handler);

  finder_type finder(graph, params.seed_len, params.gocc_threshold, params.max_mem);
  auto const& stats = finder.get_stats();
  log->info("Looking for an existing path index...");
  if (finder.load_path_index(params.pindex_path, params.context, params.step_size, params.dindex_min_ris, params.dindex_max_ris)) {
    log->info("The path index has been found and loaded.");
  }
  else if (params.path_num == 0) {
    log->info("No path has been specified. Skipping path indexing...");
  }
  else {
    log->info("No valid path index found. Creating the path index...");
    log->info("Selecting {} different path(s) in the graph...", params.path_num);
    auto info_cb = [&log](std::string const& msg) -> void { log->info(msg); };
    auto warn_cb = [&log](std::string const& msg) -> void { log->warn(msg); };
    if (params.dindex_mode == "whole")
      finder.create_path_index(params.path_num, params.patched, params.context, params.step_size, params.dindex_min_ris,
                               params.dindex_max_ris, psi::Whole{}, info_cb, warn_cb);
    else if (params.dindex_mode == "per-component")
      finder.create_path_index(params.path_num, params.patched, params.context, params.step_size, params.dindex_min_ris,
                               params.dindex_max_ris, psi::PerComponent{}, info_cb, warn_cb);
    else
      throw std::runtime_error("Unknown distance index construction mode: " + params.dindex_mode);
    log->info("Picked paths in {}.", stats.get_timer("pick-paths", tid).str());
    log->info("Indexed paths in {}.", stats.get_timer("index-paths", tid).str());
    log->info("Found uncovered loci in {}.", stats.get_timer("find-uncovered", tid).str());
    log->info("Created distance index in {}.", stats.get_timer("index-distances", tid).str());
    log->info("Saving path index...");
    if (params.pindex_path.empty()) log->warn("No path index file is specified. Skipping...");
    else if (!finder.serialize_path_index(params.pindex_path, params.step_size))
      log->warn("Specified path index file is not writable. Skipping...");
    else {
      log->info("Saved path index in {}.", stats.get_timer("save-pindex", tid).str());
      log->info("Saved distance index in {}.", stats.get_timer("save-dindex", tid).str());
    }
  }
  {
    auto c = finder.get_counters();
    log->info("Device path index: {} distinct k-mers at {} loci in {} bytes (built in {} ms on the GPU).", c.n_index_kmers,
              c.n_index_entries, c.index_bytes, c.ms_index_build);
  }
  log->info("Number of starting loci (in {} nodes of total {}): {}", finder.get_nof_uniq_nodes(),
            finder.get_graph_ptr()->get_node_count(), finder.get_starting_loci().size());

  if (params.indexonly) {
    log->info("Skipping seed finding as requested...");
    return;
  }

  unsigned long long found = 0, covered_reads = 0;
  std::vector<bool> covered;
  {
    // two chunk records alternate: while the GPU works on one chunk the next one is parsed from the reads file
    // (the stream keeps two sets of page-locked buffers for exactly this)
    decltype(finder.create_readrecord()) chunks[2] = { finder.create_readrecord(), finder.create_readrecord() };
    auto seeds = finder.create_readrecord();
    log->info("Finding seeds...");
    [[maybe_unused]] auto timer = timer_type("seed-finding");
    auto load = [&](int slot) {
      [[maybe_unused]] auto timer = timer_type("load-chunk");
      return readRecords(chunks[slot], reads_iss, params.chunk_size);
    };
    int cur = 0;
    log->info("Loading a read chunk...");
    bool have = load(cur);
    while (have) {
      auto& chunk = chunks[cur];
      log->info("Fetched {} reads with total length of {}bp in {}.", length(chunk), lengthSum(chunk), timer_type::get_duration_str("load-chunk"));
      finder.get_seeds(seeds, chunk, params.distance);
      auto seeds_index = finder.index_reads(seeds);
      log->info("Seeding done in {}.", stats.get_timer("seeding", tid).str());
      log->info("Finding all seeds...");
      finder.seeds_all_begin(seeds, seeds_index);
      // the next chunk is parsed and packed by a second host thread while this one waits for the GPU, expands the
      // records and writes them (the stream's two buffer sets alternate, so the chunk in flight is not touched)
      log->info("Loading a read chunk...");
      auto next_chunk = std::async(std::launch::async, load, cur ^ 1);
      covered.assign(chunk.size(), false);
      const uint64_t first = chunk.rec_offset;
      try {
        finder.seeds_all_wait_records([&](const uint64_t* rec, uint64_t n) {
          found += n;
          static_assert(sizeof(std::size_t) == sizeof(uint64_t), "the output format is 4 x size_t");
          if (n && std::fwrite(rec, 32, n, output_file) != n) throw std::runtime_error("could not write to the output file");
          for (uint64_t i = 0; i < n; ++i) {
            const uint64_t r = rec[4 * i + 2] - first;
            if (!covered[r]) { covered[r] = true; ++covered_reads; }
          }
        });
      }
      catch (...) { next_chunk.wait(); throw; }
      const bool have_next = next_chunk.get();
      log->info("Found seeds on paths in {}.", stats.get_timer("seeds-on-paths", tid).str());
      log->info("Found seeds off paths in {}.", stats.get_timer("seeds-off-paths", tid).str());
      log->info("Verified distance constraints in {}.", stats.get_timer("query-dindex", tid).str());
      cur ^= 1;
      have = have_next;
    }
  }
  log->info("Found seed in {}.", timer_type::get_duration_str("seed-finding"));
  log->info("Seed counts are those of the seed set: every (read, offset, node, offset) hit is written once.");
  report(finder, covered_reads, found);
}

static void startup(const Options& options)
{
  auto log = get_logger("main");
  log->info("Parameters:");
  log->info("- Seed length: {}", options.seed_len);
  log->info("- Seed distance: {}", options.distance);
  log->info("- Number of paths: {}", options.path_num);
  log->info("- Context size (used in patching): {}", options.context);
  log->info("- Patched: {}", (options.patched ? "yes" : "no"));
  log->info("- Path index file: '{}'", options.pindex_path);
  log->info("- Reads chunk size: {}", options.chunk_size);
  log->info("- Reads index type: {}", index_to_str(options.index));
  log->info("- Step size: {}", options.step_size);
  log->info("- Seed genome occurrence count threshold: {}", options.gocc_threshold);
  log->info("- Maximum number of MEMs on paths: {}", options.max_mem);
  log->info("- Distance index minimum read insert size: {}", options.dindex_min_ris);
  log->info("- Distance index maximum read insert size: {}", options.dindex_max_ris);
  log->info("- Distance index construction mode: {}", options.dindex_mode);
  const char* tmpdir = std::getenv("TMPDIR");
  log->info("- Temporary directory: '{}'", tmpdir ? tmpdir : "/tmp");
  log->info("- Output file: '{}'", options.output_path);

  log->info("Loading input graph from file '{}'...", options.rf_path);
  gum::SeqGraph<gum::Succinct> graph;
  gum::util::load(graph, options.rf_path, true);
  if (gum::util::ids_in_topological_order(graph)) log->info("Input graph node IDs are in topological sort order.");
  else log->warn("Input graph node IDs are NOT in topological sort order.");

  log->info("Opening reads file '{}'...", options.fq_path);
  SeqStreamIn reads_iss(options.fq_path.c_str());
  if (!reads_iss) {
    std::string msg = "could not open file '" + options.fq_path + "'!";
    log->error(msg);
    throw std::runtime_error(msg);
  }

  std::FILE* output_file = std::fopen(options.output_path.c_str(), "wb");
  if (!output_file) {
    std::string msg = "could not open file '" + options.output_path + "'!";
    log->error(msg);
    throw std::runtime_error(msg);
  }

  try {
    switch (options.index) {
      // the reads index type only selected the CPU index of the reference; the device hash serves both
      case IndexType::Wotd: find_seeds(graph, reads_iss, output_file, options, seqan2::IndexWotd<>()); break;
      case IndexType::Esa: find_seeds(graph, reads_iss, output_file, options, seqan2::IndexEsa<>()); break;
      default: throw std::runtime_error("Index not implemented.");
    }
  }
  catch (...) {
    std::fclose(output_file);
    throw;
  }
  if (std::fclose(output_file) != 0) throw std::runtime_error("could not close the output file");
}

int main(int argc, char* argv[])
{
  Options options;
  auto res = parse_args(options, argc, argv);
  if (res != ParseResult::Ok) return res == ParseResult::Error;
  config_logger(options);
  int rc = EXIT_SUCCESS;
  try {
    startup(options);
  }
  catch (const std::exception& e) {
    if (auto log = get_logger("main")) log->error("{}", e.what());
    std::fprintf(stderr, "psikt: %s\n", e.what());
    rc = EXIT_FAILURE;
  }
  drop_all_loggers();
  return rc;
}
