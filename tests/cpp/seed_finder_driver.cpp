// Test driver for the psi::SeedFinder mirror (psi_b200/include/psi/seed_finder.hpp): it makes the
// calls find_seeds() of the reference CLI makes (reference src/psikt.cpp:83-212), once with
// seeds_on_paths + seeds_off_paths and separate callbacks, once with seeds_all, and dumps the hits
// as 4 x u64 {read_id, read_offset, node_id (coordinate), node_offset} for the python side to compare
// with the oracle.  Usage: driver GFA READS K D N_PATHS CHUNK OUT_PREFIX
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../psi_b200/include/psi/seed_finder.hpp"

using namespace psi;

static void dump(const std::string& path, const std::vector<uint64_t>& v)
{
  std::FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { std::perror("fopen"); std::exit(2); }
  if (!v.empty() && std::fwrite(v.data(), 8, v.size(), f) != v.size()) { std::perror("fwrite"); std::exit(2); }
  std::fclose(f);
}

template <typename F>
static bool throws_runtime_error(F f, const std::string& needle)
{
  try { f(); }
  catch (const std::runtime_error& e) { return std::string(e.what()).find(needle) != std::string::npos; }
  return false;
}

int main(int argc, char** argv)
{
  if (argc < 8) { std::fprintf(stderr, "usage: %s GFA READS K D N_PATHS CHUNK OUT_PREFIX\n", argv[0]); return 2; }
  const std::string gfa = argv[1], reads_path = argv[2], out = argv[7];
  const unsigned k = std::atoi(argv[3]), d = std::atoi(argv[4]), n_paths = std::atoi(argv[5]);
  const unsigned long chunk_size = std::strtoul(argv[6], nullptr, 10);

  typedef SeedFinderTraits<gum::Succinct, Dna5QStringSet<>, seqan2::IndexWotd<>, InMemory> traits_type;
  typedef SeedFinder<NoStats, traits_type> finder_type;

  gum::SeqGraph<gum::Succinct> graph;
  gum::util::load(graph, gfa, true);
  finder_type finder(graph, k);
  std::vector<std::string> infos;
  finder.create_path_index(n_paths, true, 0, 1, 0, 0, PerComponent{}, [&](std::string const& m) { infos.push_back(m); },
                           [&](std::string const& m) { infos.push_back("W:" + m); });

  std::vector<uint64_t> on, off, all1, all2;
  auto push = [&graph](std::vector<uint64_t>& v) {
    return [&v, &graph](Seed<> const& h) {
      v.push_back(h.read_id); v.push_back(h.read_offset);
      v.push_back((uint64_t)graph.coordinate_id((int64_t)h.node_id)); v.push_back(h.node_offset);
    };
  };
  klibpp::SeqStreamIn iss(reads_path.c_str());
  if (!iss) { std::fprintf(stderr, "cannot open reads\n"); return 2; }
  auto chunk = finder.create_readrecord();
  auto seeds = finder.create_readrecord();
  auto traverser = finder.create_traverser();
  unsigned long n_chunks = 0, n_reads = 0;
  while (readRecords(chunk, iss, chunk_size)) {
    ++n_chunks;
    n_reads += length(chunk);
    finder.get_seeds(seeds, chunk, d);
    auto index = finder.index_reads(seeds);
    finder.seeds_on_paths(seeds, index, push(on));
    finder.setup_traverser(traverser, seeds, index);
    finder.seeds_off_paths(traverser, push(off));
    finder.seeds_all(seeds, index, traverser, push(all1), push(all2));
  }
  // one const finder, several host threads, one chunk each (reference seed_finder.hpp:386-399): every thread gets
  // its own pipeline over the shared index; the union of what the threads find is the same set
  std::vector<uint64_t> mt[3];
  {
    const finder_type& cf = finder;
    std::vector<std::thread> threads;
    for (int t = 0; t < 3; ++t)
      threads.emplace_back([&, t] {
        // every thread reads the file through its own stream (a stream's chunk buffers belong to one consumer) and
        // takes every third chunk
        klibpp::SeqStreamIn iss_t(reads_path.c_str());
        auto c = cf.create_readrecord();
        auto s = cf.create_readrecord();
        auto tr = cf.create_traverser();
        for (unsigned long i = 0; readRecords(c, iss_t, chunk_size ? chunk_size : 1000); ++i) {
          if (i % 3 != (unsigned long)t) continue;
          cf.get_seeds(s, c, d);
          auto ix = cf.index_reads(s);
          cf.seeds_all(s, ix, tr, push(mt[t]));
        }
      });
    for (auto& th : threads) th.join();
  }
  std::vector<uint64_t> mt_all;
  for (auto& v : mt) mt_all.insert(mt_all.end(), v.begin(), v.end());
  dump(out + ".mt", mt_all);
  // MEM mode: seeds_on_paths(sequence, callback) on the first reads of the file (seed_finder.hpp:1459-1479)
  std::vector<uint64_t> mems;
  {
    klibpp::SeqStreamIn iss_m(reads_path.c_str());
    auto c = finder.create_readrecord();
    if (readRecords(c, iss_m, 40))
      for (uint64_t r = 0; r < c.size(); ++r) {
        const std::string text = c.read(r);
        finder.seeds_on_paths(text, [&](Seed<> const& h) {
          mems.push_back(r); mems.push_back(h.read_offset); mems.push_back(h.match_len); mems.push_back(h.gocc);
          mems.push_back((uint64_t)graph.coordinate_id((int64_t)h.node_id)); mems.push_back(h.node_offset);
        });
      }
  }
  dump(out + ".mems", mems);
  dump(out + ".on", on);
  dump(out + ".off", off);
  dump(out + ".all1", all1);
  dump(out + ".all2", all2);

  // starting loci as (coordinate id, offset)
  std::vector<uint64_t> loci;
  for (auto const& l : finder.get_starting_loci()) { loci.push_back((uint64_t)graph.coordinate_id(l.node_id())); loci.push_back(l.offset()); }
  dump(out + ".loci", loci);

  // error behaviour of the reference API
  bool ok = true;
  {  // patched with a context shorter than the seed: seeds_on_paths must throw (seed_finder.hpp:1434-1437)
    finder_type f2(graph, k);
    f2.create_path_index(1, true, k - 1);
    klibpp::SeqStreamIn iss2(reads_path.c_str());
    auto c2 = f2.create_readrecord();
    auto s2 = f2.create_readrecord();
    readRecords(c2, iss2, 10);
    f2.get_seeds(s2, c2, d);
    auto i2 = f2.index_reads(s2);
    ok &= throws_runtime_error([&] { f2.seeds_on_paths(s2, i2, [](Seed<> const&) {}); }, "seed length should not be larger than context size");
  }
  {  // seeds of an older submission are rejected instead of silently using the newer chunk
    finder_type f6(graph, k);
    f6.create_path_index(1, true, 0);
    klibpp::SeqStreamIn iss6(reads_path.c_str());
    auto c6 = f6.create_readrecord();
    auto s6 = f6.create_readrecord();
    auto s7 = f6.create_readrecord();
    readRecords(c6, iss6, 10);
    f6.get_seeds(s6, c6, d);
    auto i6 = f6.index_reads(s6);
    f6.get_seeds(s7, c6, d);
    ok &= throws_runtime_error([&] { f6.seeds_all_records(s6, i6, nullptr); }, "do not belong");
  }
  ok &= throws_runtime_error([&] { finder_type f3(graph, 33); }, "seed length");
  ok &= throws_runtime_error([&] { finder_type f4(graph, k, 0, 0, 1); }, "approximate");
  {  // insert sizes given to create_path_index build the distance index too (seed_finder.hpp:1341-1343)
    finder_type f5(graph, k);
    f5.create_path_index(1, true, 0, 1, 100, 300);
    ok &= f5.get_counters().dindex_mode != 0;
  }
  std::printf("{\"chunks\": %lu, \"reads\": %lu, \"loci\": %zu, \"uniq_nodes\": %zu, \"on\": %zu, \"off\": %zu, \"all1\": %zu, \"all2\": %zu, "
              "\"infos\": %zu, \"errors_ok\": %s}\n",
              n_chunks, n_reads, finder.get_starting_loci().size(), finder.get_nof_uniq_nodes(), on.size() / 4, off.size() / 4,
              all1.size() / 4, all2.size() / 4, infos.size(), ok ? "true" : "false");
  return ok ? 0 : 1;
}
