// Test driver for the distance part of the psi::SeedFinder mirror: the calls of the reference's own scenario
// (test/src/test_seedfinder.cpp:225-312) -- create_distance_index(dmin, dmax, PerComponent{}), verify_distance on every
// pair, save_distance_index, a second finder that opens the saved index and answers again -- on locus pairs read from a
// file of 4 x u64 {coordinate id of v, offset, coordinate id of u, offset}.  Answers go to OUT as bytes: single calls,
// the bulk call, and the second finder's.  Usage: driver GFA DMIN DMAX PAIRS OUT PREFIX
#include <array>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../psi_b200/include/psi/seed_finder.hpp"

using namespace psi;

int main(int argc, char** argv)
{
  if (argc < 7) { std::fprintf(stderr, "usage: %s GFA DMIN DMAX PAIRS OUT PREFIX\n", argv[0]); return 2; }
  const unsigned dmin = std::atoi(argv[2]), dmax = std::atoi(argv[3]);
  typedef SeedFinderTraits<gum::Succinct, Dna5QStringSet<>, seqan2::IndexWotd<>, InMemory> traits_type;
  typedef SeedFinder<NoStats, traits_type> finder_type;
  gum::SeqGraph<gum::Succinct> graph;
  gum::util::load(graph, std::string(argv[1]), true);

  std::vector<std::array<uint64_t, 4>> ends;
  {
    std::FILE* f = std::fopen(argv[4], "rb");
    if (!f) { std::perror("pairs"); return 2; }
    uint64_t row[4];
    while (std::fread(row, 8, 4, f) == 4)
      ends.push_back(std::array<uint64_t, 4>{ (uint64_t)graph.id_by_coordinate((int64_t)row[0]), row[1], (uint64_t)graph.id_by_coordinate((int64_t)row[2]), row[3] });
    std::fclose(f);
  }

  finder_type::set_kokkos_handling_status(false);      // the reference's test does (test_seedfinder.cpp:255)
  finder_type finder(graph, 30);
  finder.unset_as_finaliser();
  bool threw_without_index = false;
  try { finder.verify_distance((int64_t)ends[0][0], ends[0][1], (int64_t)ends[0][2], ends[0][3]); }
  catch (const std::runtime_error&) { threw_without_index = true; }
  std::vector<std::string> infos;
  finder.create_distance_index(dmin, dmax, PerComponent{}, [&](std::string const& m) { infos.push_back(m); });
  std::vector<uint8_t> single, single_mt(ends.size());
  const size_t n_single = std::min<size_t>(ends.size(), 200);
  for (size_t i = 0; i < n_single; ++i)
    single.push_back(finder.verify_distance((int64_t)ends[i][0], ends[i][1], (int64_t)ends[i][2], ends[i][3]) ? 1 : 0);
  std::vector<uint8_t> bulk = finder.verify_distances(ends);
  // one const finder, several threads (each gets its own pipeline)
  {
    const finder_type& cf = finder;
    std::vector<std::thread> th;
    for (int t = 0; t < 3; ++t)
      th.emplace_back([&, t] {
        for (size_t i = t; i < n_single; i += 3)
          single_mt[i] = cf.verify_distance((int64_t)ends[i][0], ends[i][1], (int64_t)ends[i][2], ends[i][3]) ? 1 : 0;
      });
    for (auto& x : th) x.join();
  }
  bool mt_ok = true;
  for (size_t i = 0; i < n_single; ++i) mt_ok &= single_mt[i] == single[i];

  const std::string prefix = argv[6];
  const bool saved = finder.save_distance_index(prefix);
  finder_type finder2(graph, 30);
  const bool wrong_window = finder2.open_distance_index(prefix, dmin, dmax + 1);   // no such file
  const bool opened = finder2.open_distance_index(prefix, dmin, dmax);
  std::vector<uint8_t> again = opened ? finder2.verify_distances(ends) : std::vector<uint8_t>();

  std::FILE* o = std::fopen(argv[5], "wb");
  if (!o) { std::perror("out"); return 2; }
  std::fwrite(single.data(), 1, single.size(), o);
  std::fwrite(bulk.data(), 1, bulk.size(), o);
  std::fwrite(again.data(), 1, again.size(), o);
  std::fclose(o);
  std::printf("{\"pairs\": %zu, \"single\": %zu, \"threw_without_index\": %s, \"infos\": %zu, \"saved\": %s, \"opened\": %s, "
              "\"wrong_window\": %s, \"mt_ok\": %s, \"entries\": %llu}\n",
              ends.size(), n_single, threw_without_index ? "true" : "false", infos.size(), saved ? "true" : "false",
              opened ? "true" : "false", wrong_window ? "true" : "false", mt_ok ? "true" : "false",
              (unsigned long long)finder.get_counters().n_dindex_entries);
  return 0;
}
