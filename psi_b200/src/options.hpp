// options.hpp -- psikt command line (same option table as the reference CLI,
// src/psikt.cpp:293-471 and src/options.hpp:60-90): short and long names,
// defaults, required options, value checks and derived defaults are kept;
// parsing is done by hand (SeqAn's ArgumentParser is not part of this build).
#ifndef PSI_B200_SRC_OPTIONS_HPP
#define PSI_B200_SRC_OPTIONS_HPP

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace psi {

enum class IndexType { Sa = 1, Esa, Wotd, Dfi, QGram, FM };

inline IndexType index_from_str(const std::string& s)
{
  if (s == "SA") return IndexType::Sa;
  if (s == "ESA") return IndexType::Esa;
  if (s == "WOTD") return IndexType::Wotd;
  if (s == "DFI") return IndexType::Dfi;
  if (s == "QGRAM") return IndexType::QGram;
  if (s == "FM") return IndexType::FM;
  throw std::runtime_error("Undefined index type.");
}

inline std::string index_to_str(IndexType i)
{
  switch (i) {
    case IndexType::Sa: return "SA";
    case IndexType::Esa: return "ESA";
    case IndexType::Wotd: return "WOTD";
    case IndexType::Dfi: return "DFI";
    case IndexType::QGram: return "QGRAM";
    case IndexType::FM: return "FM";
  }
  throw std::runtime_error("Undefined index type.");
}

struct Options {
  unsigned int seed_len = 0;
  unsigned int chunk_size = 0;
  unsigned int step_size = 1;
  unsigned int distance = 0;
  unsigned int path_num = 0;
  unsigned int context = 0;
  unsigned int gocc_threshold = 0;
  unsigned int max_mem = 0;
  unsigned int dindex_min_ris = 0;
  unsigned int dindex_max_ris = 0;
  IndexType index = IndexType::Wotd;
  std::string rf_path;
  std::string fq_path;
  std::string output_path = "out.gam";
  std::string log_path = "psi.log";
  std::string pindex_path;
  std::string dindex_mode = "per-component";
  bool patched = true;
  bool indexonly = false;
  bool nologfile = false;
  bool nolog = false;
  bool quiet = false;
  bool nocolor = false;
  bool verbose = false;
};

enum class ParseResult { Ok, Help, Error };

namespace detail {

struct OptSpec {
  char short_name;          // 0 = none
  const char* long_name;
  bool takes_value;
  const char* meta;
  const char* help;
};

inline const std::vector<OptSpec>& specs()
{
  static const std::vector<OptSpec> s = {
    { 'f', "fastq", true, "FASTQ_FILE", "Reads in FASTQ format. (required; fq fastq fa fasta, optionally .gz)" },
    { 'o', "output", true, "OUTPUT_FILE", "Output file. Default: out.gam." },
    { 'I', "path-index", true, "PATH_INDEX_FILE", "Path index file." },
    { 'l', "seed-length", true, "INT", "Seed length. (required)" },
    { 'c', "chunk-size", true, "INT", "Reads chunk size. Set it to 0 to consider all reads as one chunk (default)." },
    { 'e', "step-size", true, "INT", "Minimum approximate distance allowed between two consecutive loci. Default: 1." },
    { 'd', "distance", true, "INT", "Distance between seeds. Default: seed length." },
    { 'n', "path-num", true, "INT", "Number of paths from the graph included in the path index. Default: 0." },
    { 'P', "no-patched", false, "", "Use full genome-wide paths." },
    { 't', "context", true, "INT", "Context length in patching. Default: 0." },
    { 'r', "gocc-threshold", true, "INT", "Seed genome occurrence count threshold (no threshold by default)." },
    { 'E', "max-mem", true, "INT", "Maximum number of MEMs on paths (default: find all)." },
    { 'm', "min-insert-size", true, "INT", "Distance index minimum read insert size (no distance indexing by default)." },
    { 'M', "max-insert-size", true, "INT", "Distance index maximum read insert size (minimum insert size by default)." },
    { 0, "dindex-mode", true, "MODE", "Distance index construction mode: one of per-component, whole. Default: per-component." },
    { 'i', "index", true, "INDEX", "Index type for indexing reads. One of SA, ESA, WOTD, DFI, QGRAM, FM. Default: WOTD." },
    { 'x', "index-only", false, "", "Only build path index and skip seed finding." },
    { 'L', "log-file", true, "LOG_FILE", "Sets default log file for existing and future loggers. Default: psi.log." },
    { 'Q', "no-log-file", false, "", "Disable writing logs to file (overrides -L)." },
    { 'q', "quiet", false, "", "Quiet mode. No output will be printed to console." },
    { 'C', "no-color", false, "", "Do not use a colored output." },
    { 'D', "disable-log", false, "", "Disable logging completely." },
    { 'v', "verbose", false, "", "Activates maximum verbosity." },
    { 'h', "help", false, "", "Display the help message." },
    { 0, "version", false, "", "Display version information." },
  };
  return s;
}

inline bool ends_with(const std::string& s, const std::string& suf)
{
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

inline unsigned int to_uint(const std::string& name, const std::string& v)
{
  char* end = nullptr;
  long long x = std::strtoll(v.c_str(), &end, 10);
  if (v.empty() || *end != '\0') throw std::runtime_error("the given value '" + v + "' cannot be casted to integer (option " + name + ")");
  if (x < 0 || x > 0xffffffffll) throw std::runtime_error("value out of range for option " + name);
  return (unsigned int)x;
}

}  // namespace detail

inline void print_help(std::ostream& os)
{
  os << "psikt - fully-sensitive seed finder (B200 build)\n\nSYNOPSIS\n    psikt [OPTIONS] \"GRAPH_FILE\"\n\nOPTIONS\n";
  for (const auto& s : detail::specs()) {
    os << "    ";
    if (s.short_name) os << "-" << s.short_name << ", ";
    os << "--" << s.long_name;
    if (s.takes_value) os << " " << s.meta;
    os << "\n          " << s.help << "\n";
  }
}

// Returns Ok and fills `o`, Help after printing help/version, Error after printing a message.
inline ParseResult parse_args(Options& o, int argc, char* argv[], std::ostream& out = std::cout, std::ostream& err = std::cerr)
{
  using namespace detail;
  std::vector<std::string> positional;
  bool have_f = false, have_l = false;
  std::string indexname = "WOTD";
  try {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      const OptSpec* spec = nullptr;
      std::string value;
      bool has_inline_value = false;
      if (a.size() >= 3 && a[0] == '-' && a[1] == '-') {
        std::string name = a.substr(2);
        auto eq = name.find('=');
        if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_inline_value = true; }
        for (const auto& s : specs()) if (name == s.long_name) spec = &s;
        if (!spec) throw std::runtime_error("unknown option: --" + name);
      }
      else if (a.size() >= 2 && a[0] == '-' && a != "-") {
        for (const auto& s : specs()) if (s.short_name && a[1] == s.short_name) spec = &s;
        if (!spec) throw std::runtime_error(std::string("unknown option: -") + a[1]);
        if (a.size() > 2) {
          if (!spec->takes_value) throw std::runtime_error("option -" + std::string(1, a[1]) + " takes no value");
          value = a.substr(2);
          has_inline_value = true;
        }
      }
      else { positional.push_back(a); continue; }
      if (spec->takes_value && !has_inline_value) {
        if (i + 1 >= argc) throw std::runtime_error(std::string("option --") + spec->long_name + " requires a value");
        value = argv[++i];
      }
      const std::string n = spec->long_name;
      if (n == "help") { print_help(out); return ParseResult::Help; }
      if (n == "version") { out << "psikt (psi_b200) " << psi_b200_version() << "\n"; return ParseResult::Help; }
      if (n == "fastq") { o.fq_path = value; have_f = true; }
      else if (n == "output") o.output_path = value;
      else if (n == "path-index") o.pindex_path = value;
      else if (n == "seed-length") { o.seed_len = to_uint(n, value); have_l = true; }
      else if (n == "chunk-size") o.chunk_size = to_uint(n, value);
      else if (n == "step-size") o.step_size = to_uint(n, value);
      else if (n == "distance") o.distance = to_uint(n, value);
      else if (n == "path-num") o.path_num = to_uint(n, value);
      else if (n == "no-patched") o.patched = false;
      else if (n == "context") o.context = to_uint(n, value);
      else if (n == "gocc-threshold") o.gocc_threshold = to_uint(n, value);
      else if (n == "max-mem") o.max_mem = to_uint(n, value);
      else if (n == "min-insert-size") o.dindex_min_ris = to_uint(n, value);
      else if (n == "max-insert-size") o.dindex_max_ris = to_uint(n, value);
      else if (n == "dindex-mode") {
        if (value != "per-component" && value != "whole") throw std::runtime_error("the given value '" + value + "' is not in the list of allowed values [per-component, whole] (option dindex-mode)");
        o.dindex_mode = value;
      }
      else if (n == "index") indexname = value;
      else if (n == "index-only") o.indexonly = true;
      else if (n == "log-file") o.log_path = value;
      else if (n == "no-log-file") o.nologfile = true;
      else if (n == "quiet") o.quiet = true;
      else if (n == "no-color") o.nocolor = true;
      else if (n == "disable-log") o.nolog = true;
      else if (n == "verbose") o.verbose = true;
    }
    if (positional.size() != 1) throw std::runtime_error(positional.empty() ? "Not enough arguments were provided." : "Too many arguments were provided.");
    o.rf_path = positional[0];
    if (!have_f) throw std::runtime_error("option -f, --fastq is required");
    if (!have_l) throw std::runtime_error("option -l, --seed-length is required");
    if (!(ends_with(o.rf_path, ".gfa") || ends_with(o.rf_path, ".vg") || ends_with(o.rf_path, ".gfa.gz")))
      throw std::runtime_error("invalid file extension for GRAPH_FILE (valid: vg gfa)");
    bool fq_ok = false;
    for (const char* e : { ".fq", ".fastq", ".fq.gz", ".fastq.gz", ".fa", ".fasta", ".fa.gz", ".fasta.gz" }) fq_ok |= ends_with(o.fq_path, e);
    if (!fq_ok) throw std::runtime_error("invalid file extension for option -f (valid: fq fastq fq.gz fastq.gz)");
    o.index = index_from_str(indexname);
  }
  catch (const std::exception& e) {
    err << "psikt: " << e.what() << "\n";
    return ParseResult::Error;
  }
  if (o.distance == 0) o.distance = o.seed_len;                   // src/psikt.cpp:469
  if (o.dindex_max_ris == 0) o.dindex_max_ris = o.dindex_min_ris;  // src/psikt.cpp:470
  return ParseResult::Ok;
}

}  // namespace psi
#endif
