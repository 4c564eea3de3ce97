// psi/stats.hpp -- named wall-clock timers with a process-wide registry.
//
// Mirrors the part of psi::Timer<> the CLI relies on (reference
// include/psi/stats.hpp:191-345): RAII timers keyed by name, accumulated over
// laps, `get_timers()`, `get_duration_str(name)`; durations print as seconds
// ("0.386000 s") like the reference's default CpuClock timers (stats.hpp:45,104).
// GPU phases additionally report CUDA-event milliseconds through
// psi_b200_counters(); SeedFinder folds those into the same registry.
#ifndef PSI_B200_PSI_STATS_HPP
#define PSI_B200_PSI_STATS_HPP

#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

namespace psi {

struct NoStats {};
struct WithStats {};

inline std::string get_thread_id()
{
  std::ostringstream ss;
  ss << std::this_thread::get_id();
  return ss.str();
}

class Timer {
 public:
  typedef std::chrono::steady_clock clock_type;
  struct TimePeriod {
    double seconds = 0;   // accumulated over finished laps
    double last = 0;      // last lap
    std::string str() const
    {
      char buf[64];
      std::snprintf(buf, sizeof buf, "%f s", seconds);
      return buf;
    }
    double duration() const { return seconds; }
  };
  typedef TimePeriod period_type;
  typedef Timer timer_type;

  explicit Timer(std::string name) : name_(std::move(name)), start_(clock_type::now()) {}
  Timer(Timer&& o) noexcept : name_(std::move(o.name_)), start_(o.start_), live_(o.live_) { o.live_ = false; }
  Timer(const Timer&) = delete;
  ~Timer() { stop(); }
  void stop()
  {
    if (!live_) return;
    live_ = false;
    add(name_, std::chrono::duration<double>(clock_type::now() - start_).count());
  }

  static void add(const std::string& name, double seconds)
  {
    std::lock_guard<std::mutex> lock(mutex());
    TimePeriod& p = table()[name];
    p.seconds += seconds;
    p.last = seconds;
  }
  // replace the value of a timer (used for per-chunk timers, which the CLI prints per chunk)
  static void set(const std::string& name, double seconds)
  {
    std::lock_guard<std::mutex> lock(mutex());
    TimePeriod& p = table()[name];
    p.seconds = seconds;
    p.last = seconds;
  }
  static std::map<std::string, TimePeriod> get_timers()
  {
    std::lock_guard<std::mutex> lock(mutex());
    return table();
  }
  static TimePeriod get(const std::string& name)
  {
    std::lock_guard<std::mutex> lock(mutex());
    auto it = table().find(name);
    return it == table().end() ? TimePeriod() : it->second;
  }
  static std::string get_duration_str(const std::string& name) { return get(name).str(); }
  static double get_duration(const std::string& name) { return get(name).seconds; }

 private:
  static std::map<std::string, TimePeriod>& table() { static std::map<std::string, TimePeriod> t; return t; }
  static std::mutex& mutex() { static std::mutex m; return m; }
  std::string name_;
  clock_type::time_point start_;
  bool live_ = true;
};

}  // namespace psi
#endif
