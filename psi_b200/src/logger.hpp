// logger.hpp -- the "main" logger of psikt: console and file sinks with the
// switches of the reference (src/logger.hpp:48-98): console at warn level unless
// -v, file always at info level, -q/-Q/-D/-C.  The message texts are the metrics
// interface of the reference (script/parse2csv_psikt_config.yaml), so they are
// emitted verbatim by psikt.cpp; this class only formats and routes them.
#ifndef PSI_B200_SRC_LOGGER_HPP
#define PSI_B200_SRC_LOGGER_HPP

#include <chrono>
#include <cstdio>
#include <ctime>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>

namespace psi {

class Logger {
 public:
  enum Level { Info = 0, Warn = 1, Error = 2 };
  Logger(bool nolog, bool quiet, bool nocolor, bool verbose, bool nologfile, const std::string& log_path)
  {
    if (!nolog && !quiet) { console_ = true; console_level_ = verbose ? Info : Warn; color_ = !nocolor; }
    if (!nolog && !nologfile) file_ = std::fopen(log_path.c_str(), "a");
  }
  ~Logger() { if (file_) std::fclose(file_); }

  template <typename... Args>
  void info(const std::string& fmt, Args&&... args) { log(Info, format(fmt, std::forward<Args>(args)...)); }
  template <typename... Args>
  void warn(const std::string& fmt, Args&&... args) { log(Warn, format(fmt, std::forward<Args>(args)...)); }
  template <typename... Args>
  void error(const std::string& fmt, Args&&... args) { log(Error, format(fmt, std::forward<Args>(args)...)); }

 private:
  // "{}" placeholders, left to right
  static std::string format(const std::string& fmt) { return fmt; }
  template <typename T, typename... Rest>
  static std::string format(const std::string& fmt, T&& v, Rest&&... rest)
  {
    auto p = fmt.find("{}");
    if (p == std::string::npos) return fmt;
    std::ostringstream ss;
    ss << v;
    return fmt.substr(0, p) + ss.str() + format(fmt.substr(p + 2), std::forward<Rest>(rest)...);
  }

  void log(Level lvl, const std::string& msg)
  {
    static const char* names[] = { "info", "warning", "error" };
    static const char* colors[] = { "\033[32m", "\033[33m\033[1m", "\033[31m\033[1m" };
    using namespace std::chrono;
    auto now = system_clock::now();
    std::time_t t = system_clock::to_time_t(now);
    int ms = (int)(duration_cast<milliseconds>(now.time_since_epoch()).count() % 1000);
    std::tm tm;
    localtime_r(&t, &tm);
    char ts[48];
    std::snprintf(ts, sizeof ts, "%04d-%02d-%02d %02d:%02d:%02d.%03d", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday,
                  tm.tm_hour, tm.tm_min, tm.tm_sec, ms);
    std::lock_guard<std::mutex> lock(mutex_);
    if (console_ && lvl >= console_level_) {
      if (color_) std::fprintf(stdout, "[%s] [main] [%s%s\033[m] %s\n", ts, colors[lvl], names[lvl], msg.c_str());
      else std::fprintf(stdout, "[%s] [main] [%s] %s\n", ts, names[lvl], msg.c_str());
      std::fflush(stdout);
    }
    if (file_) { std::fprintf(file_, "[%s] [main] [%s] %s\n", ts, names[lvl], msg.c_str()); std::fflush(file_); }
  }

  bool console_ = false, color_ = false;
  Level console_level_ = Warn;
  std::FILE* file_ = nullptr;
  std::mutex mutex_;
};

inline std::shared_ptr<Logger>& main_logger() { static std::shared_ptr<Logger> l; return l; }
inline std::shared_ptr<Logger> get_logger(const std::string&) { return main_logger(); }
template <typename TOptions>
inline void config_logger(const TOptions& o)
{
  main_logger() = std::make_shared<Logger>(o.nolog, o.quiet, o.nocolor, o.verbose, o.nologfile, o.log_path);
}
inline void drop_all_loggers() { main_logger().reset(); }

}  // namespace psi
#endif
