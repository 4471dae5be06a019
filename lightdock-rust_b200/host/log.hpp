// log.hpp — the two things the reference's `log` + `env_logger` 0.11 give it (src/bin/lightdock-rust.rs:89):
// `info!("Step {}", step)` per GSO step (src/lib.rs:48), `info!("Atoms read: ..")` / `warn!(..)` in src/pydock.rs,
// written to stderr as `[<UTC time>Z LEVEL target] message` and filtered by RUST_LOG (default: errors only).
// LDB200_LOG is read as an alias so the variable can be set without affecting other Rust tools of a pipeline.
// Plus NVTX ranges around the phases of a step: free when no profiler is attached.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>

#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define LDB200_NVTX 1
#else
#define LDB200_NVTX 0
#endif

namespace lightdock {

enum class LogLevel { Error = 1, Warn = 2, Info = 3, Debug = 4, Trace = 5 };

inline int parse_level(const std::string &s) {
  std::string l;
  for (char c : s) l += (char)std::tolower((unsigned char)c);
  if (l == "off") return 0;
  if (l == "error") return 1;
  if (l == "warn") return 2;
  if (l == "info") return 3;
  if (l == "debug") return 4;
  if (l == "trace") return 5;
  return -1;
}

// Maximum enabled level for targets under `lightdock`: RUST_LOG is a comma-separated list of `level` or
// `target=level` directives (env_logger); a bare target enables everything for it.
inline int log_max_level() {
  static const int level = [] {
    const char *e = std::getenv("LDB200_LOG");
    if (!e || !*e) e = std::getenv("RUST_LOG");
    int lvl = 1;  // env_logger's default filter
    if (!e) return lvl;
    std::string spec(e);
    size_t st = 0;
    while (st <= spec.size()) {
      size_t en = spec.find(',', st);
      if (en == std::string::npos) en = spec.size();
      const std::string d = spec.substr(st, en - st);
      st = en + 1;
      if (d.empty()) continue;
      const size_t eq = d.find('=');
      if (eq == std::string::npos) {
        const int l = parse_level(d);
        if (l >= 0) lvl = l;
        else if (d.rfind("lightdock", 0) == 0) lvl = 5;
      } else if (d.substr(0, eq).rfind("lightdock", 0) == 0) {
        const int l = parse_level(d.substr(eq + 1));
        if (l >= 0) lvl = l;
      }
    }
    return lvl;
  }();
  return level;
}

inline bool log_enabled(LogLevel l) { return (int)l <= log_max_level(); }

inline void log_line(LogLevel l, const char *target, const std::string &msg) {
  if (!log_enabled(l)) return;
  char ts[32];
  const std::time_t now = std::time(nullptr);
  std::tm tmv;
  gmtime_r(&now, &tmv);
  std::strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &tmv);
  static const char *names[] = {"", "ERROR", "WARN ", "INFO ", "DEBUG", "TRACE"};
  std::fprintf(stderr, "[%s %s %s] %s\n", ts, names[(int)l], target, msg.c_str());
}

struct NvtxRange {  // RAII range: `NvtxRange r("gather");`
  explicit NvtxRange(const char *name) {
#if LDB200_NVTX
    nvtxRangePushA(name);
#else
    (void)name;
#endif
  }
  ~NvtxRange() {
#if LDB200_NVTX
    nvtxRangePop();
#endif
  }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

}  // namespace lightdock
