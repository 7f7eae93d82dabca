// Error plumbing, version and launch accounting of the C-ABI library.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace pgrf {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace pgrf

extern "C" const char* pgrf_last_error(void) { return pgrf::g_err; }
extern "C" int pgrf_version(void) { return 100; }
extern "C" int64_t pgrf_launch_count(void) { return pgrf::g_launches.load(std::memory_order_relaxed); }
