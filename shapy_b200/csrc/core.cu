// Error reporting / bookkeeping shared by every entry point of libshapy_b200.so.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace shapy {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace shapy

extern "C" const char *shapy_last_error(void) { return shapy::g_err; }
extern "C" int shapy_version(void) { return 100; }
extern "C" long long shapy_launch_count(void) { return shapy::g_launches.load(); }
