// Error reporting and launch accounting shared by every C-ABI entry point.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace gldm {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace gldm

extern "C" const char* gldm_last_error(void) { return gldm::g_err; }
extern "C" int gldm_version(void) { return 100; }
extern "C" unsigned long long gldm_launch_count(void) { return gldm::g_launches.load(); }
