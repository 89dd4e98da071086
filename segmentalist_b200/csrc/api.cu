// Error reporting, launch accounting and library identity for libsegb200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace segb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace segb

extern "C" const char *segb_last_error(void) { return segb::g_err; }
extern "C" int segb_version(void) { return 100; }
extern "C" int64_t segb_launch_count(void) { return (int64_t)segb::g_launches.load(); }
