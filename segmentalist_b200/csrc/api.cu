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

// Attributes of the CURRENT device, cached per device id (one process may drive several GPUs).
int device_info(int *dev_out, int *n_sm_out, int *coop_out) {
    constexpr int MAX_DEV = 64;
    static std::atomic<int> sm[MAX_DEV], coop[MAX_DEV];       // 0 = not queried yet
    int dev = 0;
    SEGB_CUDA(cudaGetDevice(&dev));
    int n = (dev >= 0 && dev < MAX_DEV) ? sm[dev].load(std::memory_order_acquire) : 0, c = 0;
    if (n == 0) {
        SEGB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        SEGB_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrCooperativeLaunch, dev));
        if (dev >= 0 && dev < MAX_DEV) { coop[dev].store(c + 1, std::memory_order_relaxed); sm[dev].store(n, std::memory_order_release); }
    } else {
        c = coop[dev].load(std::memory_order_relaxed) - 1;
    }
    if (dev_out) *dev_out = dev;
    if (n_sm_out) *n_sm_out = n;
    if (coop_out) *coop_out = c;
    return 0;
}
}  // namespace segb

extern "C" const char *segb_last_error(void) { return segb::g_err; }
extern "C" int segb_version(void) { return 100; }
extern "C" int64_t segb_launch_count(void) { return (int64_t)segb::g_launches.load(); }
