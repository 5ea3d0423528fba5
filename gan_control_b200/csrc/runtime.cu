// Library-wide state: thread-local error string, launch counter, device properties.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace b200gan {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

}  // namespace b200gan

extern "C" {
int b200gan_version(void) { return 100; }
const char* b200gan_last_error(void) { return b200gan::g_err; }
uint64_t b200gan_launch_count(void) { return b200gan::g_launches.load(std::memory_order_relaxed); }
}
