// Host-side plumbing shared by every C-ABI entry point: error strings, CUDA status mapping, device info,
// launch accounting.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mb200 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error at %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return MB200_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace mb200

extern "C" int mb200_abi_version(void) { return MB200_ABI_VERSION; }
extern "C" const char* mb200_last_error(void) { return mb200::g_err; }
extern "C" uint64_t mb200_launch_count(void) { return mb200::g_launches.load(std::memory_order_relaxed); }
