// Error reporting + tiny device queries for the mgv_b200 C ABI.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "mgv_common.cuh"

static thread_local char g_err[512] = "";

void mgv_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int mgv_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MGV_OK;
    mgv_set_error("CUDA error %s (%d) in %s", cudaGetErrorString(e), (int)e, what);
    return MGV_ERR_CUDA;
}

extern "C" const char* mgv_last_error_string(void) { return g_err; }
extern "C" int mgv_version(void) { return 100; }
extern "C" int mgv_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

// Process-wide statistics counter: kernels launched by this library (bench.py reports it).
static std::atomic<long long> g_launches{0};
void mgv_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long mgv_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
