// Error reporting and device-attribute cache for the C ABI.
#include <stdarg.h>

#include "mf_common.cuh"

static thread_local char g_err[512] = "";

void mf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

MF_API const char* mf_last_error(void) { return g_err; }
MF_API int mf_abi_version(void) { return MF_ABI_VERSION; }

int mf_sm_count_cached() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

MF_API int mf_device_sm_count(void) {
    int dev = 0, n = 0;
    MF_CUDA(cudaGetDevice(&dev));
    MF_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}
