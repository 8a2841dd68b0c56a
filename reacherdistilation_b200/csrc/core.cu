// Error reporting and device queries of the C ABI (include/reacher_b200.h).
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace rb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RB_ERR_CUDA;
}

}  // namespace rb

extern "C" {

const char* rb_last_error(void) { return rb::g_err; }
int rb_version(void) { return 100; }

int rb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rb_sm_count(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    return v;
}

}  // extern "C"
