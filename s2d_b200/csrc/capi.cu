// C-ABI plumbing: thread-local error string, version, device query.
#include "common.cuh"

#include <stdarg.h>

namespace s2d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace s2d

extern "C" const char* s2d_last_error(void) { return s2d::g_err; }
extern "C" int s2d_version(void) { return 100; }
extern "C" int s2d_desc_size(void) { return (int)sizeof(s2d_video_desc); }

extern "C" int s2d_device_sm_count(int device) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        s2d::set_error("s2d_device_sm_count: %s", cudaGetErrorString(e));
        return -1;
    }
    return n;
}
