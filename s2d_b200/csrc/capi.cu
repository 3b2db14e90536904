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
extern "C" int s2d_version(void) { return 203; }     // round 2, kernel revision 3 (recorded beside committed ncu-derived figures)
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

// Host-side helpers of the end-to-end path (no kernels): page-locking of caller-allocated host buffers (the
// huge-page backed staging pools of s2d_b200/hostmem.py) and the PCI address of a device (NUMA placement).
extern "C" int s2d_host_register(void* ptr, int64_t bytes) {
    S2D_CHECK_ARG(ptr && bytes > 0, "s2d_host_register: bad arguments");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        s2d::set_error("s2d_host_register: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -2;
    }
    return 0;
}

extern "C" int s2d_host_unregister(void* ptr) {
    S2D_CHECK_ARG(ptr, "s2d_host_unregister: null pointer");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        s2d::set_error("s2d_host_unregister: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -2;
    }
    return 0;
}

extern "C" int s2d_device_pci_bus_id(int device, char* out, int len) {
    S2D_CHECK_ARG(out && len >= 16, "s2d_device_pci_bus_id: need a buffer of at least 16 bytes");
    cudaError_t e = cudaDeviceGetPCIBusId(out, len, device);
    if (e != cudaSuccess) {
        s2d::set_error("s2d_device_pci_bus_id: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -2;
    }
    return 0;
}
