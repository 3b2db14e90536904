// K3d: appearance events of visibility curves - moving average, threshold, morphological opening
// and (appear, disappear) transitions per row, i.e. the run-length view of the per-track
// visibility flags. Replaces extract_appearance_events (cotracker_occlusions.py:166-223, identical
// copy at cotracker_matching.py:212-269) and boolean_visibility (:226-240 / :272-286); exported by
// the reference, not called by its driver.
//
// One warp per row. The row's signals live in the warp's slice of shared memory as bytes; every
// step is a sliding window over a reflect-padded predecessor, exactly as torch pads and pools:
//   smooth[i] = sum_j w * V[refl_T(i - p1 + j)], j < sw           w = float32(1) / float32(sw)
//   bin[i]    = smooth[i] >= thresh                                (float32 compare)
//   er[i]     = min_j bin[refl_T(i - p2 + j)],  j < k,  i < T1 = T  + 2 p2 - k + 1
//   op[i]     = max_j er[refl_T1(i - p2 + j)],  j < k,  i < T2 = T1 + 2 p2 - k + 1
// (p1 = (sw-1)/2, p2 = (k-1)/2: for even k each pooling shortens the signal by one sample, as in
// the reference.) Transitions op[i] -> op[i+1] are compacted in order with ballots and popcounts
// (a warp-level prefix scan): 0->1 appends i+1 to the row's start list, 1->0 to its end list.
#include "common.cuh"

namespace s2d {

__device__ __forceinline__ int reflect_idx(int j, int n) {       // torch 'reflect' padding, |pad| < n
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    S2D_DEV_ASSERT(j >= 0 && j < n);
    return j;
}

constexpr int EV_WARPS = 4;

__global__ void __launch_bounds__(EV_WARPS * 32)
appearance_events_kernel(const float* __restrict__ V, int N, int T, int sw, float thresh, int k, int max_events,
                         int32_t* __restrict__ nstart, int32_t* __restrict__ nend,
                         int32_t* __restrict__ starts, int32_t* __restrict__ ends,
                         uint8_t* __restrict__ opened) {
    extern __shared__ uint8_t ev_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * EV_WARPS + warp;
    if (row >= N) return;
    uint8_t* bin = ev_sm + (size_t)warp * 3 * T;
    uint8_t* er = bin + T;
    uint8_t* op = er + T;
    const float* v = V + (int64_t)row * T;
    const int p1 = (sw - 1) / 2, p2 = (k - 1) / 2;
    const int T1 = T + 2 * p2 - k + 1, T2 = T1 + 2 * p2 - k + 1;
    const float w = __fdiv_rn(1.0f, (float)sw);

    for (int i = lane; i < T; i += 32) {
        float s = 0.0f;
        for (int j = 0; j < sw; ++j) s = __fmaf_rn(w, v[reflect_idx(i - p1 + j, T)], s);
        bin[i] = (sw == 1 ? v[i] : s) >= thresh ? 1 : 0;          // x * 1.0f is x: no rounding for the default window
    }
    __syncwarp();
    for (int i = lane; i < T1; i += 32) {
        uint8_t m = 1;
        for (int j = 0; j < k; ++j) m &= bin[reflect_idx(i - p2 + j, T)];
        er[i] = m;
    }
    __syncwarp();
    for (int i = lane; i < T2; i += 32) {
        uint8_t m = 0;
        for (int j = 0; j < k; ++j) m |= er[reflect_idx(i - p2 + j, T1)];
        op[i] = m;
        if (opened) opened[(int64_t)row * T + i] = m;
    }
    __syncwarp();
    int ns = 0, ne = 0;
    for (int base = 0; base < T2 - 1; base += 32) {
        const int i = base + lane;
        S2D_DEV_ASSERT(T2 <= T && T1 <= T);
        const int d = (i < T2 - 1) ? (int)op[i + 1] - (int)op[i] : 0;
        const uint32_t ms = __ballot_sync(0xffffffffu, d == 1), me = __ballot_sync(0xffffffffu, d == -1);
        const uint32_t below = (1u << lane) - 1u;
        if (d == 1) { const int p = ns + __popc(ms & below); if (p < max_events) starts[(int64_t)row * max_events + p] = i + 1; }
        if (d == -1) { const int p = ne + __popc(me & below); if (p < max_events) ends[(int64_t)row * max_events + p] = i + 1; }
        ns += __popc(ms);
        ne += __popc(me);
    }
    if (lane == 0) { nstart[row] = ns; nend[row] = ne; }
}

__global__ void boolean_visibility_kernel(const float* __restrict__ V, int64_t n, float threshold, uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = V[i] >= threshold ? 1 : 0;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_appearance_events(const float* V, int N, int T, int smoothing_window, float thresh,
                                     int min_run_length, int max_events, int32_t* nstart, int32_t* nend,
                                     int32_t* starts, int32_t* ends, uint8_t* opened, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(V && nstart && nend && starts && ends, "s2d_appearance_events: null pointer");
    S2D_CHECK_ARG(N > 0 && T > 0 && max_events > 0, "s2d_appearance_events: bad sizes");
    S2D_CHECK_ARG(smoothing_window >= 1 && (smoothing_window & 1), "s2d_appearance_events: smoothing_window must be odd and >= 1");
    S2D_CHECK_ARG(min_run_length >= 1, "s2d_appearance_events: min_run_length must be >= 1");
    const int p1 = (smoothing_window - 1) / 2, p2 = (min_run_length - 1) / 2;
    const int T1 = T + 2 * p2 - min_run_length + 1, T2 = T1 + 2 * p2 - min_run_length + 1;
    // torch's reflect padding needs pad < length at every step; the reference raises otherwise
    S2D_CHECK_ARG(p1 < T && p2 < T && T1 >= 1 && p2 < T1 && T2 >= 1,
                  "s2d_appearance_events: T=%d is too short for smoothing_window=%d / min_run_length=%d", T, smoothing_window, min_run_length);
    const size_t smem = (size_t)EV_WARPS * 3 * T;
    S2D_CHECK_ARG(smem <= 200 * 1024, "s2d_appearance_events: T=%d exceeds the shared-memory row buffers", T);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(appearance_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("appearance_events_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
    }
    appearance_events_kernel<<<(N + EV_WARPS - 1) / EV_WARPS, EV_WARPS * 32, smem, (cudaStream_t)stream>>>(
        V, N, T, smoothing_window, thresh, min_run_length, max_events, nstart, nend, starts, ends, opened);
    S2D_CHECK_LAUNCH("appearance_events_kernel");
    return 0;
}

extern "C" int s2d_boolean_visibility(const float* V, int64_t n, float threshold, uint8_t* out, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(V && out && n > 0, "s2d_boolean_visibility: bad arguments");
    boolean_visibility_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(V, n, threshold, out);
    S2D_CHECK_LAUNCH("boolean_visibility_kernel");
    return 0;
}
