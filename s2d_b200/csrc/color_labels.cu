// f1 (first "next" row): colour-coded mask frames -> per-frame label ids.
// label = 1 + rank of the pixel's (R,G,B) tuple among the frame's non-black colours in
// lexicographic order, 0 for black - what load_masks / convert_lblimg_to_maskid compute with
// np.unique(axis=0) + one np.all(rgb == colour) pass per colour (cotracker_occlusions.py:22-85,
// cotracker_matching.py:22-84, crw_utils.py:688-767; 42 % of the reference's wall time per video).
//
// Three HBM-bound passes over packed RGB (3 B/px): collect the distinct colours of each frame
// (run-length filtered, CTA-local shared hash set, then a 1024-slot global set per frame), rank them
// (lexicographic order of (R,G,B) == numeric order of R<<16|G<<8|B), map every pixel (table in
// shared memory). 7 bytes of traffic per pixel.
#include "common.cuh"

namespace s2d {

constexpr uint32_t CL_EMPTY = 0xFFFFFFFFu;
constexpr int CL_SLOTS = 1024;            // global set per frame (keys) + 1024 ranks
constexpr int CL_LOCAL = 512;             // CTA-local set
constexpr int CL_THREADS = 256;
constexpr int CL_GROUPS_PER_THREAD = 16;  // 4-pixel groups per thread

__device__ __forceinline__ uint32_t cl_hash(uint32_t k) { return k * 2654435761u; }

// keys of 4 consecutive pixels from 3 little-endian words of packed RGB
__device__ __forceinline__ void cl_keys(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&k)[4]) {
    k[0] = __byte_perm(w0, 0u, 0x4012);
    k[1] = __byte_perm(w0, w1, 0x0345) & 0x00FFFFFFu;
    k[2] = __byte_perm(w1, w2, 0x0234) & 0x00FFFFFFu;
    k[3] = __byte_perm(w2, 0u, 0x4123);
}

template <int SLOTS>
__device__ __forceinline__ bool cl_insert(uint32_t* tab, uint32_t k, int shift) {
    uint32_t h = cl_hash(k) >> shift;
    for (int probe = 0; probe < SLOTS; ++probe) {
        const uint32_t old = atomicCAS(&tab[h], CL_EMPTY, k);
        if (old == CL_EMPTY || old == k) return true;
        h = (h + 1) & (SLOTS - 1);
    }
    return false;
}

__global__ void __launch_bounds__(CL_THREADS)
color_collect_kernel(const uint8_t* __restrict__ rgb, int64_t npix, uint32_t* __restrict__ gtab, int32_t* __restrict__ overflow) {
    __shared__ uint32_t ltab[CL_LOCAL];
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < CL_LOCAL; i += CL_THREADS) ltab[i] = CL_EMPTY;
    __syncthreads();
    const uint8_t* src = rgb + (int64_t)f * npix * 3;
    const int64_t ngroups = npix >> 2;
    const int64_t g0 = (int64_t)blockIdx.x * CL_THREADS * CL_GROUPS_PER_THREAD;
    uint32_t prev = 0;
    bool ok = true;
    const bool aligned = (((uintptr_t)src) & 3) == 0;
#pragma unroll 4
    for (int i = 0; i < CL_GROUPS_PER_THREAD; ++i) {
        const int64_t g = g0 + (int64_t)i * CL_THREADS + threadIdx.x;
        if (g >= ngroups) break;
        uint32_t k[4];
        if (aligned) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + g * 3;
            cl_keys(__ldg(p), __ldg(p + 1), __ldg(p + 2), k);
        } else {
            const uint8_t* p = src + g * 12;
#pragma unroll
            for (int j = 0; j < 4; ++j) k[j] = ((uint32_t)p[3 * j] << 16) | ((uint32_t)p[3 * j + 1] << 8) | p[3 * j + 2];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k[j] != 0 && k[j] != prev) { ok &= cl_insert<CL_LOCAL>(ltab, k[j], 23); prev = k[j]; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {                  // tail pixels (npix % 4)
        for (int64_t px = ngroups * 4; px < npix; ++px) {
            const uint32_t k = ((uint32_t)src[3 * px] << 16) | ((uint32_t)src[3 * px + 1] << 8) | src[3 * px + 2];
            if (k != 0) ok &= cl_insert<CL_LOCAL>(ltab, k, 23);
        }
    }
    __syncthreads();
    uint32_t* gt = gtab + (int64_t)f * 2 * CL_SLOTS;
    for (int i = threadIdx.x; i < CL_LOCAL; i += CL_THREADS) {
        const uint32_t k = ltab[i];
        if (k != CL_EMPTY) ok &= cl_insert<CL_SLOTS>(gt, k, 22);
    }
    if (!ok) atomicExch(&overflow[f], 1);
}

__global__ void __launch_bounds__(CL_SLOTS)
color_rank_kernel(uint32_t* __restrict__ gtab, const int32_t* __restrict__ overflow, int32_t* __restrict__ ncolors) {
    __shared__ uint32_t keys[CL_SLOTS];
    const int f = blockIdx.x;
    uint32_t* gt = gtab + (int64_t)f * 2 * CL_SLOTS;
    const uint32_t mine = gt[threadIdx.x];
    keys[threadIdx.x] = mine;
    const int n = __syncthreads_count(mine != CL_EMPTY);
    uint32_t rank = 1;
    if (mine != CL_EMPTY)
        for (int i = 0; i < CL_SLOTS; ++i) rank += keys[i] < mine;       // CL_EMPTY is never smaller
    gt[CL_SLOTS + threadIdx.x] = rank;
    if (threadIdx.x == 0) ncolors[f] = overflow[f] ? 0x7FFFFFFF : n;
}

__global__ void __launch_bounds__(CL_THREADS)
color_map_kernel(const uint8_t* __restrict__ rgb, int64_t npix, const uint32_t* __restrict__ gtab, uint8_t* __restrict__ labels) {
    __shared__ uint32_t keys[CL_SLOTS];
    __shared__ uint8_t ranks[CL_SLOTS];
    const int f = blockIdx.y;
    const uint32_t* gt = gtab + (int64_t)f * 2 * CL_SLOTS;
    for (int i = threadIdx.x; i < CL_SLOTS; i += CL_THREADS) { keys[i] = gt[i]; ranks[i] = (uint8_t)gt[CL_SLOTS + i]; }
    __syncthreads();
    auto lookup = [&](uint32_t k) -> uint32_t {
        if (k == 0) return 0u;
        uint32_t h = cl_hash(k) >> 22;
        for (int probe = 0; probe < CL_SLOTS; ++probe) {
            if (keys[h] == k) return ranks[h];
            h = (h + 1) & (CL_SLOTS - 1);
        }
        return 0u;
    };
    const uint8_t* src = rgb + (int64_t)f * npix * 3;
    uint8_t* dst = labels + (int64_t)f * npix;
    const int64_t ngroups = npix >> 2;
    const int64_t g0 = (int64_t)blockIdx.x * CL_THREADS * CL_GROUPS_PER_THREAD;
    const bool aligned = ((((uintptr_t)src) | ((uintptr_t)dst)) & 3) == 0;
    uint32_t pk = 0, pl = 0;                     // last looked-up colour (label maps are piecewise constant)
#pragma unroll 4
    for (int i = 0; i < CL_GROUPS_PER_THREAD; ++i) {
        const int64_t g = g0 + (int64_t)i * CL_THREADS + threadIdx.x;
        if (g >= ngroups) break;
        uint32_t k[4];
        if (aligned) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + g * 3;
            cl_keys(__ldg(p), __ldg(p + 1), __ldg(p + 2), k);
        } else {
            const uint8_t* p = src + g * 12;
#pragma unroll
            for (int j = 0; j < 4; ++j) k[j] = ((uint32_t)p[3 * j] << 16) | ((uint32_t)p[3 * j + 1] << 8) | p[3 * j + 2];
        }
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k[j] != pk) { pk = k[j]; pl = lookup(pk); }
            out |= pl << (8 * j);
        }
        if (aligned) reinterpret_cast<uint32_t*>(dst)[g] = out;
        else { for (int j = 0; j < 4; ++j) dst[g * 4 + j] = (uint8_t)(out >> (8 * j)); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t px = ngroups * 4; px < npix; ++px)
            dst[px] = (uint8_t)lookup(((uint32_t)src[3 * px] << 16) | ((uint32_t)src[3 * px + 1] << 8) | src[3 * px + 2]);
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_color_to_labels_work_ints(int nframes, int64_t* out) {
    if (!out || nframes <= 0) return -1;
    *out = (int64_t)nframes * (2 * CL_SLOTS + 1);
    return 0;
}

extern "C" int s2d_color_to_labels(const uint8_t* rgb, int nframes, int64_t npix, uint32_t* work, uint8_t* labels,
                                   int32_t* ncolors, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(rgb && work && labels && ncolors && nframes > 0 && nframes <= 65535 && npix > 0, "s2d_color_to_labels: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* gtab = work;
    int32_t* overflow = reinterpret_cast<int32_t*>(work + (int64_t)nframes * 2 * CL_SLOTS);
    cudaMemsetAsync(gtab, 0xFF, (size_t)nframes * 2 * CL_SLOTS * sizeof(uint32_t), st);
    cudaMemsetAsync(overflow, 0, (size_t)nframes * sizeof(int32_t), st);
    const int64_t per_cta = (int64_t)CL_THREADS * CL_GROUPS_PER_THREAD * 4;
    dim3 grid((unsigned)((npix + per_cta - 1) / per_cta), nframes);
    color_collect_kernel<<<grid, CL_THREADS, 0, st>>>(rgb, npix, gtab, overflow);
    S2D_CHECK_LAUNCH("color_collect_kernel");
    color_rank_kernel<<<nframes, CL_SLOTS, 0, st>>>(gtab, overflow, ncolors);
    S2D_CHECK_LAUNCH("color_rank_kernel");
    color_map_kernel<<<grid, CL_THREADS, 0, st>>>(rgb, npix, gtab, labels);
    S2D_CHECK_LAUNCH("color_map_kernel");
    return 0;
}
