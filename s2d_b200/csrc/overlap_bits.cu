// K1 (CUDA-core variant): dense mask-overlap contraction on bit-packed planes,
//   I[a,b] = sum_w popc(A[a,w] & B[b,w]),  areaA[a] = popc(A[a,:]),  areaB[b] = popc(B[b,:])
// plus the producers of its operands: bit packing of u8 planes and rasterisation of tracks
// (pred_tracks_to_binary_masks, return_mask=False, cotracker_matching.py:453-503).
#include "common.cuh"

namespace s2d {

// ---- pack: one warp per 32 consecutive output words (1024 pixels) --------------------------
__global__ void pack_bits_kernel(const uint8_t* __restrict__ planes, int64_t nwords_total, int64_t npix,
                                 int64_t wpr, uint32_t* __restrict__ bits) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords_total) return;
    const int64_t row = w / wpr, wi = w - row * wpr;
    const uint8_t* src = planes + row * npix + wi * 32;
    const int64_t left = npix - wi * 32;
    uint32_t m = 0;
    if (left >= 32 && (((uintptr_t)src) & 15) == 0) {
        const int4 a = ld_stream(reinterpret_cast<const int4*>(src));
        const int4 b = ld_stream(reinterpret_cast<const int4*>(src) + 1);
        const uint32_t ws[8] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w,
                                (uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t nz = __vsetne4(ws[k], 0u);            // 0x01 per nonzero byte
            const uint32_t nib = (nz | (nz >> 7) | (nz >> 14) | (nz >> 21)) & 0xFu;
            m |= nib << (4 * k);
        }
    } else {
        for (int b = 0; b < 32 && b < left; ++b) m |= (uint32_t)(src[b] != 0) << b;
    }
    bits[w] = m;
}

// ---- unpack: bit rows -> u8 0/1 planes (the tensor-core operand of a Gram over bit rows); one thread per 16 output bytes
__global__ void unpack_bits_kernel(const uint32_t* __restrict__ bits, int N, int stride, int64_t ncols, uint8_t* __restrict__ planes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_row = ncols / 16;
    if (i >= (int64_t)N * per_row) return;
    const int64_t row = i / per_row, c16 = i - row * per_row;
    const int64_t w = c16 >> 1;
    uint32_t m = w < stride ? bits[row * stride + w] : 0u;
    m = (c16 & 1) ? (m >> 16) : m;
    uint4 o;
    o.x = ((m & 0xFu) * 0x00204081u) & 0x01010101u;
    o.y = (((m >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    o.z = (((m >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    o.w = (((m >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    reinterpret_cast<uint4*>(planes + row * ncols)[c16] = o;
}

// ---- rasterise: one CTA per frame, scatter 1s (duplicates collapse) ------------------------
__global__ void rasterise_kernel(const float* __restrict__ tracks, int P, int H, int W,
                                 uint8_t* __restrict__ planes) {
    const int t = blockIdx.x;
    const float2* tp = reinterpret_cast<const float2*>(tracks) + (int64_t)t * P;
    uint8_t* out = planes + (int64_t)t * H * W;
    const float Wf = (float)W, Hf = (float)H;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const float2 v = tp[p];
        const float rx = rintf(v.x), ry = rintf(v.y);
        if ((rx >= 0.f) && (rx < Wf) && (ry >= 0.f) && (ry < Hf)) out[(int64_t)(int)ry * W + (int)rx] = 1;
    }
}

// ---- overlap: 32x32 output tile per CTA, split over the word range --------------------------
constexpr int OV_TILE = 32;
constexpr int OV_WC = 64;          // words staged per step
constexpr int OV_THREADS = 1024;

__global__ void __launch_bounds__(OV_THREADS)
overlap_bits_kernel(const uint32_t* __restrict__ A, int Na, const uint32_t* __restrict__ B, int Nb,
                    int64_t nwords, int64_t words_per_split, int32_t* __restrict__ I) {
    __shared__ uint32_t As[OV_TILE][OV_WC + 1];
    __shared__ uint32_t BsT[OV_WC][OV_TILE + 1];
    const int a0 = blockIdx.y * OV_TILE, b0 = blockIdx.x * OV_TILE;
    const int64_t wbeg = (int64_t)blockIdx.z * words_per_split;
    const int64_t wend = min(nwords, wbeg + words_per_split);
    const int ta = threadIdx.x >> 5, tb = threadIdx.x & 31;
    int acc = 0;
    for (int64_t w0 = wbeg; w0 < wend; w0 += OV_WC) {
        const int cw = (int)min((int64_t)OV_WC, wend - w0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < OV_TILE * OV_WC; idx += OV_THREADS) {
            const int r = idx / OV_WC, w = idx - r * OV_WC;
            uint32_t va = 0, vb = 0;
            if (w < cw) {
                if (a0 + r < Na) va = A[(int64_t)(a0 + r) * nwords + w0 + w];
                if (b0 + r < Nb) vb = B[(int64_t)(b0 + r) * nwords + w0 + w];
            }
            As[r][w] = va;
            BsT[w][r] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int w = 0; w < OV_WC; ++w) acc += __popc(As[ta][w] & BsT[w][tb]);
    }
    if (a0 + ta < Na && b0 + tb < Nb && acc) atomicAdd(&I[(int64_t)(a0 + ta) * Nb + b0 + tb], acc);
}

__global__ void row_popc_kernel(const uint32_t* __restrict__ X, int N, int64_t nwords, int32_t* __restrict__ area) {
    const int row = blockIdx.x;
    if (row >= N) return;
    int c = 0;
    for (int64_t w = threadIdx.x; w < nwords; w += blockDim.x) c += __popc(X[(int64_t)row * nwords + w]);
    c = warp_sum(c);
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) area[row] = s;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_pack_bits(const uint8_t* planes, int N, int64_t npix, uint32_t* bits, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(planes && bits && N > 0 && npix > 0, "s2d_pack_bits: bad arguments");
    const int64_t wpr = (npix + 31) / 32, total = wpr * N;
    pack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(planes, total, npix, wpr, bits);
    S2D_CHECK_LAUNCH("pack_bits_kernel");
    return 0;
}

extern "C" int s2d_unpack_bits(const uint32_t* bits, int N, int stride_words, int64_t ncols, uint8_t* planes, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(bits && planes && N > 0 && stride_words > 0 && ncols > 0 && ncols % 16 == 0 && (((uintptr_t)planes) & 15) == 0,
                  "s2d_unpack_bits: bad arguments (ncols must be a multiple of 16, planes 16-byte aligned)");
    const int64_t n = (int64_t)N * (ncols / 16);
    unpack_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, N, stride_words, ncols, planes);
    S2D_CHECK_LAUNCH("unpack_bits_kernel");
    return 0;
}

extern "C" int s2d_rasterise_tracks(const float* tracks, int T, int P, int H, int W, uint8_t* planes, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(tracks && planes && T > 0 && P > 0 && H > 0 && W > 0, "s2d_rasterise_tracks: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(planes, 0, (size_t)T * H * W, st);
    rasterise_kernel<<<T, 256, 0, st>>>(tracks, P, H, W, planes);
    S2D_CHECK_LAUNCH("rasterise_kernel");
    return 0;
}

extern "C" int s2d_overlap_bits(const uint32_t* Abits, int Na, const uint32_t* Bbits, int Nb, int64_t nwords,
                                int32_t* I, int32_t* areaA, int32_t* areaB, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(Abits && Bbits && I && Na > 0 && Nb > 0 && nwords > 0, "s2d_overlap_bits: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(I, 0, (size_t)Na * Nb * sizeof(int32_t), st);
    const int ta = (Na + OV_TILE - 1) / OV_TILE, tb = (Nb + OV_TILE - 1) / OV_TILE;
    // enough splits to fill the 148 SMs a few times over
    int64_t splits = (148 * 4 + (int64_t)ta * tb - 1) / ((int64_t)ta * tb);
    int64_t wps = (nwords + splits - 1) / splits;
    wps = ((wps + OV_WC - 1) / OV_WC) * OV_WC;
    splits = (nwords + wps - 1) / wps;
    S2D_CHECK_ARG(splits <= 65535 && ta <= 65535, "s2d_overlap_bits: problem too large");
    dim3 grid(tb, ta, (unsigned)splits);
    overlap_bits_kernel<<<grid, OV_THREADS, 0, st>>>(Abits, Na, Bbits, Nb, nwords, wps, I);
    S2D_CHECK_LAUNCH("overlap_bits_kernel");
    if (areaA) { row_popc_kernel<<<Na, 256, 0, st>>>(Abits, Na, nwords, areaA); S2D_CHECK_LAUNCH("row_popc_kernel"); }
    if (areaB) { row_popc_kernel<<<Nb, 256, 0, st>>>(Bbits, Nb, nwords, areaB); S2D_CHECK_LAUNCH("row_popc_kernel"); }
    return 0;
}
