// K2: point-in-mask voting (the dominant, HBM-bound kernel of the path).
//
// One CTA per (query q, frame t) tile of P tracked points (8 B each, read exactly once with
// 128-bit streaming loads). Rounded, in-bounds pixels are de-duplicated in a shared-memory
// bitmap addressed relative to the tile's bounding box (XOR-swizzled so that neighbouring
// pixels fall into different banks); the first thread to set a pixel's bit gathers the pixel's
// label (u8 label map, L2/L1 resident: every query of a video re-reads the same T frames) and
// votes into a shared histogram through per-thread run-length and warp-level aggregation.
// Output per tile: hits[q,t,0..L) and uniq[q,t] - 4(L+1) bytes against 8P bytes read.
//
// Replaces pred_tracks_to_binary_masks + compute_point_mask_intersection over every mask of
// the frame (cotracker_matching.py:453-503, 640-662, 681-692): uniq = |P| = union,
// hits[lab] = |P AND mask_lab| = intersection.
#include "common.cuh"

#include <limits.h>

namespace s2d {

constexpr int PV_BM_WORDS = 4096;                 // 16 KB bitmap = 131072 pixels per band
constexpr int PV_BM_BITS = PV_BM_WORDS * 32;
constexpr uint32_t PV_INVALID = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t pv_key(float x, float y, float Wf, float Hf) {
    // torch .round() is round-half-to-even in float32; NaN/inf/huge fail the float compares
    const float rx = rintf(x), ry = rintf(y);
    const bool ok = (rx >= 0.f) && (rx < Wf) && (ry >= 0.f) && (ry < Hf);
    return ok ? (((uint32_t)(int)ry << 16) | (uint32_t)(int)rx) : PV_INVALID;
}

template <int THREADS, int PPT, bool VEC4>
__global__ void __launch_bounds__(THREADS)
point_votes_kernel(const s2d_video_desc* __restrict__ descs, const int32_t* __restrict__ rowinfo, const int32_t* __restrict__ vidinfo,
                   int32_t* __restrict__ hits, int32_t* __restrict__ uniq) {
    const s2d_video_desc d = descs[blockIdx.y];
    const int64_t rt = blockIdx.x;
    if (rt >= (int64_t)d.Nm * d.T) return;
    if (vidinfo && vidinfo[(int64_t)blockIdx.y * S2D_VIDINFO_WORDS + 1] < 0) return;
    const int q = (int)(rt / d.T), t = (int)(rt - (int64_t)q * d.T);
    if (rowinfo) {
        const int4 ri = reinterpret_cast<const int4*>(rowinfo)[d.row0 + q];
        if (ri.y < 0 || t < ri.z || t > ri.w) return;
    }
    const int n = d.npts ? min(max(d.npts[q], 0), d.P) : d.P;

    __shared__ __align__(16) uint32_t bm[PV_BM_WORDS];
    __shared__ int hist[S2D_MAX_LABELS];
    __shared__ int sbox[4];

    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < PV_BM_WORDS / 4; i += THREADS) reinterpret_cast<uint4*>(bm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < S2D_MAX_LABELS; i += THREADS) hist[i] = 0;
    if (tid == 0) { sbox[0] = INT_MAX; sbox[1] = INT_MAX; sbox[2] = -1; sbox[3] = -1; }

    const float* tp = d.tracks + rt * (int64_t)d.P * 2;
    const uint8_t* lbl = d.labels + (int64_t)t * d.H * d.W;
    const float Wf = (float)d.W, Hf = (float)d.H;

    // ---- pass 1: load, round, bounds, label gather, bounding box -------------------------
    uint32_t key[PPT];
    uint32_t lab4[(PPT + 3) / 4];
#pragma unroll
    for (int i = 0; i < (PPT + 3) / 4; ++i) lab4[i] = 0;
    int xmin = INT_MAX, ymin = INT_MAX, xmax = -1, ymax = -1;
#pragma unroll
    for (int i = 0; i < PPT / 2; ++i) {
        const int p0 = 2 * (i * THREADS + tid);
        float4 v = make_float4(-1.f, -1.f, -1.f, -1.f);
        if (VEC4) {
            if (p0 + 1 < n) {
                const int4 r = ld_stream(reinterpret_cast<const int4*>(tp) + (i * THREADS + tid));
                v = make_float4(__int_as_float(r.x), __int_as_float(r.y), __int_as_float(r.z), __int_as_float(r.w));
            } else if (p0 < n) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(tp) + p0);
                v.x = a.x; v.y = a.y;
            }
        } else {
            if (p0 < n) { const float2 a = __ldg(reinterpret_cast<const float2*>(tp) + p0); v.x = a.x; v.y = a.y; }
            if (p0 + 1 < n) { const float2 a = __ldg(reinterpret_cast<const float2*>(tp) + p0 + 1); v.z = a.x; v.w = a.y; }
        }
        key[2 * i] = pv_key(v.x, v.y, Wf, Hf);
        key[2 * i + 1] = pv_key(v.z, v.w, Wf, Hf);
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (key[k] != PV_INVALID) {
            const int ix = key[k] & 0xFFFF, iy = key[k] >> 16;
            const uint32_t lab = __ldg(lbl + (int64_t)iy * d.W + ix);
            lab4[k >> 2] |= lab << (8 * (k & 3));
            xmin = min(xmin, ix); xmax = max(xmax, ix);
            ymin = min(ymin, iy); ymax = max(ymax, iy);
        }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin);
    ymin = __reduce_min_sync(0xffffffffu, ymin);
    xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymax = __reduce_max_sync(0xffffffffu, ymax);
    __syncthreads();                      // smem init visible
    if (lane == 0 && xmax >= 0) {
        atomicMin(&sbox[0], xmin); atomicMin(&sbox[1], ymin);
        atomicMax(&sbox[2], xmax); atomicMax(&sbox[3], ymax);
    }
    __syncthreads();
    const int x0 = sbox[0], y0 = sbox[1], x1 = sbox[2], y1 = sbox[3];

    // ---- pass 2: de-duplicate in the bitmap band by band, vote ---------------------------
    if (x1 >= 0) {
        const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
        const int rpb = PV_BM_BITS / bw;                  // rows per band (bw <= 65535 -> >= 2)
        int cur = -1, cnt = 0;
        for (int ylo = y0; ylo <= y1; ylo += rpb) {
            const int yhi = min(y1, ylo + rpb - 1);
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (key[k] == PV_INVALID) continue;
                const int ix = key[k] & 0xFFFF, iy = key[k] >> 16;
                if (iy < ylo || iy > yhi) continue;
                const uint32_t kk = (uint32_t)(iy - ylo) * (uint32_t)bw + (uint32_t)(ix - x0);
                uint32_t w = kk & (PV_BM_WORDS - 1);
                w ^= (w >> 5) & 31u;                       // bank swizzle
                const uint32_t bit = 1u << (kk >> 12);
                const uint32_t old = atomicOr(&bm[w], bit);
                if (!(old & bit)) {                        // first point on this pixel
                    const int lab = (lab4[k >> 2] >> (8 * (k & 3))) & 255;
                    if (lab != cur) {
                        if (cnt) atomicAdd(&hist[cur], cnt);
                        cur = lab; cnt = 0;
                    }
                    ++cnt;
                }
            }
            if (yhi < y1) {                                // more bands: recycle the bitmap
                __syncthreads();
                for (int i = tid; i < PV_BM_WORDS / 4; i += THREADS)
                    reinterpret_cast<uint4*>(bm)[i] = make_uint4(0, 0, 0, 0);
                __syncthreads();
            }
        }
        // warp-aggregated flush of the per-thread runs (usually one label per warp)
        uint32_t remaining = __ballot_sync(0xffffffffu, cnt > 0);
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const int l0 = __shfl_sync(0xffffffffu, cur, leader);
            const bool mine = (cnt > 0) && (cur == l0);
            const int s = __reduce_add_sync(0xffffffffu, mine ? cnt : 0);
            if (lane == leader) atomicAdd(&hist[l0], s);
            remaining &= ~__ballot_sync(0xffffffffu, mine);
        }
        (void)bh;
    }
    __syncthreads();

    // ---- write out ------------------------------------------------------------------------
    int32_t* hout = hits + d.hits_off + rt * d.L;
    int part = 0;
    for (int l = tid; l < S2D_MAX_LABELS; l += THREADS) {
        const int h = hist[l];
        part += h;
        if (l < d.L) hout[l] = h;
    }
    part = warp_sum(part);
    __shared__ int s_tot;
    if (tid == 0) s_tot = 0;
    __syncthreads();
    if (lane == 0 && part) atomicAdd(&s_tot, part);
    __syncthreads();
    if (tid == 0) uniq[d.vt_off + rt] = s_tot;
}

template <int THREADS, int PPT>
static int launch_pv(bool vec4, dim3 grid, cudaStream_t st, const s2d_video_desc* descs, const int32_t* rowinfo,
                     const int32_t* vidinfo, int32_t* hits, int32_t* uniq) {
    if (vec4)
        point_votes_kernel<THREADS, PPT, true><<<grid, THREADS, 0, st>>>(descs, rowinfo, vidinfo, hits, uniq);
    else
        point_votes_kernel<THREADS, PPT, false><<<grid, THREADS, 0, st>>>(descs, rowinfo, vidinfo, hits, uniq);
    S2D_CHECK_LAUNCH("point_votes_kernel");
    return 0;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_point_votes(const s2d_video_desc* descs, int nvideos, int64_t max_rows_x_T, int max_P,
                               int vec4_ok, const int32_t* rowinfo, const int32_t* vidinfo,
                               int32_t* hits, int32_t* uniq, void* stream) {
    S2D_CHECK_ARG(descs && hits && uniq, "s2d_point_votes: null pointer");
    S2D_CHECK_ARG(nvideos > 0 && nvideos <= 65535 && max_rows_x_T > 0 && max_rows_x_T <= 2147483647LL,
                  "s2d_point_votes: bad sizes");
    S2D_CHECK_ARG(max_P >= 1 && max_P <= 32768, "s2d_point_votes: P=%d not in [1, 32768]", max_P);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)max_rows_x_T, nvideos);
    const bool v4 = vec4_ok != 0;
#define PV_ARGS v4, grid, st, descs, rowinfo, vidinfo, hits, uniq
    if (max_P <= 256 * 4) return launch_pv<256, 4>(PV_ARGS);
    if (max_P <= 256 * 8) return launch_pv<256, 8>(PV_ARGS);
    if (max_P <= 256 * 16) return launch_pv<256, 16>(PV_ARGS);
    if (max_P <= 512 * 16) return launch_pv<512, 16>(PV_ARGS);
    if (max_P <= 1024 * 16) return launch_pv<1024, 16>(PV_ARGS);
    return launch_pv<1024, 32>(PV_ARGS);
#undef PV_ARGS
}
