// K3c: majority vote per visibility cluster, run-length windows, highly-visible rows and
// candidates (identify_visibility_windows.py:134-203, get_visible_ranges :65-88,
// get_highly_visible_rows :90-105). Latency-bound (KBs per video): one launch for the column
// counts, one CTA per video for everything else.
#include "common.cuh"

namespace s2d {

// ccount[c][t] += X[row,t] for rows of cluster c; csize[c] += 1.  One thread per (row, word).
__global__ void cluster_count_kernel(const s2d_video_desc* __restrict__ descs,
                                     const uint32_t* __restrict__ xbits,
                                     const int32_t* __restrict__ labels1,
                                     int32_t* __restrict__ ccount, int32_t* __restrict__ clrow) {
    const s2d_video_desc d = descs[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)d.Nm * d.TW) return;
    const int q = (int)(i / d.TW), w = (int)(i % d.TW);
    const int c = labels1[d.row0 + q];
    if (c < 0) return;
    if (w == 0) atomicAdd(&clrow[(d.row0 + c) * 4 + 0], 1);
    uint32_t m = xbits[d.xbits_off + i];
    S2D_DEV_ASSERT(c < d.Nm);
    int32_t* dst = ccount + d.vt_off + (int64_t)c * d.T + w * 32;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        S2D_DEV_ASSERT(w * 32 + b < d.T);
        atomicAdd(&dst[b], 1);
    }
}

constexpr int WIN_THREADS = 256;

__global__ void __launch_bounds__(WIN_THREADS)
windows_kernel(const s2d_video_desc* __restrict__ descs, const uint32_t* __restrict__ xbits,
               const int32_t* __restrict__ labels1, const int32_t* __restrict__ qframe,
               const int32_t* __restrict__ ccount, float winner_fraction,
               uint32_t* __restrict__ majbits, uint32_t* __restrict__ rsbits,
               uint32_t* __restrict__ rebits, uint32_t* __restrict__ winbits,
               int32_t* __restrict__ rowinfo, int32_t* __restrict__ clrow,
               int32_t* __restrict__ vidinfo, int32_t* __restrict__ clusterinfo) {
    const int v = blockIdx.x;
    const s2d_video_desc d = descs[v];
    int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    const int k = vi[0];                       // clusters found by DBSCAN #1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = d.T, TW = d.TW;              // TW <= 32 (checked on the host)

    // phase A: one warp per cluster, lane = word of 32 frames
    for (int c = warp; c < k; c += WIN_THREADS / 32) {
        const int n = clrow[(d.row0 + c) * 4 + 0];
        uint32_t maj = 0;
        if (lane < TW) {
            const int32_t* cc = ccount + d.vt_off + (int64_t)c * T + lane * 32;
            for (int b = 0; b < 32; ++b) {
                const int t = lane * 32 + b;
                if (t < T && 2 * cc[b] > n) maj |= 1u << b;   // counts > n/2, windows.py:144
            }
        }
        // run starts / ends with carries across words (warp shuffles)
        const uint32_t prev = __shfl_up_sync(0xffffffffu, maj, 1);
        const uint32_t next = __shfl_down_sync(0xffffffffu, maj, 1);
        const uint32_t carry_in = (lane > 0) ? (prev >> 31) : 0u;
        const uint32_t carry_out = (lane < 31) ? (next & 1u) : 0u;
        const uint32_t rs = maj & ~((maj << 1) | carry_in);
        const uint32_t re = maj & ~((maj >> 1) | (carry_out << 31));
        const int nruns = warp_sum(__popc(rs));
        const uint32_t any = __ballot_sync(0xffffffffu, maj != 0);
        int v0 = -1, v1 = -1;
        if (any) {
            const int wlo = __ffs(any) - 1, whi = 31 - __clz(any);
            const uint32_t mlo = __shfl_sync(0xffffffffu, maj, wlo);
            const uint32_t mhi = __shfl_sync(0xffffffffu, maj, whi);
            v0 = wlo * 32 + __ffs(mlo) - 1;
            v1 = whi * 32 + 31 - __clz(mhi);
        }
        if (lane < TW) {
            const int64_t o = d.xbits_off + (int64_t)c * TW + lane;
            majbits[o] = maj;
            rsbits[o] = rs;
            rebits[o] = re;
        }
        if (lane == 0) {
            int32_t* cr = clrow + (d.row0 + c) * 4;
            cr[1] = nruns;
            cr[2] = v0;
            cr[3] = v1;
        }
    }
    // clusterinfo of the video: sizes / windows now, candidate counts by windows_rows_kernel (atomics), status by
    // windows_status_kernel
    __syncthreads();
    if (tid < S2D_MAX_CLUSTERS) {
        int32_t* ci = clusterinfo + ((int64_t)v * S2D_MAX_CLUSTERS + tid) * S2D_CLINFO_WORDS;
        for (int i = 0; i < S2D_CLINFO_WORDS; ++i) ci[i] = 0;
        if (tid < k) {
            const int32_t* cr = clrow + (d.row0 + tid) * 4;
            ci[0] = cr[0];
            ci[2] = cr[2];
            ci[3] = cr[3];
            ci[4] = cr[1];
        }
    }
    if (tid == 0) vi[4] = 0;
}

// phase B: one thread per row (grid over the rows of every video: a 300-frame video has 9 000 of them): winners of every
// run and the candidate run; candidate counts per cluster and per video by atomics
__global__ void __launch_bounds__(WIN_THREADS)
windows_rows_kernel(const s2d_video_desc* __restrict__ descs, const uint32_t* __restrict__ xbits,
                    const int32_t* __restrict__ labels1, const int32_t* __restrict__ qframe, float winner_fraction,
                    const uint32_t* __restrict__ majbits, uint32_t* __restrict__ winbits, int32_t* __restrict__ rowinfo,
                    const int32_t* __restrict__ clrow, int32_t* __restrict__ vidinfo, int32_t* __restrict__ clusterinfo) {
    const int v = blockIdx.y;
    const s2d_video_desc d = descs[v];
    const int q = blockIdx.x * WIN_THREADS + threadIdx.x;
    if (q >= d.Nm) return;
    const int T = d.T, TW = d.TW;
    const int c = labels1[d.row0 + q];
    int cand = -1, v0 = -1, v1 = -1;
    uint32_t* wb = winbits + d.xbits_off + (int64_t)q * TW;
    for (int w = 0; w < TW; ++w) wb[w] = 0;
    if (c >= 0) {
        const uint32_t* xr = xbits + d.xbits_off + (int64_t)q * TW;
        const uint32_t* mj = majbits + d.xbits_off + (int64_t)c * TW;
        const int f = qframe[d.row0 + q];
        v0 = clrow[(d.row0 + c) * 4 + 2];
        v1 = clrow[(d.row0 + c) * 4 + 3];
        int run = -1, cntv = 0, start = 0;
        bool inrun = false;
        for (int t = 0; t <= T; ++t) {
            const bool m = t < T && ((mj[t >> 5] >> (t & 31)) & 1u);
            if (m) {
                if (!inrun) { inrun = true; start = t; cntv = 0; ++run; }
                cntv += (xr[t >> 5] >> (t & 31)) & 1u;
            } else if (inrun) {
                inrun = false;
                const int end = t - 1, len = end - start + 1;
                // frac = counts / length in float32, frac > 0.3 (float32), windows.py:100-103
                if (__fdiv_rn((float)cntv, (float)len) > winner_fraction) {
                    S2D_DEV_ASSERT((run >> 5) < TW);
                    wb[run >> 5] |= 1u << (run & 31);
                    if (f >= start && f <= end) cand = run;
                }
            }
        }
        if (cand >= 0) {
            atomicAdd(&vidinfo[(int64_t)v * S2D_VIDINFO_WORDS + 4], 1);
            if (c < S2D_MAX_CLUSTERS) atomicAdd(&clusterinfo[((int64_t)v * S2D_MAX_CLUSTERS + c) * S2D_CLINFO_WORDS + 1], 1);
        }
    }
    reinterpret_cast<int4*>(rowinfo)[d.row0 + q] = make_int4(c, cand, v0, v1);
}

// phase C: stage-B status of every video
__global__ void windows_status_kernel(int nvideos, int32_t* __restrict__ vidinfo, const int32_t* __restrict__ clusterinfo) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvideos) return;
    int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    const int k = vi[0];
    // load_cluster_masks sorts `cluster_*` folders lexicographically and drops empty ones:
    // the video survives stage D only with 1..10 clusters that all own a candidate
    // (cotracker_matching.py:89-90, 937-938, 1014-1017, 1042-1051; Appendix A.7 quirks 2, 3)
    int ok = (k >= 1 && k <= 10);
    for (int c = 0; c < k && c < S2D_MAX_CLUSTERS; ++c)
        ok &= clusterinfo[((int64_t)v * S2D_MAX_CLUSTERS + c) * S2D_CLINFO_WORDS + 1] > 0;
    vi[1] = ok ? 1 : -1;
    vi[2] = -1;
    vi[3] = -1;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_windows(const s2d_video_desc* descs, int nvideos, int64_t max_rows_x_TW, int max_TW,
                           int64_t total_rows, int64_t total_vt, const uint32_t* xbits,
                           const int32_t* labels1, const int32_t* qframe, float winner_fraction,
                           int32_t* ccount, int32_t* clrow, uint32_t* majbits, uint32_t* rsbits,
                           uint32_t* rebits, uint32_t* winbits, int32_t* rowinfo, int32_t* vidinfo,
                           int32_t* clusterinfo, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(descs && xbits && labels1 && qframe && ccount && clrow && majbits && rsbits && rebits &&
                      winbits && rowinfo && vidinfo && clusterinfo, "s2d_windows: null pointer");
    S2D_CHECK_ARG(nvideos > 0 && nvideos <= 65535, "s2d_windows: bad nvideos %d", nvideos);
    S2D_CHECK_ARG(max_TW >= 1 && max_TW <= 32, "s2d_windows: videos longer than 1024 frames are not supported (TW=%d)", max_TW);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ccount, 0, (size_t)total_vt * sizeof(int32_t), st);
    cudaMemsetAsync(clrow, 0, (size_t)total_rows * 4 * sizeof(int32_t), st);
    dim3 grid((unsigned)((max_rows_x_TW + 255) / 256), nvideos);
    cluster_count_kernel<<<grid, 256, 0, st>>>(descs, xbits, labels1, ccount, clrow);
    S2D_CHECK_LAUNCH("cluster_count_kernel");
    windows_kernel<<<nvideos, WIN_THREADS, 0, st>>>(descs, xbits, labels1, qframe, ccount, winner_fraction,
                                                   majbits, rsbits, rebits, winbits, rowinfo, clrow, vidinfo,
                                                   clusterinfo);
    S2D_CHECK_LAUNCH("windows_kernel");
    const int64_t max_rows = max_rows_x_TW;      // an upper bound of every video's row count (rows x words per row, words >= 1)
    dim3 rgrid((unsigned)((max_rows + WIN_THREADS - 1) / WIN_THREADS), nvideos);
    windows_rows_kernel<<<rgrid, WIN_THREADS, 0, st>>>(descs, xbits, labels1, qframe, winner_fraction, majbits, winbits, rowinfo,
                                                     clrow, vidinfo, clusterinfo);
    S2D_CHECK_LAUNCH("windows_rows_kernel");
    windows_status_kernel<<<(nvideos + 127) / 128, 128, 0, st>>>(nvideos, vidinfo, clusterinfo);
    S2D_CHECK_LAUNCH("windows_status_kernel");
    return 0;
}
