// K4: coherence scoring, keymask selection and temporal-correspondence grouping.
//   select : iou = hits/uniq (float64 like the reference's python ints), match bits, one-to-many
//            flags (cotracker_matching.py:692-717, 1082-1111)
//   group  : per visibility cluster, bounding-box crop of the match matrix, Hamming DBSCAN #2
//            with the reference's eps/min_samples table, zero rows -> -1, factor, coverage and
//            one2x sums (cotracker_matching.py:764-840, 843-921)
#include "dbscan.cuh"

#include <limits.h>

namespace s2d {

// ------------------------------------------------------------------------------------------
// select: one warp per candidate query
// ------------------------------------------------------------------------------------------
constexpr int SEL_WARPS = 4;

// One warp per candidate query, lane = label: for every frame of the window the lanes load the frame's hit counts
// (one coalesced row of hits[q, t, :]) and global ids, the union count is a broadcast load. The number of masks with
// iou > one2x_iou in a frame is a ballot + popcount (no shared memory, no atomics), match bits are rare atomicOr's
// into the query's bit row. Four frames are in flight per step (independent loads). Frames for which no tracks are
// stored (windowed storage) count as intersection 0 / union 0.
__global__ void __launch_bounds__(SEL_WARPS * 32)
select_kernel(const s2d_video_desc* __restrict__ descs, const int32_t* __restrict__ hits,
              const int32_t* __restrict__ uniq, const int32_t* __restrict__ gid_of,
              const int32_t* __restrict__ rowinfo, double match_thr, double one2x_iou,
              int one2x_frames, uint32_t* __restrict__ mbits, int32_t* __restrict__ one2x,
              int32_t* __restrict__ nmatch, int32_t* __restrict__ vidinfo) {
    const int v = blockIdx.y;
    const s2d_video_desc d = descs[v];
    const int q = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
    if (q >= d.Nm) return;
    const int lane = threadIdx.x & 31;
    int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    const int4 ri = reinterpret_cast<const int4*>(rowinfo)[d.row0 + q];
    if (vi[1] < 0 || ri.y < 0) {
        if (lane == 0) { one2x[d.row0 + q] = 0; nmatch[d.row0 + q] = 0; }
        return;
    }
    uint32_t* mrow = mbits + d.mbits_off + (int64_t)q * d.NW;
    const int L = d.L;
    const int ts0 = d.tstart ? d.tstart[q] : 0, ts1 = ts0 + d.Ttr;        // frames that carry votes
    const int t0 = max(ri.z, 0), t1 = min(ri.w, d.T - 1);
    const int32_t* hq = hits + d.hits_off + (int64_t)q * d.T * L;
    const int32_t* uq = uniq + d.vt_off + (int64_t)q * d.T;
    const int32_t* gq = gid_of + d.frame0 * S2D_MAX_LABELS;
    int warn = 0, nm = 0, maxg = -1;
    // iou = intersection / union as python floats, 0.0 when union == 0 (matching.py:659-662). The comparison iou > thr
    // is decided without the division whenever I - thr * U is clearly away from zero (one FMA; a gap of 1e-12 * U is
    // thousands of ulps of the quotient), and by the reference's own expression otherwise.
    auto score = [&](int I, int U, int gid, int& c25) {
        if (gid < 0 || gid >= d.Nm) return;       // not an object (or a label map that enumerates more objects than Nm rows)
        const double fI = (double)I, fU = (double)U, tol = 1e-12 * fU;
        const double dm = fma(-match_thr, fU, fI), d2 = fma(-one2x_iou, fU, fI);
        bool m1, m2;
        if (U == 0) { m1 = 0.0 > match_thr; m2 = 0.0 > one2x_iou; }
        else {
            m1 = dm > tol ? true : (dm < -tol ? false : (fI / fU > match_thr));
            m2 = d2 > tol ? true : (d2 < -tol ? false : (fI / fU > one2x_iou));
        }
        if (m1) {
            S2D_DEV_ASSERT((gid >> 5) < d.NW);
            atomicOr(&mrow[gid >> 5], 1u << (gid & 31));
            ++nm;
            maxg = max(maxg, gid);
        }
        c25 += m2 ? 1 : 0;
    };
    if (L <= 32) {
        constexpr int FR = 4;                      // frames in flight
        for (int tb = t0; tb <= t1; tb += FR) {
            int I[FR], U[FR], g[FR];
#pragma unroll
            for (int k = 0; k < FR; ++k) {         // all loads of the step first (clamped: always inside the query's rows)
                const int t = min(tb + k, t1);
                I[k] = lane < L ? hq[(int64_t)t * L + lane] : 0;
                g[k] = lane < L ? gq[(int64_t)t * S2D_MAX_LABELS + lane] : -1;
                U[k] = uq[t];
            }
#pragma unroll
            for (int k = 0; k < FR; ++k) {
                const int t = tb + k;
                if (t > t1) break;                 // warp-uniform
                if (t < ts0 || t >= ts1) { I[k] = 0; U[k] = 0; }
                int c25 = 0;
                score(I[k], U[k], g[k], c25);
                warn += (__popc(__ballot_sync(0xffffffffu, c25 != 0)) > 1) ? 1 : 0;     // same value in every lane
            }
        }
        warn *= (lane == 0) ? 1 : 0;               // counted once
    } else {
        for (int t = t0; t <= t1; ++t) {
            const bool stored = t >= ts0 && t < ts1;
            const int U = stored ? uq[t] : 0;
            int c25 = 0;
            for (int l = lane; l < L; l += 32)
                score(stored ? hq[(int64_t)t * L + l] : 0, U, gq[(int64_t)t * S2D_MAX_LABELS + l], c25);
            c25 = warp_sum(c25);
            warn += (lane == 0 && c25 > 1) ? 1 : 0;
        }
    }
    warn = warp_sum(warn);
    nm = warp_sum(nm);
    maxg = __reduce_max_sync(0xffffffffu, maxg);
    if (lane == 0) {
        one2x[d.row0 + q] = warn >= one2x_frames ? 1 : 0;
        nmatch[d.row0 + q] = nm;
        if (maxg >= 0) atomicMax(&vi[2], maxg);
    }
}

// ------------------------------------------------------------------------------------------
// group prep: one CTA per (video, cluster): crop box, DBSCAN #2 parameters, problem descriptor
// per-video scratch layout in `work` (int32): for cluster c, base = 5*(16*row0 + c*Nm):
//   core[Nm] parent[Nm] aux[Nm] labels[Nm] valid[Nm bytes, padded to Nm ints]
// ------------------------------------------------------------------------------------------
constexpr int GP_THREADS = 256;

__device__ __forceinline__ int block_reduce(int v, bool is_min, int* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_min ? __reduce_min_sync(0xffffffffu, v) : __reduce_max_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    int r = sm[0];
    for (int w = 1; w < GP_THREADS / 32; ++w) r = is_min ? min(r, sm[w]) : max(r, sm[w]);
    return r;
}

__global__ void __launch_bounds__(GP_THREADS)
group_prep_kernel(const s2d_video_desc* __restrict__ descs, const uint32_t* __restrict__ mbits,
                  const int32_t* __restrict__ rowinfo, int32_t* __restrict__ work,
                  int32_t* __restrict__ vidinfo, int32_t* __restrict__ clusterinfo,
                  DbProblem* __restrict__ problems, const int32_t* __restrict__ gram, const int64_t* __restrict__ gram_off) {
    const int v = blockIdx.y, c = blockIdx.x;
    const s2d_video_desc d = descs[v];
    const int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    int32_t* ci = clusterinfo + ((int64_t)v * S2D_MAX_CLUSTERS + c) * S2D_CLINFO_WORDS;
    DbProblem* prob = problems + (int64_t)v * S2D_MAX_CLUSTERS + c;
    __shared__ int sm[GP_THREADS / 32];
    const int tid = threadIdx.x;
    const bool active = vi[1] > 0 && c < vi[0];
    const int maxid = vi[2];

    int32_t* base = work + 5 * (16 * d.row0 + (int64_t)c * d.Nm);
    uint8_t* valid = reinterpret_cast<uint8_t*>(base + 4 * (int64_t)d.Nm);

    int rmin = INT_MAX, rmax = -1, cmin = INT_MAX, cmax = -1;
    if (active) {
        for (int q = tid; q < d.Nm; q += GP_THREADS) {
            const int4 ri = reinterpret_cast<const int4*>(rowinfo)[d.row0 + q];
            // rows above the largest matched id are skipped with a warning (matching.py:784-786)
            const bool mine = ri.x == c && ri.y >= 0 && q <= maxid;
            valid[q] = mine ? 1 : 0;
            if (!mine) continue;
            const uint32_t* mrow = mbits + d.mbits_off + (int64_t)q * d.NW;
            int first = -1, last = -1;
            for (int w = 0; w < d.NW; ++w) {
                const uint32_t m = mrow[w];
                if (m) {
                    if (first < 0) first = w * 32 + __ffs(m) - 1;
                    last = w * 32 + 31 - __clz(m);
                }
            }
            if (first >= 0) {
                rmin = min(rmin, q); rmax = max(rmax, q);
                cmin = min(cmin, first); cmax = max(cmax, last);
            }
        }
    }
    rmin = block_reduce(rmin, true, sm);
    rmax = block_reduce(rmax, false, sm);
    cmin = block_reduce(cmin, true, sm);
    cmax = block_reduce(cmax, false, sm);
    if (tid == 0) {
        DbProblem p;
        p.bits = nullptr; p.valid = nullptr; p.core = base; p.parent = base + d.Nm; p.aux = base + 2 * (int64_t)d.Nm;
        p.labels = base + 3 * (int64_t)d.Nm; p.nclusters = nullptr;
        p.gram = nullptr; p.gstride = 0; p.pad = 0;
        p.stride = d.NW; p.w0 = 0; p.nw = 0; p.N = 0; p.kmax = 0; p.min_samples = 1;
        ci[5] = -1; ci[6] = -1; ci[7] = -1; ci[8] = -1; ci[9] = 0; ci[10] = 0;
        if (active && rmax >= 0) {
            const int ncols = cmax - cmin + 1;
            double eps; int ms;                      // matching.py:795-803
            if (ncols > 50) { eps = 0.05; ms = 5; }
            else if (ncols < 10) { eps = 0.1; ms = 3; }
            else { eps = 0.1; ms = 5; }
            p.bits = mbits + d.mbits_off + (int64_t)rmin * d.NW;
            p.valid = valid + rmin;
            p.core += rmin; p.parent += rmin; p.aux += rmin; p.labels += rmin;   // keep absolute row indexing
            if (gram && gram_off && gram_off[v] >= 0) {      // the video's Gram matrix [Nm][Nm] of its match rows
                p.gram = gram + gram_off[v] + (int64_t)rmin * d.Nm + rmin;
                p.gstride = d.Nm;
            }
            p.w0 = cmin >> 5;
            p.nw = (cmax >> 5) - p.w0 + 1;
            p.N = rmax - rmin + 1;
            p.kmax = hamming_kmax(ncols, eps);
            p.min_samples = ms;
            ci[5] = rmin; ci[6] = rmax; ci[7] = cmin; ci[8] = cmax; ci[9] = p.kmax; ci[10] = ms;
        }
        *prob = p;
    }
}

// ------------------------------------------------------------------------------------------
// group finalize: one CTA per (video, cluster): zero rows -> -1, factor, coverage, one2x sums,
// per-group sums; then (cluster 0's CTA, last) the video status.
// grp arrays: [16*row0 + c*Nm + label] -> grp_n, grp_one2x
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GP_THREADS)
group_finalize_kernel(const s2d_video_desc* __restrict__ descs, const uint32_t* __restrict__ mbits,
                      const int32_t* __restrict__ rowinfo, const int32_t* __restrict__ one2x,
                      const int32_t* __restrict__ work, int32_t* __restrict__ glabel,
                      int32_t* __restrict__ grp_n, int32_t* __restrict__ grp_one2x,
                      int32_t* __restrict__ clusterinfo, const int32_t* __restrict__ vidinfo) {
    const int v = blockIdx.y, c = blockIdx.x;
    const s2d_video_desc d = descs[v];
    const int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    if (!(vi[1] > 0 && c < vi[0])) return;
    int32_t* ci = clusterinfo + ((int64_t)v * S2D_MAX_CLUSTERS + c) * S2D_CLINFO_WORDS;
    const int rmin = ci[5], rmax = ci[6];
    const int32_t* base = work + 5 * (16 * d.row0 + (int64_t)c * d.Nm);
    const int32_t* labels = base + 3 * (int64_t)d.Nm;      // indexed by absolute row
    const uint8_t* valid = reinterpret_cast<const uint8_t*>(base + 4 * (int64_t)d.Nm);
    int32_t* gn = grp_n + 16 * d.row0 + (int64_t)c * d.Nm;
    int32_t* go = grp_one2x + 16 * d.row0 + (int64_t)c * d.Nm;
    __shared__ int s_factor, s_matched, s_o2sum, s_nq;
    if (threadIdx.x == 0) { s_factor = 0; s_matched = 0; s_o2sum = 0; s_nq = 0; }
    __syncthreads();
    for (int q = threadIdx.x; q < d.Nm; q += GP_THREADS) {
        const int4 ri = reinterpret_cast<const int4*>(rowinfo)[d.row0 + q];
        if (ri.x != c) continue;
        int lab = -1;
        if (ri.y >= 0) {
            atomicAdd(&s_nq, 1);
            atomicAdd(&s_o2sum, one2x[d.row0 + q]);
            if (rmax >= 0 && q >= rmin && q <= rmax && valid[q]) {
                const uint32_t* mrow = mbits + d.mbits_off + (int64_t)q * d.NW;
                bool any = false;
                for (int w = 0; w < d.NW && !any; ++w) any = mrow[w] != 0;
                if (any) lab = labels[q];                       // zero rows are forced to -1 (:813-815)
            }
        }
        glabel[d.row0 + q] = lab;
        S2D_DEV_ASSERT(lab < d.Nm);
        if (lab >= 0) {
            atomicAdd(&s_matched, 1);
            if (atomicAdd(&gn[lab], 1) == 0) atomicAdd(&s_factor, 1);
            atomicAdd(&go[lab], one2x[d.row0 + q]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ci[11] = s_factor;
        ci[12] = s_matched;
        ci[13] = s_o2sum;
        ci[14] = s_nq;
    }
}

__global__ void video_status_kernel(const s2d_video_desc* __restrict__ descs, int nvideos,
                                    const int32_t* __restrict__ clusterinfo, int32_t* __restrict__ vidinfo) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvideos) return;
    int32_t* vi = vidinfo + (int64_t)v * S2D_VIDINFO_WORDS;
    int ok = vi[1] > 0;
    if (ok) {
        // any cluster whose cropped match matrix is empty fails the video (matching.py:806-807)
        for (int c = 0; c < vi[0]; ++c)
            ok &= clusterinfo[((int64_t)v * S2D_MAX_CLUSTERS + c) * S2D_CLINFO_WORDS + 6] >= 0;
    }
    vi[3] = ok ? 1 : -1;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_select(const s2d_video_desc* descs, int nvideos, int max_Nm, int64_t total_mbits_words,
                          const int32_t* hits, const int32_t* uniq, const int32_t* gid_of,
                          const int32_t* rowinfo, double matching_threshold, double one2x_iou,
                          int one2x_frames, uint32_t* mbits, int32_t* one2x, int32_t* nmatch,
                          int32_t* vidinfo, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(descs && hits && uniq && gid_of && rowinfo && mbits && one2x && nmatch && vidinfo,
                  "s2d_select: null pointer");
    S2D_CHECK_ARG(nvideos > 0 && nvideos <= 65535 && max_Nm > 0, "s2d_select: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(mbits, 0, (size_t)total_mbits_words * sizeof(uint32_t), st);
    dim3 grid((max_Nm + SEL_WARPS - 1) / SEL_WARPS, nvideos);
    select_kernel<<<grid, SEL_WARPS * 32, 0, st>>>(descs, hits, uniq, gid_of, rowinfo, matching_threshold,
                                                  one2x_iou, one2x_frames, mbits, one2x, nmatch, vidinfo);
    S2D_CHECK_LAUNCH("select_kernel");
    return 0;
}

extern "C" int s2d_group_work_ints(int64_t total_rows, int nvideos, int64_t* out) {
    if (!out) return -1;
    *out = 5 * 16 * total_rows + 2 + (int64_t)nvideos * S2D_MAX_CLUSTERS * (int64_t)(sizeof(DbProblem) / 4) +
           db_worklist_ints(nvideos * S2D_MAX_CLUSTERS);
    return 0;
}

extern "C" int s2d_group(const s2d_video_desc* descs, int nvideos, int max_Nm, int max_NW, int64_t total_rows,
                         const uint32_t* mbits, const int32_t* rowinfo, const int32_t* one2x,
                         int32_t* work, int32_t* glabel, int32_t* grp_n, int32_t* grp_one2x,
                         int32_t* vidinfo, int32_t* clusterinfo, void* stream) {
    return s2d_group_gram(descs, nvideos, max_Nm, max_NW, total_rows, mbits, rowinfo, one2x, work, glabel, grp_n, grp_one2x,
                          vidinfo, clusterinfo, nullptr, nullptr, stream);
}

extern "C" int s2d_group_gram(const s2d_video_desc* descs, int nvideos, int max_Nm, int max_NW, int64_t total_rows,
                              const uint32_t* mbits, const int32_t* rowinfo, const int32_t* one2x,
                              int32_t* work, int32_t* glabel, int32_t* grp_n, int32_t* grp_one2x,
                              int32_t* vidinfo, int32_t* clusterinfo, const int32_t* gram, const int64_t* gram_off,
                              void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(descs && mbits && rowinfo && one2x && work && glabel && grp_n && grp_one2x && vidinfo && clusterinfo,
                  "s2d_group: null pointer");
    S2D_CHECK_ARG(nvideos > 0 && nvideos <= 65535 && max_Nm > 0, "s2d_group: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t off = 5 * 16 * total_rows;
    off += off & 1;
    DbProblem* problems = reinterpret_cast<DbProblem*>(work + off);
    S2D_CHECK_ARG((((uintptr_t)problems) & 7) == 0, "s2d_group: work must be 8-byte aligned");
    cudaMemsetAsync(glabel, 0xFF, (size_t)total_rows * sizeof(int32_t), st);   // -1: noise / not grouped
    cudaMemsetAsync(grp_n, 0, (size_t)total_rows * 16 * sizeof(int32_t), st);
    cudaMemsetAsync(grp_one2x, 0, (size_t)total_rows * 16 * sizeof(int32_t), st);
    dim3 grid(S2D_MAX_CLUSTERS, nvideos);
    group_prep_kernel<<<grid, GP_THREADS, 0, st>>>(descs, mbits, rowinfo, work, vidinfo, clusterinfo, problems, gram, gram_off);
    S2D_CHECK_LAUNCH("group_prep_kernel");
    int rc = launch_dbscan(problems, nvideos * S2D_MAX_CLUSTERS, max_Nm, max_NW,
                           reinterpret_cast<int32_t*>(problems + (int64_t)nvideos * S2D_MAX_CLUSTERS), st);
    if (rc) return rc;
    group_finalize_kernel<<<grid, GP_THREADS, 0, st>>>(descs, mbits, rowinfo, one2x, work, glabel, grp_n,
                                                      grp_one2x, clusterinfo, vidinfo);
    S2D_CHECK_LAUNCH("group_finalize_kernel");
    video_status_kernel<<<(nvideos + 127) / 128, 128, 0, st>>>(descs, nvideos, clusterinfo, vidinfo);
    S2D_CHECK_LAUNCH("video_status_kernel");
    return 0;
}
