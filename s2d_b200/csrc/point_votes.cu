// K2: point-in-mask voting (the dominant, HBM-bound kernel of the path).
//
// One CTA per (query q, frame t) tile of P tracked points (8 B each, read exactly once with
// 128-bit streaming loads). Rounded, in-bounds pixels are de-duplicated in a shared-memory
// bitmap that covers a band of whole frame rows starting at the tile's first occupied row
// (word index XOR-swizzled so that regular grids of points spread over the banks); the first
// thread to set a pixel's bit votes the pixel's label (u8 label map, L2/L1 resident: every
// query of a video re-reads the same T frames) into a shared histogram through per-thread
// run-length and warp-level aggregation.
// Output per tile: hits[q,t,0..L) and uniq[q,t] - 4(L+1) bytes against 8P bytes read.
//
// Replaces pred_tracks_to_binary_masks + compute_point_mask_intersection over every mask of
// the frame (cotracker_matching.py:453-503, 640-662, 681-692): uniq = |P| = union,
// hits[lab] = |P AND mask_lab| = intersection.
#include "common.cuh"

#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

namespace s2d {

constexpr int PV_BM_WORDS = 8192;                 // 32 KB bitmap = 262144 pixels per band
constexpr int PV_BM_BITS = PV_BM_WORDS * 32;
constexpr int PV_WORD_SHIFT = 13;                 // log2(PV_BM_WORDS)
constexpr uint32_t PV_INVALID = 0xFFFFFFFFu;

// Rounded pixel of one track point as a frame-linear index (iy*W + ix), PV_INVALID if the point
// does not land inside the frame. cvt.rni is round-half-to-even like torch.round(). fmaxf(v,-1)
// maps NaN (and everything below -1) to -1, i.e. out of bounds; +inf / huge values saturate to
// INT_MAX and fail the unsigned compare (the reference's int64 conversion drops all of them too).
__device__ __forceinline__ uint32_t pv_lin(float x, float y, uint32_t W, uint32_t H) {
    const uint32_t ix = (uint32_t)__float2int_rn(fmaxf(x, -1.0f));
    const uint32_t iy = (uint32_t)__float2int_rn(fmaxf(y, -1.0f));
    return (ix < W && iy < H) ? iy * W + ix : PV_INVALID;
}

// PTX shl clamps shift amounts above 31: the result is 0 for them, which is exactly "not in
// this band" below.
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}

// Try to claim pixel `lin` in the bitmap band that starts at linear index `base` and covers
// PV_BM_BITS pixels. Returns nonzero iff this call set the pixel's bit (first point on it).
// Pixels outside the band (kk >= 2^18, including PV_INVALID and wrapped negatives) OR in 0.
__device__ __forceinline__ uint32_t pv_claim(uint32_t* bm, uint32_t lin, uint32_t base) {
    const uint32_t kk = lin - base;
    const uint32_t w = (kk ^ (kk >> 5)) & (PV_BM_WORDS - 1);   // word index, XOR bank swizzle (bijective)
    const uint32_t bit = shl_clamp(1u, kk >> PV_WORD_SHIFT);
    const uint32_t old = atomicOr(&bm[w], bit);
    return bit & ~old;
}
// same with the bitmap given as a 32-bit shared-window address (no generic -> shared conversion)
__device__ __forceinline__ uint32_t pv_claim_s(uint32_t bm_saddr, uint32_t lin, uint32_t base) {
    const uint32_t kk = lin - base;
    const uint32_t w4 = ((kk ^ (kk >> 5)) & (PV_BM_WORDS - 1)) << 2;
    const uint32_t bit = shl_clamp(1u, kk >> PV_WORD_SHIFT);
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(bm_saddr + w4), "r"(bit) : "memory");
    return bit & ~old;
}

// same for a bitmap of 2^LOGW words (the label-table kernel's fallback sizes its bitmap to its buffer)
template <int LOGW>
__device__ __forceinline__ uint32_t pv_claim_t(uint32_t bm_saddr, uint32_t lin, uint32_t base) {
    const uint32_t kk = lin - base;
    const uint32_t w4 = ((kk ^ (kk >> 5)) & ((1u << LOGW) - 1u)) << 2;
    const uint32_t bit = shl_clamp(1u, kk >> LOGW);
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(bm_saddr + w4), "r"(bit) : "memory");
    return bit & ~old;
}

// expand the low 4 bits of m into a byte mask (bit i -> byte i = 0xFF)
__device__ __forceinline__ uint32_t nib2bytes(uint32_t m) {
    return (((m & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
}

// Votes of one thread: `first` has bit k set when point k was the first on its pixel, lab4 holds
// the points' labels (4 per word). Runs of equal labels are accumulated in (cur, cnt) and only
// flushed to the shared histogram when the label changes.
template <int PPT>
__device__ __forceinline__ void pv_vote(uint32_t first, const uint32_t (&lab4)[PPT / 4], int& cur, int& cnt, int* hist) {
#pragma unroll
    for (int k4 = 0; k4 < PPT / 4; ++k4) {
        const uint32_t f = (first >> (4 * k4)) & 0xFu;
        if (f == 0) continue;
        const uint32_t l4 = lab4[k4];
        if (cur < 0) cur = (l4 >> (8 * (__ffs(f) - 1))) & 255;
        const uint32_t diff = (l4 ^ ((uint32_t)cur * 0x01010101u)) & nib2bytes(f);
        if (diff == 0) {
            cnt += __popc(f);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!((f >> j) & 1u)) continue;
                const int lab = (l4 >> (8 * j)) & 255;
                if (lab != cur) {
                    if (cnt) atomicAdd(&hist[cur], cnt);
                    cur = lab; cnt = 0;
                }
                ++cnt;
            }
        }
    }
}

// warp-aggregated flush of the per-thread runs (usually one label per warp)
__device__ __forceinline__ void pv_flush(int cur, int cnt, int lane, int* hist) {
    uint32_t remaining = __ballot_sync(0xffffffffu, cnt > 0);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int l0 = __shfl_sync(0xffffffffu, cur, leader);
        const bool mine = (cnt > 0) && (cur == l0);
        const int sm = __reduce_add_sync(0xffffffffu, mine ? cnt : 0);
        if (lane == leader) atomicAdd(&hist[l0], sm);
        remaining &= ~__ballot_sync(0xffffffffu, mine);
    }
}

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// ==========================================================================================
// One CTA per (query, frame) tile. Fallback for odd P / unaligned tracks / P > 8192, and the
// path used when the caller passes no work buffer.
// ==========================================================================================
template <int THREADS, int PPT, bool VEC4>
__global__ void __launch_bounds__(THREADS)
point_votes_kernel(const s2d_video_desc* __restrict__ descs, const int32_t* __restrict__ rowinfo,
                   const int32_t* __restrict__ vidinfo, int32_t* __restrict__ hits,
                   int32_t* __restrict__ uniq) {
    const s2d_video_desc* dp = descs + blockIdx.z;
    const int T = dp->T, Nm = dp->Nm;
    const int q = blockIdx.y, t = blockIdx.x;
    if (q >= Nm || t >= T) return;
    if (vidinfo && vidinfo[(int64_t)blockIdx.z * S2D_VIDINFO_WORDS + 1] < 0) return;
    const int64_t rt = (int64_t)q * T + t;
    if (rowinfo) {
        const int4 ri = reinterpret_cast<const int4*>(rowinfo)[dp->row0 + q];
        if (ri.y < 0 || t < ri.z || t > ri.w) return;
    }
    const int P = dp->P;
    const uint32_t W = dp->W, H = dp->H;
    const int32_t* np = dp->npts;
    const int n = np ? min(max(np[q], 0), P) : P;

    __shared__ __align__(16) uint32_t bm[PV_BM_WORDS];
    __shared__ int hist[S2D_MAX_LABELS + 1];          // [256] = number of unique pixels
    __shared__ uint32_t sbox[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int32_t* tsp = dp->tstart;
    const int ts = t - (tsp ? tsp[q] : 0);                 // frame index inside the stored track window
    if (ts < 0 || ts >= dp->Ttr) return;
    const float* tp = dp->tracks + ((int64_t)q * dp->Ttr + ts) * P * 2;
    const uint8_t* lbl = dp->labels + (int64_t)t * (int64_t)(W * H);

    // ---- pass 1: all loads of the thread in flight at once, then round / bounds ----------
    uint32_t lin[PPT];
    if (VEC4) {
        int4 raw[PPT / 2];
        const int last4 = max(P / 2 - 1, 0);               // clamp: every load stays inside the tile
#pragma unroll
        for (int i = 0; i < PPT / 2; ++i)
            raw[i] = ld_stream(reinterpret_cast<const int4*>(tp) + min(i * THREADS + tid, last4));
#pragma unroll
        for (int i = 0; i < PPT / 2; ++i) {
            const int p0 = 2 * (i * THREADS + tid);
            const uint32_t a = pv_lin(__int_as_float(raw[i].x), __int_as_float(raw[i].y), W, H);
            const uint32_t b = pv_lin(__int_as_float(raw[i].z), __int_as_float(raw[i].w), W, H);
            lin[2 * i] = (p0 < n) ? a : PV_INVALID;
            lin[2 * i + 1] = (p0 + 1 < n) ? b : PV_INVALID;
        }
    } else {
        float2 raw[PPT];
        const int last = max(P - 1, 0);
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int p = 2 * ((i >> 1) * THREADS + tid) + (i & 1);
            raw[i] = __ldg(reinterpret_cast<const float2*>(tp) + min(p, last));
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int p = 2 * ((i >> 1) * THREADS + tid) + (i & 1);
            const uint32_t a = pv_lin(raw[i].x, raw[i].y, W, H);
            lin[i] = (p < n) ? a : PV_INVALID;
        }
    }
    uint32_t labr[PPT];
    uint32_t lmin = PV_INVALID, lmax1 = 0;       // min of lin, max of lin+1 (invalid -> 0)
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const uint32_t l = lin[k];
        labr[k] = 0;
        if (l != PV_INVALID) labr[k] = __ldg(lbl + l);
        lmin = min(lmin, l);
        lmax1 = max(lmax1, l + 1u);
    }
#pragma unroll
    for (int i = 0; i < PV_BM_WORDS / 4 / THREADS; ++i)
        reinterpret_cast<uint4*>(bm)[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < S2D_MAX_LABELS + 1; i += THREADS) hist[i] = 0;
    if (tid == 0) { sbox[0] = PV_INVALID; sbox[1] = 0; }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax1 = __reduce_max_sync(0xffffffffu, lmax1);
    __syncthreads();                      // smem init visible
    if (lane == 0 && lmax1 != 0) { atomicMin(&sbox[0], lmin); atomicMax(&sbox[1], lmax1); }
    __syncthreads();
    lmin = sbox[0];
    lmax1 = sbox[1];
    uint32_t lab4[PPT / 4];
#pragma unroll
    for (int k4 = 0; k4 < PPT / 4; ++k4) lab4[k4] = pack4(labr[4 * k4], labr[4 * k4 + 1], labr[4 * k4 + 2], labr[4 * k4 + 3]);

    // ---- pass 2: de-duplicate in the bitmap band by band, vote -----------------------------
    if (lmax1 != 0) {
        int cur = -1, cnt = 0, nfirst = 0;
        for (uint32_t base = lmin;;) {
            uint32_t first = 0;
#pragma unroll
            for (int k = 0; k < PPT; ++k) first |= pv_claim(bm, lin[k], base) ? (1u << k) : 0u;
            nfirst += __popc(first);
            pv_vote<PPT>(first, lab4, cur, cnt, hist);
            if (lmax1 - base <= (uint32_t)PV_BM_BITS) break;       // uniform: extent is CTA-wide
            base += PV_BM_BITS;
            __syncthreads();                                       // recycle the bitmap for the next band
#pragma unroll
            for (int i = 0; i < PV_BM_WORDS / 4 / THREADS; ++i)
                reinterpret_cast<uint4*>(bm)[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
        pv_flush(cur, cnt, lane, hist);
        nfirst = __reduce_add_sync(0xffffffffu, nfirst);
        if (lane == 0 && nfirst) atomicAdd(&hist[S2D_MAX_LABELS], nfirst);
    }
    __syncthreads();

    // ---- write out ------------------------------------------------------------------------
    const int L = dp->L;
    int32_t* hout = hits + dp->hits_off + rt * L;
    for (int l = tid; l < S2D_MAX_LABELS + 1; l += THREADS) {
        const int h = hist[l];
        if (l < L) hout[l] = h;
        if (l == S2D_MAX_LABELS) uniq[dp->vt_off + rt] = h;
    }
}

// ==========================================================================================
// Persistent, TMA-fed bitmap variant (round-1 first design; compiled only with -DS2D_EXPERIMENTS).
//
// A small plan pass turns (candidate rows x window frames) into a flat list of tiles; 2 CTAs per
// SM pull chunks of consecutive tiles from an atomic counter. Each CTA has one producer warp that
// keeps a two-stage ring of 8P-byte track tiles in shared memory filled by cp.async.bulk (one
// instruction per tile, completion on an mbarrier) and THREADS consumer threads that software-
// pipeline over tiles: while tile i is de-duplicated and voted, tile i+1 has already been pulled
// out of its stage into registers and its label gathers are in flight.
// ==========================================================================================
struct PvTile {             // written by the producer thread, read by everybody after the mbarrier wait
    const uint8_t* lbl;     // label map of the target frame
    int32_t* hout;          // hits[q,t,:]
    int32_t* uout;          // &uniq[q,t]
    uint32_t W, H;
    int32_t n, L;
    int32_t valid, pad;     // pad: P of the tile in the label-table kernel
    const uint8_t* tm;      // label-table kernel: the video's TMA descriptors (s2d_point_votes_tmaps) or null
    uint32_t ybase, pad2;   // t * H: first row of the frame in the descriptors' [T*H][W] view
    const float* src;       // label-table kernel: tracks[q, t] of the tile
};

struct PvOut { int32_t* hout; int32_t* uout; int32_t L, pad; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: park instead of spinning
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// plan: per batch row (tile0 filled by the scan, ntiles, first frame, video)
__global__ void pv_plan_rows_kernel(const s2d_video_desc* __restrict__ descs, int nvideos, int64_t total_rows,
                                    const int32_t* __restrict__ rowinfo, const int32_t* __restrict__ vidinfo,
                                    int4* __restrict__ rowplan) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total_rows) return;
    int lo = 0, hi = nvideos;                       // last video with row0 <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (descs[mid].row0 <= r) lo = mid; else hi = mid;
    }
    const int T = descs[lo].T;
    int nt = T, t0 = 0;
    const int32_t* tsp = descs[lo].tstart;
    const int ts0 = tsp ? tsp[r - descs[lo].row0] : 0;     // frames [ts0, ts0 + Ttr) have tracks
    if (vidinfo && vidinfo[(int64_t)lo * S2D_VIDINFO_WORDS + 1] < 0) nt = 0;
    if (rowinfo) {
        const int4 ri = reinterpret_cast<const int4*>(rowinfo)[r];
        if (ri.y < 0) nt = 0;
        else { t0 = max(ri.z, 0); nt = nt ? max(min(ri.w, T - 1) - t0 + 1, 0) : 0; }
    }
    if (nt) {                                              // clip to the stored track window
        const int a = max(t0, ts0), b = min(t0 + nt, ts0 + descs[lo].Ttr);
        t0 = a; nt = max(b - a, 0);
    }
    rowplan[r] = make_int4(0, nt, t0, lo);
}

// exclusive scan of ntiles over all batch rows (one CTA), total -> ctrl[1], work counter ctrl[0] = 0.
// Each thread owns a contiguous run of rows: one pass to sum it, one block scan, one pass to write the offsets
// (three barriers in all instead of three per 1024 rows).
__global__ void __launch_bounds__(1024) pv_scan_kernel(int4* __restrict__ rowplan, int64_t total_rows, int32_t* __restrict__ ctrl) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (total_rows + 1023) / 1024;
    const int64_t rb = min(total_rows, (int64_t)tid * per), re = min(total_rows, rb + per);
    int s = 0;
#pragma unroll 4
    for (int64_t r = rb; r < re; ++r) s += rowplan[r].y;
    int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int t = wsum[w]; if (w < warp) before += t; total += t; }
    int run = before + x - s;
    for (int64_t r = rb; r < re; ++r) { const int v = rowplan[r].y; rowplan[r].x = run; run += v; }
    if (tid == 0) { ctrl[0] = 0; ctrl[1] = total; }
}

constexpr int PV_CHUNK = 16;       // consecutive tiles claimed per atomic

__device__ __forceinline__ void consumer_sync(int nthreads) {   // named barrier 1: consumer warps only
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

#ifdef S2D_EXPERIMENTS   // superseded round-1 design (bitmap + global gathers), kept for A/B runs: `make exp`
// THREADS consumer threads + one producer warp (the last warp of the CTA)
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS + 32, (THREADS * PPT <= 1024) ? 4 : ((THREADS * PPT <= 4096) ? 2 : 1))
point_votes_tma_kernel(const s2d_video_desc* __restrict__ descs, const int4* __restrict__ rowplan,
                       int total_rows, int32_t* __restrict__ ctrl, int32_t* __restrict__ hits,
                       int32_t* __restrict__ uniq) {
    constexpr int STAGE_BYTES = THREADS * PPT * 8;
    extern __shared__ __align__(128) uint8_t dsm[];
    uint32_t* bm = reinterpret_cast<uint32_t*>(dsm + 2 * STAGE_BYTES);
    __shared__ int hist[S2D_MAX_LABELS + 1];          // [256] = number of unique pixels
    __shared__ uint32_t sbox[3][2];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ __align__(8) uint64_t empty[2];
    __shared__ PvTile tinfo[2];
    __shared__ PvOut tout[3];
    static_assert(THREADS >= S2D_MAX_LABELS, "output phase uses one thread per histogram bin");

    const int tid = threadIdx.x, lane = tid & 31;
    const int total = ctrl[1];

    if (tid < THREADS) {
        for (int i = tid; i < PV_BM_WORDS / 4; i += THREADS)
            reinterpret_cast<uint4*>(bm)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < S2D_MAX_LABELS + 1; i += THREADS) hist[i] = 0;
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { sbox[i][0] = PV_INVALID; sbox[i][1] = 0; }
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], THREADS / 32);
        mbar_init(&empty[1], THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= THREADS) {
        // =========================== producer warp ===========================================
        int pi = 0, pi_end = 0, prow = 0;          // next tile index, end of claimed chunk, its row
        int4 prp = make_int4(0, 0, 0, 0);          // rowplan[prow]
        for (int j = 0;; ++j) {
            const int s = j & 1;
            if (j >= 2) mbar_wait(&empty[s], ((j >> 1) - 1) & 1);    // stage s drained by the consumers
            bool done = false;
            if (pi == pi_end) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&ctrl[0], PV_CHUNK);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= total) {
                    done = true;
                } else {
                    pi = c;
                    pi_end = min(c + PV_CHUNK, total);
                    int lo = 0, hi = total_rows;     // 32-ary search: last row with tile0 <= pi
                    while (hi - lo > 32) {
                        const int step = (hi - lo + 31) >> 5;
                        const int probe = lo + lane * step;
                        const bool ok = probe < hi && rowplan[probe].x <= pi;
                        const uint32_t b = __ballot_sync(0xffffffffu, ok);
                        const int k = 31 - __clz(b);
                        lo += k * step;
                        hi = min(hi, lo + step);
                    }
                    const int probe = lo + lane;
                    const bool ok = probe < hi && rowplan[probe].x <= pi;
                    const uint32_t b = __ballot_sync(0xffffffffu, ok);
                    prow = lo + 31 - __clz(b);
                    prp = rowplan[prow];
                }
            }
            if (done) {
                if (lane == 0) { tinfo[s].valid = 0; mbar_arrive(&full[s]); }
                break;
            }
            while (pi >= prp.x + prp.y) { ++prow; prp = rowplan[prow]; }   // rows without tiles are skipped
            if (lane == 0) {
                const s2d_video_desc* dp = descs + prp.w;
                const int q = prow - (int)dp->row0;
                const int t = prp.z + (pi - prp.x);
                const int P = dp->P, T = dp->T, L = dp->L;
                const int64_t rt = (int64_t)q * T + t;
                PvTile ti;
                ti.lbl = dp->labels + (int64_t)t * dp->H * dp->W;
                ti.hout = hits + dp->hits_off + rt * L;
                ti.uout = uniq + dp->vt_off + rt;
                *ti.uout = 0;        // consumers accumulate per-warp partial sums (ordered by the mbarrier release/acquire)
                ti.W = dp->W; ti.H = dp->H; ti.L = L;
                const int32_t* np = dp->npts;
                ti.n = np ? min(max(np[q], 0), P) : P;
                ti.valid = 1; ti.pad = 0;
                tinfo[s] = ti;
                const uint32_t bytes = (uint32_t)P * 8u;
                mbar_expect_tx(&full[s], bytes);
                const int32_t* tsp = dp->tstart;
                const int64_t ts = t - (tsp ? tsp[q] : 0);           // frame index inside the stored track window
                bulk_g2s(dsm + s * STAGE_BYTES, dp->tracks + ((int64_t)q * dp->Ttr + ts) * P * 2, bytes, &full[s]);
            }
            ++pi;
        }
        return;
    }

    // =============================== consumer warps ==========================================
    // Two register sets (A/B) ping-pong between "being gathered" and "being voted".
    uint32_t linA[PPT], labA[PPT], linB[PPT], labB[PPT];
    bool okA, okB;

    // stage A of tile j: stage -> registers, round/bounds, issue label gathers, extent
    auto stage_a = [&](int j, uint32_t (&lin_n)[PPT], uint32_t (&lab_n)[PPT]) -> bool {
        const int s = j & 1;
        mbar_wait(&full[s], (j >> 1) & 1);
        if (!tinfo[s].valid) return false;
        const uint32_t W = tinfo[s].W, H = tinfo[s].H;
        const int n = tinfo[s].n;
        const uint8_t* lbl = tinfo[s].lbl;
        if (tid == 0) { tout[j % 3].hout = tinfo[s].hout; tout[j % 3].uout = tinfo[s].uout; tout[j % 3].L = tinfo[s].L; }
        // thread tid takes points tid, tid + THREADS, ...: the 32 lanes of a warp hold 32 consecutive
        // points, i.e. (for grid-ordered tracks) neighbouring pixels -> few cache lines per label gather
        const float2* sp = reinterpret_cast<const float2*>(dsm + s * STAGE_BYTES);
        if (n >= THREADS * PPT) {             // full tile: no per-point tail checks
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const float2 v = sp[k * THREADS + tid];
                lin_n[k] = pv_lin(v.x, v.y, W, H);
            }
        } else {
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const float2 v = sp[k * THREADS + tid];
                const uint32_t a = pv_lin(v.x, v.y, W, H);
                lin_n[k] = (k * THREADS + tid < n) ? a : PV_INVALID;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);     // this warp no longer reads stage s / tinfo[s]
        uint32_t lmin = PV_INVALID, lmax1 = 0;
        const uint32_t lastpx = W * H - 1u;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const uint32_t l = lin_n[k];
            // unconditional gather (invalid points read the last pixel; they never claim a bit)
            lab_n[k] = __ldg(lbl + min(l, lastpx));                // consumed one iteration later
            lmin = min(lmin, l);
            lmax1 = max(lmax1, l + 1u);
        }
        lmin = __reduce_min_sync(0xffffffffu, lmin);
        lmax1 = __reduce_max_sync(0xffffffffu, lmax1);
        uint32_t* sb = sbox[j % 3];
        if (lane == 0 && lmax1 != 0) { atomicMin(&sb[0], lmin); atomicMax(&sb[1], lmax1); }
        return true;
    };

    const uint32_t bm_s = (uint32_t)__cvta_generic_to_shared(bm);
    // stage B of tile `it`: de-duplicate, vote, write out, reset
    auto stage_b = [&](int it, const uint32_t (&lin)[PPT], const uint32_t (&lab)[PPT]) {
        consumer_sync(THREADS);                // S1: extents complete; previous tile's resets visible
        const uint32_t lmin = sbox[it % 3][0];
        const uint32_t lmax1 = sbox[it % 3][1];
        if (tid == 0) { sbox[(it + 2) % 3][0] = PV_INVALID; sbox[(it + 2) % 3][1] = 0; }
        if (lmax1 != 0) {
            // common case: every first point of the thread carries the label of its point 0; those
            // are counted in a register, the rest (object borders, other masks) vote one by one
            const uint32_t cur = lab[0];
            int cnt = 0;
            for (uint32_t base = lmin;;) {
                uint32_t odd = 0;
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const bool fst = pv_claim_s(bm_s, lin[k], base) != 0;
                    const bool same = lab[k] == cur;
                    cnt += (fst && same) ? 1 : 0;
                    odd |= (fst && !same) ? (1u << k) : 0u;
                }
                if (odd) {
#pragma unroll
                    for (int k = 0; k < PPT; ++k)
                        if ((odd >> k) & 1u) atomicAdd(&hist[lab[k]], 1);
                }
                if (lmax1 - base <= (uint32_t)PV_BM_BITS) break;
                base += PV_BM_BITS;
                consumer_sync(THREADS);
#pragma unroll
                for (int i = 0; i < PV_BM_WORDS / 4 / THREADS; ++i)
                    reinterpret_cast<uint4*>(bm)[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
                consumer_sync(THREADS);
            }
            pv_flush((int)cur, cnt, lane, hist);
        }
        consumer_sync(THREADS);                // S2: histogram complete
        if (tid < S2D_MAX_LABELS) {            // 8 warps: write hits, uniq = sum of the histogram
            const int h = hist[tid];
            const PvOut o = tout[it % 3];
            if (tid < o.L) o.hout[tid] = h;
            hist[tid] = 0;
            const int ws = __reduce_add_sync(0xffffffffu, h);
            if (lane == 0 && ws) atomicAdd(o.uout, ws);        // *uout was cleared by the producer-side plan
        }
        if (lmax1 != 0) {
#pragma unroll
            for (int i = 0; i < PV_BM_WORDS / 4 / THREADS; ++i)
                reinterpret_cast<uint4*>(bm)[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
        }
        // the next tile's S1 orders these resets before its bitmap / histogram atomics
    };

    okA = stage_a(0, linA, labA);
    for (int it = 0;; it += 2) {
        if (!okA) break;
        okB = stage_a(it + 1, linB, labB);
        stage_b(it, linA, labA);
        if (!okB) break;
        okA = stage_a(it + 2, linA, labA);
        stage_b(it + 1, linB, labB);
    }
}

#endif  // S2D_EXPERIMENTS

// ==========================================================================================
// Label-table variant: ONE shared-memory atomic per point does both the de-duplication and the
// label lookup (the production path when it applies: even P, 16-byte aligned tracks, P <= 8192).
//
// A tile lives in ONE shared-memory buffer that is first the landing zone of its 8P bytes of
// tracks (cp.async.bulk) and then, once the points are rounded into registers (phase A) and their
// bounding box is known, the table: exactly that region of the target frame's u8 label map, one
// cp.async.bulk per bbox row issued by as many threads in parallel, completion on one mbarrier.
// Rows keep their global 16-byte phase: the row pitch is chosen congruent to W modulo 16, so pixel
// (x, y) always sits at (y - y0) * pitch + (x - x0) + const and unaligned widths (W = 854) cost
// nothing extra. Phase B claims each point's pixel with atomicOr(word, 0xFF << 8*(offset & 3)):
// the returned byte is the pixel's label if this point is the first on the pixel, 0xFF if the
// pixel was already taken. No bitmap, no zeroing (the next copy overwrites the buffer), no global
// gather. Bounding boxes taller than the buffer are processed in bands of rows. Tiles the table
// cannot serve (label id 255 in use, or a bbox that would need more than PV_MAX_BANDS bands) fall
// back, inside the same kernel, to the bitmap + global-gather method with the buffer as bitmap.
// Nothing overlaps inside a CTA - the chain tracks -> bbox -> table -> votes is serial by nature -
// so the buffer is kept small enough for 4-5 CTAs per SM, which overlap each other's waits.
// Warp 0 also plans: it claims chunks of tiles and prepares the next tile's record while the
// CTA waits for its table.
// ==========================================================================================
constexpr int PV_MAX_BANDS = 6;
constexpr uint32_t PV_PK_INVALID = 0xFFFFFFFFu;

// Shared memory of the label-table kernel, all of it dynamic and carved by hand so that phase B can address it as
// (one base register | compile-time offset):
//   [0, 1024)            vote histogram, 256 int32 - the base is 1 KB aligned, so a bin's address is (4 * label) | base
//   [1024, 1152)         dummy words (all ones): target of points outside the band
//   [1152, 1664)         PvCtl: mbarriers, tile records, scheduler state, per-warp bounding boxes
//   [1664, +TRK)         SPLIT only: the tile's tracks (their own region: the next tile's tracks are fetched as soon as
//                        phase A has read this tile's, i.e. during the table fetch and phase B)
//   [.., +BUF + 64)      the buffer: the tile's tracks (not SPLIT), then its label table (+ 64 B slack for row copies)
constexpr int PV_FIXED_BYTES = 1664;
// `nch`: a tile's tracks arrive in nch chunks through a ring of TWO chunk slots at the start of the buffer (nch = 2: the
// whole tile is resident, as for 4096-point tiles; nch = 4: 128 KB tiles stream through 64 KB, so two CTAs fit an SM)
__host__ __device__ constexpr int pv_buf_bytes(int threads, int ppt, int ctas, bool split, int nch = 2) {
    // 228 KB per SM, 1 KB reserved by the system per CTA
    const int avail = 233472 / ctas - 1024 - PV_FIXED_BYTES - 64 - (split ? threads * ppt * 8 : 0);
    const int cap = (avail / 128) * 128;
    const int ring = threads * ppt * 8 * 2 / nch;
    return (!split && cap < ring) ? ring : cap;
}
__host__ __device__ constexpr int pv_smem_bytes(int threads, int ppt, int ctas, bool split, int nch = 2) {
    return PV_FIXED_BYTES + (split ? threads * ppt * 8 : 0) + pv_buf_bytes(threads, ppt, ctas, split, nch) + 64;
}

// (iy << 16 | ix) of a point that lands inside the frame, PV_PK_INVALID otherwise (W, H <= 65535)
__device__ __forceinline__ uint32_t pv_pack(float x, float y, uint32_t W, uint32_t H) {
    const uint32_t ix = (uint32_t)__float2int_rn(fmaxf(x, -1.0f));
    const uint32_t iy = (uint32_t)__float2int_rn(fmaxf(y, -1.0f));
    return (ix < W && iy < H) ? (iy << 16) + ix : PV_PK_INVALID;
}

// L2 policies: the tracks are read exactly once (evict first), the label maps are re-read by every
// query of the video (evict last) - without them the 55 GB track stream pushes the label maps out of
// L2 and 40 % of the table bytes come from DRAM again (ncu: 78.7 GB read for 54.9 GB algorithmic).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_addr(uint32_t smem_dst, uint64_t gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_box_2d(uint32_t smem_dst, const void* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct PvPlan {              // warp-0 state of the tile scheduler (identical in all its lanes)
    int pi, pi_end, prow;    // next tile index, end of the claimed chunk, row of tile pi
    int4 prp;                // rowplan[prow]
};

// Claim / locate the next tile (warp 0, all lanes) and let lane 0 write its record. Returns false
// when the work list is exhausted (the record is then marked invalid).
__device__ __forceinline__ bool pv_plan_next(PvPlan& pl, PvTile* rec, int lane, const s2d_video_desc* __restrict__ descs,
                                             const int4* __restrict__ rowplan, int total_rows, int total,
                                             int32_t* __restrict__ ctrl, int32_t* __restrict__ hits,
                                             int32_t* __restrict__ uniq, const uint8_t* __restrict__ tmaps,
                                             const float** src) {
    if (pl.pi == pl.pi_end) {
        // consecutive tiles per claim: up to PV_CHUNK (same query, neighbouring frames: label maps stay hot
        // in L2), fewer when the launch is small so that every CTA gets work
#ifdef S2D_EXPERIMENTS
        const int chunk = max(1, min(ctrl[2] > 0 ? ctrl[2] : PV_CHUNK, total / (int)(gridDim.x * 2u)));      // ctrl[2]: S2D_PV_CHUNK
#else
        const int chunk = max(1, min(PV_CHUNK, total / (int)(gridDim.x * 2u)));
#endif
        int c = 0;
        if (lane == 0) c = atomicAdd(&ctrl[0], chunk);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= total) {
            if (lane == 0) rec->valid = 0;
            return false;
        }
        pl.pi = c;
        pl.pi_end = min(c + chunk, total);
        int lo = 0, hi = total_rows;     // 32-ary search: last row with tile0 <= pi
        while (hi - lo > 32) {
            const int step = (hi - lo + 31) >> 5;
            const int probe = lo + lane * step;
            const bool ok = probe < hi && rowplan[probe].x <= pl.pi;
            const uint32_t b = __ballot_sync(0xffffffffu, ok);
            const int k = 31 - __clz(b);
            lo += k * step;
            hi = min(hi, lo + step);
        }
        const int probe = lo + lane;
        const bool ok = probe < hi && rowplan[probe].x <= pl.pi;
        const uint32_t b = __ballot_sync(0xffffffffu, ok);
        pl.prow = lo + 31 - __clz(b);
        pl.prp = rowplan[pl.prow];
    }
    while (pl.pi >= pl.prp.x + pl.prp.y) { ++pl.prow; pl.prp = rowplan[pl.prow]; }   // rows without tiles are skipped
    if (lane == 0) {
        const s2d_video_desc* dp = descs + pl.prp.w;
        const int q = pl.prow - (int)dp->row0;
        const int t = pl.prp.z + (pl.pi - pl.prp.x);
        const int P = dp->P, T = dp->T, L = dp->L;
        const int64_t rt = (int64_t)q * T + t;
        PvTile ti;
        ti.lbl = dp->labels + (int64_t)t * dp->H * dp->W;
        ti.hout = hits + dp->hits_off + rt * L;
        ti.uout = uniq + dp->vt_off + rt;
        *ti.uout = 0;        // the output phase accumulates per-warp partial sums (ordered by the barriers in between)
        ti.W = dp->W; ti.H = dp->H; ti.L = L;
        const int32_t* np = dp->npts;
        ti.n = np ? min(max(np[q], 0), P) : P;
        ti.valid = 1; ti.pad = P;
        const uint8_t* tm = tmaps ? tmaps + (size_t)pl.prp.w * S2D_PV_TMAP_BYTES : nullptr;
        if (tm && *reinterpret_cast<const int32_t*>(tm + S2D_PV_TMAPS * 128) == 0) tm = nullptr;     // this video has no descriptors
        ti.tm = tm;
        ti.ybase = (uint32_t)t * (uint32_t)dp->H; ti.pad2 = 0;
        const int32_t* tsp = dp->tstart;
        const int64_t ts = t - (tsp ? tsp[q] : 0);           // frame index inside the stored track window
        ti.src = dp->tracks + ((int64_t)q * dp->Ttr + ts) * P * 2;
        *rec = ti;
        *src = ti.src;
    }
    ++pl.pi;
    return true;
}

struct PvCtl {
    uint2 wred[16];                  // per-warp packed (min, max + 1) of (iy, ix)
    uint64_t full, full2, tabbar;    // tracks arrive in two halves: phase A starts on the first
    PvTile tinfo[2];                 // current tile, next tile (planned during the current one)
    PvPlan plan;                     // scheduler state (kept out of the registers)
    const float* nsrc;               // tracks of the planned tile
    int more;
};
static_assert(sizeof(PvCtl) <= 512, "PvCtl must fit its slot of the shared-memory layout");

template <int THREADS, int PPT, int CTAS, bool SPLIT, int NCH>
__global__ void __launch_bounds__(THREADS, CTAS)
point_votes_tab_kernel(const s2d_video_desc* __restrict__ descs, const int4* __restrict__ rowplan,
                       int total_rows, int32_t* __restrict__ ctrl, int32_t* __restrict__ hits,
                       int32_t* __restrict__ uniq, const uint8_t* __restrict__ tmaps) {
    constexpr int BUF_BYTES = pv_buf_bytes(THREADS, PPT, CTAS, SPLIT, NCH);
    constexpr int SLOT_BYTES = THREADS * PPT * 8 / NCH;          // one chunk of a tile's tracks
    constexpr int KCH = PPT / 2 / NCH;                           // 16-byte point pairs of a thread per chunk
    static_assert(NCH >= 2 && NCH % 2 == 0 && PPT % (2 * NCH) == 0 && (!SPLIT || NCH == 2), "chunking of the tracks");
    constexpr int TRK_BYTES = SPLIT ? THREADS * PPT * 8 : 0;
    constexpr int NWARPS = THREADS / 32;
    static_assert(NWARPS <= 16, "PvCtl::wred");
    static_assert(BUF_BYTES >= 8192, "label table buffer too small");
    // the fallback bitmap lives in the buffer: the largest power of two of words that fits (at most 2^13 = 32 KB)
    constexpr int FB_LOGW = BUF_BYTES >= 32768 ? 13 : (BUF_BYTES >= 16384 ? 12 : 11);
    constexpr int FB_WORDS = 1 << FB_LOGW;
    constexpr uint32_t FB_BITS = (uint32_t)FB_WORDS * 32u;
    static_assert(BUF_BYTES >= FB_WORDS * 4, "the fallback bitmap lives in the buffer");
    static_assert(S2D_MAX_LABELS % THREADS == 0 || THREADS % S2D_MAX_LABELS == 0, "output phase: whole warps per pass");
    static_assert(PPT % 4 == 0, "points are read two at a time, in two halves");
    // the address of a static __shared__ variable costs ptxas seven uniform instructions (S2UR SR_CgaCtaId, UMOV, ULEA ...)
    // and it re-materialised them in every group of four points (ncu: 6 % of the kernel's instructions): everything
    // lives in dynamic shared memory at compile-time offsets from one base register (layout: pv_buf_bytes above)
    extern __shared__ __align__(1024) uint8_t dsm[];
    int* const hist = reinterpret_cast<int*>(dsm);
    uint32_t* const dummy = reinterpret_cast<uint32_t*>(dsm + 1024);
    PvCtl& ctl = *reinterpret_cast<PvCtl*>(dsm + 1152);
    uint8_t* const trk = dsm + PV_FIXED_BYTES;                  // landing zone of the tile's tracks
    uint8_t* const buf = dsm + PV_FIXED_BYTES + TRK_BYTES;      // label table (not SPLIT: the same bytes as trk)
    uint2* const wred = ctl.wred;
    uint64_t& full = ctl.full; uint64_t& full2 = ctl.full2; uint64_t& tabbar = ctl.tabbar;
    PvTile* const tinfo = ctl.tinfo;
    PvPlan& plan_s = ctl.plan;
    const float*& nsrc_s = ctl.nsrc;
    int& more_s = ctl.more;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = ctrl[1];

    for (int i = tid; i < S2D_MAX_LABELS; i += THREADS) hist[i] = 0;
    if (tid < 32) dummy[tid] = 0xFFFFFFFFu;
    if (tid == 0) {
        plan_s.pi = plan_s.pi_end = plan_s.prow = 0;
        plan_s.prp = make_int4(0, 0, 0, 0);
        mbar_init(&full, 1);
        mbar_init(&full2, 1);
        mbar_init(&tabbar, THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0: plan one tile (state and results live in shared memory)
    auto plan = [&](PvTile* rec) {
        PvPlan pl = plan_s;
        const float* src = nullptr;
        const bool ok = pv_plan_next(pl, rec, lane, descs, rowplan, total_rows, total, ctrl, hits, uniq, tmaps, &src);
        __syncwarp();
        if (lane == 0) { plan_s = pl; nsrc_s = src; more_s = ok ? 1 : 0; }
    };
    // tracks of the planned tile -> buffer, as two bulk copies on two mbarriers (the halves of the k loop of phase A)
    // chunk c of a tile (P points at src) -> ring slot c & 1, completion on that slot's mbarrier (full / full2)
    auto issue_chunk = [&](const float* src, int P, int c) {
        const uint32_t bytes = (uint32_t)P * 8u, off = (uint32_t)c * SLOT_BYTES;
        const uint32_t nb = bytes > off ? min(bytes - off, (uint32_t)SLOT_BYTES) : 0u;
        uint64_t* bar = (c & 1) ? &full2 : &full;
        mbar_expect_tx(bar, nb);
        if (nb) bulk_g2s_hint(trk + (c & 1) * SLOT_BYTES, reinterpret_cast<const uint8_t*>(src) + off, nb, bar, l2_policy_evict_first());
    };
    // the first two chunks of the planned tile (NCH = 2: the whole tile, in two halves: phase A starts on the first)
    auto issue_tracks = [&](int P) {
        issue_chunk(nsrc_s, P, 0);
        issue_chunk(nsrc_s, P, 1);
    };
    if (warp == 0) {
        plan(&tinfo[0]);
        __syncwarp();
        if (lane == 0 && more_s) issue_tracks(tinfo[0].pad);
    }
    __syncthreads();

    // the shuffle makes the base an ordinary register value that ptxas cannot re-derive from SR_CgaCtaId at every use
    const uint32_t hist_s = __shfl_sync(0xffffffffu, smem_u32(dsm), 0);
    if (hist_s & 1023u) __trap();                     // a bin's address is formed as (4 * label) | hist_s
    const uint32_t tab_s = hist_s + (uint32_t)(PV_FIXED_BYTES + TRK_BYTES);
    const uint32_t dummy_s = hist_s + 1024u + 4u * (uint32_t)lane;
    uint32_t tabphase = 0;
    int scur = 0, snxt = 1;                           // slots of tile j and of tile j + 1 (planned during tile j)
    for (int j = 0;; ++j) {
        const PvTile* ti = &tinfo[scur];
        if (!ti->valid) break;                        // written before the barrier that precedes this read
        const uint32_t W = ti->W, H = ti->H;
        if (W > 65535u || H > 65535u) __trap();      // packed 16-bit coordinates (documented limit)
        const int n = ti->n;

        // ---- phase A: ring slots -> registers, round / bounds / pack, bounding box ----------------
        uint32_t pk[PPT];
        uint32_t mn = 0xFFFFFFFFu, mx = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            // slot c & 1 completes NCH / 2 times per tile; one warp polls the mbarrier, the rest park at the hardware barrier
            if (warp == 0) mbar_wait((c & 1) ? &full2 : &full, (uint32_t)(j * (NCH / 2) + (c >> 1)) & 1u);
            __syncthreads();
            const float4* sp = reinterpret_cast<const float4*>(trk + (c & 1) * SLOT_BYTES);
            if (n >= THREADS * PPT) {
#pragma unroll
                for (int kk = 0; kk < KCH; ++kk) {
                    const int k = c * KCH + kk;
                    const float4 v = sp[kk * THREADS + tid];
                    pk[2 * k] = pv_pack(v.x, v.y, W, H);
                    pk[2 * k + 1] = pv_pack(v.z, v.w, W, H);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < KCH; ++kk) {
                    const int k = c * KCH + kk;
                    const float4 v = sp[kk * THREADS + tid];
                    const int p0 = 2 * (k * THREADS + tid);
                    const uint32_t a = pv_pack(v.x, v.y, W, H), b = pv_pack(v.z, v.w, W, H);
                    pk[2 * k] = (p0 < n) ? a : PV_PK_INVALID;
                    pk[2 * k + 1] = (p0 + 1 < n) ? b : PV_PK_INVALID;
                }
            }
            if (c + 2 < NCH) {                        // the slot is free once everybody has read it: chunk c + 2 streams in
                __syncthreads();
                if (tid == 0) issue_chunk(ti->src, ti->pad, c + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < PPT / 2; ++k) {
            // per-halfword min / max; an invalid point is (0xFFFF, 0xFFFF) for the min and, after the
            // +0x00010001, (1, 0) for the max: harmless, any valid point has iy + 1 >= 1
            mn = __vimin3_u16x2(mn, pk[2 * k], pk[2 * k + 1]);
            mx = __vimax3_u16x2(mx, pk[2 * k] + 0x00010001u, pk[2 * k + 1] + 0x00010001u);
        }
        {
            const uint32_t mnx = __reduce_min_sync(0xffffffffu, mn & 0xFFFFu), mny = __reduce_min_sync(0xffffffffu, mn >> 16);
            const uint32_t mxx = __reduce_max_sync(0xffffffffu, mx & 0xFFFFu), mxy = __reduce_max_sync(0xffffffffu, mx >> 16);
            if (lane == 0) wred[warp] = make_uint2((mny << 16) | mnx, (mxy << 16) | mxx);
        }
        __syncthreads();                              // S1: boxes complete, nobody reads the tracks any more
        uint32_t bmn = 0xFFFFFFFFu, bmx = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            const uint2 v = wred[w];
            bmn = __vminu2(bmn, v.x);
            bmx = __vmaxu2(bmx, v.y);
        }
        const int L = ti->L;
        bool planned = false;                         // warp 0: next tile's record prepared
        if (bmn != 0xFFFFFFFFu) {                     // at least one point inside the frame
            const uint32_t x0 = bmn & 0xFFFFu, y0 = bmn >> 16;
            const uint32_t bw = (bmx & 0xFFFFu) - x0, bh = (bmx >> 16) - y0;      // max holds coordinate + 1
            // With TMA descriptors (row pitch and base of the label maps are multiples of 16) the table is
            // fetched as 2D boxes of 16 rows x 16..512 pixels, one instruction each; otherwise row by row.
            // (a box must start on a 16-byte boundary of its row: it starts at x0 & ~15)
            const uint8_t* tm = (bw + (x0 & 15u) <= 16u * S2D_PV_TMAPS) ? ti->tm : nullptr;
            uint32_t pitch, R;
            if (tm) {
                pitch = (bw + (x0 & 15u) + 15u) & ~15u;
                R = ((uint32_t)BUF_BYTES / pitch) & ~15u;                         // rows per band, whole boxes
            } else {
                pitch = bw + 15u;
                pitch += (W - pitch) & 15u;           // pitch = W (mod 16), pitch >= bw + 15
                R = (uint32_t)BUF_BYTES / pitch;
            }
            const uint8_t* lbl = ti->lbl;
            if (L <= 255 && R * PV_MAX_BANDS >= bh) {
                // ---- table mode ---------------------------------------------------------------
                for (uint32_t b0 = 0; b0 < bh; b0 += R) {
                    const uint32_t rows = min(R, bh - b0);
                    uint32_t a15 = 0;
                    if (tm) {
                        a15 = x0 & 15u;
                        if (tid == 0) {
                            const uint32_t nbox = (rows + 15u) >> 4;
                            mbar_expect_tx(&tabbar, nbox * 16u * pitch);
                            const uint8_t* map = tm + ((pitch >> 4) - 1u) * 128u;
#ifdef S2D_PV_BOUNDS_CHECK
                            if (nbox * 16u * pitch > (uint32_t)BUF_BYTES || (pitch & 15u) || pitch > 16u * S2D_PV_TMAPS) __trap();
#endif
                            for (uint32_t i = 0; i < nbox; ++i)
                                tma_box_2d(tab_s + i * 16u * pitch, map, &tabbar, (int)((x0 & ~15u) >> 1), (int)(ti->ybase + y0 + b0 + 16u * i), l2_policy_evict_last());
                        } else {
                            mbar_arrive(&tabbar);
                        }
                    } else {
                        const uint64_t A = (uint64_t)(uintptr_t)lbl + (uint64_t)(y0 + b0) * W + x0;
                        a15 = (uint32_t)A & 15u;
                        uint32_t bytes = 0;
                        const uint64_t pol_labels = l2_policy_evict_last();
                        for (uint32_t r = tid; r < rows; r += THREADS) {
                            const uint32_t ph = (uint32_t)(A + (uint64_t)r * W) & 15u;
                            bytes += (ph + bw + 15u) & ~15u;
                        }
                        mbar_expect_tx(&tabbar, bytes);                            // one arrival per thread
                        for (uint32_t r = tid; r < rows; r += THREADS) {
                            const uint64_t g = A + (uint64_t)r * W;
                            const uint32_t ph = (uint32_t)g & 15u;
#ifdef S2D_PV_BOUNDS_CHECK
                            {
                                const uint32_t d0 = r * pitch + a15 - ph, len = (ph + bw + 15u) & ~15u;
                                if ((d0 & 15u) || d0 + len > (uint32_t)BUF_BYTES + 64u || ((g - ph) & 15u)) __trap();
                            }
#endif
                            bulk_g2s_addr(tab_s + r * pitch + a15 - ph, g - ph, (ph + bw + 15u) & ~15u, &tabbar, pol_labels);
                        }
                    }
                    if (warp == 0 && !planned) {      // overlap the plan's dependent loads with the table's flight
                        plan(&tinfo[snxt]);
                        planned = true;
                        if (SPLIT) {                  // the tracks region is free since S1: the next tile's tracks fly from here on
                            __syncwarp();
                            if (lane == 0 && more_s) issue_tracks(tinfo[snxt].pad);
                        }
                    }
                    if (warp == 0) mbar_wait(&tabbar, tabphase);
                    __syncthreads();
                    tabphase ^= 1u;

                    // phase B: e = (dy << 16) + dx for points of this band, >= lim otherwise (rows above
                    // wrap around, rows below and PV_PK_INVALID keep a high half >= rows: y0 + b0 + rows <= H)
                    const uint32_t pk0 = ((y0 + b0) << 16) + x0;
                    const uint32_t tabc = tab_s + a15;
                    const uint32_t lim = rows << 16;
                    const uint32_t negc = pitch - 65536u;
                    // Groups of G points: all G atomics are issued before the first result is used.
                    // Points outside the band (and invalid ones) hit a per-lane dummy word that is all
                    // ones, so they read back 0xFF = "not first" without a branch. Every point then
                    // votes for the byte it read with one shared-memory reduction (SASS ATOMS.POPC.INC:
                    // the lanes of a warp that hit the same bin are merged); bin 255 collects the
                    // points that were not first and is dropped in the output phase (L <= 255 here).
                    constexpr int G = (PPT < 8 || CTAS >= 5) ? 4 : 8;
                    const bool banded = bh > R;        // several bands: most groups of a pass are outside its band
#pragma unroll
                    for (int g = 0; g < PPT; g += G) {
                        uint32_t old[G], sh[G];
                        if (banded) {                  // points are close to raster order: skip groups with no point in the band
                            bool any = false;
#pragma unroll
                            for (int k = 0; k < G; ++k) any |= (pk[g + k] - pk0) < lim;
                            if (!__any_sync(0xffffffffu, any)) continue;
                        }
#pragma unroll
                        for (int k = 0; k < G; ++k) {
                            const uint32_t e = pk[g + k] - pk0;
                            const uint32_t off = (e >> 16) * negc + e + tabc;      // table + dy * pitch + dx + a15
                            sh[k] = (off << 3) - 2u;                               // rotation counts, used modulo 32
                            const uint32_t addr = (e < lim) ? (off & ~3u) : dummy_s;
#ifdef S2D_PV_BOUNDS_CHECK      // debug build (make check): every claimed byte lies inside the fetched table
                            if (e < lim && (off - tab_s >= (uint32_t)BUF_BYTES || off - tabc >= rows * pitch)) __trap();
#endif
                            // 0xFF in byte (off & 3) = 0x3FC rotated left by 8 * (off & 3) - 2
                            asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old[k]) : "r"(addr), "r"(__funnelshift_l(0x3FCu, 0x3FCu, sh[k])));
                        }
#pragma unroll
                        for (int k = 0; k < G; ++k) {
                            // 4 * label = the byte rotated down to bits 2..9 (label 0xFF: not first / not in band)
                            const uint32_t bin = (__funnelshift_r(old[k], old[k], sh[k]) & 0x3FCu) | hist_s;     // one LOP3
                            asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(bin) : "memory");
                        }
                    }
                    if (b0 + R < bh) {                 // the next band's copies overwrite the table
                        fence_proxy_async();
                        __syncthreads();
                    }
                }
            } else {
                // ---- fallback: bitmap in the buffer + global label gather ----------------------
                uint32_t* bm = reinterpret_cast<uint32_t*>(buf);
                const uint32_t lend = (y0 + bh - 1u) * W + x0 + bw;            // last pixel + 1
                for (uint32_t base = y0 * W + x0;; base += FB_BITS) {
                    for (int i = tid; i < FB_WORDS / 4; i += THREADS)
                        reinterpret_cast<uint4*>(bm)[i] = make_uint4(0, 0, 0, 0);
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        const uint32_t p = pk[k];
                        const uint32_t lin = (p == PV_PK_INVALID) ? PV_INVALID : (p >> 16) * W + (p & 0xFFFFu);
                        if (pv_claim_t<FB_LOGW>(tab_s, lin, base)) atomicAdd(&hist[__ldg(lbl + lin)], 1);
                    }
                    if (lend - base <= FB_BITS) break;
                    __syncthreads();
                }
            }
        }
        if (warp == 0 && !planned) {
            plan(&tinfo[snxt]);
            if (SPLIT) {
                __syncwarp();
                if (lane == 0 && more_s) issue_tracks(tinfo[snxt].pad);
            }
        }
        fence_proxy_async();                          // buffer atomics before the next tile's bulk copy
        __syncthreads();                              // S2: histogram complete, buffer free, next record visible
        if (!SPLIT && tid == 0 && more_s) {           // the next tile's tracks fly during the output phase
            issue_tracks(tinfo[snxt].pad);
        }
#pragma unroll
        for (int b = tid; b < S2D_MAX_LABELS; b += THREADS) {     // write hits, uniq = sum of the histogram
            const int h = (b == 255 && L <= 255) ? 0 : hist[b];           // table mode parks non-first points in bin 255
            if (b < L) ti->hout[b] = h;
            hist[b] = 0;
            const int ws = __reduce_add_sync(0xffffffffu, h);
            if (lane == 0 && ws) atomicAdd(ti->uout, ws);      // *uout was cleared when the tile was planned
        }
        // the next tile's S1 orders the histogram resets before its votes; this tile's record is
        // rewritten only after that S1 as well (the next plan runs behind it)
        scur ^= 1; snxt ^= 1;
    }
}

template <int THREADS, int PPT, int CTAS, bool SPLIT = false, int NCH = 2>
static int launch_pv_tab(cudaStream_t st, int nsm, const s2d_video_desc* descs, const int4* rowplan,
                         int total_rows, int32_t* ctrl, int32_t* hits, int32_t* uniq, const uint8_t* tmaps) {
    const int smem = pv_smem_bytes(THREADS, PPT, CTAS, SPLIT, NCH);
    auto kfn = point_votes_tab_kernel<THREADS, PPT, CTAS, SPLIT, NCH>;
    static bool configured[S2D_MAX_DEVICES] = {};
    {
        cudaError_t e = opt_in_smem(kfn, smem, configured);
        if (e != cudaSuccess) { set_error("point_votes_tab_kernel: cannot opt in to %d B of shared memory: %s", smem, cudaGetErrorString(e)); return -2; }
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    kfn<<<nsm * per_sm, THREADS, smem, st>>>(descs, rowplan, total_rows, ctrl, hits, uniq, tmaps);
    S2D_CHECK_LAUNCH("point_votes_tab_kernel");
    return 0;
}

// ==========================================================================================
// K2 for sparse tiles (P <= 1024 tracked points per query): ONE WARP PER TILE.
//
// The label-table kernel above spends about 2 200 warp-instructions of fixed work on every tile (plan by one warp while
// three wait, five block-wide barriers, a 256-bin output pass by four warps, the table fetch): for a 4096-point tile that
// is 40 % on top of the per-point work, for a 1024-point tile it is twice the per-point work, and the kernel ends up
// bound by instruction issue at a third of the HBM roofline (ncu: 3 504 instructions per 8 KB tile). Here a tile belongs
// to one warp of a one-warp CTA - no block barriers, no idle warps, ~15 tiles in flight per SM:
//   * the tile's 8 P bytes of tracks land in the warp's 8 KB of shared memory by one cp.async.bulk (mbarrier), the next
//     tile's copy is issued as soon as the points sit in registers (32 per lane, packed (iy << 16) | ix);
//   * the bounding box comes from packed 16-bit min / max and four warp reductions;
//   * de-duplication is a bitmap of the bounding box in shared memory (4 KB = 32 768 pixels per band, bit index =
//     dy * bw + dx - band base; larger boxes take several bands), claimed with one shared atomic per point;
//   * the first point on a pixel reads the pixel's label straight from the label map (L2): eight independent loads in
//     flight per lane, and points arrive close to raster order, so a warp-wide load touches a handful of sectors;
//   * votes are shared-memory reductions into the warp's own 256-bin histogram, which one pass writes to hits / uniq.
// Any label id (0..255) and any box size are handled by this one path; results are identical to the other kernels'.
// Measured (profiles/r02_k2_small_ab_warp_v1.json): 480p videos 31 -> 40 % of the HBM copy peak against the label-table
// kernel, but 32 -> 25 % on the 720p / 1080p mixture, whose boxes take 2-5 bitmap bands with all 32 points of a lane
// re-tested per band - the product dispatch therefore uses it for frames of up to PVW_MAX_FRAME_PIXELS only. The kernel is
// 56 KB of code (all points of a lane unrolled in registers) and its first stall reason is instruction fetch; a version
// with real loops and synchronous track loads (profiles/r02_k2_small_ab_warp_v2_sync_loads_rejected.json) removed that
// stall and lost more on the exposed loads.
// ==========================================================================================
constexpr int PVW_BITS = 32768;                                     // bitmap bits per band
constexpr int64_t PVW_MAX_FRAME_PIXELS = 16 * PVW_BITS;             // product dispatch: frames up to 512 Ki pixels (480 x 854: yes, 720p: no)
constexpr int PVW_TRK_BYTES = 8192;                                 // 1024 points
constexpr int PVW_CTL_OFF = 1152, PVW_TRK_OFF = 1536, PVW_BITS_OFF = PVW_TRK_OFF + PVW_TRK_BYTES;
constexpr int PVW_SMEM_BYTES = PVW_BITS_OFF + PVW_BITS / 8;         // hist 1 KB | dummy 128 B | control | tracks | bitmap
constexpr int PVW_G = 8;                                            // points per lane whose atomics / loads are issued together

struct PvwCtl { uint64_t full; PvTile tinfo[2]; };
static_assert(sizeof(PvwCtl) <= PVW_TRK_OFF - PVW_CTL_OFF, "PvwCtl must fit its slot");

__global__ void __launch_bounds__(32, 15)
point_votes_warp_kernel(const s2d_video_desc* __restrict__ descs, const int4* __restrict__ rowplan, int total_rows,
                        int32_t* __restrict__ ctrl, int32_t* __restrict__ hits, int32_t* __restrict__ uniq) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    int* const hist = reinterpret_cast<int*>(dsm);
    uint32_t* const dummy = reinterpret_cast<uint32_t*>(dsm + 1024);
    PvwCtl& ctl = *reinterpret_cast<PvwCtl*>(dsm + PVW_CTL_OFF);
    uint8_t* const trk = dsm + PVW_TRK_OFF;
    uint32_t* const bits = reinterpret_cast<uint32_t*>(dsm + PVW_BITS_OFF);
    const int lane = threadIdx.x;
    const int total = ctrl[1];

#pragma unroll
    for (int i = lane; i < S2D_MAX_LABELS; i += 32) hist[i] = 0;
    dummy[lane] = 0xFFFFFFFFu;
    if (lane == 0) {
        mbar_init(&ctl.full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue_tracks = [&](const PvTile* rec) {        // lane 0: the tile's tracks -> shared memory, completion on `full`
        const uint32_t bytes = (uint32_t)rec->pad * 8u;
        mbar_expect_tx(&ctl.full, bytes);
        bulk_g2s_hint(trk, rec->src, bytes, &ctl.full, l2_policy_evict_first());
    };
    PvPlan pl;
    pl.pi = pl.pi_end = pl.prow = 0;
    pl.prp = make_int4(0, 0, 0, 0);
    const float* unused_src = nullptr;
    bool more = pv_plan_next(pl, &ctl.tinfo[0], lane, descs, rowplan, total_rows, total, ctrl, hits, uniq, nullptr, &unused_src);
    __syncwarp();
    if (lane == 0 && more) issue_tracks(&ctl.tinfo[0]);

    const uint32_t hist_s = __shfl_sync(0xffffffffu, smem_u32(dsm), 0);
    if (hist_s & 1023u) __trap();                       // a bin's address is formed as (4 * label) | hist_s
    const uint32_t bits_s = hist_s + (uint32_t)PVW_BITS_OFF;
    const uint32_t dummy_s = hist_s + 1024u + 4u * (uint32_t)lane;
    int scur = 0;
    for (int j = 0;; ++j) {
        const PvTile* ti = &ctl.tinfo[scur];
        if (!ti->valid) break;
        const uint32_t W = ti->W, H = ti->H;
        if (W > 65535u || H > 65535u) __trap();        // packed 16-bit coordinates (documented limit)
        const int n = ti->n, L = ti->L;
        const uint8_t* const lbl = ti->lbl;
        int32_t* const hout = ti->hout;
        int32_t* const uout = ti->uout;
        S2D_DEV_ASSERT(ti->pad >= 1 && ti->pad <= 1024 && n <= ti->pad);

        // the next tile's record is prepared while this tile's tracks are in flight (the plan is a chain of global loads)
        more = pv_plan_next(pl, &ctl.tinfo[scur ^ 1], lane, descs, rowplan, total_rows, total, ctrl, hits, uniq, nullptr, &unused_src);

        // ---- phase A: shared memory -> registers, round / bounds / pack ------------------------------
        mbar_wait(&ctl.full, (uint32_t)j & 1u);
        // lane l holds points 64 k + 2 l, + 1 (k = 0..15): rows k < kfull are complete, row kfull may be ragged (n points)
        const int kfull = n >> 6, kmax = (n + 63) >> 6;
        uint32_t pk[32];
        const float4* sp = reinterpret_cast<const float4*>(trk);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t a = PV_PK_INVALID, b = PV_PK_INVALID;
            if (k < kfull) {
                const float4 v = sp[k * 32 + lane];
                a = pv_pack(v.x, v.y, W, H);
                b = pv_pack(v.z, v.w, W, H);
            } else if (k < kmax) {
                const float4 v = sp[k * 32 + lane];
                const int p0 = 2 * (k * 32 + lane);
                if (p0 < n) a = pv_pack(v.x, v.y, W, H);
                if (p0 + 1 < n) b = pv_pack(v.z, v.w, W, H);
            }
            pk[2 * k] = a;
            pk[2 * k + 1] = b;
        }
        __syncwarp();                                   // every lane has read the tracks: the region is free again
        if (lane == 0 && more) issue_tracks(&ctl.tinfo[scur ^ 1]);

        // ---- bounding box ------------------------------------------------------------------------------
        uint32_t mn = 0xFFFFFFFFu, mx = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // per-halfword min / max; an invalid point is (0xFFFF, 0xFFFF) for the min and, after the
            // +0x00010001, (1, 0) for the max: harmless, any valid point has iy + 1 >= 1
            mn = __vimin3_u16x2(mn, pk[2 * k], pk[2 * k + 1]);
            mx = __vimax3_u16x2(mx, pk[2 * k] + 0x00010001u, pk[2 * k + 1] + 0x00010001u);
        }
        const uint32_t x0 = __reduce_min_sync(0xffffffffu, mn & 0xFFFFu), y0 = __reduce_min_sync(0xffffffffu, mn >> 16);
        const uint32_t x1 = __reduce_max_sync(0xffffffffu, mx & 0xFFFFu), y1 = __reduce_max_sync(0xffffffffu, mx >> 16);
        if (((y0 << 16) | x0) != 0xFFFFFFFFu) {         // at least one point inside the frame
            const uint32_t bw = x1 - x0, bh = y1 - y0;  // the max holds coordinate + 1
            S2D_DEV_ASSERT(bw >= 1 && bh >= 1 && x0 + bw <= W && y0 + bh <= H);
            const uint32_t pk0 = (y0 << 16) + x0, lim = bh << 16;
            const uint32_t npx = bw * bh;               // <= W * H < 2^32
            const uint32_t org = y0 * W + x0;           // pixel index of the box's corner
            for (uint32_t base = 0; base < npx; base += (uint32_t)PVW_BITS) {
                // an opaque copy of pk0 per band: without it the compiler hoists the 32 points' band-invariant terms (bit
                // index, label address: ~100 registers) out of this loop, which has one iteration for most tiles
                uint32_t pk0b;
                asm volatile("mov.u32 %0, %1;" : "=r"(pk0b) : "r"(pk0));
                const uint32_t nq = (min((uint32_t)PVW_BITS, npx - base) + 127u) >> 7;      // 128-bit words to clear
                for (uint32_t i = lane; i < nq; i += 32) reinterpret_cast<uint4*>(bits)[i] = make_uint4(0, 0, 0, 0);
                __syncwarp();
#pragma unroll
                for (int g = 0; g < 32; g += PVW_G) {
                    if (g / 2 >= kmax) break;           // no points beyond (uniform)
                    // (1) claim the pixels: points outside the frame / the band hit a per-lane dummy word that is all ones
                    uint32_t first = 0;
#pragma unroll
                    for (int k = 0; k < PVW_G; ++k) {
                        const uint32_t e = pk[g + k] - pk0b;                   // (dy << 16) + dx of a valid point
                        const uint32_t lin = (e >> 16) * bw + (e & 0xFFFFu) - base;
                        const bool ok = (e < lim) && (lin < (uint32_t)PVW_BITS);
                        S2D_DEV_ASSERT(!(e < lim) || (e & 0xFFFFu) < bw);
                        const uint32_t m = 1u << (lin & 31u);
                        const uint32_t addr = ok ? bits_s + ((lin >> 5) << 2) : dummy_s;
                        uint32_t old;
                        asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(m) : "memory");
                        first |= ((old & m) ? 0u : 1u) << k;
                    }
                    // (2) labels of the first points: independent loads, all issued before the first is used
                    uint32_t lab[PVW_G];
#pragma unroll
                    for (int k = 0; k < PVW_G; ++k) {
                        lab[k] = 0;
                        if ((first >> k) & 1u) {
                            const uint32_t e = pk[g + k] - pk0b;
                            S2D_DEV_ASSERT((e >> 16) < bh && (e & 0xFFFFu) < bw);
                            lab[k] = __ldg(lbl + (size_t)(org + (e >> 16) * W + (e & 0xFFFFu)));
                        }
                    }
                    // (3) votes
#pragma unroll
                    for (int k = 0; k < PVW_G; ++k)
                        if ((first >> k) & 1u) asm volatile("red.shared.add.u32 [%0], 1;" :: "r"((lab[k] << 2) | hist_s) : "memory");
                }
                if (base + (uint32_t)PVW_BITS < npx) __syncwarp();             // the next band clears the bitmap
            }
        }
        __syncwarp();                                   // histogram complete

        // ---- output: hits[q, t, :] and uniq[q, t] = number of distinct pixels --------------------------------
        int sum = 0;
#pragma unroll
        for (int b = lane; b < S2D_MAX_LABELS; b += 32) {
            const int h = hist[b];
            if (b < L) hout[b] = h;
            hist[b] = 0;
            sum += h;
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) *uout = sum;
        __syncwarp();                                   // histogram resets before the next tile's votes
        scur ^= 1;
    }
}

static int launch_pv_warp(cudaStream_t st, int nsm, const s2d_video_desc* descs, const int4* rowplan, int total_rows,
                          int32_t* ctrl, int32_t* hits, int32_t* uniq) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, point_votes_warp_kernel, 32, PVW_SMEM_BYTES);
    if (per_sm < 1) per_sm = 1;
    point_votes_warp_kernel<<<nsm * per_sm, 32, PVW_SMEM_BYTES, st>>>(descs, rowplan, total_rows, ctrl, hits, uniq);
    S2D_CHECK_LAUNCH("point_votes_warp_kernel");
    return 0;
}

#ifdef S2D_EXPERIMENTS
template <int THREADS, int PPT>
static int launch_pv_tma(cudaStream_t st, int nsm, const s2d_video_desc* descs, const int4* rowplan,
                         int total_rows, int32_t* ctrl, int32_t* hits, int32_t* uniq) {
    const int smem = 2 * THREADS * PPT * 8 + PV_BM_WORDS * 4;
    auto kfn = point_votes_tma_kernel<THREADS, PPT>;
    static bool configured[S2D_MAX_DEVICES] = {};
    {
        cudaError_t e = opt_in_smem(kfn, smem, configured);
        if (e != cudaSuccess) { set_error("point_votes_tma_kernel: cannot opt in to %d B of shared memory: %s", smem, cudaGetErrorString(e)); return -2; }
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, THREADS + 32, smem);
    if (per_sm < 1) per_sm = 1;
    kfn<<<nsm * per_sm, THREADS + 32, smem, st>>>(descs, rowplan, total_rows, ctrl, hits, uniq);
    S2D_CHECK_LAUNCH("point_votes_tma_kernel");
    return 0;
}

#endif

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

static int g_pv_variant = 0;

extern "C" int s2d_point_votes_variant(int variant) {
#ifdef S2D_EXPERIMENTS
    S2D_CHECK_ARG(variant >= 0 && variant <= 4, "s2d_point_votes_variant: %d not in {0 product dispatch, 1 bitmap, 2 one CTA per tile, 3 label table for every P, 4 one warp per tile for every frame size}", variant);
#else
    S2D_CHECK_ARG(variant == 0 || variant == 2 || variant == 3 || variant == 4, "s2d_point_votes_variant: %d not in {0 product dispatch, 2 one CTA per tile, 3 label table for every P, 4 one warp per tile for every frame size} "
                  "(1, the superseded bitmap kernel, exists only in the experiments build: make -C s2d_b200/csrc exp)", variant);
#endif
    g_pv_variant = variant;
    return 0;
}

extern "C" int s2d_point_votes_work_ints(int64_t total_rows, int64_t* out) {
    if (!out) return -1;
    *out = 4 * total_rows + 8;
    return 0;
}

// Host side: TMA descriptors of the videos' label maps for the label-table kernel. Per video
// S2D_PV_TMAP_BYTES: S2D_PV_TMAPS CUtensorMap (boxes of 16 rows x 16, 32, ... 512 pixels of the [T*H][W] u8 map)
// and a 128-byte trailer whose first int is 1 when the descriptors are usable (W and the base
// address are multiples of 16, W >= 16), 0 otherwise (that video's tables are fetched row by row).
typedef CUresult (*PvEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int s2d_point_votes_tmaps(const s2d_video_desc* host_descs, int nvideos, void* host_out) {
    S2D_CHECK_ARG(host_descs && host_out && nvideos > 0, "s2d_point_votes_tmaps: bad arguments");
    static PvEncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            enc = (PvEncodeTiledFn)p;
    }
    uint8_t* out = static_cast<uint8_t*>(host_out);
    memset(out, 0, (size_t)nvideos * S2D_PV_TMAP_BYTES);
    if (!enc) return 0;                       // no driver entry point: every video falls back to row copies
    for (int v = 0; v < nvideos; ++v) {
        const s2d_video_desc& d = host_descs[v];
        uint8_t* blk = out + (size_t)v * S2D_PV_TMAP_BYTES;
        if (!d.labels || d.W < 16 || (d.W & 15) || (((uintptr_t)d.labels) & 15) || d.T <= 0 || d.H <= 0) continue;
        bool ok = true;
        for (int k = 0; k < S2D_PV_TMAPS && ok; ++k) {
            if (16 * (k + 1) > d.W) {         // box wider than the frame: never selected (bw <= W); keep the slot well formed
                memcpy(blk + k * 128, blk + (k - 1) * 128, 128);
                continue;
            }
            // the u8 map is described as u16 [T*H][W/2]: a box dimension is limited to 256 ELEMENTS, so
            // 16-bit elements give boxes of up to 512 pixels (same bytes, coordinates in pixel pairs)
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)d.W / 2, (cuuint64_t)d.T * (cuuint64_t)d.H};
            cuuint64_t strides[1] = {(cuuint64_t)d.W};
            cuuint32_t box[2] = {(cuuint32_t)(8 * (k + 1)), 16u};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)d.labels, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) ok = false;
            else memcpy(blk + k * 128, &map, 128);
        }
        if (ok) *reinterpret_cast<int32_t*>(blk + S2D_PV_TMAPS * 128) = 1;
    }
    return 0;
}

extern "C" int s2d_point_votes(const s2d_video_desc* descs, int nvideos, int max_T, int max_Nm, int max_P,
                               int vec4_ok, int64_t total_rows, const int32_t* rowinfo,
                               const int32_t* vidinfo, int32_t* work, const void* label_tmaps,
                               int32_t* hits, int32_t* uniq, void* stream) {
    return s2d_point_votes_sized(descs, nvideos, max_T, max_Nm, max_P, vec4_ok, total_rows, rowinfo, vidinfo, work, label_tmaps,
                                 0, hits, uniq, stream);
}

extern "C" int s2d_point_votes_sized(const s2d_video_desc* descs, int nvideos, int max_T, int max_Nm, int max_P,
                                     int vec4_ok, int64_t total_rows, const int32_t* rowinfo,
                                     const int32_t* vidinfo, int32_t* work, const void* label_tmaps,
                                     int64_t max_frame_pixels, int32_t* hits, int32_t* uniq, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG((((uintptr_t)label_tmaps) & 63) == 0, "s2d_point_votes: label_tmaps must be 64-byte aligned");
    const uint8_t* tm = static_cast<const uint8_t*>(label_tmaps);
    S2D_CHECK_ARG(descs && hits && uniq, "s2d_point_votes: null pointer");
    S2D_CHECK_ARG(nvideos > 0 && nvideos <= 65535 && max_T > 0 && max_Nm > 0 && max_Nm <= 65535,
                  "s2d_point_votes: bad sizes");
    S2D_CHECK_ARG(max_P >= 1 && max_P <= 32768, "s2d_point_votes: P=%d not in [1, 32768]", max_P);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = vec4_ok != 0;
    // tiles of <= 1024 points: one warp per tile when the frames are small enough for its bitmap bands (variant 4: always,
    // variant 3: never - the label-table kernel for every P)
    const bool warp_tiles = g_pv_variant == 4 || (g_pv_variant == 0 && max_frame_pixels > 0 && max_frame_pixels <= PVW_MAX_FRAME_PIXELS);
    const int variant = (g_pv_variant == 3 || g_pv_variant == 4) ? 0 : g_pv_variant;
    if (variant != 2 && v4 && work && max_P <= (variant == 0 ? 16384 : 8192) && total_rows > 0 && total_rows <= 2147483647LL) {
        // persistent TMA path
        S2D_CHECK_ARG((((uintptr_t)work) & 15) == 0, "s2d_point_votes: work must be 16-byte aligned");
        int dev = 0, nsm = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        int32_t* ctrl = work;
        int4* rowplan = reinterpret_cast<int4*>(work + 8);
        pv_plan_rows_kernel<<<(unsigned)((total_rows + 255) / 256), 256, 0, st>>>(descs, nvideos, total_rows, rowinfo, vidinfo, rowplan);
        S2D_CHECK_LAUNCH("pv_plan_rows_kernel");
        pv_scan_kernel<<<1, 1024, 0, st>>>(rowplan, total_rows, ctrl);
        S2D_CHECK_LAUNCH("pv_scan_kernel");
#ifdef S2D_EXPERIMENTS
        {
            static int32_t chunk_h = 0;
            chunk_h = getenv("S2D_PV_CHUNK") ? atoi(getenv("S2D_PV_CHUNK")) : 0;
            cudaMemcpyAsync(ctrl + 2, &chunk_h, sizeof(int32_t), cudaMemcpyHostToDevice, st);
        }
#endif
        const int tr = (int)total_rows;
        if (variant == 0) {      // label-table kernels
#ifdef S2D_EXPERIMENTS   // 8 KB tiles: CTAs per SM / split tracks region (A/B runs)
            if (max_P <= 128 * 8 && getenv("S2D_PV_SMALL")) {
                const int v = atoi(getenv("S2D_PV_SMALL"));
                if (v == 60) return launch_pv_tab<128, 8, 6, false>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (v == 6) return launch_pv_tab<128, 8, 6, true>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (v == 7) return launch_pv_tab<128, 8, 7, true>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (v == 8) return launch_pv_tab<128, 8, 8, true>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (v == 9) return launch_pv_tab<128, 8, 9, true>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
            }
#endif
            if (max_P <= 128 * 8 && warp_tiles) return launch_pv_warp(st, nsm, descs, rowplan, tr, ctrl, hits, uniq);
            if (max_P <= 128 * 8) return launch_pv_tab<128, 8, 7, true>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
            if (max_P <= 128 * 16) return launch_pv_tab<128, 16, 6>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
#ifdef S2D_EXPERIMENTS   // CTAs per SM of the 4096-point configuration (A/B runs; 6 x 128 threads is the measured best)
            if (max_P <= 256 * 16) {
                const int ctas = getenv("S2D_PV_CTAS") ? atoi(getenv("S2D_PV_CTAS")) : 6;
                if (ctas == 3) return launch_pv_tab<256, 16, 3>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (ctas == 4) return launch_pv_tab<256, 16, 4>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
                if (ctas == 5) return launch_pv_tab<128, 32, 5>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
            }
#endif
            if (max_P <= 128 * 32) return launch_pv_tab<128, 32, 6>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
            if (max_P <= 256 * 32) return launch_pv_tab<256, 32, 3>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
#ifdef S2D_EXPERIMENTS
            if (getenv("S2D_PV_BIG") && atoi(getenv("S2D_PV_BIG")) == 1)
                return launch_pv_tab<512, 32, 1>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);     // round 1: 128 KB tiles resident, one CTA per SM
#endif
            // 128 KB tiles stream through a ring of two 32 KB slots: two CTAs of 256 threads per SM
            return launch_pv_tab<256, 64, 2, false, 4>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq, tm);
        }
#ifdef S2D_EXPERIMENTS
        if (max_P <= 256 * 4) return launch_pv_tma<256, 4>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq);   // 4 CTAs/SM
        if (max_P <= 512 * 4) return launch_pv_tma<512, 4>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq);
        if (max_P <= 4096) return launch_pv_tma<512, 8>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq);
        return launch_pv_tma<512, 16>(st, nsm, descs, rowplan, tr, ctrl, hits, uniq);
#endif
    }
    dim3 grid((unsigned)max_T, (unsigned)max_Nm, nvideos);
#define PV_ARGS v4, grid, st, descs, rowinfo, vidinfo, hits, uniq
    if (max_P <= 256 * 4) return launch_pv<256, 4>(PV_ARGS);
    if (max_P <= 256 * 8) return launch_pv<256, 8>(PV_ARGS);
    if (max_P <= 256 * 16) return launch_pv<256, 16>(PV_ARGS);
    if (max_P <= 512 * 16) return launch_pv<512, 16>(PV_ARGS);
    if (max_P <= 1024 * 16) return launch_pv<1024, 16>(PV_ARGS);
    return launch_pv<1024, 32>(PV_ARGS);
#undef PV_ARGS
}
