// K1 (tensor-core variant): dense mask-overlap contraction on the 5th-gen tensor cores.
//   I[a,b] = sum_px A[a,px] * B[b,px]      A, B: u8 planes holding 0/1, int32 accumulation
//
//   s2d_overlap_i8            operands are u8 planes in HBM, staged by TMA (cp.async.bulk.tensor,
//                             SWIZZLE_128B) into a 4-stage shared-memory ring; tcgen05.mma
//                             kind::i8 (M128 x N x K32) accumulates into TMEM; split-K over the
//                             pixel range, int32 atomics into I.
//   s2d_overlap_gram_labels   Gram form G = X X^T for the one-hot expansion X[(t,l), px] =
//                             (labels[t][px] == l) of the label maps of a window: the operand
//                             tiles are synthesised in shared memory (already in the 128B-swizzled
//                             K-major layout the MMA descriptors expect) from 1 B/px label bytes, so
//                             HBM traffic is T*H*W bytes while the MMA work is 2*(T*L)^2*H*W ops.
//
// Reference semantics: cotracker_matching.py:653-657 (point raster x mask) and
// model_training/mask2former_video/engine/train_loop.py:378-388 (mask x mask, x @ y.T).
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

namespace s2d {

constexpr int GM_BLOCK_M = 128;
constexpr int GM_BLOCK_K = 128;      // bytes of K per stage row = one 128B swizzle atom
constexpr int GM_UMMA_K = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int GM_THREADS = 192;      // warp 0: TMA / operand producer control, warp 1: MMA, warps 2-5: epilogue

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(n), "r"(s_u32(b))); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra LAB_DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "LAB_DONE:\n\t}" ::"r"(s_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups of 128 B rows, 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

// kind::i8 instruction descriptor: unsigned 8-bit A and B (K-major), S32 accumulate, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_i8(int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(GM_BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ------------------------------------------------------------------------------------------
// s2d_overlap_i8: TMA-staged planes
// ------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(GM_THREADS, 1)
overlap_i8_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  int Na, int Nb, int kblocks_total, int kblocks_per_split, int32_t* __restrict__ I) {
    constexpr int A_BYTES = GM_BLOCK_M * GM_BLOCK_K;
    constexpr int B_BYTES = BN * GM_BLOCK_K;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    uint8_t* gsm = gsm_raw + ((1024u - (s_u32(gsm_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1024 B alignment
    uint8_t* sA = gsm;                                  // STAGES x A tile
    uint8_t* sB = gsm + STAGES * A_BYTES;               // STAGES x B tile
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accum_full;
    __shared__ uint32_t tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GM_BLOCK_M, n0 = blockIdx.y * BN;
    const int kb0 = blockIdx.z * kblocks_per_split;
    const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        bar_init(&accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc<TCOLS>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (nkb > 0) {
        if (warp == 0 && lane == 0) {
            // ---- TMA producer ----
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) bar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                bar_expect(&full[s], A_BYTES + B_BYTES);
                tma_load_2d(sA + s * A_BYTES, &mapA, &full[s], (kb0 + i) * GM_BLOCK_K, m0);
                tma_load_2d(sB + s * B_BYTES, &mapB, &full[s], (kb0 + i) * GM_BLOCK_K, n0);
            }
        } else if (warp == 1 && lane == 0) {
            // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_i8(BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                bar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint64_t ad = umma_desc(s_u32(sA + s * A_BYTES));
                const uint64_t bd = umma_desc(s_u32(sB + s * B_BYTES));
#pragma unroll
                for (int k = 0; k < GM_BLOCK_K / GM_UMMA_K; ++k)
                    umma_i8(tmem, ad + (uint64_t)(k * GM_UMMA_K >> 4), bd + (uint64_t)(k * GM_UMMA_K >> 4), idesc, (i | k) != 0);
                tc_commit(&empty[s]);          // frees the stage when these MMAs have read it
            }
            tc_commit(&accum_full);
        } else if (warp >= 2) {
            // ---- epilogue: TMEM -> registers -> int32 atomics (split-K reduction) ----
            bar_wait(&accum_full, 0);
            tc_fence_after();
            const int qd = warp & 3;                    // TMEM lane quarter this warp may access
            const int row = m0 + qd * 32 + lane;
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + c0, v);
                if (row < Na) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = n0 + c0 + j;
                        S2D_DEV_ASSERT(col >= Nb || ((int64_t)row * Nb + col < (int64_t)Na * Nb));
                        if (col < Nb && v[j]) atomicAdd(&I[(int64_t)row * Nb + col], (int)v[j]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TCOLS>(tmem);
}

// ------------------------------------------------------------------------------------------
// s2d_overlap_gram_labels: one-hot operands synthesised on-chip from label bytes.
//   rows r = f * L + l  (f < F frames, l < L labels);  G[r, r'] = sum_px [lab_f[px]==l] [lab_f'[px]==l']
// 8 producer warps build the A (128 rows) and B (BN rows) tiles of a 128-pixel k-block directly in
// the SWIZZLE_128B K-major layout (16-byte chunk c of row r lives at chunk c ^ (r & 7)), warp 8
// issues the MMAs, warps 0-3 drain TMEM at the end (split-K atomics).
// ------------------------------------------------------------------------------------------
constexpr int GR_PF = 3;                 // label k-blocks in flight per producer group (cp.async groups)
constexpr int GR_STAGES = 4;             // operand stages
constexpr int GR_GROUPS = 2;             // independent producer groups working on alternate k-blocks
static_assert(GR_STAGES == 2 * GR_GROUPS, "each producer group alternates between exactly two stages");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One producer thread per operand row (128 A rows + BN B rows): the row's frame / label are fixed
// for the whole kernel, so a k-block costs 8 x (LDS.128 of the frame's label bytes, 4 byte-wise
// compares, STS.128 into the SWIZZLE_128B K-major slot: chunk c of row r lives at chunk c ^ (r & 7)).
// A warp covers 32 consecutive rows of one chunk: 1-3 distinct label addresses (broadcast) and
// 512 B of conflict-free stores per instruction.
template <int BN>
__global__ void __launch_bounds__(GR_GROUPS * (GM_BLOCK_M + BN) + 32, 1)
gram_labels_kernel(const uint8_t* __restrict__ labels, int F, int L, int64_t npix, int kblocks_total,
                   int kblocks_per_split, int nfr_max, int Rp, int mt, int32_t* __restrict__ part, int64_t part_ints) {
    constexpr int STAGES = GR_STAGES;
    constexpr int PRODUCERS = GM_BLOCK_M + BN;
    constexpr int A_BYTES = GM_BLOCK_M * GM_BLOCK_K;
    constexpr int B_BYTES = BN * GM_BLOCK_K;
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    uint8_t* gsm = gsm_raw + ((1024u - (s_u32(gsm_raw) & 1023u)) & 1023u);
    uint8_t* sOps = gsm;                                         // STAGES x (A tile | B tile)
    // per producer group: (GR_PF + 1) label slots x nfr_max x 128 B, then 2 slots x nfr_max x 8 chunk descriptors
    const int grp = threadIdx.x / PRODUCERS;                     // 0, 1: producer groups; 2: MMA warp
    uint8_t* sLab = gsm + STAGES * (A_BYTES + B_BYTES) + (grp & 1) * ((GR_PF + 1) * nfr_max * 128 + 2 * nfr_max * 8);
    uint8_t* sDesc = sLab + (GR_PF + 1) * nfr_max * 128;
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accum_full;
    __shared__ uint32_t tmem_base;

    const int R = F * L;
    const int tid = threadIdx.x % PRODUCERS, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // tid: index inside the group
    // G is symmetric: only the tiles that touch the upper triangle are computed (blockIdx.x walks them
    // column block by column block; column block nj holds the row blocks 0 .. min(mt - 1, nj * BN / 128 + BN / 128 - 1)),
    // gram_reduce_kernel mirrors them into the lower triangle.
    int mi = blockIdx.x, nj = 0;
    for (;;) {
        const int rows_here = min(mt, (nj + 1) * (BN / GM_BLOCK_M));
        if (mi < rows_here) break;
        mi -= rows_here; ++nj;
    }
    const int m0 = mi * GM_BLOCK_M, n0 = nj * BN;
    const int kb0 = blockIdx.z * kblocks_per_split;
    const int nkb = min(kblocks_total, kb0 + kblocks_per_split) - kb0;
    // frames whose labels the two operand tiles need
    const int fa0 = m0 / L, fa1 = min(R - 1, m0 + GM_BLOCK_M - 1) / L;
    const int fb0 = min(n0, R - 1) / L, fb1 = min(R - 1, n0 + BN - 1) / L;
    const int nfa = fa1 - fa0 + 1, nfb = fb1 - fb0 + 1;
    const int slot_bytes = nfr_max * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], PRODUCERS / 32); bar_init(&empty[s], 1); }
        bar_init(&accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<BN>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (nkb > 0) {
        if (grp < GR_GROUPS) {
            // ---- operand producers: this thread's row; group g builds k-blocks g, g+2, g+4, ... ----
            const bool isA = tid < GM_BLOCK_M;
            const int rl = isA ? tid : tid - GM_BLOCK_M;                 // row inside the A or B tile
            const int r = (isA ? m0 : n0) + rl;
            const bool rvalid = r < R;
            const int f = rvalid ? r / L : (isA ? fa0 : fb0);        // padding rows: any frame slot of their operand (output is zero)
            const uint32_t sp = rvalid ? (uint32_t)(r - f * L) * 0x01010101u : 0xFEFEFEFEu;   // 0xFE never matches (L <= 254)
            const int lab_off = (isA ? (f - fa0) : nfa + (f - fb0)) * 128;
            const int row_off = (isA ? 0 : A_BYTES) + rl * 128;
            const int r7 = rl & 7;

            auto issue_labels = [&](int ii) {       // label bytes of local k-block ii -> ring slot ii % (GR_PF + 1)
                const int i = ii * GR_GROUPS + grp;
                if (i < nkb) {
                    uint8_t* slot = sLab + (ii % (GR_PF + 1)) * slot_bytes;
                    const int64_t px0 = (int64_t)(kb0 + i) * GM_BLOCK_K;
                    for (int j = tid; j < (nfa + nfb) * 8; j += PRODUCERS) {
                        const int fr = j >> 3, c = j & 7;
                        const int ff = fr < nfa ? fa0 + fr : fb0 + (fr - nfa);
                        const int64_t px = px0 + 16 * c;
                        S2D_DEV_ASSERT(fr < nfr_max && ff >= 0 && ff < F && slot + fr * 128 + 16 * c + 16 <= sDesc);
                        if (px < npix) cp_async16(slot + fr * 128 + 16 * c, labels + (int64_t)ff * npix + px);
                        else *reinterpret_cast<uint4*>(slot + fr * 128 + 16 * c) = make_uint4(~0u, ~0u, ~0u, ~0u);   // 0xFF: no label
                    }
                }
                cp_async_commit();
            };
            // chunk descriptors of k-block i: the single label of a 16-pixel chunk, 0xFF when it is mixed.
            // Label maps are piecewise constant, so almost every chunk is uniform and a row's 16 output
            // bytes are all-ones or all-zeros without looking at the pixels.
            auto make_desc = [&](int ii) {
                if (ii * GR_GROUPS + grp < nkb) {
                    const uint8_t* slot = sLab + (ii % (GR_PF + 1)) * slot_bytes;
                    for (int j = tid; j < (nfa + nfb) * 8; j += PRODUCERS) {
                        const uint4 w = *reinterpret_cast<const uint4*>(slot + j * 16);
                        const bool uni = (w.x == w.y) & (w.y == w.z) & (w.z == w.w) & (w.x == __byte_perm(w.x, 0, 0x0000));
                        S2D_DEV_ASSERT(j < nfr_max * 8);
                        sDesc[(ii & 1) * nfr_max * 8 + j] = uni ? (uint8_t)(w.x & 255u) : (uint8_t)0xFF;
                    }
                }
            };
            // local sequence number ii <-> global k-block i = ii * GR_GROUPS + grp
            const int nloc = (nkb - grp + GR_GROUPS - 1) / GR_GROUPS;
            for (int j = 0; j < GR_PF; ++j) issue_labels(j);
            cp_async_wait<GR_PF - 1>();
            if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(PRODUCERS) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"(PRODUCERS) : "memory");
            make_desc(0);
            const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            const uint4 zeros = make_uint4(0, 0, 0, 0);
            const uint32_t mylab = sp & 255u;
            const int desc_off = (isA ? (f - fa0) : nfa + (f - fb0)) * 8;
            for (int ii = 0; ii < nloc; ++ii) {
                const int i = ii * GR_GROUPS + grp;
                const int s = i % STAGES;
                cp_async_wait<GR_PF - 2>();          // this thread's copies of local k-blocks <= ii+1 have landed
                // everybody's; desc(ii) visible; local k-block ii-1 consumed
                if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(PRODUCERS) : "memory");
                else asm volatile("bar.sync 2, %0;" ::"n"(PRODUCERS) : "memory");
                issue_labels(ii + GR_PF);            // refills the slot local k-block ii-1 used
                make_desc(ii + 1);
                if (i >= STAGES) bar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                const uint8_t* lab = sLab + (ii % (GR_PF + 1)) * slot_bytes + lab_off;
                uint8_t* dst = sOps + s * (A_BYTES + B_BYTES) + row_off;
                S2D_DEV_ASSERT(s < STAGES && row_off + 128 <= A_BYTES + B_BYTES && lab_off + 128 <= slot_bytes && desc_off + 8 <= nfr_max * 8);
                const uint2 d8 = *reinterpret_cast<const uint2*>(sDesc + (ii & 1) * nfr_max * 8 + desc_off);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t u = ((c < 4 ? d8.x : d8.y) >> (8 * (c & 3))) & 255u;
                    uint4* out = reinterpret_cast<uint4*>(dst + ((c ^ r7) << 4));
                    if (u != 0xFFu) {
                        *out = (u == mylab && rvalid) ? ones : zeros;
                    } else {                          // mixed chunk (object border): compare the 16 pixels
                        const uint4 w = *reinterpret_cast<const uint4*>(lab + 16 * c);
                        uint4 o;
                        o.x = __vcmpeq4(w.x, sp) & 0x01010101u;
                        o.y = __vcmpeq4(w.y, sp) & 0x01010101u;
                        o.z = __vcmpeq4(w.z, sp) & 0x01010101u;
                        o.w = __vcmpeq4(w.w, sp) & 0x01010101u;
                        *out = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (MMA) reads
                __syncwarp();
                if (lane == 0) bar_arrive(&full[s]);                           // one arrival per producer warp
            }
            // ---- epilogue (first four producer warps own the four TMEM lane quarters): the split's
            // partial tile goes out with plain 128-bit stores; gram_reduce_kernel sums the splits ----
            if (warp < 4) {
                bar_wait(&accum_full, 0);
                tc_fence_after();
                const int row = m0 + warp * 32 + lane;
                int32_t* prow = part + ((int64_t)blockIdx.z * R + row) * Rp;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                    if (row < R) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int col = n0 + c0 + j;
                            S2D_DEV_ASSERT(col >= Rp || (((int64_t)blockIdx.z * R + row) * Rp + col + 4 <= part_ints));
                            if (col < Rp) *reinterpret_cast<uint4*>(prow + col) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            }
        } else if (lane == 0) {
            // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_i8(BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                bar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint64_t ad = umma_desc(s_u32(sOps + s * (A_BYTES + B_BYTES)));
                const uint64_t bd = umma_desc(s_u32(sOps + s * (A_BYTES + B_BYTES) + A_BYTES));
#pragma unroll
                for (int k = 0; k < GM_BLOCK_K / GM_UMMA_K; ++k)
                    umma_i8(tmem, ad + (uint64_t)(k * GM_UMMA_K >> 4), bd + (uint64_t)(k * GM_UMMA_K >> 4), idesc, (i | k) != 0);
                tc_commit(&empty[s]);
            }
            tc_commit(&accum_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<BN>(tmem);
}

// ------------------------------------------------------------------------------------------
// Two m-tiles per CTA. The operand producers pace the kernel above (ncu: tensor pipe 39 % active): a
// k-block costs 128 + 256 rows of synthesis for one M128 x N256 MMA group. Here a CTA owns a 256 x 256
// output block: it synthesises 256 A rows + 256 B rows per k-block and issues TWO MMA groups (rows 0-127
// and 128-255 of the A tile against the same B tile) into the two halves of the 512 TMEM columns - a third
// less synthesis per MMA. One producer group of 512 threads (one per operand row), three 64 KB stages.
// ------------------------------------------------------------------------------------------
constexpr int G2_AM = 256;               // A rows per CTA (two M = 128 instruction tiles)
constexpr int G2_BN = 256;
constexpr int G2_STAGES = 3;
constexpr int G2_GROUPS = 2;             // producer groups working on alternate k-blocks
constexpr int G2_GTHREADS = 256;         // threads per group: thread t builds A row t and B row t
constexpr int G2_PRODUCERS = G2_GROUPS * G2_GTHREADS;

template <int STAGES>     // 3 operand stages; 2 when the label ring of a shape with few labels per frame needs the room
__global__ void __launch_bounds__(G2_PRODUCERS + 32, 1)
gram_labels2_kernel(const uint8_t* __restrict__ labels, int F, int L, int64_t npix, int kblocks_total,
                    int nt, int s_off, int s_diag, int per_off, int per_diag, int nfr_max, int Rp,
                    int32_t* __restrict__ part, int64_t part_ints, const int2* __restrict__ blist) {
    constexpr int GT = G2_GTHREADS;
    constexpr int A_BYTES = G2_AM * GM_BLOCK_K;
    constexpr int B_BYTES = G2_BN * GM_BLOCK_K;
    static_assert(G2_AM == GT && G2_BN == GT, "one A row and one B row per producer thread");
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    uint8_t* gsm = gsm_raw + ((1024u - (s_u32(gsm_raw) & 1023u)) & 1023u);
    uint8_t* sOps = gsm;                                         // STAGES x (A tile | B tile)
    const int grp = threadIdx.x / GT;                            // 0, 1: producer groups; 2: MMA warp
    // per producer group: (GR_PF + 1) label slots x nfr_max x 128 B, then 2 slots x nfr_max x 8 chunk descriptors
    uint8_t* sLab = gsm + STAGES * (A_BYTES + B_BYTES) + (grp & 1) * ((GR_PF + 1) * nfr_max * 128 + 2 * nfr_max * 8);
    uint8_t* sDesc = sLab + (GR_PF + 1) * nfr_max * 128;
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accum_full;
    __shared__ uint32_t tmem_base;

    const int R = F * L;
    const int tid = threadIdx.x % GT, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // tid: index inside the group
    // symmetric: block column nj holds the block rows 0 .. nj (256 x 256 blocks). A diagonal block's A rows ARE its
    // B rows: it synthesises the B tile only and points both A descriptors into it.
    // banded form (blist != nullptr): an explicit list of blocks (only those whose frames lie within the band), every block
    // split s_off times; its partial tiles are stored tile by tile ([cta][256][256]) instead of in the dense [split][R][Rp]
    int x = blockIdx.x, mi = 0, nj = 0;
    if (blist) {
        const int2 b = blist[x / s_off];
        mi = b.x; nj = b.y; x = x % s_off;
    } else {
        for (;;) {
            const int cnt = (mi == nj) ? s_diag : s_off;
            if (x < cnt) break;
            x -= cnt;
            if (++mi > nj) { mi = 0; ++nj; }
        }
    }
    const bool diag = mi == nj;
    const int split = x, kblocks_per_split = diag ? per_diag : per_off;
    const int m0 = mi * G2_AM, n0 = nj * G2_BN;
    const int kb0 = split * kblocks_per_split;
    const int nkb = max(min(kblocks_total, kb0 + kblocks_per_split) - kb0, 0);
    const int fa0 = m0 / L, fa1 = min(R - 1, m0 + G2_AM - 1) / L;
    const int fb0 = min(n0, R - 1) / L, fb1 = min(R - 1, n0 + G2_BN - 1) / L;
    const int nfa = diag ? 0 : fa1 - fa0 + 1, nfb = fb1 - fb0 + 1;       // diagonal: no separate A frames
    const int slot_bytes = nfr_max * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], GT / 32); bar_init(&empty[s], 1); }
        bar_init(&accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;

    if (nkb > 0) {
        if (grp < G2_GROUPS) {
            // ---- operand producers: thread t of a group builds A row t and B row t; group g takes k-blocks g, g+2, ... ----
            const int rA = m0 + tid, rB = n0 + tid;
            const bool vA = rA < R && !diag, vB = rB < R;
            const int fA = rA < R ? rA / L : fa0, fB = vB ? rB / L : fb0;    // padding rows: frame slot 0 of their operand (output is zero)
            const uint32_t spA = vA ? (uint32_t)(rA - fA * L) * 0x01010101u : 0xFEFEFEFEu;   // 0xFE never matches (L <= 254)
            const uint32_t spB = vB ? (uint32_t)(rB - fB * L) * 0x01010101u : 0xFEFEFEFEu;
            const int frA = diag ? 0 : fA - fa0, frB = nfa + (fB - fb0);                      // frame slots in the label ring
            const int r7 = tid & 7;

            auto issue_labels = [&](int ii) {       // label bytes of local k-block ii -> ring slot ii % (GR_PF + 1)
                const int i = ii * G2_GROUPS + grp;
                if (i < nkb) {
                    uint8_t* slot = sLab + (ii % (GR_PF + 1)) * slot_bytes;
                    const int64_t px0 = (int64_t)(kb0 + i) * GM_BLOCK_K;
                    for (int j = tid; j < (nfa + nfb) * 8; j += GT) {
                        const int fr = j >> 3, c = j & 7;
                        const int ff = fr < nfa ? fa0 + fr : fb0 + (fr - nfa);
                        const int64_t px = px0 + 16 * c;
                        S2D_DEV_ASSERT(fr < nfr_max && ff >= 0 && ff < F && slot + fr * 128 + 16 * c + 16 <= sDesc);
                        if (px < npix) cp_async16(slot + fr * 128 + 16 * c, labels + (int64_t)ff * npix + px);
                        else *reinterpret_cast<uint4*>(slot + fr * 128 + 16 * c) = make_uint4(~0u, ~0u, ~0u, ~0u);   // 0xFF: no label
                    }
                }
                cp_async_commit();
            };
            auto make_desc = [&](int ii) {          // per 16-pixel chunk: its single label, 0xFF when mixed
                if (ii * G2_GROUPS + grp < nkb) {
                    const uint8_t* slot = sLab + (ii % (GR_PF + 1)) * slot_bytes;
                    for (int j = tid; j < (nfa + nfb) * 8; j += GT) {
                        const uint4 w = *reinterpret_cast<const uint4*>(slot + j * 16);
                        const bool uni = (w.x == w.y) & (w.y == w.z) & (w.z == w.w) & (w.x == __byte_perm(w.x, 0, 0x0000));
                        S2D_DEV_ASSERT(j < nfr_max * 8);
                        sDesc[(ii & 1) * nfr_max * 8 + j] = uni ? (uint8_t)(w.x & 255u) : (uint8_t)0xFF;
                    }
                }
            };
            auto group_sync = [&]() {
                if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
                else asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory");
            };
            const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            const uint4 zeros = make_uint4(0, 0, 0, 0);
            // one operand row of the k-block: 8 chunks of 16 pixels
            auto build_row = [&](uint8_t* dst, const uint8_t* lab, uint2 d8, uint32_t sp, bool rvalid) {
                const uint32_t mylab = sp & 255u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t u = ((c < 4 ? d8.x : d8.y) >> (8 * (c & 3))) & 255u;
                    uint4* out = reinterpret_cast<uint4*>(dst + ((c ^ r7) << 4));
                    if (u != 0xFFu) {
                        *out = (u == mylab && rvalid) ? ones : zeros;
                    } else {                          // mixed chunk (object border): compare the 16 pixels
                        const uint4 w = *reinterpret_cast<const uint4*>(lab + 16 * c);
                        uint4 o;
                        o.x = __vcmpeq4(w.x, sp) & 0x01010101u;
                        o.y = __vcmpeq4(w.y, sp) & 0x01010101u;
                        o.z = __vcmpeq4(w.z, sp) & 0x01010101u;
                        o.w = __vcmpeq4(w.w, sp) & 0x01010101u;
                        *out = o;
                    }
                }
            };
            const int nloc = (nkb - grp + G2_GROUPS - 1) / G2_GROUPS;
            for (int j = 0; j < GR_PF; ++j) issue_labels(j);
            cp_async_wait<GR_PF - 1>();
            group_sync();
            make_desc(0);
            for (int ii = 0; ii < nloc; ++ii) {
                const int i = ii * G2_GROUPS + grp;
                const int s = i % STAGES;
                cp_async_wait<GR_PF - 2>();          // this thread's copies of local k-blocks <= ii+1 have landed
                group_sync();                        // everybody's; desc(ii) visible; local k-block ii-1 consumed
                issue_labels(ii + GR_PF);            // refills the slot local k-block ii-1 used
                make_desc(ii + 1);
                if (i >= STAGES) bar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                const uint8_t* slot = sLab + (ii % (GR_PF + 1)) * slot_bytes;
                const uint8_t* dsc = sDesc + (ii & 1) * nfr_max * 8;
                uint8_t* stage = sOps + s * (A_BYTES + B_BYTES);
                S2D_DEV_ASSERT(s < STAGES && frA >= 0 && frA < nfr_max && frB >= 0 && frB < nfr_max && (frB + 1) * 128 <= slot_bytes);
                if (!diag)
                    build_row(stage + tid * 128, slot + frA * 128, *reinterpret_cast<const uint2*>(dsc + frA * 8), spA, vA);
                build_row(stage + A_BYTES + tid * 128, slot + frB * 128, *reinterpret_cast<const uint2*>(dsc + frB * 8), spB, vB);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (MMA) reads
                __syncwarp();
                if (lane == 0) bar_arrive(&full[s]);                           // one arrival per producer warp of the group
            }
            // ---- epilogue: all 16 producer warps drain TMEM. A warp may only touch the lane quarter (warp % 4); the four
            // warps of a quarter take every fourth 32-column block. Two accumulators (A rows 0-127 and 128-255) ----
            {
                const int qd = warp & 3, cgrp = warp >> 2;                 // warps 0-15: both producer groups
                bar_wait(&accum_full, 0);
                tc_fence_after();
#pragma unroll 1
                for (int acc = 0; acc < 2; ++acc) {
                    const int row = m0 + acc * GM_BLOCK_M + qd * 32 + lane;
                    int32_t* prow = blist ? part + ((int64_t)blockIdx.x * G2_AM + (row - m0)) * G2_BN - n0      // tile-addressed
                                          : part + ((int64_t)split * R + row) * Rp;
#pragma unroll 1
                    for (int c0 = cgrp * 32; c0 < G2_BN; c0 += 32 * (G2_PRODUCERS / 128)) {
                        uint32_t v[32];
                        tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + acc * G2_BN + c0, v);
                        if (row < R) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const int col = n0 + c0 + j;
                                S2D_DEV_ASSERT(col >= Rp || (prow + col >= part && prow + col + 4 <= part + part_ints));
                                if (col < Rp) *reinterpret_cast<uint4*>(prow + col) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                        }
                    }
                }
            }
        } else if (lane == 0) {
            // ---- MMA issuer: two M = 128 groups per k-block against the same B tile ----
            constexpr uint32_t idesc = umma_idesc_i8(G2_BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                bar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint8_t* abase = sOps + s * (A_BYTES + B_BYTES) + (diag ? A_BYTES : 0);     // diagonal: A = B
                const uint64_t ad0 = umma_desc(s_u32(abase));
                const uint64_t ad1 = umma_desc(s_u32(abase + GM_BLOCK_M * GM_BLOCK_K));
                const uint64_t bd = umma_desc(s_u32(sOps + s * (A_BYTES + B_BYTES) + A_BYTES));
#pragma unroll
                for (int k = 0; k < GM_BLOCK_K / GM_UMMA_K; ++k) {
                    const uint64_t ko = (uint64_t)(k * GM_UMMA_K >> 4);
                    umma_i8(tmem, ad0 + ko, bd + ko, idesc, (i | k) != 0);
                    umma_i8(tmem + G2_BN, ad1 + ko, bd + ko, idesc, (i | k) != 0);
                }
                tc_commit(&empty[s]);
            }
            tc_commit(&accum_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// G[r][c] = sum over splits of part[s][r][c]  (row pitch Rp in part, R in G)
// `splits_diag` > 0: 256 x 256 blocks on the diagonal were split `splits_diag` times, the others `splits` times
__global__ void gram_reduce_kernel(const int32_t* __restrict__ part, int splits, int splits_diag, int R, int Rp, int32_t* __restrict__ G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)R * R) return;
    const int r = (int)(i / R), c = (int)(i - (int64_t)r * R);
    const int rr = min(r, c), cc = max(r, c);          // only upper-triangle tiles were computed
    const int ns = (splits_diag > 0 && (rr >> 8) == (cc >> 8)) ? splits_diag : splits;
    int acc = 0;
    S2D_DEV_ASSERT(cc < Rp);
    for (int s = 0; s < ns; ++s) acc += part[((int64_t)s * R + rr) * Rp + cc];
    G[i] = acc;
}

// ------------------------------------------------------------------------------------------
// Banded form: only pairs of rows whose frames are at most `band` apart ("all mask pairs within a frame window").
// Blocks (mi <= nj) of 256 x 256 rows are kept when the closest frames of their row / column ranges are within the band.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline bool gram_band_block(int mi, int nj, int R, int L, int band) {
    const int fa1 = (min(R - 1, mi * G2_AM + G2_AM - 1)) / L;      // last frame of the block's rows
    const int fb0 = (nj * G2_BN) / L;                               // first frame of its columns (nj >= mi)
    return fb0 - fa1 <= band;
}

// index[mi * nt + nj] = position of block (mi, nj) in blist, -1 when it is outside the band; one thread (<= a few thousand blocks)
__global__ void gram_band_plan_kernel(int nt, int R, int L, int band, int32_t* __restrict__ index, int2* __restrict__ blist) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0;
    for (int nj = 0; nj < nt; ++nj)
        for (int mi = 0; mi < nt; ++mi) {
            int v = -1;
            if (mi <= nj && gram_band_block(mi, nj, R, L, band)) { blist[n] = make_int2(mi, nj); v = n++; }
            index[mi * nt + nj] = v;
        }
}

// Gband[r][(d + band) * L + l2] = overlap of row r = (f, l) with row (f + d, l2), 0 when frame f + d does not exist
__global__ void gram_band_reduce_kernel(const int32_t* __restrict__ part, const int32_t* __restrict__ index, int nt, int splits,
                                        int F, int L, int band, int32_t* __restrict__ Gband) {
    const int Wb = (2 * band + 1) * L;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)F * L * Wb) return;
    const int r = (int)(i / Wb), c = (int)(i - (int64_t)r * Wb);
    const int f2 = r / L + c / L - band;
    int acc = 0;
    if (f2 >= 0 && f2 < F) {
        const int r2 = f2 * L + c % L;
        const int rr = min(r, r2), cc = max(r, r2);
        const int bi = index[(rr >> 8) * nt + (cc >> 8)];
        S2D_DEV_ASSERT(bi >= 0);
        if (bi >= 0)
            for (int s = 0; s < splits; ++s) acc += part[(((int64_t)bi * splits + s) * G2_AM + (rr & 255)) * G2_BN + (cc & 255)];
    }
    Gband[i] = acc;
}

static int gram_band_blocks(int R, int L, int band) {
    const int nt = (R + G2_BN - 1) / G2_BN;
    int n = 0;
    for (int nj = 0; nj < nt; ++nj)
        for (int mi = 0; mi <= nj; ++mi) n += gram_band_block(mi, nj, R, L, band) ? 1 : 0;
    return n;
}

// splits of the pixel range per block: the count (<= 8) that fills whole waves of 148 CTAs best
static int gram_band_splits(int nblk, int kblocks) {
    int best = 1; double beff = -1.0;
    for (int sp = 1; sp <= 8 && sp <= kblocks; ++sp) {
        const double waves = (double)nblk * sp / 148.0;
        const double eff = waves / (double)((int64_t)(waves + 0.999999));
        if (eff > beff + 1e-9) { beff = eff; best = sp; }
    }
    return best;
}

static int gram_tiles(int mt, int nt, int BN) {        // tiles that touch the upper triangle (see gram_labels_kernel)
    int n = 0;
    for (int nj = 0; nj < nt; ++nj) { const int rows_here = (nj + 1) * (BN / GM_BLOCK_M); n += rows_here < mt ? rows_here : mt; }
    return n;
}

static void gram_plan(int R, int BN, int64_t npix, int* mt, int* nt, int* kblocks, int* per, int* splits, int* Rp) {
    *mt = (R + GM_BLOCK_M - 1) / GM_BLOCK_M;
    *nt = (R + BN - 1) / BN;
    *kblocks = (int)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K);
    int sp = 148 / gram_tiles(*mt, *nt, BN);  // one wave of CTAs (1 CTA per SM)
    if (sp > *kblocks) sp = *kblocks;
    if (sp < 1) sp = 1;
    *per = (*kblocks + sp - 1) / sp;
    *splits = (*kblocks + *per - 1) / *per;
    *Rp = *nt * BN;                          // every column an epilogue may store exists
}

// splits of the pixel range per block so that one wave of <= 148 CTAs finishes together: a diagonal block costs
// G2_DIAG_COST of an off-diagonal one (it synthesises half the operand rows, but see the constant)
// measured (tools/r02_k1_diag.sh, C2 video): 1.0 -> 0.310 ms, 0.85 -> 0.303, 0.7 -> 0.279, 0.55 -> 0.288, 0.4 -> 0.356. A diagonal
// block synthesises half the operand rows; with the cost at 1.0 its CTAs finished early and their SMs idled for a fifth of the kernel
constexpr double G2_DIAG_COST = 0.7;
static void gram2_plan(int R, int64_t npix, int* nt, int* kblocks, int* s_off, int* s_diag, int* per_off, int* per_diag, int* Rp) {
    *nt = (R + G2_BN - 1) / G2_BN;
    const int nd = *nt, no = *nt * (*nt - 1) / 2;
    *kblocks = (int)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K);
    int best_o = no ? 1 : 0, best_d = 1;
    double best = 1e300;
    double diag_cost = G2_DIAG_COST;
#ifdef S2D_EXPERIMENTS
    if (getenv("S2D_GRAM_DIAG_COST")) diag_cost = atof(getenv("S2D_GRAM_DIAG_COST"));
#endif
    for (int sd = 1; sd <= 148; ++sd) {
        const int so = no ? (148 - nd * sd) / no : 0;
        if (no && so < 1) break;
        const double t = std::max(no ? 1.0 / so : 0.0, diag_cost / sd);       // time of the slowest CTA
        if (t < best) { best = t; best_o = so; best_d = sd; }
    }
    if (best_d > *kblocks) best_d = *kblocks;
    if (best_o > *kblocks) best_o = *kblocks;
    *per_off = best_o ? (*kblocks + best_o - 1) / best_o : 0;
    *per_diag = (*kblocks + best_d - 1) / best_d;
    *s_off = best_o ? (*kblocks + *per_off - 1) / *per_off : 0;       // no empty splits
    *s_diag = (*kblocks + *per_diag - 1) / *per_diag;
    *Rp = *nt * G2_BN;
}

static bool gram_use_v2(int R) {
#ifdef S2D_EXPERIMENTS   // A/B runs of the one-m-tile kernel on every shape (make exp)
    static const int forced = getenv("S2D_GRAM_V2") ? atoi(getenv("S2D_GRAM_V2")) : -1;
    if (forced >= 0) return forced != 0 && R > 128;
#endif
    return R > 256;
}

// shared memory of the two-m-tile kernel: the label ring grows with the frames a 256-row tile touches (256 / L + 2)
static int gram2_smem(int F, int L, int stages, int* nfr_max_out) {
    const int per_tile = G2_AM / L + 2;
    const int nfr_max = 2 * (F < per_tile ? F : per_tile);
    if (nfr_max_out) *nfr_max_out = nfr_max;
    return stages * (G2_AM + G2_BN) * GM_BLOCK_K + G2_GROUPS * ((GR_PF + 1) * nfr_max * 128 + 2 * nfr_max * 8) + 1024;
}
// operand stages of the 256 x 256 kernel for this shape: 3 if the label ring fits beside them, else 2, else 0 (narrower tilings)
static int gram2_stages(int F, int L) {
    constexpr int SMEM_MAX = 227 * 1024;
#ifdef S2D_EXPERIMENTS
    if (getenv("S2D_GRAM_STAGES") && atoi(getenv("S2D_GRAM_STAGES")) == 2) return gram2_smem(F, L, 2, nullptr) <= SMEM_MAX ? 2 : 0;
#endif
    if (gram2_smem(F, L, G2_STAGES, nullptr) <= SMEM_MAX) return G2_STAGES;
    if (gram2_smem(F, L, 2, nullptr) <= SMEM_MAX) return 2;
    return 0;
}

template <int STAGES>
static int launch_gram2(const uint8_t* labels, int F, int L, int64_t npix, int32_t* work, int32_t* G, cudaStream_t st) {
    const int R = F * L;
    int nfr_max;
    const int smem = gram2_smem(F, L, STAGES, &nfr_max);
    if (smem > 227 * 1024) { set_error("s2d_overlap_gram_labels: nlab=%d is too small for the label ring (needs %d B of shared memory)", L, smem); return -1; }
    cudaError_t e = cudaFuncSetAttribute(gram_labels2_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("gram_labels2_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
    int nt, kblocks, s_off, s_diag, per_off, per_diag, Rp;
    gram2_plan(R, npix, &nt, &kblocks, &s_off, &s_diag, &per_off, &per_diag, &Rp);
    const int nctas = nt * s_diag + nt * (nt - 1) / 2 * s_off;
    gram_labels2_kernel<STAGES><<<nctas, G2_PRODUCERS + 32, smem, st>>>(labels, F, L, npix, kblocks, nt, s_off, s_diag, per_off, per_diag,
                                                              nfr_max, Rp, work,
                                                              (int64_t)(s_off > s_diag ? s_off : s_diag) * R * Rp, nullptr);
    S2D_CHECK_LAUNCH("gram_labels2_kernel");
    gram_reduce_kernel<<<(unsigned)(((int64_t)R * R + 255) / 256), 256, 0, st>>>(work, s_off, s_diag, R, Rp, G);
    S2D_CHECK_LAUNCH("gram_reduce_kernel");
    return 0;
}

static int gram1_smem(int BN, int F, int L, int* nfr_max_out) {
    const int nfr_max = (F < GM_BLOCK_M / L + 2 ? F : GM_BLOCK_M / L + 2) + (F < BN / L + 2 ? F : BN / L + 2);
    if (nfr_max_out) *nfr_max_out = nfr_max;
    return GR_STAGES * (GM_BLOCK_M + BN) * GM_BLOCK_K + GR_GROUPS * ((GR_PF + 1) * nfr_max * 128 + 2 * nfr_max * 8) + 1024;
}

template <int BN>
static int launch_gram(const uint8_t* labels, int F, int L, int64_t npix, int32_t* work, int32_t* G, cudaStream_t st) {
    const int R = F * L;
    int nfr_max;
    const int smem = gram1_smem(BN, F, L, &nfr_max);
    if (smem > 227 * 1024) { set_error("s2d_overlap_gram_labels: nlab=%d is too small for the label ring (needs %d B of shared memory)", L, smem); return -1; }
    auto kfn = gram_labels_kernel<BN>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("gram_labels_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
    int mt, nt, kblocks, per, splits, Rp;
    gram_plan(R, BN, npix, &mt, &nt, &kblocks, &per, &splits, &Rp);
    dim3 grid(gram_tiles(mt, nt, BN), 1, splits);
    kfn<<<grid, GR_GROUPS * (GM_BLOCK_M + BN) + 32, smem, st>>>(labels, F, L, npix, kblocks, per, nfr_max, Rp, mt, work, (int64_t)splits * R * Rp);
    S2D_CHECK_LAUNCH("gram_labels_kernel");
    gram_reduce_kernel<<<(unsigned)(((int64_t)R * R + 255) / 256), 256, 0, st>>>(work, splits, 0, R, Rp, G);
    S2D_CHECK_LAUNCH("gram_reduce_kernel");
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_plane_map(CUtensorMap* map, const uint8_t* base, int rows, int64_t npix, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return -3; }
    cuuint64_t dims[2] = {(cuuint64_t)npix, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)npix};
    cuuint32_t box[2] = {(cuuint32_t)GM_BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

template <int BN>
static int launch_overlap_i8(const uint8_t* A, int Na, const uint8_t* B, int Nb, int64_t npix, int32_t* I, cudaStream_t st) {
    constexpr int STAGES = 4;
    CUtensorMap mapA, mapB;
    int rc = make_plane_map(&mapA, A, Na, npix, GM_BLOCK_M);
    if (rc) return rc;
    rc = make_plane_map(&mapB, B, Nb, npix, BN);
    if (rc) return rc;
    const int smem = STAGES * (GM_BLOCK_M + BN) * GM_BLOCK_K + 1024;
    auto kfn = overlap_i8_kernel<BN, STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("overlap_i8_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
    const int mt = (Na + GM_BLOCK_M - 1) / GM_BLOCK_M, nt = (Nb + BN - 1) / BN;
    const int kblocks = (int)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K);
    int splits = (148 * 2 + mt * nt - 1) / (mt * nt);
    if (splits > kblocks) splits = kblocks;
    if (splits < 1) splits = 1;
    const int per = (kblocks + splits - 1) / splits;
    splits = (kblocks + per - 1) / per;
    cudaMemsetAsync(I, 0, (size_t)Na * Nb * sizeof(int32_t), st);
    dim3 grid(mt, nt, splits);
    kfn<<<grid, GM_THREADS, smem, st>>>(mapA, mapB, Na, Nb, kblocks, per, I);
    S2D_CHECK_LAUNCH("overlap_i8_kernel");
    return 0;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_overlap_i8(const uint8_t* A, int Na, const uint8_t* B, int Nb, int64_t npix, int32_t* I, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(A && B && I && Na > 0 && Nb > 0 && npix > 0, "s2d_overlap_i8: bad arguments");
    S2D_CHECK_ARG(npix % 16 == 0 && (((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0,
                  "s2d_overlap_i8: planes must be 16-byte aligned with a pixel count that is a multiple of 16 (TMA); "
                  "use s2d_overlap_bits otherwise");
    cudaStream_t st = (cudaStream_t)stream;
    if (Nb <= 32) return launch_overlap_i8<32>(A, Na, B, Nb, npix, I, st);
    if (Nb <= 64) return launch_overlap_i8<64>(A, Na, B, Nb, npix, I, st);
    if (Nb <= 128) return launch_overlap_i8<128>(A, Na, B, Nb, npix, I, st);
    return launch_overlap_i8<256>(A, Na, B, Nb, npix, I, st);
}

extern "C" int s2d_overlap_gram_work_ints(int nframes, int nlab, int64_t npix, int64_t* out) {
    if (!out || nframes <= 0 || nlab <= 0 || npix <= 0) return -1;
    const int R = nframes * nlab;
    int mt, nt, kblocks, per, splits, Rp;
    gram_plan(R, 128, npix, &mt, &nt, &kblocks, &per, &splits, &Rp);
    int64_t need = (int64_t)splits * R * Rp + 4;
    if (R > 128) {                            // whichever tiling s2d_overlap_gram_labels ends up using
        gram_plan(R, 256, npix, &mt, &nt, &kblocks, &per, &splits, &Rp);
        const int64_t need1 = (int64_t)splits * R * Rp + 4;
        if (need1 > need) need = need1;
    }
    if (R > 128) {                            // the two-m-tile kernel uses more, shorter splits
        int s_off, s_diag, per_off, per_diag;
        gram2_plan(R, npix, &nt, &kblocks, &s_off, &s_diag, &per_off, &per_diag, &Rp);
        const int64_t need2 = (int64_t)(s_off > s_diag ? s_off : s_diag) * R * Rp + 4;
        if (need2 > need) need = need2;
    }
    *out = need;
    return 0;
}

extern "C" int s2d_overlap_gram_labels(const uint8_t* labels, int nframes, int nlab, int64_t npix, int32_t* work,
                                       int32_t* G, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(labels && G && work && nframes > 0 && nlab > 0 && nlab <= 254 && npix > 0, "s2d_overlap_gram_labels: bad arguments");
    S2D_CHECK_ARG(npix % 16 == 0 && (((uintptr_t)labels) & 15) == 0 && (((uintptr_t)work) & 15) == 0,
                  "s2d_overlap_gram_labels: label maps / work must be 16-byte aligned with a pixel count that is a multiple of 16");
    S2D_CHECK_ARG((int64_t)nframes * nlab <= 46340, "s2d_overlap_gram_labels: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = nframes * nlab;
    int tiling = 0;
    s2d_overlap_gram_tiling(nframes, nlab, &tiling);
    if (tiling == 2) return gram2_stages(nframes, nlab) == 2 ? launch_gram2<2>(labels, nframes, nlab, npix, work, G, st)
                                                             : launch_gram2<G2_STAGES>(labels, nframes, nlab, npix, work, G, st);
    if (tiling == 1) return launch_gram<256>(labels, nframes, nlab, npix, work, G, st);
    return launch_gram<128>(labels, nframes, nlab, npix, work, G, st);
}

extern "C" int s2d_overlap_gram_band_work_ints(int nframes, int nlab, int64_t npix, int band_frames, int64_t* out) {
    if (!out || nframes <= 0 || nlab <= 0 || npix <= 0 || band_frames < 0) return -1;
    const int R = nframes * nlab, nt = (R + G2_BN - 1) / G2_BN;
    const int nblk = gram_band_blocks(R, nlab, band_frames);
    const int kblocks = (int)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K);
    *out = (int64_t)nt * nt + 2 * (int64_t)nblk + 4 + (int64_t)nblk * gram_band_splits(nblk, kblocks) * G2_AM * G2_BN;
    return 0;
}

extern "C" int s2d_overlap_gram_labels_banded(const uint8_t* labels, int nframes, int nlab, int64_t npix, int band_frames,
                                              int32_t* work, int32_t* Gband, void* stream) {
    S2D_ENTER(stream);
    S2D_CHECK_ARG(labels && Gband && work && nframes > 0 && nlab > 0 && nlab <= 254 && npix > 0 && band_frames >= 0,
                  "s2d_overlap_gram_labels_banded: bad arguments");
    S2D_CHECK_ARG(npix % 16 == 0 && (((uintptr_t)labels) & 15) == 0 && (((uintptr_t)work) & 15) == 0,
                  "s2d_overlap_gram_labels_banded: label maps / work must be 16-byte aligned with a pixel count that is a multiple of 16");
    S2D_CHECK_ARG((int64_t)nframes * nlab <= 46340, "s2d_overlap_gram_labels_banded: too many rows");
    const int stages = gram2_stages(nframes, nlab);
    S2D_CHECK_ARG(stages > 0, "s2d_overlap_gram_labels_banded: nlab=%d is too small for the label ring of the 256 x 256 tiling", nlab);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = nframes * nlab, nt = (R + G2_BN - 1) / G2_BN;
    const int nblk = gram_band_blocks(R, nlab, band_frames);
    const int kblocks = (int)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K);
    int splits = gram_band_splits(nblk, kblocks);
    const int per = (kblocks + splits - 1) / splits;
    splits = (kblocks + per - 1) / per;                                   // no empty splits
    int32_t* index = work;
    int64_t off = (int64_t)nt * nt;
    off += off & 1;
    int2* blist = reinterpret_cast<int2*>(work + off);
    off += 2 * (int64_t)nblk;
    off = (off + 3) & ~(int64_t)3;
    int32_t* part = work + off;
    gram_band_plan_kernel<<<1, 32, 0, st>>>(nt, R, nlab, band_frames, index, blist);
    S2D_CHECK_LAUNCH("gram_band_plan_kernel");
    int nfr_max;
    const int smem = gram2_smem(nframes, nlab, stages, &nfr_max);
    const int64_t part_ints = (int64_t)nblk * splits * G2_AM * G2_BN;
    if (stages == 2) {
        cudaError_t e = cudaFuncSetAttribute(gram_labels2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gram_labels2_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
        gram_labels2_kernel<2><<<nblk * splits, G2_PRODUCERS + 32, smem, st>>>(labels, nframes, nlab, npix, kblocks, nt, splits, splits, per, per,
                                                                           nfr_max, G2_BN * nt, part, part_ints, blist);
    } else {
        cudaError_t e = cudaFuncSetAttribute(gram_labels2_kernel<G2_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gram_labels2_kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return -2; }
        gram_labels2_kernel<G2_STAGES><<<nblk * splits, G2_PRODUCERS + 32, smem, st>>>(labels, nframes, nlab, npix, kblocks, nt, splits, splits, per, per,
                                                                                   nfr_max, G2_BN * nt, part, part_ints, blist);
    }
    S2D_CHECK_LAUNCH("gram_labels2_kernel (banded)");
    const int64_t nout = (int64_t)R * (2 * band_frames + 1) * nlab;
    gram_band_reduce_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(part, index, nt, splits, nframes, nlab, band_frames, Gband);
    S2D_CHECK_LAUNCH("gram_band_reduce_kernel");
    return 0;
}

extern "C" int s2d_overlap_gram_executed_ops(int nframes, int nlab, int64_t npix, double* out) {
    if (!out || nframes <= 0 || nlab <= 0 || npix <= 0) return -1;
    const int R = nframes * nlab;
    int tiling = 0;
    s2d_overlap_gram_tiling(nframes, nlab, &tiling);
    const double kpix = (double)((npix + GM_BLOCK_K - 1) / GM_BLOCK_K) * GM_BLOCK_K;
    if (tiling == 2) {
        const int nt = (R + G2_BN - 1) / G2_BN;
        *out = (double)(nt * (nt + 1) / 2) * 2.0 * G2_AM * G2_BN * kpix;
    } else {
        const int BN = tiling == 1 ? 256 : 128;
        const int mt = (R + GM_BLOCK_M - 1) / GM_BLOCK_M, nt = (R + BN - 1) / BN;
        *out = (double)gram_tiles(mt, nt, BN) * 2.0 * GM_BLOCK_M * BN * kpix;
    }
    return 0;
}

extern "C" int s2d_overlap_gram_tiling(int nframes, int nlab, int* out) {
    if (!out || nframes <= 0 || nlab <= 0) return -1;
    const int R = nframes * nlab;
    *out = 0;
    if (R <= 128) return 0;
    // few labels per frame = many frames per operand tile = a bigger label ring: fall back to the narrower tilings
    // (256 x 256 two-m-tile -> 128 x 256 -> 128 x 128) until the ring fits beside the operand stages
    constexpr int SMEM_MAX = 227 * 1024;
    if (gram_use_v2(R) && gram2_stages(nframes, nlab) > 0) *out = 2;
    else if (gram1_smem(256, nframes, nlab, nullptr) <= SMEM_MAX) *out = 1;
    return 0;
}
