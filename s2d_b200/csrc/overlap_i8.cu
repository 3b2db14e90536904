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
