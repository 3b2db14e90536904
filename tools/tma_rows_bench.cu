// Microbenchmark: how fast can one SM pull a 2D region of a u8 label map into shared memory as
// one cp.async.bulk per row (rows are too narrow / unaligned for a tensor map: W = 854)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_rows_bench tma_rows_bench.cu
// Each CTA loops over "tiles": ROWS row copies of ROWB bytes from a random position of a 33 MB
// frame stack (L2 resident), completion on one mbarrier, all threads wait. Reports copies per
// microsecond per SM and bytes per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// MODE 0: every thread r < rows issues its own row (arrive.expect_tx per issuing thread)
// MODE 1: one warp issues all rows (lane-strided)
template <int MODE>
__global__ void __launch_bounds__(256) k(const uint8_t* __restrict__ frames, int W, int H, int T, int rows, int rowb,
                                         int iters, uint32_t* out) {
    extern __shared__ __align__(128) uint8_t tab[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, MODE == 0 ? 256 : 32); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint32_t acc = 0;
    const int pitch = rowb + 16;
    for (int it = 0; it < iters; ++it) {
        const uint32_t h = hash(blockIdx.x * 7919u + it);
        const int t = h % T, y0 = (h >> 8) % (H - rows), x0 = ((h >> 16) % (W - rowb - 16)) & ~15;
        const uint8_t* src = frames + ((size_t)t * H + y0) * W + x0;
        if (MODE == 0) {
            if (tid < rows) {
                mbar_expect_tx(&bar, rowb);
                bulk_g2s(tab + tid * pitch, src + (size_t)tid * W, rowb, &bar);
            } else {
                mbar_arrive(&bar);
            }
        } else if (tid < 32) {
            int n = 0;
            for (int r = tid; r < rows; r += 32) ++n;
            mbar_expect_tx(&bar, n * rowb);
            for (int r = tid; r < rows; r += 32) bulk_g2s(tab + r * pitch, src + (size_t)r * W, rowb, &bar);
        }
        mbar_wait(&bar, it & 1);
        acc += tab[(tid * 37) % (rows * pitch)];
        __syncthreads();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    out[blockIdx.x * 256 + tid] = acc;
}

template <int MODE>
void run(const uint8_t* frames, int W, int H, int T, int rows, int rowb, int ctas_per_sm, uint32_t* out) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * ctas_per_sm, iters = 400;
    const int smem = rows * (rowb + 16);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 256, smem>>>(frames, W, H, T, rows, rowb, iters, out);
    cudaEventRecord(a);
    k<MODE><<<grid, 256, smem>>>(frames, W, H, T, rows, rowb, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double copies = (double)grid * iters * rows;
    printf("mode=%d rows=%3d rowb=%4d ctas/SM=%d  %.3f ms  %.1f copies/us/SM  %.2f GB/s/SM  tile %.2f us  total %.2f TB/s (%s)\n", MODE, rows, rowb,
           ctas_per_sm, ms, copies / nsm / (ms * 1e3), copies * rowb / nsm / (ms * 1e6), ms * 1e3 / iters * 1.0,
           copies * rowb / (ms * 1e9), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int W = 1280, H = 720, T = 36;
    uint8_t* frames; cudaMalloc(&frames, (size_t)W * H * T);
    cudaMemset(frames, 1, (size_t)W * H * T);
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    for (int c : {1, 2, 3, 4}) {
        for (int rowb : {64, 128, 256, 512}) {
            const int rows = (rowb == 512) ? 80 : 144;
            run<0>(frames, W, H, T, rows, rowb, c, out);
            run<1>(frames, W, H, T, rows, rowb, c, out);
        }
    }
    return 0;
}
