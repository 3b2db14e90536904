// Microbenchmark: shared-memory de-duplication primitives on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_atom_bench smem_atom_bench.cu
// Each CTA (256 threads) does ITER rounds of 16 operations per thread on a 32 KB table with
// pseudo-random word addresses. Reports cycles per warp-level operation per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WORDS = 8192;
constexpr int ITER = 64;

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int spread) {
    __shared__ uint32_t bm[WORDS];
    __shared__ uint16_t own[WORDS];
    for (int i = threadIdx.x; i < WORDS; i += 256) { bm[i] = 0; own[i] = 0; }
    __syncthreads();
    uint32_t acc = 0;
    const uint32_t seed = blockIdx.x * 256 + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t h = hash(seed * 1315423911u + it * 16 + k);
            uint32_t kk = spread ? (h & 0x3FFFF) : (((seed * 16 + k) * 3 + it * 977) & 0x3FFFF);
            uint32_t w = (kk ^ (kk >> 5)) & (WORDS - 1);
            uint32_t bit = 1u << (kk >> 13);
            if (MODE == 0) {            // atomicOr with return
                uint32_t old = atomicOr(&bm[w], bit);
                acc += (bit & ~old) != 0;
            } else if (MODE == 1) {     // atomicOr, result unused (RED)
                atomicOr(&bm[w], bit);
            } else if (MODE == 2) {     // plain load + store (racy, cost reference)
                uint32_t old = bm[w];
                bm[w] = old | bit;
                acc += (bit & ~old) != 0;
            } else if (MODE == 3) {     // 16-bit owner election: store id, (sync), load
                own[kk & (WORDS - 1)] = (uint16_t)threadIdx.x;
            } else if (MODE == 4) {     // atomicAdd with return
                uint32_t old = atomicAdd(&bm[w], 1u);
                acc += old;
            } else if (MODE == 5) {     // atomicCAS
                uint32_t old = atomicCAS(&bm[w], 0u, bit);
                acc += old;
            }
        }
        if (MODE == 3) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint32_t h = hash(seed * 1315423911u + it * 16 + k);
                uint32_t kk = spread ? (h & 0x3FFFF) : (((seed * 16 + k) * 3 + it * 977) & 0x3FFFF);
                acc += own[kk & (WORDS - 1)] == (uint16_t)threadIdx.x;
            }
            __syncthreads();
        }
    }
    if (acc == 0xdeadbeef) out[0] = bm[threadIdx.x];
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int spread) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int grid = nsm * ctas_per_sm;
    uint32_t* out; cudaMalloc(&out, grid * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 256>>>(out, spread);
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) k<MODE><<<grid, 256>>>(out, spread);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    double warp_ops_per_sm = (double)ctas_per_sm * 8 * ITER * 16;
    printf("%-28s ctas/SM=%d spread=%d  %.3f ms  %.2f cycles per warp-op per SM (%s)\n", name, ctas_per_sm, spread, ms,
           cycles / warp_ops_per_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int spread = 0; spread < 2; ++spread)
        for (int c : {1, 2, 4}) {
            run<0>("atomicOr+ret", c, spread);
            run<1>("atomicOr noret", c, spread);
            run<2>("ld+st (racy)", c, spread);
            run<3>("u16 election st+sync+ld", c, spread);
            run<4>("atomicAdd+ret", c, spread);
            run<5>("atomicCAS", c, spread);
        }
    return 0;
}
