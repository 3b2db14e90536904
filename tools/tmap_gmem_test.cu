// Does cp.async.bulk.tensor accept a CUtensorMap that lives in global memory (host-encoded, cudaMemcpy'd)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmap_gmem_test tmap_gmem_test.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const uint8_t* maps, int which, int x0, int y0, int bw, uint32_t* out, int fence) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const void* map = maps + which * 128;
    if (threadIdx.x == 0) {
        if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(map) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(16 * bw), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(buf)), "l"(map), "r"(smem_u32(&bar)), "r"(x0), "r"(y0) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    uint32_t s = 0;
    for (int i = threadIdx.x; i < 16 * bw; i += blockDim.x) s += buf[i];
    atomicAdd(out, s);
}
int main() {
    const int W = 1280, H = 720, T = 4;
    uint8_t* lab; cudaMalloc(&lab, (size_t)W * H * T);
    uint8_t* h = new uint8_t[(size_t)W * H * T];
    for (size_t i = 0; i < (size_t)W * H * T; ++i) h[i] = (uint8_t)((i % W) & 7);
    cudaMemcpy(lab, h, (size_t)W * H * T, cudaMemcpyHostToDevice);
    uint8_t hm[4 * 128];
    for (int kk = 0; kk < 4; ++kk) {
        CUtensorMap m;
        cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)T * H};
        cuuint64_t strides[1] = {(cuuint64_t)W};
        cuuint32_t box[2] = {(cuuint32_t)(64 * (kk + 1)), 16u};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, lab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode box %d: %d\n", 64 * (kk + 1), (int)r);
        memcpy(hm + kk * 128, &m, 128);
    }
    uint8_t* dm; cudaMalloc(&dm, 4 * 128); cudaMemcpy(dm, hm, 4 * 128, cudaMemcpyHostToDevice);
    uint32_t* out; cudaMalloc(&out, 4);
    for (int fence = 0; fence < 2; ++fence)
        for (int kk = 0; kk < 4; ++kk) {
            const int bw = 64 * (kk + 1);
            for (int x0 : {0, 16, 1104, 1264}) {
                cudaMemset(out, 0, 4);
                k<<<1, 128, 16 * 256>>>(dm, kk, x0, 700, bw, out, fence);
                cudaError_t e = cudaDeviceSynchronize();
                uint32_t got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
                uint32_t want = 0;
                for (int r = 0; r < 16; ++r) for (int c = 0; c < bw; ++c) if (x0 + c < W) want += (uint32_t)(((x0 + c) % W) & 7);
                printf("fence=%d box=%3d x0=%4d: %s got %u want %u\n", fence, bw, x0, cudaGetErrorString(e), got, want);
                if (e != cudaSuccess) return 1;
            }
        }
    return 0;
}
