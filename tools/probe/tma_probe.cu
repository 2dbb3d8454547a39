// dev probe: which TMA 2-D box shapes / coordinates does the hardware accept?  usage: tma_probe boxw boxh c0 c1 cols rows pitch
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int nfloats, float* out) {
    extern __shared__ __align__(128) uint8_t raw[];
    uint8_t* sm = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
    uint64_t* bar = (uint64_t*)sm;
    float* dst = (float*)(sm + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nfloats * 4));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(dst)), "l"((uint64_t)&m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0; int spins = 0;
    while (!done && spins++ < (1 << 22))
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(smem_u32(bar)));
    if (threadIdx.x == 0 && !done) printf("timeout\n");
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = dst[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int boxw = atoi(argv[1]), boxh = atoi(argv[2]), c0 = atoi(argv[3]), c1 = atoi(argv[4]);
    long cols = atol(argv[5]), rows = atol(argv[6]), pitch = atol(argv[7]);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    float* g; cudaMalloc(&g, rows * pitch * 4);
    std::vector<float> h(rows * pitch);
    for (long i = 0; i < rows * pitch; ++i) h[i] = (float)(i / pitch) * 1000.f + (float)(i % pitch);
    cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t bx[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d ", (int)r);
    int n = boxw * boxh; float* out; cudaMalloc(&out, n * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 + n * 4);
    k<<<1, 128, 256 + n * 4>>>(m, c0, c1, n, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box=%dx%d c=(%d,%d) dims=%ldx%ld pitch=%ld -> %s", boxw, boxh, c0, c1, cols, rows, pitch, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < boxh; ++rr) for (int cc = 0; cc < boxw; ++cc) {
            long gr = c1 + rr, gc = c0 + cc; float exp = (gr >= 0 && gr < rows && gc >= 0 && gc < cols) ? gr * 1000.f + gc : 0.f;
            if (o[rr * boxw + cc] != exp) bad++;
        }
        printf(" mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
