// mma_probe.cu - microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, cta_group::1, SS mode) as a function of N, operand
// swizzle, A start-address shift, accumulator alternation and CTAs per SM.  Dev tool (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
    int N, iters, a_layout, a_lbo, a_sbo, a_shift, a_kstep, b_layout, b_lbo, b_sbo, b_kstep, n_acc, ksteps, a_tiles, a_tile_bytes, b_tiles, b_tile_bytes;
    int alt_N;      // 2CTA probe: != 0 -> every second MMA uses this N (alternating instruction shapes, as the stacked + lo pair of tc_conv does)
    int alt_grouped;   // != 0 -> the 4 MMAs of shape N first, then the 4 of shape alt_N (same work, shapes grouped)
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

template <int KS>
__global__ void __launch_bounds__(128) probe(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // zero operands
    for (int i = threadIdx.x; i < (c.a_tiles * c.a_tile_bytes + c.b_tiles * c.b_tile_bytes) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u / (c.N <= 128 && gridDim.x > 148 ? 2 : 1)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0;
    const bool leader = elect_one();
    if (threadIdx.x == 0) {
        t0 = clock64();
    }
    {
        // warp-uniform issue loop (descriptors stay in uniform registers); one elected lane issues
        uint32_t a_lo = (uint32_t)mk_desc(smem_u32(sm) + c.a_shift, c.a_lbo, c.a_sbo, c.a_layout);
        const uint32_t a_hi = (uint32_t)(mk_desc(0, c.a_lbo, c.a_sbo, c.a_layout) >> 32);
        uint32_t b_lo = (uint32_t)mk_desc(smem_u32(sm) + c.a_tiles * c.a_tile_bytes, c.b_lbo, c.b_sbo, c.b_layout);
        const uint32_t b_hi = (uint32_t)(mk_desc(0, c.b_lbo, c.b_sbo, c.b_layout) >> 32);
        const uint32_t a_step = c.a_kstep >> 4, b_step = c.b_kstep >> 4;
        const uint32_t a_flip = c.a_tile_bytes >> 4, b_flip = c.b_tile_bytes >> 4;
        uint32_t d = tmem;
        const uint32_t d_flip = c.n_acc > 1 ? c.N : 0;
        if (threadIdx.x < 32) {
            for (int i = 0; i < c.iters; i += 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t aa = a_lo + (h ? a_flip : 0), bb = b_lo + (h ? b_flip : 0), dd = d + (h ? d_flip : 0);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t da = ((uint64_t)a_hi << 32) | (aa + ks * a_step);
                            const uint64_t db = ((uint64_t)b_hi << 32) | (bb + ks * b_step);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd), "l"(da), "l"(db), "r"(idesc), "r"(1u)
                                : "memory");
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u / (c.N <= 128 && gridDim.x > 148 ? 2 : 1)) : "memory");
    }
}

static void run(const char* name, Cfg c, int grid) {
    long long* d;
    cudaMalloc(&d, grid * sizeof(long long));
    const int smem = c.a_tiles * c.a_tile_bytes + c.b_tiles * c.b_tile_bytes + 2048;
    auto kern = c.ksteps == 4 ? probe<4> : probe<2>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    kern<<<grid, 128, smem>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    kern<<<grid, 128, smem>>>(c, d);
    cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double n = (double)c.iters * c.ksteps;
    printf("%-44s N=%3d grid=%3d  cyc/mma min %.1f med %.1f max %.1f   (floor %d, smem B/mma %d)\n", name, c.N, grid, h[0] / n, h[grid / 2] / n,
           h[grid - 1] / n, c.N / 2, 4096 + 32 * c.N);
    cudaFree(d);
}


// ---- cta_group::2: M = 256 over a CTA pair; each CTA holds its own 128 A rows and N/2 rows of B ----------------------
template <int KS, int ALT>      // ALT: 0 = one instruction shape, 1 = shapes N / alt_N alternating, 2 = grouped (compile time: the plain loop stays as measured in round 1)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (c.a_tiles * c.a_tile_bytes + c.b_tiles * c.b_tile_bytes) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t idesc_alt = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((c.alt_N ? c.alt_N : c.N) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    long long t0 = 0;
    const bool leader = elect_one();
    if (threadIdx.x == 0 && rank == 0) {
        t0 = clock64();
    }
    {
        // warp-uniform issue loop (descriptors stay in uniform registers); one elected lane issues
        uint32_t a_lo = (uint32_t)mk_desc(smem_u32(sm) + c.a_shift, c.a_lbo, c.a_sbo, c.a_layout);
        const uint32_t a_hi = (uint32_t)(mk_desc(0, c.a_lbo, c.a_sbo, c.a_layout) >> 32);
        uint32_t b_lo = (uint32_t)mk_desc(smem_u32(sm) + c.a_tiles * c.a_tile_bytes, c.b_lbo, c.b_sbo, c.b_layout);
        const uint32_t b_hi = (uint32_t)(mk_desc(0, c.b_lbo, c.b_sbo, c.b_layout) >> 32);
        const uint32_t a_step = c.a_kstep >> 4, b_step = c.b_kstep >> 4;
        const uint32_t a_flip = c.a_tile_bytes >> 4, b_flip = c.b_tile_bytes >> 4;
        uint32_t d = tmem;
        const uint32_t d_flip = c.n_acc > 1 ? c.N : 0;
        if (threadIdx.x < 32 && rank == 0) {
            for (int i = 0; i < c.iters; i += 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t aa = a_lo + (h ? a_flip : 0), bb = b_lo + (h ? b_flip : 0), dd = d + (h ? d_flip : 0);
                    if (leader) {
                        if (ALT == 2) {
#pragma unroll
                            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint64_t da = ((uint64_t)a_hi << 32) | (aa + ks * a_step);
                                    const uint64_t db = ((uint64_t)b_hi << 32) | (bb + ks * b_step);
                                    asm volatile(
                                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd), "l"(da), "l"(db), "r"(pass ? idesc_alt : idesc), "r"(1u)
                                        : "memory");
                                }
                        } else {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t da = ((uint64_t)a_hi << 32) | (aa + ks * a_step);
                            const uint64_t db = ((uint64_t)b_hi << 32) | (bb + ks * b_step);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd), "l"(da), "l"(db), "r"(idesc), "r"(1u)
                                : "memory");
                            if (ALT == 1) {
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd), "l"(da), "l"(db), "r"(idesc_alt), "r"(1u)
                                    : "memory");
                            }
                        }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (threadIdx.x == 0 && rank == 0) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
        long long t1 = clock64();
        out[blockIdx.x / 2] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

static void run2(const char* name, Cfg c, int grid) {
    long long* d;
    cudaMalloc(&d, grid * sizeof(long long));
    const int smem = c.a_tiles * c.a_tile_bytes + c.b_tiles * c.b_tile_bytes + 2048;
    auto kern = c.alt_N == 0 ? probe2<4, 0> : (c.alt_grouped ? probe2<4, 2> : probe2<4, 1>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int r = 0; r < 2; ++r) {
        kern<<<grid, 128, smem>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    const int np = grid / 2;
    std::vector<long long> h(np);
    cudaMemcpy(h.data(), d, np * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double n = (double)c.iters * c.ksteps;
    printf("%-44s N=%3d grid=%3d  cyc/mma min %.1f med %.1f max %.1f   (M=256 pair; floor %d, smem B/mma/SM %d)\n", name, c.N, grid, h[0] / n,
           h[np / 2] / n, h[np - 1] / n, c.N / 2, 4096 + 16 * c.N);
    cudaFree(d);
}

int main() {
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int grid : {148, 296}) {
        for (int N : Ns) {
            // A: 128 rows x 128 B, 128B swizzle; 4 k-steps of 32 B inside the row.  a_tiles distinct tiles cycled.
            Cfg c{};
            c.N = N; c.iters = 2000; c.ksteps = 4; c.n_acc = 1;
            c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_shift = 0; c.a_kstep = 32; c.a_tiles = 2; c.a_tile_bytes = 144 * 128;
            // B: no-swizzle K-major core matrices: k-step plane = N x 32 B; LBO = N*16, SBO = 128
            c.b_layout = 0; c.b_lbo = N * 16; c.b_sbo = 128; c.b_kstep = N * 32; c.b_tiles = 2; c.b_tile_bytes = 4 * N * 32;
            if (grid == 296 && N == 256) continue;
            run("A sw128 aligned, B noswz", c, grid);
            c.a_shift = 5 * 128;
            run("A sw128 shifted 5 rows, B noswz", c, grid);
            c.a_shift = 0;
            c.n_acc = 512 / (grid == 296 ? 2 : 1) / N >= 2 ? 2 : 1;
            run("A sw128 aligned, B noswz, 2 acc", c, grid);
            c.n_acc = 1;
            // B 128B swizzle: N rows x 128 B
            c.b_layout = 2; c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tile_bytes = N * 128;
            run("A sw128 aligned, B sw128", c, grid);
            // both unswizzled (interleaved core matrices): A k-step plane = 128 x 32 B
            c.a_layout = 0; c.a_lbo = 128 * 16; c.a_sbo = 128; c.a_kstep = 128 * 32; c.a_tile_bytes = 4 * 128 * 32 + 2048;
            c.b_layout = 0; c.b_lbo = N * 16; c.b_sbo = 128; c.b_kstep = N * 32; c.b_tile_bytes = 4 * N * 32;
            {
                // A no-swizzle: LBO must be the K-direction core-matrix stride = 128 rows * 16 B
                Cfg c2 = c;
                run("A noswz, B noswz", c2, grid);
            }
            // A 64B swizzle rows (C=16 layout): row = 64 B, 2 k-steps (hi, lo) of 32 B
            c.a_layout = 4; c.a_lbo = 16; c.a_sbo = 512; c.a_kstep = 32; c.a_tile_bytes = 144 * 64; c.ksteps = 2;
            c.b_tile_bytes = 2 * N * 32;
            run("A sw64, B noswz", c, grid);
        }
    }
    for (int N : {32, 64, 128, 256}) {
        Cfg c{};
        c.N = N; c.iters = 2000; c.ksteps = 4; c.n_acc = 1;
        c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_shift = 0; c.a_kstep = 32; c.a_tiles = 2; c.a_tile_bytes = 144 * 128;
        const int Nh = N / 2;
        c.b_layout = 0; c.b_lbo = Nh * 16; c.b_sbo = 128; c.b_kstep = Nh * 32; c.b_tiles = 2; c.b_tile_bytes = 4 * Nh * 32;
        run2("2CTA: A sw128, B noswz (N/2 rows per CTA)", c, 148);
        c.b_layout = 2; c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32; c.b_tile_bytes = Nh * 128;
        if (Nh >= 8) run2("2CTA: A sw128, B sw128", c, 148);
    }
    {   // the stacked + lo pair of tc_conv's CTA-pair path: N = 256 and N = 128 MMAs alternating vs grouped (cycles per PAIR of MMAs / 2)
        Cfg c{};
        c.N = 256; c.iters = 2000; c.ksteps = 8; c.n_acc = 1;      // 8 MMAs per 4 k-steps
        c.a_layout = 2; c.a_lbo = 16; c.a_sbo = 1024; c.a_shift = 0; c.a_kstep = 32; c.a_tiles = 2; c.a_tile_bytes = 144 * 128;
        c.b_layout = 0; c.b_lbo = 128 * 16; c.b_sbo = 128; c.b_kstep = 128 * 32; c.b_tiles = 2; c.b_tile_bytes = 4 * 128 * 32;
        c.alt_N = 128; c.alt_grouped = 0;
        run2("2CTA: N=256 / N=128 ALTERNATING (avg of the two)", c, 148);
        c.alt_grouped = 1;
        run2("2CTA: N=256 x4 then N=128 x4 GROUPED (avg)", c, 148);
        c.N = 128; c.alt_N = 64; c.alt_grouped = 0;
        c.b_lbo = 64 * 16; c.b_kstep = 64 * 32; c.b_tile_bytes = 4 * 64 * 32;
        run2("2CTA: N=128 / N=64 ALTERNATING (avg)", c, 148);
        c.alt_grouped = 1;
        run2("2CTA: N=128 x4 then N=64 x4 GROUPED (avg)", c, 148);
    }
    return 0;
}
