// tc_conv.cu - implicit-GEMM Conv1d / polyphase ConvTranspose1d on the 5th-gen tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged by TMA) for sm_100a.  Replaces, on the tensor-core path of the vocoder,
//   F.conv1d + leaky_relu + residual       reference utils/vocoder/models.py:31-38,96
//   F.conv_transpose1d                     reference utils/vocoder/models.py:98-99
//   xs += rb(x); x = xs / num_kernels      reference utils/vocoder/models.py:100-106
//
// Kernel anatomy (one persistent CTA per SM, 256 threads):
//   warp 0 / lane 0   producer: TMA 3-D tile loads of the activations (128B / 64B swizzle, out-of-range rows zero-filled by
//                     the TMA unit = the conv's zero padding) and 1-D bulk copies of the pre-packed weights
//   warp 1 / lane 0   MMA issuer: per tap one shifted view of the SAME activation tile (descriptor start address advanced by
//                     tap*dilation rows), 3 tcgen05.mma per 16-channel k-step (hi.hi + lo.hi + hi.lo), fp32 accumulate in TMEM
//   warp 2            TMEM allocation / release
//   warps 4..7        epilogue: tcgen05.ld of the accumulator (double buffered in TMEM so it overlaps the next tile's MMAs),
//                     bias + residual + resblock mean + leaky-relu + length mask, split to bf16 hi/lo, 32-byte stores
// Weights of the narrow layers (<= 100 KB packed) stay resident in shared memory for the whole launch; the wide layers
// stream one (channel chunk, tap) block at a time through a ring.
#include <cuda_bf16.h>

#include "pb_tc.cuh"

namespace pb {

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;                 // time rows per tile (UMMA M, cta_group::1)
constexpr int TC_MAX_STAGES = 8;
constexpr uint32_t TC_SMEM_LIMIT = 232448;  // 227 KB
constexpr uint32_t TC_B_RESIDENT_MAX = 100 * 1024;

struct TcParams {
    const uint8_t* w;
    const float* bias;
    const __nv_bfloat16* res;
    const __nv_bfloat16* accin;
    __nv_bfloat16* out;
    float* out_f32;
    const int32_t* lens;
    int len_mul;
    int B, L, Cin, N, NT, n_tiles_n, out_C, U;
    int K, D, pad;
    int nchunks, ksteps, cin_t;
    int a_loads, a_row_bytes, a_tile_bytes, a_stage_bytes, lo_off, a_layout, a_sbo;
    int n_a_stages, n_b_stages, b_block_bytes, b_resident, w_total_bytes;
    uint32_t smem_a_off, smem_b_off;
    int tiles_t, total_tiles;
    float post_slope, scale, res_unslope;
    int tanh_out;
    uint32_t idesc;
    int single_pass;
    int tmem_cols;
};

// ---- tcgen05 / TMEM wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 operands, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (sm_100 "version 1"): start address, leading / stride byte offsets in 16-byte units
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;            // 0 none, 2 128B, 4 64B, 6 32B swizzle
    return d;
}

__device__ __forceinline__ void st_global_v4(void* p, uint4 v) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap xmap, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint64_t* full_a = bars;
    uint64_t* empty_a = bars + TC_MAX_STAGES;
    uint64_t* full_b = bars + 2 * TC_MAX_STAGES;
    uint64_t* empty_b = bars + 3 * TC_MAX_STAGES;
    uint64_t* acc_full = bars + 4 * TC_MAX_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* b_res_full = acc_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sm_base = smem_u32(sm);

    if (threadIdx.x == 0) {
        prefetch_tmap(&xmap);
        for (int s = 0; s < TC_MAX_STAGES; ++s) {
            mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1);
            mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1);
        }
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_empty[0], 128); mbar_init(&acc_empty[1], 128);
        mbar_init(b_res_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_b = p.tiles_t * p.n_tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            // =========================== producer ===========================
            if (p.b_resident) {
                mbar_expect_tx(b_res_full, (uint32_t)p.w_total_bytes);
                for (int off = 0; off < p.w_total_bytes; off += 32768) {
                    const int n = p.w_total_bytes - off < 32768 ? p.w_total_bytes - off : 32768;
                    bulk_load_1d(sm + p.smem_b_off + off, p.w + off, (uint32_t)n, b_res_full);
                }
            }
            uint32_t a_it = 0, b_it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int b = tile / tiles_per_b;
                const int rem = tile - b * tiles_per_b;
                const int tt = rem / p.n_tiles_n, nt = rem - tt * p.n_tiles_n;
                const int row0 = tt * TC_M - p.pad;
                for (int c = 0; c < p.nchunks; ++c) {
                    const uint32_t s = a_it % p.n_a_stages, ph = (a_it / p.n_a_stages) & 1;
                    mbar_wait(&empty_a[s], ph ^ 1);
                    mbar_expect_tx(&full_a[s], (uint32_t)(p.a_loads * p.a_tile_bytes));
                    uint8_t* dst = sm + p.smem_a_off + s * p.a_stage_bytes;
                    tma_load_3d(dst, &xmap, c * p.cin_t, row0, b, &full_a[s]);
                    if (p.a_loads == 2) tma_load_3d(dst + p.a_tile_bytes, &xmap, p.Cin + c * p.cin_t, row0, b, &full_a[s]);
                    ++a_it;
                    if (!p.b_resident) {
                        for (int j = 0; j < p.K; ++j) {
                            const uint32_t sb = b_it % p.n_b_stages, phb = (b_it / p.n_b_stages) & 1;
                            mbar_wait(&empty_b[sb], phb ^ 1);
                            mbar_expect_tx(&full_b[sb], (uint32_t)p.b_block_bytes);
                            bulk_load_1d(sm + p.smem_b_off + sb * p.b_block_bytes,
                                         p.w + (size_t)((nt * p.nchunks + c) * p.K + j) * p.b_block_bytes, (uint32_t)p.b_block_bytes,
                                         &full_b[sb]);
                            ++b_it;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // =========================== MMA issuer ===========================
            if (p.b_resident) { mbar_wait(b_res_full, 0); tc_fence_after(); }
            uint32_t a_it = 0, b_it = 0, tile_i = 0;
            const uint32_t b_lbo = (uint32_t)p.NT * 16u;          // between the two 8-element K halves of a k-step
            const uint32_t kstep_b_bytes = (uint32_t)p.NT * 32u;  // one k-step (16 channels) of one precision plane
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_i) {
                const int rem = tile % tiles_per_b;
                const int nt = rem % p.n_tiles_n;
                const uint32_t as = tile_i & 1, aph = (tile_i >> 1) & 1;
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * (uint32_t)p.NT;
                uint32_t accumulate = 0;
                for (int c = 0; c < p.nchunks; ++c) {
                    const uint32_t s = a_it % p.n_a_stages, ph = (a_it / p.n_a_stages) & 1;
                    mbar_wait(&full_a[s], ph);
                    tc_fence_after();
                    const uint32_t a_base = sm_base + p.smem_a_off + s * p.a_stage_bytes;
                    for (int j = 0; j < p.K; ++j) {
                        uint32_t b_base;
                        uint32_t sb = 0;
                        if (p.b_resident) {
                            b_base = sm_base + p.smem_b_off + (uint32_t)((nt * p.nchunks + c) * p.K + j) * (uint32_t)p.b_block_bytes;
                        } else {
                            sb = b_it % p.n_b_stages;
                            const uint32_t phb = (b_it / p.n_b_stages) & 1;
                            mbar_wait(&full_b[sb], phb);
                            tc_fence_after();
                            b_base = sm_base + p.smem_b_off + sb * (uint32_t)p.b_block_bytes;
                        }
                        const uint32_t a_tap = a_base + (uint32_t)(j * p.D) * (uint32_t)p.a_row_bytes;
                        for (int ks = 0; ks < p.ksteps; ++ks) {
                            const uint32_t a_hi = a_tap + ks * 32u, a_lo = a_hi + (uint32_t)p.lo_off;
                            // The swizzle is a function of the absolute shared-memory address (measured on B200: a start
                            // address advanced by whole rows inside a 1024-byte-aligned tile needs base_offset = 0), so the
                            // K taps are K views of one staged tile.
                            const uint64_t da_hi = smem_desc(a_hi, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0);
                            const uint64_t db_hi = smem_desc(b_base + ks * kstep_b_bytes, b_lbo, 128, 0, 0);
                            umma_bf16(d_tmem, da_hi, db_hi, p.idesc, accumulate);
                            accumulate = 1;
                            if (!p.single_pass) {
                                const uint64_t da_lo = smem_desc(a_lo, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0);
                                const uint64_t db_lo = smem_desc(b_base + (p.ksteps + ks) * kstep_b_bytes, b_lbo, 128, 0, 0);
                                umma_bf16(d_tmem, da_lo, db_hi, p.idesc, 1);
                                umma_bf16(d_tmem, da_hi, db_lo, p.idesc, 1);
                            }
                        }
                        if (!p.b_resident) { umma_commit(&empty_b[sb]); ++b_it; }
                    }
                    umma_commit(&empty_a[s]);
                    ++a_it;
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int Lout = p.L * p.U;
        const size_t out_pitch = (size_t)2 * p.out_C;
        uint32_t tile_i = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_i) {
            const int b = tile / tiles_per_b;
            const int rem = tile - b * tiles_per_b;
            const int tt = rem / p.n_tiles_n, nt = rem - tt * p.n_tiles_n;
            const int t = tt * TC_M + row;
            const uint32_t as = tile_i & 1, aph = (tile_i >> 1) & 1;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            int valid = Lout;
            if (p.lens) { const int v = p.lens[b] * p.len_mul; valid = v < valid ? v : valid; }
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.NT;
            for (int n0 = 0; n0 < p.NT; n0 += 16) {
                float v[16];
                tmem_ld16(taddr0 + (uint32_t)n0, v);
                if (t >= p.L) continue;
                const int ng = nt * p.NT + n0;
                const int phs = ng / p.out_C, co = ng - phs * p.out_C;
                const int orow = t * p.U + phs;
                const size_t rowoff = ((size_t)b * Lout + orow) * out_pitch;
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + co + i);
                }
                if (p.res) {
                    const uint4* rh = reinterpret_cast<const uint4*>(p.res + rowoff + co);
                    const uint4* rl = reinterpret_cast<const uint4*>(p.res + rowoff + p.out_C + co);
                    const uint4 h0 = rh[0], h1 = rh[1], l0 = rl[0], l1 = rl[1];
                    const uint32_t hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    const uint32_t ll[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float r0 = bf16lo(hh[i]) + bf16lo(ll[i]);
                        float r1 = bf16hi(hh[i]) + bf16hi(ll[i]);
                        r0 = r0 < 0.f ? r0 * p.res_unslope : r0;
                        r1 = r1 < 0.f ? r1 * p.res_unslope : r1;
                        v[2 * i] += r0; v[2 * i + 1] += r1;
                    }
                }
                if (p.accin) {
                    const uint4* rh = reinterpret_cast<const uint4*>(p.accin + rowoff + co);
                    const uint4* rl = reinterpret_cast<const uint4*>(p.accin + rowoff + p.out_C + co);
                    const uint4 h0 = rh[0], h1 = rh[1], l0 = rl[0], l1 = rl[1];
                    const uint32_t hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    const uint32_t ll[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[2 * i] = (bf16lo(hh[i]) + bf16lo(ll[i])) + v[2 * i];
                        v[2 * i + 1] = (bf16hi(hh[i]) + bf16hi(ll[i])) + v[2 * i + 1];
                    }
                }
                const bool live = orow < valid;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float u = v[i] * p.scale;
                    u = fmaxf(u, u * p.post_slope);     // leaky relu for slope in [0,1]; slope 1 = identity
                    if (p.tanh_out) u = tanhf(u);
                    v[i] = live ? u : 0.f;
                }
                if (p.out) {
                    uint32_t hw[8], lw[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
                        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
                        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
                        hw[i] = pack_bf16(h0, h1);
                        lw[i] = pack_bf16(l0, l1);
                    }
                    __nv_bfloat16* oh = p.out + rowoff + co;
                    __nv_bfloat16* ol = p.out + rowoff + p.out_C + co;
                    st_global_v4(oh, make_uint4(hw[0], hw[1], hw[2], hw[3]));
                    st_global_v4(oh + 8, make_uint4(hw[4], hw[5], hw[6], hw[7]));
                    st_global_v4(ol, make_uint4(lw[0], lw[1], lw[2], lw[3]));
                    st_global_v4(ol + 8, make_uint4(lw[4], lw[5], lw[6], lw[7]));
                }
                if (p.out_f32) {
                    float* of = p.out_f32 + ((size_t)b * Lout + orow) * p.out_C + co;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(of + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------
int tc_n_tile(int N) {
    if (N <= 0 || N % 16) return 0;
    if (N <= 256) return N;
    if (N % 256 == 0) return 256;
    if (N % 128 == 0) return 128;
    return 0;
}

static int tc_cin_tile(int Cin) {
    if (Cin >= 64) return Cin % 64 == 0 ? 64 : 0;
    if (Cin == 32 || Cin == 16) return Cin;
    return 0;
}

size_t tc_packed_weight_bytes(int Cin, int N, int K) { return (size_t)Cin * N * K * 4; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// 3-D map over split-bf16 activations [B][L][2C]: dims (2C, L, B), box (box_c, rows, 1), 128B / 64B swizzle, zero OOB fill
static int make_tmap_act(CUtensorMap* map, const __nv_bfloat16* base, int C, int L, int B, int box_c, int rows, bool sw128) {
    EncodeTiledFn fn = encode_fn();
    PB_REQUIRE(fn != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    PB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, PB_ERR_INVALID, "tc_conv: activation base must be 16-byte aligned");
    cuuint64_t gdim[3] = {(cuuint64_t)2 * C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)4 * C, (cuuint64_t)4 * C * L};
    cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed (%d): C=%d L=%d B=%d box=%dx%d", (int)r, C, L, B, box_c, rows);
    return PB_OK;
}

int launch_tc_conv(const TcConvArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.x && a.w && (a.out || a.out_f32), PB_ERR_INVALID, "tc_conv: null tensor");
    PB_REQUIRE(a.B > 0 && a.L > 0 && a.K >= 1 && a.D >= 1 && a.U >= 1, PB_ERR_INVALID, "tc_conv: bad shape");
    const int cin_t = tc_cin_tile(a.Cin);
    const int NT = tc_n_tile(a.N);
    PB_REQUIRE(cin_t > 0, PB_ERR_INVALID, "tc_conv: Cin=%d unsupported (16, 32 or a multiple of 64)", a.Cin);
    PB_REQUIRE(NT > 0, PB_ERR_INVALID, "tc_conv: N=%d unsupported", a.N);
    PB_REQUIRE(a.out_C > 0 && a.out_C % 16 == 0 && a.N == a.U * a.out_C, PB_ERR_INVALID, "tc_conv: N=%d != U=%d * out_C=%d", a.N, a.U, a.out_C);
    PB_REQUIRE((a.K - 1) * a.D % 2 == 0, PB_ERR_INVALID, "tc_conv: (K-1)*D must be even ('same' padding)");
    TcParams p{};
    p.w = a.w; p.bias = a.bias; p.res = a.res; p.accin = a.accin; p.out = a.out; p.out_f32 = a.out_f32; p.lens = a.lens;
    p.len_mul = a.len_mul;
    p.B = a.B; p.L = a.L; p.Cin = a.Cin; p.N = a.N; p.NT = NT; p.n_tiles_n = a.N / NT; p.out_C = a.out_C; p.U = a.U;
    p.K = a.K; p.D = a.D; p.pad = (a.K - 1) * a.D / 2;
    p.cin_t = cin_t; p.nchunks = a.Cin / cin_t; p.ksteps = cin_t / 16;
    p.a_loads = a.Cin >= 64 ? 2 : 1;
    p.a_row_bytes = cin_t == 16 ? 64 : 128;
    const int a_rows = (int)round_up(TC_M + (a.K - 1) * a.D, 8);
    PB_REQUIRE(a_rows <= 256, PB_ERR_INVALID, "tc_conv: receptive field (K-1)*D=%d too wide", (a.K - 1) * a.D);
    p.a_tile_bytes = a_rows * p.a_row_bytes;
    p.a_stage_bytes = (int)round_up((int64_t)p.a_loads * p.a_tile_bytes, 1024);
    p.lo_off = a.Cin >= 64 ? p.a_tile_bytes : (a.Cin == 32 ? 64 : 32);
    p.a_layout = cin_t == 16 ? 4 : 2;
    p.a_sbo = 8 * p.a_row_bytes;
    p.b_block_bytes = 2 * p.ksteps * NT * 32;
    p.w_total_bytes = p.n_tiles_n * p.nchunks * a.K * p.b_block_bytes;
    const uint32_t budget = TC_SMEM_LIMIT - 2048;   // barriers + alignment slack
    p.smem_a_off = 1024;
    p.b_resident = (uint32_t)p.w_total_bytes <= TC_B_RESIDENT_MAX ? 1 : 0;
    uint32_t b_bytes;
    if (p.b_resident) {
        b_bytes = (uint32_t)round_up(p.w_total_bytes, 1024);
        p.n_b_stages = 1;
        p.n_a_stages = (int)((budget - b_bytes) / (uint32_t)p.a_stage_bytes);
    } else {
        p.n_a_stages = 2;
        const uint32_t left = budget - 2u * (uint32_t)p.a_stage_bytes;
        p.n_b_stages = (int)(left / (uint32_t)p.b_block_bytes);
        if (p.n_b_stages > TC_MAX_STAGES) p.n_b_stages = TC_MAX_STAGES;
        // spend what is left on deeper activation prefetch
        b_bytes = (uint32_t)p.n_b_stages * (uint32_t)p.b_block_bytes;
        p.n_a_stages = (int)((budget - b_bytes) / (uint32_t)p.a_stage_bytes);
    }
    if (p.n_a_stages > TC_MAX_STAGES) p.n_a_stages = TC_MAX_STAGES;
    PB_REQUIRE(p.n_a_stages >= 1 && p.n_b_stages >= 1, PB_ERR_INVALID, "tc_conv: tile does not fit in shared memory (Cin=%d N=%d K=%d D=%d)", a.Cin, a.N, a.K, a.D);
    p.smem_b_off = p.smem_a_off + (uint32_t)p.n_a_stages * (uint32_t)p.a_stage_bytes;
    const uint32_t smem_bytes = p.smem_b_off + b_bytes + 1024;
    PB_REQUIRE(smem_bytes <= TC_SMEM_LIMIT, PB_ERR_INVALID, "tc_conv: shared memory plan %u exceeds the limit", smem_bytes);
    p.tiles_t = (int)ceil_div(a.L, TC_M);
    const int64_t total = (int64_t)a.B * p.tiles_t * p.n_tiles_n;
    PB_REQUIRE(total < (1LL << 31), PB_ERR_INVALID, "tc_conv: too many tiles");
    p.total_tiles = (int)total;
    p.post_slope = a.post_slope; p.scale = a.scale; p.res_unslope = a.res_unslope; p.tanh_out = a.tanh_out;
    // instruction descriptor: D f32, A/B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    p.single_pass = a.single_pass;
    int cols = 32;
    while (cols < 2 * NT) cols *= 2;
    p.tmem_cols = cols;

    CUtensorMap xmap;
    PB_TRY(make_tmap_act(&xmap, a.x, a.Cin, a.L, a.B, cin_t == 16 ? 32 : 64, a_rows, cin_t != 16));

    static int n_sm = 0;
    static bool attr_set = false;
    if (!attr_set) {
        int dev = 0;
        PB_CUDA(cudaGetDevice(&dev));
        PB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        PB_CUDA(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT));
        attr_set = true;
    }
    const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
    tc_conv_kernel<<<grid, TC_THREADS, smem_bytes, st>>>(xmap, p);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// weight packing: B operand blocks in the no-swizzle K-major canonical layout (8x16-byte core matrices)
//   packed[((nt*nchunks + c)*K + j)] = block of 2 planes (hi, lo) x ksteps x [2 K-halves][NT/8 row groups][8 rows][8 bf16]
//   element (n, k) of a k-step: byte (kc*(NT/8) + n/8)*128 + (n%8)*16 + (k%8)*2, kc = k/8
// ------------------------------------------------------------------------------------------------------------------------
struct PackDev {
    const float* w;
    int64_t s_n, s_ci, s_j;   // strides of the source for (n, ci, tap); convtr handled by the tap table below
    int Cin, N, K, NT, cin_t, ksteps, nchunks;
    // transposed conv: N = U*Cout, 3 taps; source tap index for (phase, tap) or -1
    int convtr, Cout, Kt, U;
};

__global__ void tc_pack_kernel(const PackDev d, uint8_t* __restrict__ packed) {
    const int64_t total = (int64_t)d.Cin * d.N * d.K;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(packed);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // decode destination-major index: [nt][c][j][ks][kc][ng][r][e]  (per plane)
        int64_t r_ = i;
        const int e = (int)(r_ % 8); r_ /= 8;
        const int r = (int)(r_ % 8); r_ /= 8;
        const int ng = (int)(r_ % (d.NT / 8)); r_ /= (d.NT / 8);
        const int kc = (int)(r_ % 2); r_ /= 2;
        const int ks = (int)(r_ % d.ksteps); r_ /= d.ksteps;
        const int j = (int)(r_ % d.K); r_ /= d.K;
        const int c = (int)(r_ % d.nchunks); r_ /= d.nchunks;
        const int nt = (int)r_;
        const int n = nt * d.NT + ng * 8 + r;
        const int ci = c * d.cin_t + ks * 16 + kc * 8 + e;
        float v;
        if (!d.convtr) {
            v = d.w[n * d.s_n + ci * d.s_ci + j * d.s_j];
        } else {
            // n = phase*Cout + co; tap j <-> x[q + j - 1]; source tap jj = r0 + m*U with shift a - m = j - 1
            const int ph = n / d.Cout, co = n - ph * d.Cout;
            const int sft = ph + (d.Kt - d.U) / 2;
            const int a = sft / d.U, r0 = sft % d.U;
            const int m = a - (j - 1);
            const int jj = r0 + m * d.U;
            v = (m >= 0 && jj < d.Kt) ? d.w[((int64_t)ci * d.Cout + co) * d.Kt + jj] : 0.f;
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        // block base (in bf16 elements): block index * (2 planes * ksteps * NT * 16)
        const int64_t blk = ((int64_t)(nt * d.nchunks + c) * d.K + j) * (2LL * d.ksteps * d.NT * 16);
        const int64_t within = ((int64_t)(kc * (d.NT / 8) + ng) * 8 + r) * 8 + e;
        out[blk + (int64_t)ks * d.NT * 16 + within] = hi;
        out[blk + (int64_t)(d.ksteps + ks) * d.NT * 16 + within] = lo;
    }
}

static int pack_common(PackDev d, uint8_t* packed, cudaStream_t st) {
    d.cin_t = tc_cin_tile(d.Cin);
    d.NT = tc_n_tile(d.N);
    PB_REQUIRE(d.cin_t > 0 && d.NT > 0, PB_ERR_INVALID, "tc_pack: unsupported Cin=%d / N=%d", d.Cin, d.N);
    d.ksteps = d.cin_t / 16;
    d.nchunks = d.Cin / d.cin_t;
    const int64_t total = (int64_t)d.Cin * d.N * d.K;
    const int blocks = (int)(ceil_div(total, 256) < 2048 ? ceil_div(total, 256) : 2048);
    tc_pack_kernel<<<blocks, 256, 0, st>>>(d, packed);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int tc_pack_conv_weight(const float* w, uint8_t* packed, int Cout, int Cin, int K, cudaStream_t st) {
    PackDev d{};
    d.w = w; d.s_n = (int64_t)Cin * K; d.s_ci = K; d.s_j = 1; d.Cin = Cin; d.N = Cout; d.K = K;
    return pack_common(d, packed, st);
}

int tc_pack_convtr_weight(const float* w, uint8_t* packed, int Cin, int Cout, int Kt, int U, cudaStream_t st) {
    PB_REQUIRE((Kt - U) % 2 == 0 && Kt <= 3 * U, PB_ERR_INVALID, "tc_pack_convtr: kernel %d / stride %d unsupported", Kt, U);
    // every phase must only reach x[q-1], x[q], x[q+1]
    for (int ph = 0; ph < U; ++ph) {
        const int sft = ph + (Kt - U) / 2, a = sft / U, r0 = sft % U;
        for (int m = 0; r0 + m * U < Kt; ++m)
            PB_REQUIRE(a - m >= -1 && a - m <= 1, PB_ERR_INVALID, "tc_pack_convtr: kernel %d / stride %d needs more than 3 input taps", Kt, U);
    }
    PackDev d{};
    d.w = w; d.Cin = Cin; d.N = U * Cout; d.K = 3; d.convtr = 1; d.Cout = Cout; d.Kt = Kt; d.U = U;
    return pack_common(d, packed, st);
}

// ------------------------------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------------------------------
__global__ void f32_to_split_kernel(const float* __restrict__ x, int64_t pitch, __nv_bfloat16* __restrict__ y, int C, int L, float pre_slope) {
    // block: 32 positions x 32 channels, transposed through shared memory
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, t = t0 + tx;
        float v = (c < C && t < L) ? x[((size_t)b * C + c) * pitch + t] : 0.f;
        tile[cc][tx] = v > 0.f ? v : v * pre_slope;
    }
    __syncthreads();
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt, c = c0 + tx;
        if (t < L && c < C) {
            const float v = tile[tx][tt];
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            __nv_bfloat16* row = y + ((size_t)b * L + t) * 2 * C;
            row[c] = hi;
            row[C + c] = lo;
        }
    }
}

int launch_f32_to_split(const float* x, int64_t pitch, __nv_bfloat16* y, int B, int C, int L, float pre_slope, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    f32_to_split_kernel<<<grid, 256, 0, st>>>(x, pitch, y, C, L, pre_slope);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void split_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t pitch, int C, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt, c = c0 + tx;
        float v = 0.f;
        if (t < L && c < C) {
            const __nv_bfloat16* row = x + ((size_t)b * L + t) * 2 * C;
            v = __bfloat162float(row[c]) + __bfloat162float(row[C + c]);
        }
        tile[tt][tx] = v;
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, t = t0 + tx;
        if (c < C && t < L) y[((size_t)b * C + c) * pitch + t] = tile[tx][cc];
    }
}

int launch_split_to_f32(const __nv_bfloat16* x, float* y, int64_t pitch, int B, int C, int L, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    split_to_f32_kernel<<<grid, 256, 0, st>>>(x, y, pitch, C, L);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
