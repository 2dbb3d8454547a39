// attn_tc.cu - fused multi-head attention of the FFT block on the 5th-gen tensor cores (sm_100a).
// Replaces, on the tensor-core path of the TTE, the nn.MultiheadAttention core of the reference
//   softmax(q.k^T / sqrt(d_head) + key_padding_mask) . v        modules/fft.py:53-59; torch F.multi_head_attention_forward
// (the projections around it are GEMMs of tc_conv.cu).  Scores are never written to HBM.
//
// One CTA = one (batch item, head, 128-query tile), 192 threads:
//   warp 0      producer: TMA loads of the Q tile and of 64-key K / V tiles (128-byte swizzle) from the split-fp16 qkv rows
//   warp 1      MMA issuer + TMEM owner:  S = Q.K^T  (operands from shared memory),  O += P.V  (P from TENSOR MEMORY, V from
//               shared memory as an MN-major operand - no transpose of V anywhere)
//   warps 2..5  softmax / epilogue: thread = query row = TMEM lane
// Products are evaluated on split-fp16 operands (x = hi + lo): a.b ~ a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, with [b_hi; b_lo]
// stacked along N so that a k-step costs two MMAs; accumulation is fp32 in TMEM.
// Softmax is exact two-pass (pass 1: row maximum and denominator; pass 2: normalised probabilities -> P.V), which keeps the
// reference's order of operations (normalise, then multiply by V) and needs no rescaling of the output accumulator.
#include <cuda_fp16.h>

#include "pb_tte.cuh"
#include "pb_umma.cuh"

namespace pb {

namespace {

constexpr int AT_THREADS = 192;
constexpr int AT_MQ = 128;      // queries per CTA
constexpr int AT_NK = 64;       // keys per tile
constexpr int AT_DH = 128;      // head width (d_model 256 / 2 heads)
constexpr int AT_STAGES = 2;
constexpr uint32_t AT_Q_BYTES = 4 * AT_MQ * 128;          // [hi c0][hi c1][lo c0][lo c1], each 128 rows x 128 bytes
constexpr uint32_t AT_K_BYTES = 4 * AT_NK * 128;          // per channel half: [hi rows][lo rows]
constexpr uint32_t AT_V_BYTES = 4 * AT_NK * 128;          // [Vhi d0][Vhi d1][Vlo d0][Vlo d1], each 64 keys x 128 bytes
constexpr uint32_t AT_STAGE_BYTES = AT_K_BYTES + AT_V_BYTES;
constexpr uint32_t AT_SMEM = 1024 + AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES + 1024;
// tensor memory columns
constexpr uint32_t TM_S = 0;      // 128: [0,64) hi.hi + lo.hi, [64,128) hi.lo
constexpr uint32_t TM_P = 128;    // 64:  [128,160) P_hi (64 keys, 2 per column), [160,192) P_lo
constexpr uint32_t TM_O = 256;    // 256: [256,384) P.V_hi, [384,512) P_hi.V_lo

struct AttnParams {
    const uint8_t* key_valid;     // [B][L] 1 = attend, or null
    uint16_t* out;                // split-fp16 rows [B*L][2*D]: head h writes hi at h*128, lo at D + h*128
    int B, L, H, D;
    float scale_log2e;            // (1 / sqrt(d_head)) * log2(e)
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

}  // namespace

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kvmap, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;                  // [AT_STAGES]
    uint64_t* kv_empty = kv_full + AT_STAGES;      // [AT_STAGES]
    uint64_t* s_full = kv_empty + AT_STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_ready = s_free + 1;
    uint64_t* p_free = p_ready + 1;
    uint64_t* o_full = p_free + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
    uint8_t* q_s = sm + 1024;
    uint8_t* kv_s = q_s + AT_Q_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * AT_MQ;
    const int n_tiles = (p.L + AT_NK - 1) / AT_NK;
    const int C3 = 3 * p.D;                 // channels per plane of a qkv row: [q | k | v]

    if (threadIdx.x == 0) {
        prefetch_tmap(&qmap);
        prefetch_tmap(&kvmap);
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_ready, 4); mbar_init(p_free, 1); mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // =========================== producer ===========================
        if (elect_one()) {
            mbar_expect_tx(q_full, AT_Q_BYTES);
            for (int pl = 0; pl < 2; ++pl)
                for (int c = 0; c < 2; ++c)
                    tma_load_3d(q_s + (pl * 2 + c) * (AT_MQ * 128), &qmap, pl * C3 + h * AT_DH + c * 64, q0, b, q_full);
        }
        __syncwarp();
        uint32_t st = 0, ph = 0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int t = 0; t < n_tiles; ++t) {
                mbar_wait(&kv_empty[st], ph ^ 1);
                if (elect_one()) {
                    uint8_t* ks = kv_s + st * AT_STAGE_BYTES;
                    uint8_t* vs = ks + AT_K_BYTES;
                    mbar_expect_tx(&kv_full[st], pass == 0 ? AT_K_BYTES : AT_STAGE_BYTES);
                    const int k0 = t * AT_NK;
                    for (int c = 0; c < 2; ++c)
                        for (int pl = 0; pl < 2; ++pl)   // per channel half: hi rows then lo rows (one stacked 128-row operand)
                            tma_load_3d(ks + (c * 2 + pl) * (AT_NK * 128), &kvmap, pl * C3 + p.D + h * AT_DH + c * 64, k0, b, &kv_full[st]);
                    if (pass == 1)
                        for (int pl = 0; pl < 2; ++pl)
                            for (int c = 0; c < 2; ++c)  // [Vhi d0][Vhi d1][Vlo d0][Vlo d1]: four 64-wide MN groups
                                tma_load_3d(vs + (pl * 2 + c) * (AT_NK * 128), &kvmap, pl * C3 + 2 * p.D + h * AT_DH + c * 64, k0, b, &kv_full[st]);
                }
                __syncwarp();
                if (++st == AT_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        // instruction descriptors: D f32, A/B f16; S: N = 128 (stacked) / 64, K-major B; O: N = 256 (stacked) / 128, MN-major B
        const uint32_t id_base = (1u << 4) | ((uint32_t)(AT_MQ >> 4) << 24);
        const uint32_t id_s2 = id_base | ((uint32_t)((2 * AT_NK) >> 3) << 17);
        const uint32_t id_s1 = id_base | ((uint32_t)(AT_NK >> 3) << 17);
        const uint32_t id_o2 = id_base | (1u << 16) | ((uint32_t)((2 * AT_DH) >> 3) << 17);
        const uint32_t id_o1 = id_base | (1u << 16) | ((uint32_t)(AT_DH >> 3) << 17);
        const uint32_t q_base = smem_u32(q_s), kv_base = smem_u32(kv_s);
        // K-major, 128-byte swizzle: rows 128 bytes apart, 8-row groups 1024 bytes apart
        const uint64_t kdesc0 = smem_desc(0, 16, 1024, 2, 0);
        // MN-major, 128-byte swizzle: 64-element MN groups LBO = one 64-key box apart, 8-key groups 1024 bytes apart
        const uint64_t vdesc0 = smem_desc(0, AT_NK * 128, 1024, 2, 0);
        auto addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3fffu); };
        mbar_wait(q_full, 0);
        tc_fence_after();
        uint32_t st = 0, ph = 0;
        auto issue_qk = [&](uint32_t stage) {
            const uint32_t kb = kv_base + stage * AT_STAGE_BYTES;
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t a_hi = addr(kdesc0, q_base + (0 * 2 + c) * (AT_MQ * 128) + ks * 32);
                    const uint64_t a_lo = addr(kdesc0, q_base + (1 * 2 + c) * (AT_MQ * 128) + ks * 32);
                    const uint64_t bd = addr(kdesc0, kb + c * (2 * AT_NK * 128) + ks * 32);
                    umma_bf16(tmem + TM_S, a_hi, bd, id_s2, acc);     // q_hi . [k_hi; k_lo]
                    umma_bf16(tmem + TM_S, a_lo, bd, id_s1, 1);       // q_lo . k_hi
                    acc = 1;
                }
        };
        uint32_t sfree_ph = 0, pready_ph = 0;
        // ---- pass 1: scores only (row maximum / denominator)
        for (int t = 0; t < n_tiles; ++t) {
            mbar_wait(&kv_full[st], ph);
            if (t > 0) { mbar_wait(s_free, sfree_ph); sfree_ph ^= 1; }
            tc_fence_after();
            if (elect_one()) {
                issue_qk(st);
                umma_commit(s_full);
                umma_commit(&kv_empty[st]);
            }
            __syncwarp();
            if (++st == AT_STAGES) { st = 0; ph ^= 1; }
        }
        // ---- pass 2: S(t+1) is issued as soon as the softmax warps have read S(t); P(t).V(t) follows when P(t) is in TMEM
        mbar_wait(&kv_full[st], ph);
        mbar_wait(s_free, sfree_ph); sfree_ph ^= 1;         // last tile of pass 1 consumed
        tc_fence_after();
        if (elect_one()) { issue_qk(st); umma_commit(s_full); }
        __syncwarp();
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t cur = st;
            uint32_t nst = st + 1, nph = ph;
            if (nst == AT_STAGES) { nst = 0; nph ^= 1; }
            if (t + 1 < n_tiles) {
                mbar_wait(&kv_full[nst], nph);
                mbar_wait(s_free, sfree_ph); sfree_ph ^= 1;
                tc_fence_after();
                if (elect_one()) { issue_qk(nst); umma_commit(s_full); }
                __syncwarp();
            }
            mbar_wait(p_ready, pready_ph); pready_ph ^= 1;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t vb = kv_base + cur * AT_STAGE_BYTES + AT_K_BYTES;
#pragma unroll
                for (int ks = 0; ks < AT_NK / 16; ++ks) {
                    const uint64_t bd = addr(vdesc0, vb + ks * (16 * 128));
                    umma_f16_ts(tmem + TM_O, tmem + TM_P + ks * 8, bd, id_o2, (t > 0 || ks > 0) ? 1u : 0u);   // p_hi . [v_hi; v_lo]
                    umma_f16_ts(tmem + TM_O, tmem + TM_P + 32 + ks * 8, bd, id_o1, 1);                          // p_lo . v_hi
                }
                umma_commit(&kv_empty[cur]);
                umma_commit(p_free);
                if (t + 1 == n_tiles) umma_commit(o_full);
            }
            __syncwarp();
            st = nst; ph = nph;
        }
    } else {
        // =========================== softmax / epilogue (thread = query row) ===========================
        const int q = warp & 3;                              // TMEM lane quarter of this warp
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = q * 32 + lane;
        const int qrow = q0 + row;
        const uint8_t* kval = p.key_valid ? p.key_valid + (size_t)b * p.L : nullptr;
        const float c2 = p.scale_log2e;
        float m = -INFINITY, l = 0.f;
        uint32_t sfull_ph = 0, pfree_ph = 0;
        // 64-bit validity mask of the tile's keys, identical in every lane
        auto tile_mask = [&](int t, uint32_t& lo32, uint32_t& hi32) {
            const int k0 = t * AT_NK;
            const int ka = k0 + lane, kb2 = k0 + 32 + lane;
            const bool va = ka < p.L && (!kval || kval[ka] != 0);
            const bool vb = kb2 < p.L && (!kval || kval[kb2] != 0);
            lo32 = __ballot_sync(0xffffffffu, va);
            hi32 = __ballot_sync(0xffffffffu, vb);
        };
        // ---- pass 1
        for (int t = 0; t < n_tiles; ++t) {
            uint32_t mk[2];
            tile_mask(t, mk[0], mk[1]);
            mbar_wait(s_full, sfull_ph); sfull_ph ^= 1;
            tc_fence_after();
            float s[AT_NK];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t a[32], c[32];
                tmem_ld32_nowait(tmem + lane_off + TM_S + half * 32, a);
                tmem_ld32_nowait(tmem + lane_off + TM_S + AT_NK + half * 32, c);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) s[half * 32 + i] = __uint_as_float(a[i]) + __uint_as_float(c[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            float tm = m;
#pragma unroll
            for (int i = 0; i < AT_NK; ++i)
                if ((mk[i >> 5] >> (i & 31)) & 1u) tm = fmaxf(tm, s[i]);
            if (tm > -INFINITY) {
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < AT_NK; ++i)
                    if ((mk[i >> 5] >> (i & 31)) & 1u) sum += exp2f((s[i] - tm) * c2);
                l = l * exp2f((m - tm) * c2) + sum;     // m = -inf on the first valid tile: exp2(-inf) = 0
                m = tm;
            }
        }
        const float inv_l = 1.0f / l;
        // ---- pass 2
        for (int t = 0; t < n_tiles; ++t) {
            uint32_t mk[2];
            tile_mask(t, mk[0], mk[1]);
            mbar_wait(s_full, sfull_ph); sfull_ph ^= 1;
            tc_fence_after();
            float s[AT_NK];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t a[32], c[32];
                tmem_ld32_nowait(tmem + lane_off + TM_S + half * 32, a);
                tmem_ld32_nowait(tmem + lane_off + TM_S + AT_NK + half * 32, c);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) s[half * 32 + i] = __uint_as_float(a[i]) + __uint_as_float(c[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);          // S may be overwritten by the next tile's Q.K^T
#pragma unroll
            for (int i = 0; i < AT_NK; ++i)
                s[i] = ((mk[i >> 5] >> (i & 31)) & 1u) ? exp2f((s[i] - m) * c2) * inv_l : 0.f;
            // P(t-1) must have been consumed by its MMAs before it is overwritten
            if (t > 0) { mbar_wait(p_free, pfree_ph); pfree_ph ^= 1; }
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < AT_NK / 32; ++g) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = s[g * 32 + 2 * i], c = s[g * 32 + 2 * i + 1];
                    const __half2 hh = __floats2half2_rn(a, c);
                    const float2 hf = __half22float2(hh);
                    hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo[i] = pack_h2(a - hf.x, c - hf.y);
                }
                tmem_st16(tmem + lane_off + TM_P + g * 16, hi);
                tmem_st16(tmem + lane_off + TM_P + 32 + g * 16, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
        }
        // ---- epilogue: O = P.V_hi + P_hi.V_lo -> split-fp16 row [hi(128) | lo(128)] of head h
        mbar_wait(o_full, 0);
        tc_fence_after();
        uint16_t* orow = p.out + ((size_t)b * p.L + qrow) * 2 * p.D + h * AT_DH;
#pragma unroll 1
        for (int g = 0; g < AT_DH / 32; ++g) {
            uint32_t a[32], c[32];
            tmem_ld32_nowait(tmem + lane_off + TM_O + g * 32, a);
            tmem_ld32_nowait(tmem + lane_off + TM_O + AT_DH + g * 32, c);
            tmem_ld_wait();
            if (qrow < p.L) {
                uint32_t hw[16], lw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x0 = __uint_as_float(a[2 * i]) + __uint_as_float(c[2 * i]);
                    const float x1 = __uint_as_float(a[2 * i + 1]) + __uint_as_float(c[2 * i + 1]);
                    const __half2 hh = __floats2half2_rn(x0, x1);
                    const float2 hf = __half22float2(hh);
                    hw[i] = *reinterpret_cast<const uint32_t*>(&hh);
                    lw[i] = pack_h2(x0 - hf.x, x1 - hf.y);
                }
                uint4* oh = reinterpret_cast<uint4*>(orow + g * 32);
                uint4* ol = reinterpret_cast<uint4*>(orow + p.D + g * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    oh[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
                    ol[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// 3-D map over split-fp16 qkv rows [B][L][2*3D]: dims (6D, L, B), box (64, rows, 1), 128-byte swizzle, zero fill past L
static int make_tmap_qkv(CUtensorMap* map, const uint16_t* base, int D, int L, int B, int rows) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        PB_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess,
                   PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    cuuint64_t gdim[3] = {(cuuint64_t)6 * D, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)12 * D, (cuuint64_t)12 * D * L};
    cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed (%d): D=%d L=%d B=%d", (int)r, D, L, B);
    return PB_OK;
}

int launch_attention_tc(const uint16_t* qkv, const uint8_t* key_valid, uint16_t* out, int B, int L, int H, int D, cudaStream_t st) {
    PB_REQUIRE(qkv && out && B > 0 && L > 0, PB_ERR_INVALID, "attention_tc: bad argument");
    PB_REQUIRE(H * AT_DH == D, PB_ERR_INVALID, "attention_tc: head width %d unsupported (128 expected)", H > 0 ? D / H : 0);
    PB_REQUIRE(B <= 65535 && H <= 65535, PB_ERR_INVALID, "attention_tc: batch too large");
    CUtensorMap qmap, kvmap;
    PB_TRY(make_tmap_qkv(&qmap, qkv, D, L, B, AT_MQ));
    PB_TRY(make_tmap_qkv(&kvmap, qkv, D, L, B, AT_NK));
    AttnParams p{};
    p.key_valid = key_valid; p.out = out; p.B = B; p.L = L; p.H = H; p.D = D;
    p.scale_log2e = sqrtf(1.0f / (float)AT_DH) * 1.4426950408889634f;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div(L, AT_MQ), (unsigned)H, (unsigned)B);
    attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(qmap, kvmap, p);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
