// attn_tc.cu - fused multi-head attention of the FFT block on the 5th-gen tensor cores (sm_100a).
// Replaces, on the tensor-core path of the TTE, the nn.MultiheadAttention core of the reference
//   softmax(q.k^T / sqrt(d_head) + key_padding_mask) . v        modules/fft.py:53-59; torch F.multi_head_attention_forward
// (the projections around it are GEMMs of tc_conv.cu).  Scores are never written to HBM.
//
// One CTA = one (batch item, head, 128-query tile), 19 warps:
//   warp 0      producer: TMA loads of the Q tile and of 64-key K / V tiles (128-byte swizzle) from the split-fp16 qkv rows
//   warp 1      Q.K^T issuer + TMEM owner: copies Q into TENSOR MEMORY once (tcgen05.cp), then S = Q.K^T with the A operand from
//               tensor memory and K from shared memory
//   warp 18     P.V issuer:  O += P.V  (P from tensor memory, V from shared memory as an MN-major operand - no transpose of V)
//   warps 2..17 softmax / epilogue: thread = (query row = TMEM lane, a quarter of a tile's keys); scores and probabilities are
//               double-buffered in tensor memory so the tensor pipe and the exponentials overlap
// Products are evaluated on split-fp16 operands (x = hi + lo): a.b ~ a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, three MMAs per k-step
// accumulating in place (fp32, TMEM).
// Softmax is exact two-pass (pass 1: row maximum; pass 2: exp(s - max) -> P.V and the denominator, applied to O at the end):
// no rescaling of the output accumulator is ever needed and every exponential is evaluated once.
// Measured limits on B200 (PB_AT_ABLATE): barrier / TMEM-load skeleton 0.27 ms, + exponentials 0.34 ms alone, + MMAs and K / V
// loads 0.28 ms alone per decoder-layer launch (B = 32, L = 2048); they overlap to 0.49 ms.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pb_tte.cuh"
#include "pb_umma.cuh"

namespace pb {

namespace {

constexpr int AT_PARTS = 4;      // softmax warps per TMEM lane quarter: each takes 64 / AT_PARTS keys of a tile
constexpr int AT_KP = 64 / AT_PARTS;
constexpr int AT_SW = 4 * AT_PARTS;  // softmax warps
constexpr int AT_THREADS = 96 + 32 * AT_SW;  // producer warp, Q.K^T issuer, softmax warps, P.V issuer
constexpr int AT_MQ = 128;       // queries per CTA
constexpr int AT_NK = 64;        // keys per tile
constexpr int AT_DH = 128;       // head width (d_model 256 / 2 heads)
constexpr int AT_KST = 4;        // K ring: 3 stages of their own + the first half of the Q tile's shared memory, free once Q is in TMEM
constexpr int AT_VST = 3;        // V ring: 2 stages + the second half of the Q tile's shared memory
// (the kernel is bound by the latency of the K / V tile loads from L2: bytes in flight per SM are what buys throughput)
constexpr uint32_t AT_Q_BYTES = 4 * AT_MQ * 128;          // [hi c0][hi c1][lo c0][lo c1], each 128 rows x 128 bytes
constexpr uint32_t AT_K_BYTES = 4 * AT_NK * 128;          // per channel half: [hi rows][lo rows]
constexpr uint32_t AT_V_BYTES = 4 * AT_NK * 128;          // [Vhi d0][Vhi d1][Vlo d0][Vlo d1], each 64 keys x 128 bytes
constexpr uint32_t AT_SMEM = 2048 + AT_Q_BYTES + (AT_KST - 1) * AT_K_BYTES + (AT_VST - 1) * AT_V_BYTES + 1024;   // = 232448, the 227 KB per-CTA limit
// tensor memory columns (512).  Q lives in tensor memory for the whole CTA (copied once with tcgen05.cp), so the Q.K^T MMAs
// read only K from shared memory (the A-operand reads were 60 % of that kernel's shared-memory traffic).
constexpr uint32_t TM_S = 0;      // score buffers of 64 columns (one per key of the tile; the three products of the split
                                  // accumulate in place - with Q in tensor memory a 64-wide MMA costs N/2 = 32 cycles, so the
                                  // stacked [k_hi; k_lo] form would only double the accumulator read-back): 2 in pass 2, 4 in pass 1
constexpr uint32_t TM_P = 128;    // 2 x 64:  [0,32) P_hi (64 keys, 2 per column), [32,64) P_lo
constexpr uint32_t TM_O = 256;    // 128
constexpr uint32_t TM_Q = 384;    // 128: [0,64) Q_hi (8 k-steps x 8 columns), [64,128) Q_lo
constexpr uint32_t AT_S1 = 4;     // pass-1 score buffers (columns [0,256): the P region is still unused)
constexpr uint32_t AT_S2 = 2;     // pass-2 score buffers (columns [0,128))

struct AttnParams {
    const uint8_t* key_valid;     // [B][L] 1 = attend, or null
    uint16_t* out;                // split-fp16 rows [B*L][2*D]: head h writes hi at h*128, lo at D + h*128
    int B, L, H, D;
    float scale_log2e;            // (1 / sqrt(d_head)) * log2(e)
    int ablate;                   // dev only (PB_AT_ABLATE): 1 = skip the exponentials of pass 2, 2 = skip the P stores, 4 = skip the pass-1 exponentials, 8 = skip the MMAs, 16 = skip the K / V loads
};

// 2^x on the SFU (ex2.approx.ftz: 2 ulp, flushes results below 2^-126 to zero - the probabilities it would lose are < 1e-38)
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

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
    uint64_t* k_full = bars + 1;                   // [AT_KST]
    uint64_t* k_empty = k_full + AT_KST;
    uint64_t* v_full = k_empty + AT_KST;           // [AT_VST]
    uint64_t* v_empty = v_full + AT_VST;
    uint64_t* s_full = v_empty + AT_VST;           // [2] score buffers
    uint64_t* s_free = s_full + 2;
    uint64_t* p_ready = s_free + 2;                // [2] probability buffers
    uint64_t* p_free = p_ready + 2;
    uint64_t* o_full = p_free + 2;
    uint64_t* s1_full = o_full + 1;                // [4] pass 1: four score buffers (the P / O columns are still unused)
    uint64_t* s1_free = s1_full + 4;
    uint64_t* q_done = s1_free + 4;                // Q has been copied to tensor memory: its shared memory may be reused
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_done + 1);
    uint32_t* kmask = reinterpret_cast<uint32_t*>(sm + 512);   // [n_tiles][2] validity bits of each tile's keys (L <= 4096)
    float2* stat_x = reinterpret_cast<float2*>(sm + 1024);     // [128] (m, l) exchange buffer for the merge after pass 1
    uint8_t* q_s = sm + 2048;
    uint8_t* k_s = q_s + AT_Q_BYTES;
    uint8_t* v_s = k_s + (AT_KST - 1) * AT_K_BYTES;
    // stage -> shared memory: the last K stage and the last V stage alias the two halves of the Q tile
    auto k_stage = [&](uint32_t st) { return st + 1 < AT_KST ? k_s + st * AT_K_BYTES : q_s; };
    auto v_stage = [&](uint32_t st) { return st + 1 < AT_VST ? v_s + st * AT_V_BYTES : q_s + AT_K_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * AT_MQ;
    const int n_tiles = (p.L + AT_NK - 1) / AT_NK;
    const int C3 = 3 * p.D;                 // channels per plane of a qkv row: [q | k | v]

    if (threadIdx.x == 0) {
        prefetch_tmap(&qmap);
        prefetch_tmap(&kvmap);
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_KST; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < AT_VST; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_free[s], AT_SW); mbar_init(&p_ready[s], AT_SW); mbar_init(&p_free[s], 1); }
        mbar_init(o_full, 1);
        mbar_init(q_done, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(&s1_full[s], 1); mbar_init(&s1_free[s], AT_SW); }
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
        uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int t = 0; t < n_tiles; ++t) {
                const int k0 = t * AT_NK;
                mbar_wait(&k_empty[ks], kph ^ 1);
                if (ks + 1 == AT_KST && kph == 0) mbar_wait(q_done, 0);      // first use of the stage that aliases the Q tile
                if (elect_one()) {
                    uint8_t* dst = k_stage(ks);
                    if (p.ablate & 16) mbar_arrive(&k_full[ks]); else {
                    mbar_expect_tx(&k_full[ks], AT_K_BYTES);
                    for (int c = 0; c < 2; ++c)
                        for (int pl = 0; pl < 2; ++pl)   // per channel half: hi rows then lo rows (one stacked 128-row operand)
                            tma_load_3d(dst + (c * 2 + pl) * (AT_NK * 128), &kvmap, pl * C3 + p.D + h * AT_DH + c * 64, k0, b, &k_full[ks]);
                    }
                }
                __syncwarp();
                if (++ks == AT_KST) { ks = 0; kph ^= 1; }
                if (pass == 1) {
                    mbar_wait(&v_empty[vs], vph ^ 1);
                    if (vs + 1 == AT_VST && vph == 0) mbar_wait(q_done, 0);
                    if (elect_one()) {
                        uint8_t* dst = v_stage(vs);
                        if (p.ablate & 16) mbar_arrive(&v_full[vs]); else {
                        mbar_expect_tx(&v_full[vs], AT_V_BYTES);
                        for (int pl = 0; pl < 2; ++pl)
                            for (int c = 0; c < 2; ++c)  // [Vhi d0][Vhi d1][Vlo d0][Vlo d1]: 64-wide MN groups
                                tma_load_3d(dst + (pl * 2 + c) * (AT_NK * 128), &kvmap, pl * C3 + 2 * p.D + h * AT_DH + c * 64, k0, b, &v_full[vs]);
                        }
                    }
                    __syncwarp();
                    if (++vs == AT_VST) { vs = 0; vph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        // instruction descriptors: D f32, A/B f16; S: N = 128 (stacked) / 64, K-major B; O: N = 128, MN-major B
        const uint32_t id_base = (1u << 4) | ((uint32_t)(AT_MQ >> 4) << 24);
        const uint32_t id_s1 = id_base | ((uint32_t)(AT_NK >> 3) << 17);
        const uint32_t q_base = smem_u32(q_s), k_base = smem_u32(k_s);
        // K-major, 128-byte swizzle: rows 128 bytes apart, 8-row groups 1024 bytes apart
        const uint64_t kdesc0 = smem_desc(0, 16, 1024, 2, 0);
        auto addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3fffu); };
        mbar_wait(q_full, 0);
        tc_fence_after();
        uint32_t ks = 0, kph = 0;
        // S of tile t of pass 1 -> one of FOUR score buffers (the P / O columns are unused until pass 2), of pass 2 -> one of two;
        // the buffer's previous use must have been read by the softmax warps
        // Q: shared memory -> tensor memory, one 128-row x 32-byte k-step per tcgen05.cp; copies and MMAs execute in issue order
        if (elect_one()) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t sd = addr(kdesc0, q_base + (pl * 2 + c) * (AT_MQ * 128) + k4 * 32);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem + TM_Q + pl * 64 + (c * 4 + k4) * 8), "l"(sd) : "memory");
                    }
            umma_commit(q_done);
        }
        __syncwarp();
        auto qk_tile = [&](int pass, int t) {
            const uint32_t nb = pass == 0 ? AT_S1 : AT_S2;
            const uint32_t sb = (uint32_t)t % nb, n = (uint32_t)t / nb;
            uint64_t* full = pass == 0 ? &s1_full[sb] : &s_full[sb];
            uint64_t* free_ = pass == 0 ? &s1_free[sb] : &s_free[sb];
            mbar_wait(&k_full[ks], kph);
            if (n > 0) mbar_wait(free_, (n - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t kb = ks + 1 < AT_KST ? k_base + ks * AT_K_BYTES : q_base;
                const uint32_t d = tmem + TM_S + sb * 64u;
                uint32_t acc = 0;
                if (!(p.ablate & 8))
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint32_t qa = tmem + TM_Q + (uint32_t)(c * 4 + k4) * 8u;
                        const uint64_t b_hi = addr(kdesc0, kb + c * (2 * AT_NK * 128) + k4 * 32);
                        const uint64_t b_lo = addr(kdesc0, kb + c * (2 * AT_NK * 128) + AT_NK * 128 + k4 * 32);
                        umma_f16_ts(d, qa, b_hi, id_s1, acc);        // q_hi . k_hi
                        umma_f16_ts(d, qa + 64u, b_hi, id_s1, 1);    // q_lo . k_hi
                        umma_f16_ts(d, qa, b_lo, id_s1, 1);          // q_hi . k_lo
                        acc = 1;
                    }
                umma_commit(full);
                umma_commit(&k_empty[ks]);
            }
            __syncwarp();
            if (++ks == AT_KST) { ks = 0; kph ^= 1; }
        };
        // ---- pass 1: scores only (row maximum / denominator), up to four tiles ahead of the softmax warps
        for (int t = 0; t < n_tiles; ++t) qk_tile(0, t);
        // every pass-1 buffer must have been read before pass 2 reuses the columns
        for (int j = 0; j < (int)AT_S1 && j < n_tiles; ++j) {
            const uint32_t uses = ((uint32_t)(n_tiles - j) + AT_S1 - 1u) / AT_S1;
            mbar_wait(&s1_free[j], (uses - 1) & 1u);
        }
        // ---- pass 2: S(t) runs up to two tiles ahead; P(t).V(t) is issued by its own warp as soon as P(t) is in tensor memory
        for (int t = 0; t < n_tiles; ++t) qk_tile(1, t);
    } else if (warp == 2 + AT_SW) {
        // =========================== P.V issuer ===========================
        const uint32_t id_o = (1u << 4) | ((uint32_t)(AT_MQ >> 4) << 24) | (1u << 16) | ((uint32_t)(AT_DH >> 3) << 17);
        const uint32_t v_base = smem_u32(v_s);
        // MN-major, 128-byte swizzle: 64-element MN groups LBO = one 64-key box apart, 8-key groups 1024 bytes apart
        const uint64_t vdesc0 = smem_desc(0, AT_NK * 128, 1024, 2, 0);
        auto addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3fffu); };
        uint32_t vs = 0, vph = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t pb = (uint32_t)t & 1u, n = (uint32_t)t >> 1;
            mbar_wait(&v_full[vs], vph);
            mbar_wait(&p_ready[pb], n & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t vb = vs + 1 < AT_VST ? v_base + vs * AT_V_BYTES : smem_u32(q_s) + AT_K_BYTES;
                const uint32_t pa = tmem + TM_P + pb * 64u;
                if (!(p.ablate & 8))
#pragma unroll
                for (int k4 = 0; k4 < AT_NK / 16; ++k4) {
                    const uint64_t b_hi = addr(vdesc0, vb + k4 * (16 * 128));
                    const uint64_t b_lo = addr(vdesc0, vb + 2 * (AT_NK * 128) + k4 * (16 * 128));
                    umma_f16_ts(tmem + TM_O, pa + k4 * 8, b_hi, id_o, (t > 0 || k4 > 0) ? 1u : 0u);   // p_hi . v_hi
                    umma_f16_ts(tmem + TM_O, pa + 32 + k4 * 8, b_hi, id_o, 1);                        // p_lo . v_hi
                    umma_f16_ts(tmem + TM_O, pa + k4 * 8, b_lo, id_o, 1);                             // p_hi . v_lo
                }
                umma_commit(&v_empty[vs]);
                umma_commit(&p_free[pb]);
                if (t + 1 == n_tiles) umma_commit(o_full);
            }
            __syncwarp();
            if (++vs == AT_VST) { vs = 0; vph ^= 1; }
        }
    } else if (warp >= 2 && warp < 2 + AT_SW) {
        // =========================== softmax / epilogue ===========================
        // thread = (query row, key part): AT_PARTS warps per TMEM lane quarter, each taking AT_KP of a tile's 64 keys
        const int q = warp & 3;                              // TMEM lane quarter of this warp
        const int part = (warp - 2) >> 2;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = q * 32 + lane;
        const int qrow = q0 + row;
        const uint8_t* kval = p.key_valid ? p.key_valid + (size_t)b * p.L : nullptr;
        const float c2 = p.scale_log2e;
        float m = -INFINITY, l = 0.f;
        // validity bits of every tile's keys, computed once (a global load per tile would sit in the critical path of each tile)
        for (int t = warp - 2; t < n_tiles; t += AT_SW) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k = t * AT_NK + hh * 32 + lane;
                const uint32_t bits = __ballot_sync(0xffffffffu, k < p.L && (!kval || kval[k] != 0));
                if (lane == 0) kmask[2 * t + hh] = bits;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SW) : "memory");
        auto tile_mask = [&](int t) {     // this thread's AT_KP keys of tile t
            const uint32_t w = kmask[2 * t + (part * AT_KP) / 32];
            return (w >> ((part * AT_KP) & 31)) & (AT_KP == 32 ? 0xffffffffu : ((1u << AT_KP) - 1u));
        };
        auto load_scores = [&](int pass, int t, float* s) {
            const uint32_t nb = pass == 0 ? AT_S1 : AT_S2;
            const uint32_t sb = (uint32_t)t % nb, n = (uint32_t)t / nb;
            mbar_wait(pass == 0 ? &s1_full[sb] : &s_full[sb], n & 1u);
            tc_fence_after();
            uint32_t a[AT_KP];
            const uint32_t base = tmem + lane_off + TM_S + sb * 64u + (uint32_t)(part * AT_KP);
            if (AT_KP == 32) tmem_ld32_nowait(base, a); else tmem_ld16_nowait(base, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < AT_KP; ++i) s[i] = __uint_as_float(a[i]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pass == 0 ? &s1_free[sb] : &s_free[sb]);   // the buffer may take another tile's scores
        };
        // ---- pass 1: row maximum over this thread's keys (the denominator is accumulated in pass 2 from the very exponentials
        //      that are multiplied with V, and applied to O at the end: normalise-after instead of the reference's
        //      normalise-before - the same value up to fp32 rounding, and half the exponentials)
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t mk = tile_mask(t);
            float s[AT_KP];
            load_scores(0, t, s);
            constexpr uint32_t FULL = AT_KP == 32 ? 0xffffffffu : ((1u << AT_KP) - 1u);
            if (mk == FULL) {      // every key of the tile is valid (all tiles but the last of an unpadded item): no per-key tests
#pragma unroll
                for (int i = 0; i < AT_KP; ++i) m = fmaxf(m, s[i]);
            } else {
#pragma unroll
                for (int i = 0; i < AT_KP; ++i)
                    if ((mk >> i) & 1u) m = fmaxf(m, s[i]);
            }
        }
        // merge the key parts of every row through one 1 KB exchange buffer: parts 1.. hand their value to part 0 in turn
#define AT_BAR() asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SW) : "memory")
        auto merge_parts = [&](float v, bool is_max) {
            float* xbuf = reinterpret_cast<float*>(stat_x);
            for (int j = 1; j < AT_PARTS; ++j) {
                if (part == j) xbuf[row] = v;
                AT_BAR();
                if (part == 0) v = is_max ? fmaxf(v, xbuf[row]) : v + xbuf[row];
                AT_BAR();
            }
            if (part == 0) xbuf[row] = v;
            AT_BAR();
            if (part != 0) v = xbuf[row];
            AT_BAR();
            return v;
        };
        m = merge_parts(m, true);
        const float moff = m * c2;
        // ---- pass 2: normalised probabilities -> tensor memory (fp16 hi + lo)
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t mk = tile_mask(t);
            constexpr uint32_t FULL2 = AT_KP == 32 ? 0xffffffffu : ((1u << AT_KP) - 1u);
            float s[AT_KP];
            load_scores(1, t, s);
#pragma unroll
            for (int i = 0; i < AT_KP; ++i) {
                const float e = ex2f(fmaf(s[i], c2, -moff));
                s[i] = (p.ablate & 1) ? 0.01f : ((mk == FULL2 || ((mk >> i) & 1u)) ? e : 0.f);
                l += s[i];
            }
            uint32_t hi[AT_KP / 2], lo[AT_KP / 2];
#pragma unroll
            for (int i = 0; i < AT_KP / 2; ++i) {
                const float a = s[2 * i], c = s[2 * i + 1];
                const __half2 hh = __floats2half2_rn(a, c);
                const float2 hf = __half22float2(hh);
                hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[i] = pack_h2(a - hf.x, c - hf.y);
            }
            const uint32_t pb = (uint32_t)t & 1u, n = (uint32_t)t >> 1;
            if (n > 0) mbar_wait(&p_free[pb], (n - 1) & 1u);   // P(t-2) consumed by its MMAs
            tc_fence_after();
            const uint32_t pbase = tmem + lane_off + TM_P + pb * 64u + (uint32_t)(part * (AT_KP / 2));
            if (!(p.ablate & 2)) {
                if (AT_KP == 32) { tmem_st16(pbase, hi); tmem_st16(pbase + 32, lo); }
                else { tmem_st8(pbase, hi); tmem_st8(pbase + 32, lo); }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[pb]);
        }
        // ---- epilogue: O / l -> split-fp16 row [hi(128) | lo(128)] of head h; the parts share the 128 columns
        l = merge_parts(l, false);
#undef AT_BAR
        const float inv_l = 1.0f / l;
        mbar_wait(o_full, 0);
        tc_fence_after();
        uint16_t* orow = p.out + ((size_t)b * p.L + qrow) * 2 * p.D + h * AT_DH;
#pragma unroll 1
        for (int g = part * (4 / AT_PARTS); g < (part + 1) * (4 / AT_PARTS); ++g) {
            uint32_t a[32];
            tmem_ld32_nowait(tmem + lane_off + TM_O + g * 32, a);
            tmem_ld_wait();
            if (qrow < p.L) {
                uint32_t hw[16], lw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x0 = __uint_as_float(a[2 * i]) * inv_l, x1 = __uint_as_float(a[2 * i + 1]) * inv_l;
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
    PB_REQUIRE(L <= 4096, PB_ERR_INVALID, "attention_tc: L=%d exceeds the 4096 keys the mask table holds", L);
    CUtensorMap qmap, kvmap;
    PB_TRY(make_tmap_qkv(&qmap, qkv, D, L, B, AT_MQ));
    PB_TRY(make_tmap_qkv(&kvmap, qkv, D, L, B, AT_NK));
    AttnParams p{};
    p.key_valid = key_valid; p.out = out; p.B = B; p.L = L; p.H = H; p.D = D;
    p.scale_log2e = sqrtf(1.0f / (float)AT_DH) * 1.4426950408889634f;
    { static const int abl = getenv("PB_AT_ABLATE") ? atoi(getenv("PB_AT_ABLATE")) : 0; p.ablate = abl; }
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
