// tc_pair.cu - one fused launch for a ResBlock1 step of the narrow HiFi-GAN stages (C = 16 / 32) on sm_100a:
//     xt  = leaky_relu(conv1(leaky_relu(x), dilation d))            reference utils/vocoder/models.py:31-38 (c1)
//     out = x + conv2(xt)  (+ running sum, x 1/3, next activation)  (c2, and models.py:100-106 for the last step of a block)
// The intermediate xt never leaves the SM: conv1's accumulator is converted to split-bf16 rows in SHARED MEMORY (the swizzled
// K-major layout TMA would have produced) and is the A operand of conv2's MMAs.  HBM traffic per pair drops from five tensor
// passes (read x, write xt, read xt, read x, write out) to ~2.2 (x with halo, out; the residual re-read hits L2).
//
// Two kernels.  tc_pair_kernel (C = 32, and C = 16 at k = 3) takes the taps one by one:
// Tile: conv1 produces 128 rows of xt, conv2 the 128 - (K-1) output rows they support (tiles advance by that many rows; the
// K-1 garbage rows an M = 128 MMA computes past them are never stored).  Both weight sets stay resident in shared memory.
//   warp 0       producer: TMA loads of the x tile (128 + (K-1) d rows, zero fill = conv padding)
//   warps 1, 3   MMA issuers: warp 1 issues conv1 of every tile (running ahead), warp 3 conv2, so the tensor pipe works on the
//                next tiles while the epilogue warps turn conv1(i)'s accumulator into the shared-memory operand
//   warp 2       TMEM owner (NG conv1 + NG conv2 accumulator stages)
//   warps 4..    NG = 3 or 4 epilogue groups taking the tiles round-robin: epilogue A (TMEM -> bias, leaky relu, edge zeros -> swizzled
//                smem), then epilogue B (TMEM -> bias, residual - at C = 32 from the x tile still in shared memory -, running sum, scale,
//                activation, length mask -> global, at C = 32 through per-warp staging tiles + TMA stores)
// tc_pair16_kernel (C = 16, k >= 7) evaluates conv2 time-grouped against a block-Toeplitz operand: see its own header below.
// Both produce the same bits (tests/test_switches_gpu.py); launch_tc_pair chooses per layer.
#include <stdlib.h>

#include "pb_tc.cuh"
#include "pb_umma.cuh"

namespace pb {

namespace {

constexpr int TP_M = 128;
constexpr int TP_MAX_GROUPS = 4;  // NG epilogue warpgroups take the tiles round-robin, each with its own conv1 accumulator / xt tile /
                                  // conv2 accumulator stage: NG tiles in flight (template parameter: 3 or 4, chosen per layer on the host)
constexpr int TP_MAX_XST = 6;
constexpr uint32_t TP_SMEM_LIMIT = 232448;

struct PairParams {
    const uint8_t *w1, *w2;
    const float *bias1, *bias2;
    const __nv_bfloat16* res;      // x itself (stored leaky-relu'd): residual of conv2
    const __nv_bfloat16* accin;
    __nv_bfloat16* out;
    const int32_t* lens;
    int len_mul;
    int B, L, C, K, D;
    int p1, p2, m_out;             // conv1 / conv2 half widths, output rows per tile
    int ksteps, row_bytes, lo_off, a_layout, a_sbo;
    int x_rows, x_stage_bytes, n_x_stages, mid_bytes, w_bytes;
    uint32_t smem_bias_off, smem_w1_off, smem_w2_off, smem_x_off, smem_mid_off;
    int tiles_t, total_tiles;
    float mid_slope, post_slope, scale, res_unslope;
    uint32_t idesc, idesc_stack;
    int acc_cols, tmem_cols;
    int tma_out;                   // epilogue B writes through per-warp staging tiles + TMA stores
    uint32_t smem_stage_off;
    int res_smem;                  // epilogue B reads the residual rows from the x tile in shared memory (held until then)
    int ablate;                    // profiling only (PB_TP_ABLATE): 1 no MMAs, 2 no output stores, 4 no residual / running-sum loads,
                                   // 8 no x loads, 16 no xt stores, 32 no epilogue arithmetic, 64 no accumulator reads - results are wrong by design
};

}  // namespace

template <bool ACC, int KSTEPS, int NG>
__global__ void __launch_bounds__(128 + 128 * NG, 1)
tc_pair_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap_tail,
               const PairParams p) {
    constexpr int TP_GROUPS = NG, TP_STAGES = NG;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint64_t* full_x = bars;                         // [TP_MAX_XST]
    uint64_t* empty_x = bars + TP_MAX_XST;
    uint64_t* a_full = bars + 2 * TP_MAX_XST;        // [TP_STAGES] conv1 accumulator ready
    uint64_t* a_empty = a_full + TP_STAGES;
    uint64_t* mid_full = a_empty + TP_STAGES;        // [TP_STAGES] xt tile in shared memory
    uint64_t* mid_empty = mid_full + TP_STAGES;
    uint64_t* b_full = mid_empty + TP_STAGES;        // [TP_STAGES] conv2 accumulator ready
    uint64_t* b_empty = b_full + TP_STAGES;
    uint64_t* w_full = b_empty + TP_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sm_base = smem_u32(sm);

    if (threadIdx.x == 0) {
        prefetch_tmap(&xmap);
        if (p.tma_out) { prefetch_tmap(&omap); prefetch_tmap(&omap_tail); }
        for (int s = 0; s < TP_MAX_XST; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], p.res_smem ? 5 : 1); }
        for (int s = 0; s < TP_STAGES; ++s) {
            mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 4);
            mbar_init(&mid_full[s], 4); mbar_init(&mid_empty[s], 1);
            mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 4);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    {   // biases -> shared memory; the K-1 rows past each xt tile are read (never used) by conv2: keep them finite
        float* bw = reinterpret_cast<float*>(sm + p.smem_bias_off);
        for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) bw[i] = i < p.C ? p.bias1[i] : p.bias2[i - p.C];
        uint4* mz = reinterpret_cast<uint4*>(sm + p.smem_mid_off);
        for (int i = threadIdx.x; i < TP_STAGES * p.mid_bytes / 16; i += blockDim.x) mz[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_b = p.tiles_t;

    if (warp == 0) {
        // =========================== producer ===========================
        if (elect_one()) {
            mbar_expect_tx(w_full, (uint32_t)(2 * p.w_bytes));
            for (int off = 0; off < p.w_bytes; off += 32768) {
                const int n = p.w_bytes - off < 32768 ? p.w_bytes - off : 32768;
                bulk_load_1d(sm + p.smem_w1_off + off, p.w1 + off, (uint32_t)n, w_full);
                bulk_load_1d(sm + p.smem_w2_off + off, p.w2 + off, (uint32_t)n, w_full);
            }
        }
        __syncwarp();
        asm volatile("griddepcontrol.wait;" ::: "memory");
        uint32_t xs = 0, xph = 0;
        int b = blockIdx.x / tiles_per_b, rem = blockIdx.x - b * tiles_per_b;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int row0 = rem * p.m_out - p.p2 - p.p1;
            mbar_wait(&empty_x[xs], xph ^ 1);
            if (elect_one()) {
                uint8_t* dst = sm + p.smem_x_off + xs * p.x_stage_bytes;
                // the box is at most 256 rows: the tile is loaded as two boxes of x_rows / 2 rows
                if (p.ablate & 8) mbar_arrive(&full_x[xs]);
                else {
                    mbar_expect_tx(&full_x[xs], (uint32_t)(p.x_rows * p.row_bytes));
                    tma_load_3d(dst, &xmap, 0, row0, b, &full_x[xs]);
                    tma_load_3d(dst + (p.x_rows / 2) * p.row_bytes, &xmap, 0, row0 + p.x_rows / 2, b, &full_x[xs]);
                }
            }
            __syncwarp();
            if (++xs == (uint32_t)p.n_x_stages) { xs = 0; xph ^= 1; }
            rem += (int)gridDim.x;
            while (rem >= tiles_per_b) { rem -= tiles_per_b; ++b; }
        }
    } else if (warp == 1 || warp == 3) {
        // =========================== MMA issuers: warp 1 issues every conv1, warp 3 every conv2 ===========================
        // (one issuing thread sustains ~70 cycles per MMA here; the narrow MMAs need ~40, so two issuers keep the pipe fed)
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint32_t a_desc_hi = (uint32_t)(smem_desc(0, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0) >> 32);
        const uint32_t a_desc_lo = (uint32_t)smem_desc(0, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0);
        const uint32_t b_desc_hi = (uint32_t)(smem_desc(0, (uint32_t)p.C * 32u, 128, 0, 0) >> 32);
        const uint32_t b_desc_lo = (uint32_t)smem_desc(0, (uint32_t)p.C * 32u, 128, 0, 0);
        const uint32_t lo_step = (uint32_t)p.lo_off >> 4;
        const uint32_t kstep_step = (uint32_t)p.C * 4u;                 // 2 planes x 2 K halves x C x 16 bytes
        const uint32_t block_step = kstep_step * (uint32_t)KSTEPS;
        const uint32_t tap1 = (uint32_t)(p.D * p.row_bytes) >> 4, tap2 = (uint32_t)p.row_bytes >> 4;
        const uint32_t w1_lo = b_desc_lo | (((sm_base + p.smem_w1_off) >> 4) & 0x3fffu);
        const uint32_t w2_lo = b_desc_lo | (((sm_base + p.smem_w2_off) >> 4) & 0x3fffu);
        const int K = p.K;
        const uint32_t idesc = p.idesc, idesc_stack = p.idesc_stack, acc_cols = (uint32_t)p.acc_cols;
        // one convolution of one tile: K taps x ksteps x {a_hi.[w_hi; w_lo], a_lo.w_hi}
        auto conv = [&](uint32_t d_tmem, uint32_t a_lo0, uint32_t tap_step, uint32_t w_lo) {
            uint32_t acc = 0;
            uint32_t a_tap = a_lo0, b_tap = w_lo;
            if (p.ablate & 1) return;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint64_t da_hi = ((uint64_t)a_desc_hi << 32) | (a_tap + 2u * ks);
                    const uint64_t da_lo = ((uint64_t)a_desc_hi << 32) | (a_tap + 2u * ks + lo_step);
                    const uint64_t db = ((uint64_t)b_desc_hi << 32) | (b_tap + kstep_step * ks);
                    umma_bf16(d_tmem, da_hi, db, idesc_stack, acc);
                    umma_bf16(d_tmem, da_lo, db, idesc, 1);
                    acc = 1;
                }
                a_tap += tap_step;
                b_tap += block_step;
            }
        };
        uint32_t xs = 0, xph = 0;
        int n_my = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) ++n_my;
        auto conv1 = [&](int i) {
            const uint32_t st = (uint32_t)i % TP_STAGES, n = (uint32_t)i / TP_STAGES;
            mbar_wait(&full_x[xs], xph);
            if (n > 0) mbar_wait(&a_empty[st], (n - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_base = sm_base + p.smem_x_off + xs * (uint32_t)p.x_stage_bytes;
                conv(tmem_base + st * acc_cols, a_desc_lo | ((a_base >> 4) & 0x3fffu), tap1, w1_lo);
                umma_commit(&empty_x[xs]);
                umma_commit(&a_full[st]);
            }
            __syncwarp();
            if (++xs == (uint32_t)p.n_x_stages) { xs = 0; xph ^= 1; }
        };
        auto conv2 = [&](int i) {
            const uint32_t st = (uint32_t)i % TP_STAGES, n = (uint32_t)i / TP_STAGES;
            mbar_wait(&mid_full[st], n & 1u);
            if (n > 0) mbar_wait(&b_empty[st], (n - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t m_base = sm_base + p.smem_mid_off + st * (uint32_t)p.mid_bytes;
                conv(tmem_base + (TP_STAGES + st) * acc_cols, a_desc_lo | ((m_base >> 4) & 0x3fffu), tap2, w2_lo);
                umma_commit(&mid_empty[st]);
                umma_commit(&b_full[st]);
            }
            __syncwarp();
        };
        if (warp == 1) {
            for (int i = 0; i < n_my; ++i) conv1(i);      // runs ahead of conv2 by up to TP_STAGES tiles (a_empty / full_x gate it)
        } else {
            for (int i = 0; i < n_my; ++i) conv2(i);
        }
    } else if (warp >= 4) {
        // =========================== epilogue groups ===========================
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int C = p.C;
        const size_t pitch = (size_t)2 * C;
        const float* bias1 = reinterpret_cast<const float*>(sm + p.smem_bias_off);
        const float* bias2 = bias1 + C;
        const int stride = (int)gridDim.x * TP_GROUPS;
        int tile = blockIdx.x + grp * (int)gridDim.x;
        int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
        auto valid_of = [&](int bb) {
            int v = p.L;
            if (p.lens) { const int u = __ldg(p.lens + bb) * p.len_mul; v = u < v ? u : v; }
            return v;
        };
        int valid = tile < p.total_tiles ? valid_of(b) : 0;
        const bool sw128 = p.a_layout == 2;
        for (int i = grp; tile < p.total_tiles; i += TP_GROUPS, tile += stride) {
            const uint32_t st = (uint32_t)i % TP_STAGES, n = (uint32_t)i / TP_STAGES;
            const int o0 = rem * p.m_out;
            // ---------------- epilogue A: conv1 accumulator -> xt tile in shared memory ----------------
            {
                const int tm = o0 - p.p2 + row;                    // sequence position of this xt row
                const bool in_seq = tm >= 0 && tm < valid;         // conv2 sees zeros outside the (valid) sequence
                const bool edge = __any_sync(0xffffffffu, !in_seq);   // (warp-uniform: the common tile skips the zeroing altogether)
                if (ACC) {      // running-sum row this thread adds in epilogue B: pull it into L2 a conv ahead (row <= 128 B)
                    const int to = o0 + row;
                    if (row < p.m_out && to < p.L)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.accin + ((size_t)b * p.L + to) * pitch));
                }
                mbar_wait(&a_full[st], n & 1u);
                if (n > 0) mbar_wait(&mid_empty[st], (n - 1) & 1u);   // conv2(i-2) has consumed this xt buffer
                tc_fence_after();
                uint8_t* mrow = sm + p.smem_mid_off + st * p.mid_bytes + row * p.row_bytes;
                const uint32_t swz = sw128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                for (int n0 = 0; n0 < C; n0 += 16) {
                    float v[16];
                    const uint32_t taddr = tmem_base + lane_off + st * (uint32_t)p.acc_cols + (uint32_t)n0;
                    if (p.ablate & 64) { for (int z = 0; z < 16; ++z) v[z] = 1.0f; } else tmem_ld16x2_sum(taddr, taddr + (uint32_t)C, v);
                    const float4* bp = reinterpret_cast<const float4*>(bias1 + n0);
                    uint32_t hw[8], lw[8];
                    if (p.ablate & 32) {
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) { hw[k2] = __float_as_uint(v[2 * k2]); lw[k2] = __float_as_uint(v[2 * k2 + 1]); }
                    } else {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 bv = bp[k4];
                        add2(v[4 * k4], v[4 * k4 + 1], bv.x, bv.y);
                        add2(v[4 * k4 + 2], v[4 * k4 + 3], bv.z, bv.w);
                    }
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) {
                        float a0 = v[2 * k2], a1 = v[2 * k2 + 1];
                        float t0 = a0, t1 = a1;
                        mul2(t0, t1, p.mid_slope, p.mid_slope);
                        a0 = fmaxf(a0, t0);
                        a1 = fmaxf(a1, t1);
                        split2(a0, a1, 0, hw[k2], lw[k2]);
                    }
                    if (edge && !in_seq) {                         // rows outside the sequence: only in a sequence's first / last tiles
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) hw[k2] = lw[k2] = 0u;
                    }
                    }
                    if (p.ablate & 16) continue;
                    // 16-byte chunks of the row: hi channels n0.. -> chunks n0/8, n0/8 + 1; lo -> lo_off/16 + the same
                    const uint32_t ch = (uint32_t)n0 >> 3, cl = ((uint32_t)p.lo_off >> 4) + ch;
                    *reinterpret_cast<uint4*>(mrow + ((ch ^ swz) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(mrow + (((ch + 1) ^ swz) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    *reinterpret_cast<uint4*>(mrow + ((cl ^ swz) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    *reinterpret_cast<uint4*>(mrow + (((cl + 1) ^ swz) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                }
                tc_fence_before();
                fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a_empty[st]); mbar_arrive(&mid_full[st]); }
            }
            // ---------------- epilogue B: conv2 accumulator -> output rows ----------------
            {
                const int to = o0 + row;
                const bool row_ok = row < p.m_out && to < p.L;
                const bool live = to < valid;
                const size_t row_el = ((size_t)b * p.L + to) * pitch;
                uint4 R[4], A[4];
                // The residual is x itself, and the x tile of this tile is still in shared memory (res_smem: its stage is released by
                // conv1's commit AND these four warps): a row-per-lane global load costs 32 L1 wavefronts, its 16-byte chunks here 4.
                const uint32_t xs_i = (uint32_t)i % (uint32_t)p.n_x_stages;
                const int xr = p.p1 + p.p2 + row;                                  // this output row inside the x tile
                const uint8_t* xrow = sm + p.smem_x_off + xs_i * p.x_stage_bytes + xr * p.row_bytes;
                const uint32_t xsw = sw128 ? (uint32_t)(xr & 7) : (uint32_t)((xr >> 1) & 3);
                // (the TMA writes of that tile are ordered before these reads through full_x -> conv1's issuer -> its commit on a_full -> this
                // warp's wait on a_full in epilogue A: no wait on full_x here)
                auto request = [&](int n0) {
                    if (!row_ok || (p.ablate & 4)) return;
                    if (p.res_smem) {
                        const uint32_t ch = (uint32_t)n0 >> 3, cl = ((uint32_t)p.lo_off >> 4) + ch;
                        R[0] = *reinterpret_cast<const uint4*>(xrow + ((ch ^ xsw) << 4));
                        R[1] = *reinterpret_cast<const uint4*>(xrow + (((ch + 1) ^ xsw) << 4));
                        R[2] = *reinterpret_cast<const uint4*>(xrow + ((cl ^ xsw) << 4));
                        R[3] = *reinterpret_cast<const uint4*>(xrow + (((cl + 1) ^ xsw) << 4));
                    } else {
                        ld_global_nc_v8(p.res + row_el + n0, R[0], R[1]);
                        ld_global_nc_v8(p.res + row_el + C + n0, R[2], R[3]);
                    }
                    if (ACC) {
                        ld_global_v8(p.accin + row_el + n0, A[0], A[1]);
                        ld_global_v8(p.accin + row_el + C + n0, A[2], A[3]);
                    }
                };
                request(0);
                mbar_wait(&b_full[st], n & 1u);
                tc_fence_after();
                for (int n0 = 0; n0 < C; n0 += 16) {
                    float v[16];
                    uint32_t hw[8], lw[8];
                    const uint32_t taddr = tmem_base + lane_off + (TP_STAGES + st) * (uint32_t)p.acc_cols + (uint32_t)n0;
                    if (p.ablate & 64) { for (int z = 0; z < 16; ++z) v[z] = 1.0f; } else tmem_ld16x2_sum(taddr, taddr + (uint32_t)C, v);
                    if (n0 + 16 >= C) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&b_empty[st]);
                    }
                    if (row_ok && (p.ablate & 32)) {
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) { hw[k2] = __float_as_uint(v[2 * k2]); lw[k2] = __float_as_uint(v[2 * k2 + 1]); }
                    }
                    if (row_ok && !(p.ablate & 32)) {
                        const float4* bp = reinterpret_cast<const float4*>(bias2 + n0);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const float4 bv = bp[k4];
                            add2(v[4 * k4], v[4 * k4 + 1], bv.x, bv.y);
                            add2(v[4 * k4 + 2], v[4 * k4 + 3], bv.z, bv.w);
                        }
                        const uint32_t hh[8] = {R[0].x, R[0].y, R[0].z, R[0].w, R[1].x, R[1].y, R[1].z, R[1].w};
                        const uint32_t ll[8] = {R[2].x, R[2].y, R[2].z, R[2].w, R[3].x, R[3].y, R[3].z, R[3].w};
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) {
                            const float2 r = join2(hh[k2], ll[k2], 0);
                            float tx = r.x, ty = r.y;
                            mul2(tx, ty, p.res_unslope, p.res_unslope);
                            add2(v[2 * k2], v[2 * k2 + 1], fminf(r.x, tx), fminf(r.y, ty));      // undo the stored leaky relu
                        }
                        if (ACC) {
                            const uint32_t ah[8] = {A[0].x, A[0].y, A[0].z, A[0].w, A[1].x, A[1].y, A[1].z, A[1].w};
                            const uint32_t al[8] = {A[2].x, A[2].y, A[2].z, A[2].w, A[3].x, A[3].y, A[3].z, A[3].w};
#pragma unroll
                            for (int k2 = 0; k2 < 8; ++k2) {
                                const float2 r = join2(ah[k2], al[k2], 0);
                                add2(v[2 * k2], v[2 * k2 + 1], r.x, r.y);
                            }
                        }
                        const float sc = live ? p.scale : 0.f;
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) {
                            float a0 = v[2 * k2], a1 = v[2 * k2 + 1];
                            mul2(a0, a1, sc, sc);
                            float t0 = a0, t1 = a1;
                            mul2(t0, t1, p.post_slope, p.post_slope);
                            a0 = fmaxf(a0, t0);
                            a1 = fmaxf(a1, t1);
                            split2(a0, a1, 0, hw[k2], lw[k2]);
                        }
                        if (n0 + 16 < C) request(n0 + 16);
                    }
                    if (row_ok) {
                        if (!p.tma_out && !(p.ablate & 2)) {
                            __nv_bfloat16* oh = p.out + row_el + n0;
                            __nv_bfloat16* ol = oh + C;
                            st_global_v8(oh, make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]));
                            st_global_v8(ol, make_uint4(lw[0], lw[1], lw[2], lw[3]), make_uint4(lw[4], lw[5], lw[6], lw[7]));
                        }
                    }
                    if (p.tma_out && !(p.ablate & 2)) {
                        // the warp parks its [32 rows x 16 ch] hi / lo tiles in shared memory (32B swizzle) and one lane hands them to
                        // the TMA unit: row-per-thread 32-byte global stores cost ~50 L1 wavefronts per instruction.  The last warp of a
                        // tile owns fewer than 32 output rows (m_out < 128): it uses the map whose box has exactly that many.
                        uint8_t* sbw = sm + p.smem_stage_off + (uint32_t)(warp - 4) * 2048u;
                        if (lane == 0) bulk_wait_group_read<0>();
                        __syncwarp();
                        const uint32_t sx = ((uint32_t)lane >> 2) & 1u;
                        uint8_t* rp = sbw + lane * 32;
                        *reinterpret_cast<uint4*>(rp + (sx << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(rp + ((sx ^ 1u) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                        *reinterpret_cast<uint4*>(rp + 1024 + (sx << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        *reinterpret_cast<uint4*>(rp + 1024 + ((sx ^ 1u) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                        fence_proxy_async();
                        __syncwarp();
                        const int t0 = o0 + q * 32;
                        if (lane == 0 && t0 < p.L) {
                            const CUtensorMap* om = q == 3 ? &omap_tail : &omap;
                            tma_store_3d(om, sbw, n0, t0, b);
                            tma_store_3d(om, sbw + 1024, C + n0, t0, b);
                            bulk_commit_group();
                        }
                    }
                }
                if (p.res_smem) {      // the residual rows have been consumed: this warp's share of the x stage's release
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_x[xs_i]);
                }
            }
            rem += stride;
            if (rem >= tiles_per_b) {
                do { rem -= tiles_per_b; ++b; } while (rem >= tiles_per_b);
                if (tile + stride < p.total_tiles) valid = valid_of(b);
            }
        }
    }

    if (p.tma_out && warp >= 4 && lane == 0) bulk_wait_group<0>();     // this lane's TMA stores have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// tc_pair16_kernel - the same ResBlock1 step for C = 16 with conv2 TIME-GROUPED: an M = 128 MMA reads its 4 KB A operand whatever N
// is, so a 16-channel conv2 tap (N = 32 stacked + 16) uses 1/8 of the tensor cycles it pays for.  Here an accumulator ROW is a group
// of G = 4 consecutive output positions and a column is (position in the group, channel): conv2 becomes G + K - 1 k-steps (one per
// input position a group can see) against a block-Toeplitz operand  T[s][ci][(g, co)] = W2[tap s - g][ci][co]  (zero outside
// 0 <= s - g < K), N = 64 (stacked [hi; lo]: 128), 512 positions per MMA tile instead of 128:
//     tensor cycles per output position at K = 11:  conv2 11 x 79 / 118 = 7.4  ->  14 x 112 / 500 = 3.1   (conv1 unchanged)
// The operand is expanded from the ordinary packed conv2 weights into shared memory at kernel start (16-byte vectors, no arithmetic);
// the products and their order per output are those of tc_pair_kernel (zero blocks add exact zeros): same bits.
//   super tile = 4 conv1 M tiles (512 xt positions, each with its own x tile and halo) -> one conv2 tile (4 R output positions,
//   R = (512 - (K-1)) / 4 rows; the rest of the 128 rows is never stored)
//   warp 0 producer (x tiles), warp 1 conv1 issuer, warp 2 TMEM owner, warp 3 conv2 issuer, warps 4..19 four epilogue groups:
//   epilogue A: group m turns conv1 M tile m into xt rows (hi / lo planes, 4 positions = one 128-byte swizzled row);
//   epilogue B: group g takes column chunk g (position g of every row's group): bias, residual, running sum, activation -> global.
//   Each group runs A(S + 1) BEFORE B(S): conv2(S) executes while the group is busy with the next tile's xt (two stages of everything).
namespace {
constexpr int T16_G = 4, T16_MT = 4, T16_C = 16;
constexpr int T16_XT_PLANE = 17408;              // 544 positions x 32 bytes (conv2's last rows read up to position 521)
constexpr int T16_THREADS = 128 + 128 * 4;

struct Pair16Params {
    const uint8_t *w1, *w2;
    const float *bias1, *bias2;
    const __nv_bfloat16* res;
    const __nv_bfloat16* accin;
    __nv_bfloat16* out;
    const int32_t* lens;
    int len_mul;
    int B, L, K, D;
    int p1, p2, rows_out, m_out, n_ksteps2;
    int x_rows, x_stage_bytes, n_x_stages, w1_bytes;
    uint32_t smem_bias_off, smem_w1_off, smem_w2_off, smem_xt_off, smem_rz_off, smem_x_off;
    int tiles_t, total_tiles;
    float mid_slope, post_slope, scale, res_unslope;
    int ablate;                    // PB_TP_ABLATE (profiling): 1 no MMAs, 2 no output stores, 4 no residual / running-sum loads, 64 no accumulator reads
};
}  // namespace

template <bool ACC, bool TMAE>
__global__ void __launch_bounds__(T16_THREADS, 1)
tc_pair16_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap rmap, const __grid_constant__ CUtensorMap amap,
                 const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap_tail, const Pair16Params p) {
    constexpr int C = T16_C;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint64_t* full_x = bars;                         // [TP_MAX_XST]
    uint64_t* empty_x = bars + TP_MAX_XST;
    uint64_t* a_full = bars + 2 * TP_MAX_XST;        // [2][4] conv1 accumulator of (stage, M tile) ready
    uint64_t* a_empty = a_full + 8;                  // [2][4] ... read by its epilogue group
    uint64_t* mid_full = a_empty + 8;                // [2] the stage's xt planes are complete
    uint64_t* b_full = mid_full + 2;                 // [2] conv2 accumulator ready
    uint64_t* w_full = b_full + 2;
    uint64_t* rz_full = w_full + 1;                  // (TMAE) the residual rows of a super tile are in shared memory
    uint64_t* rz_empty = rz_full + 1;                // (TMAE) ... and its output rows have left it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rz_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sm_base = smem_u32(sm);

    if (threadIdx.x == 0) {
        prefetch_tmap(&xmap);
        for (int s = 0; s < TP_MAX_XST; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], 1); }
        for (int s = 0; s < 8; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 4); }
        for (int s = 0; s < 2; ++s) { mbar_init(&mid_full[s], 16); mbar_init(&b_full[s], 1); }
        mbar_init(w_full, 1);
        mbar_init(rz_full, 1); mbar_init(rz_empty, 16);
        if (TMAE) { prefetch_tmap(&rmap); prefetch_tmap(&omap); prefetch_tmap(&omap_tail); if (ACC) prefetch_tmap(&amap); }
        fence_mbar_init();
    }
    {   // biases; zeroed xt planes (the positions past 511 are read, never used: keep them finite); the Toeplitz operand of conv2
        float* bw = reinterpret_cast<float*>(sm + p.smem_bias_off);
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) bw[i] = i < C ? p.bias1[i] : p.bias2[i - C];
        uint4* mz = reinterpret_cast<uint4*>(sm + p.smem_xt_off);
        for (int i = threadIdx.x; i < 4 * T16_XT_PLANE / 16; i += blockDim.x) mz[i] = make_uint4(0, 0, 0, 0);
        // k-step s, K half h, N row n (n < 64: w_hi of (g, co) = (n / 16, n % 16); n >= 64: w_lo): one 16-byte vector = 8 input channels.
        // Source: the packed conv2 weights, tap j = [K half h][row: co (hi) | 16 + co (lo)][16 bytes]  (tc_pack_conv_weight, N = 16)
        uint4* wt = reinterpret_cast<uint4*>(sm + p.smem_w2_off);
        const uint4* src = reinterpret_cast<const uint4*>(p.w2);
        for (int i = threadIdx.x; i < p.n_ksteps2 * 256; i += blockDim.x) {
            const int s = i >> 8, h = (i >> 7) & 1, n = i & 127;
            const int g = (n & 63) >> 4, co = n & 15, j = s - g;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (j >= 0 && j < p.K) v = __ldg(src + (j * 2 + h) * 32 + (n >> 6) * 16 + co);
            wt[i] = v;
        }
    }
    fence_proxy_async();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 2) tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_b = p.tiles_t;
    int n_my = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) ++n_my;

    if (warp == 0) {
        // =========================== producer: one x tile (with its halo) per conv1 M tile ===========================
        if (elect_one()) {
            mbar_expect_tx(w_full, (uint32_t)p.w1_bytes);
            bulk_load_1d(sm + p.smem_w1_off, p.w1, (uint32_t)p.w1_bytes, w_full);
        }
        __syncwarp();
        asm volatile("griddepcontrol.wait;" ::: "memory");
        uint32_t xs = 0, xph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
            for (int m = 0; m < T16_MT; ++m) {
                const int row0 = rem * p.m_out - p.p2 - p.p1 + 128 * m;
                mbar_wait(&empty_x[xs], xph ^ 1);
                if (elect_one()) {
                    uint8_t* dst = sm + p.smem_x_off + xs * p.x_stage_bytes;
                    if (p.ablate & 8) mbar_arrive(&full_x[xs]);
                    else {
                        mbar_expect_tx(&full_x[xs], (uint32_t)(p.x_rows * 64));
                        tma_load_3d(dst, &xmap, 0, row0, b, &full_x[xs]);
                        tma_load_3d(dst + (p.x_rows / 2) * 64, &xmap, 0, row0 + p.x_rows / 2, b, &full_x[xs]);
                    }
                }
                __syncwarp();
                if (++xs == (uint32_t)p.n_x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================== conv1: K taps x {x_hi.[w_hi; w_lo], x_lo.w_hi} per M tile ===========================
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint32_t a_desc_hi = (uint32_t)(smem_desc(0, 16, 512, 4, 0) >> 32);
        const uint32_t a_desc_lo = (uint32_t)smem_desc(0, 16, 512, 4, 0);
        const uint32_t b_desc_hi = (uint32_t)(smem_desc(0, (uint32_t)C * 32u, 128, 0, 0) >> 32);
        const uint32_t b_desc_lo = (uint32_t)smem_desc(0, (uint32_t)C * 32u, 128, 0, 0);
        const uint32_t w1_lo = b_desc_lo | (((sm_base + p.smem_w1_off) >> 4) & 0x3fffu);
        const uint32_t ab_fmt = (1u << 7) | (1u << 10);
        const uint32_t idesc = (1u << 4) | ab_fmt | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
        const uint32_t idesc_stack = (1u << 4) | ab_fmt | ((uint32_t)((2 * C) >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
        const uint32_t tap1 = (uint32_t)(p.D * 64) >> 4;
        uint32_t xs = 0, xph = 0;
        for (int i = 0; i < n_my; ++i) {
            const uint32_t st = (uint32_t)i & 1u, n = (uint32_t)i >> 1;
            for (int m = 0; m < T16_MT; ++m) {
                mbar_wait(&full_x[xs], xph);
                if (n > 0) mbar_wait(&a_empty[st * 4 + m], (n - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = sm_base + p.smem_x_off + xs * (uint32_t)p.x_stage_bytes;
                    const uint32_t d_tmem = tmem_base + (st * 4u + (uint32_t)m) * 32u;
                    uint32_t a_tap = a_desc_lo | ((a_base >> 4) & 0x3fffu), b_tap = w1_lo, acc = 0;
#pragma unroll 1
                    for (int j = 0; j < ((p.ablate & 1) ? 0 : p.K); ++j) {
                        const uint64_t da_hi = ((uint64_t)a_desc_hi << 32) | a_tap;
                        const uint64_t da_lo = ((uint64_t)a_desc_hi << 32) | (a_tap + 2u);       // lo plane: +32 bytes in the row
                        const uint64_t db = ((uint64_t)b_desc_hi << 32) | b_tap;
                        umma_bf16(d_tmem, da_hi, db, idesc_stack, acc);
                        umma_bf16(d_tmem, da_lo, db, idesc, 1);
                        acc = 1;
                        a_tap += tap1;
                        b_tap += (uint32_t)C * 4u;
                    }
                    umma_commit(&empty_x[xs]);
                    umma_commit(&a_full[st * 4 + m]);
                }
                __syncwarp();
                if (++xs == (uint32_t)p.n_x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // =========================== (TMAE) residual rows of every super tile -> shared memory ===========================
        // Four boxes, one per position g of a group: `rows` consecutive groups x 64 bytes (hi | lo), 64B-swizzled - the layout in which
        // a lane's row is a conflict-free 16-byte access.  Single buffer: refilled once every epilogue warp's output store has read it.
        if (TMAE) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            int i = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
                const int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
                if (i > 0) mbar_wait(rz_empty, (uint32_t)(i - 1) & 1u);
                if (elect_one()) {
                    mbar_expect_tx(rz_full, (ACC ? 8u : 4u) * 128u * 64u);
                    for (int g = 0; g < 4; ++g) {
                        tma_load_4d(sm + p.smem_rz_off + g * 8192, &rmap, 0, g, rem * p.rows_out, b, rz_full);
                        if (ACC) tma_load_4d(sm + p.smem_rz_off + 32768 + g * 8192, &amap, 0, g, rem * p.rows_out, b, rz_full);   // running sum
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // =========================== conv2, time-grouped: G + K - 1 k-steps x {xt_hi.[T_hi; T_lo], xt_lo.T_hi} ===========================
        const uint32_t a_desc_hi = (uint32_t)(smem_desc(0, 16, 1024, 2, 0) >> 32);
        const uint32_t a_desc_lo = (uint32_t)smem_desc(0, 16, 1024, 2, 0);
        const uint32_t b_desc_hi = (uint32_t)(smem_desc(0, 2048, 128, 0, 0) >> 32);
        const uint32_t b_desc_lo = (uint32_t)smem_desc(0, 2048, 128, 0, 0);
        const uint32_t w2_lo = b_desc_lo | (((sm_base + p.smem_w2_off) >> 4) & 0x3fffu);
        const uint32_t ab_fmt = (1u << 7) | (1u << 10);
        const uint32_t idesc64 = (1u << 4) | ab_fmt | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
        const uint32_t idesc128 = (1u << 4) | ab_fmt | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
        for (int i = 0; i < n_my; ++i) {
            const uint32_t st = (uint32_t)i & 1u, n = (uint32_t)i >> 1;
            mbar_wait(&mid_full[st], n & 1u);      // (every group has read conv2(i - 2)'s accumulator before it wrote this xt: program order)
            tc_fence_after();
            if (elect_one()) {
                const uint32_t hi_base = sm_base + p.smem_xt_off + st * 2u * (uint32_t)T16_XT_PLANE;
                uint32_t a_hi = a_desc_lo | ((hi_base >> 4) & 0x3fffu);
                uint32_t a_lo = a_desc_lo | (((hi_base + (uint32_t)T16_XT_PLANE) >> 4) & 0x3fffu);
                uint32_t b_s = w2_lo, acc = 0;
                const uint32_t d_tmem = tmem_base + 256u + st * 128u;
#pragma unroll 1
                for (int s = 0; s < ((p.ablate & 1) ? 0 : p.n_ksteps2); ++s) {
                    const uint64_t db = ((uint64_t)b_desc_hi << 32) | b_s;
                    umma_bf16(d_tmem, ((uint64_t)a_desc_hi << 32) | a_hi, db, idesc128, acc);
                    umma_bf16(d_tmem, ((uint64_t)a_desc_hi << 32) | a_lo, db, idesc64, 1);
                    acc = 1;
                    a_hi += 2u; a_lo += 2u;        // next input position: +32 bytes (4 positions per 128-byte row, rows contiguous)
                    b_s += 256u;                   // 4096 bytes per k-step
                }
                umma_commit(&b_full[st]);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =========================== epilogue groups ===========================
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float* bias1 = reinterpret_cast<const float*>(sm + p.smem_bias_off);
        const float* bias2 = bias1 + C;
        auto valid_of = [&](int bb) {
            int v = p.L;
            if (p.lens) { const int u = __ldg(p.lens + bb) * p.len_mul; v = u < v ? u : v; }
            return v;
        };
        // epilogue A of this CTA's i-th super tile: conv1 M tile `grp` -> xt positions 128 grp .. 128 grp + 127 of stage i & 1
        auto epi_a = [&](int i) {
            const int tile = blockIdx.x + i * (int)gridDim.x;
            const int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
            const int o0 = rem * p.m_out;
            const uint32_t st = (uint32_t)i & 1u, n = (uint32_t)i >> 1;
            const int tau = 128 * grp + row;
            const int tm = o0 - p.p2 + tau;
            const bool in_seq = tm >= 0 && tm < valid_of(b);
            const bool edge = __any_sync(0xffffffffu, !in_seq);
            if (ACC && !TMAE) {      // the running-sum row this thread adds in epilogue B of the same tile: pull it into L2 now
                const int to = o0 + 4 * row + grp;
                if (row < p.rows_out && to < p.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.accin + ((size_t)b * p.L + to) * 32));
            }
            mbar_wait(&a_full[st * 4 + grp], n & 1u);
            tc_fence_after();
            float v[16];
            const uint32_t taddr = tmem_base + lane_off + (st * 4u + (uint32_t)grp) * 32u;
            if (p.ablate & 64) { for (int z = 0; z < 16; ++z) v[z] = 1.0f; } else tmem_ld16x2_sum(taddr, taddr + 16u, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[st * 4 + grp]);
            uint32_t hw[8], lw[8];
            const float4* bp = reinterpret_cast<const float4*>(bias1);
            if (p.ablate & 32) {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) { hw[k2] = __float_as_uint(v[2 * k2]); lw[k2] = __float_as_uint(v[2 * k2 + 1]); }
            } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const float4 bv = bp[k4];
                add2(v[4 * k4], v[4 * k4 + 1], bv.x, bv.y);
                add2(v[4 * k4 + 2], v[4 * k4 + 3], bv.z, bv.w);
            }
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                float a0 = v[2 * k2], a1 = v[2 * k2 + 1];
                float t0 = a0, t1 = a1;
                mul2(t0, t1, p.mid_slope, p.mid_slope);
                a0 = fmaxf(a0, t0);
                a1 = fmaxf(a1, t1);
                split2(a0, a1, 0, hw[k2], lw[k2]);
            }
            if (edge && !in_seq) {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) hw[k2] = lw[k2] = 0u;
            }
            }
            // position tau -> 128-byte row tau / 4, 16-byte chunks 2 (tau % 4), + 1; 128-byte swizzle: chunk ^= row % 8
            uint8_t* prow = sm + p.smem_xt_off + st * 2u * (uint32_t)T16_XT_PLANE + (uint32_t)(tau >> 2) * 128u;
            const uint32_t c0 = (uint32_t)(tau & 3) * 2u, swz = (uint32_t)(tau >> 2) & 7u;
            *reinterpret_cast<uint4*>(prow + ((c0 ^ swz) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(prow + (((c0 + 1u) ^ swz) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            *reinterpret_cast<uint4*>(prow + T16_XT_PLANE + ((c0 ^ swz) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            *reinterpret_cast<uint4*>(prow + T16_XT_PLANE + (((c0 + 1u) ^ swz) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mid_full[st]);
        };
        // epilogue B: accumulator row r = output positions 4 r .. 4 r + 3; this group takes position 4 r + grp (columns 16 grp ..)
        // (the residual / running-sum rows are requested right before epilogue B: earlier - before epilogue A of the next tile - cost
        // 5-10 % per launch, measured)
        uint4 R[4], A[4];
        auto request_b = [&](int i) {
            const int tile = blockIdx.x + i * (int)gridDim.x;
            const int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
            const int to = rem * p.m_out + 4 * row + grp;
            if (TMAE || !(row < p.rows_out && to < p.L) || (p.ablate & 4)) return;
            const size_t row_el = ((size_t)b * p.L + to) * 32;
            ld_global_nc_v8(p.res + row_el, R[0], R[1]);
            ld_global_nc_v8(p.res + row_el + C, R[2], R[3]);
            if (ACC) {
                ld_global_v8(p.accin + row_el, A[0], A[1]);
                ld_global_v8(p.accin + row_el + C, A[2], A[3]);
            }
        };
        auto epi_b = [&](int i) {
            const int tile = blockIdx.x + i * (int)gridDim.x;
            const int b = tile / tiles_per_b, rem = tile - b * tiles_per_b;
            const int o0 = rem * p.m_out;
            const uint32_t st = (uint32_t)i & 1u, n = (uint32_t)i >> 1;
            const int to = o0 + 4 * row + grp;
            const bool row_ok = row < p.rows_out && to < p.L;
            const bool live = to < valid_of(b);
            const size_t row_el = ((size_t)b * p.L + to) * 32;
            uint8_t* rrow = sm + p.smem_rz_off + grp * 8192 + row * 64;      // (TMAE) this thread's residual row, then its output row
            const uint32_t rsw = ((uint32_t)row >> 1) & 3u;                     // 64B swizzle: 16-byte chunk ^= (row / 2) % 4
            if (TMAE) {
                mbar_wait(rz_full, (uint32_t)i & 1u);
                R[0] = *reinterpret_cast<const uint4*>(rrow + ((0u ^ rsw) << 4));
                R[1] = *reinterpret_cast<const uint4*>(rrow + ((1u ^ rsw) << 4));
                R[2] = *reinterpret_cast<const uint4*>(rrow + ((2u ^ rsw) << 4));
                R[3] = *reinterpret_cast<const uint4*>(rrow + ((3u ^ rsw) << 4));
                if (ACC) {
                    A[0] = *reinterpret_cast<const uint4*>(rrow + 32768 + ((0u ^ rsw) << 4));
                    A[1] = *reinterpret_cast<const uint4*>(rrow + 32768 + ((1u ^ rsw) << 4));
                    A[2] = *reinterpret_cast<const uint4*>(rrow + 32768 + ((2u ^ rsw) << 4));
                    A[3] = *reinterpret_cast<const uint4*>(rrow + 32768 + ((3u ^ rsw) << 4));
                }
            }
            mbar_wait(&b_full[st], n & 1u);
            tc_fence_after();
            float v[16];
            const uint32_t taddr = tmem_base + lane_off + 256u + st * 128u + 16u * (uint32_t)grp;
            if (p.ablate & 64) { for (int z = 0; z < 16; ++z) v[z] = 1.0f; } else tmem_ld16x2_sum(taddr, taddr + 64u, v);
            tc_fence_before();          // (ordered before conv2(i + 2) by this warp's mid_full arrival of A(i + 2))
            if (!TMAE && (!row_ok || (p.ablate & 32))) return;
            uint32_t hw[8], lw[8];
            if (row_ok) {
            const float4* bp = reinterpret_cast<const float4*>(bias2);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const float4 bv = bp[k4];
                add2(v[4 * k4], v[4 * k4 + 1], bv.x, bv.y);
                add2(v[4 * k4 + 2], v[4 * k4 + 3], bv.z, bv.w);
            }
            const uint32_t hh[8] = {R[0].x, R[0].y, R[0].z, R[0].w, R[1].x, R[1].y, R[1].z, R[1].w};
            const uint32_t ll[8] = {R[2].x, R[2].y, R[2].z, R[2].w, R[3].x, R[3].y, R[3].z, R[3].w};
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const float2 r = join2(hh[k2], ll[k2], 0);
                float tx = r.x, ty = r.y;
                mul2(tx, ty, p.res_unslope, p.res_unslope);
                add2(v[2 * k2], v[2 * k2 + 1], fminf(r.x, tx), fminf(r.y, ty));      // undo the stored leaky relu
            }
            if (ACC) {
                const uint32_t ah[8] = {A[0].x, A[0].y, A[0].z, A[0].w, A[1].x, A[1].y, A[1].z, A[1].w};
                const uint32_t al[8] = {A[2].x, A[2].y, A[2].z, A[2].w, A[3].x, A[3].y, A[3].z, A[3].w};
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const float2 r = join2(ah[k2], al[k2], 0);
                    add2(v[2 * k2], v[2 * k2 + 1], r.x, r.y);
                }
            }
            const float sc = live ? p.scale : 0.f;
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                float a0 = v[2 * k2], a1 = v[2 * k2 + 1];
                mul2(a0, a1, sc, sc);
                float t0 = a0, t1 = a1;
                mul2(t0, t1, p.post_slope, p.post_slope);
                a0 = fmaxf(a0, t0);
                a1 = fmaxf(a1, t1);
                split2(a0, a1, 0, hw[k2], lw[k2]);
            }
            }
            if (TMAE) {
                // the output row replaces the residual row in place; one lane hands the warp's 32 rows (the last warp: rows_out - 96) to
                // the TMA unit, which clips the groups past the end of the sequence
                if (row_ok) {
                    *reinterpret_cast<uint4*>(rrow + ((0u ^ rsw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(rrow + ((1u ^ rsw) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    *reinterpret_cast<uint4*>(rrow + ((2u ^ rsw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    *reinterpret_cast<uint4*>(rrow + ((3u ^ rsw) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (!(p.ablate & 2) && o0 + 128 * q < p.L) {
                        tma_store_4d(q == 3 ? &omap_tail : &omap, sm + p.smem_rz_off + grp * 8192 + q * 2048, 0, grp, rem * p.rows_out + q * 32, b);
                        bulk_commit_group();
                        bulk_wait_group_read<0>();
                    }
                    mbar_arrive(rz_empty);
                }
                __syncwarp();
                return;
            }
            if (p.ablate & 2 || !row_ok) return;
            __nv_bfloat16* oh = p.out + row_el;
            st_global_v8(oh, make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]));
            st_global_v8(oh + C, make_uint4(lw[0], lw[1], lw[2], lw[3]), make_uint4(lw[4], lw[5], lw[6], lw[7]));
        };
        if (n_my > 0) epi_a(0);
        for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) epi_a(i + 1);
            request_b(i);
            epi_b(i);
        }
    }

    if (TMAE && warp >= 4 && lane == 0) bulk_wait_group<0>();      // this lane's TMA stores have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512u);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------
bool tc_pair_supported(int C, int K, int D) {
    if (C != 16 && C != 32) return false;
    if (K < 1 || K % 2 == 0 || K > 15 || D < 1) return false;
    const int x_rows = (int)round_up(TP_M + (K - 1) * D, 16);
    return x_rows <= 512;
}

// C = 16: conv2 time-grouped (tc_pair16_kernel)
static bool pair16_tma_epilogue(const TcPairArgs& a) {      // residual (+ running-sum) staging fits beside two x stages, and L % 4 == 0
    static const int tmae_env = getenv("PB_TP_TMAE") ? atoi(getenv("PB_TP_TMAE")) : 1;
    const uint32_t fixed = 2048u + (uint32_t)round_up(a.K * 1024, 1024) + (uint32_t)(T16_G + a.K - 1) * 4096u + 4u * (uint32_t)T16_XT_PLANE;
    const uint32_t x_stage = (uint32_t)round_up(round_up(TP_M + (a.K - 1) * a.D, 16) * 64, 1024);
    return tmae_env != 0 && a.L % 4 == 0 && fixed + (a.accin ? 65536u : 32768u) + 2u * x_stage + 2048u <= TP_SMEM_LIMIT;
}

static int launch_tc_pair16(const TcPairArgs& a, cudaStream_t st) {
    Pair16Params p{};
    p.w1 = a.w1; p.w2 = a.w2; p.bias1 = a.bias1; p.bias2 = a.bias2; p.res = a.x; p.accin = a.accin; p.out = a.out;
    p.lens = a.lens; p.len_mul = a.len_mul;
    p.B = a.B; p.L = a.L; p.K = a.K; p.D = a.D;
    p.p1 = (a.K - 1) * a.D / 2; p.p2 = (a.K - 1) / 2;
    p.rows_out = (T16_MT * TP_M - (a.K - 1)) / T16_G;
    p.m_out = p.rows_out * T16_G;
    p.n_ksteps2 = T16_G + a.K - 1;
    p.x_rows = (int)round_up(TP_M + (a.K - 1) * a.D, 16);
    p.x_stage_bytes = (int)round_up((int64_t)p.x_rows * 64, 1024);
    p.w1_bytes = a.K * 1024;
    p.smem_bias_off = 1024;
    p.smem_w1_off = 2048;
    p.smem_w2_off = p.smem_w1_off + (uint32_t)round_up(p.w1_bytes, 1024);
    p.smem_xt_off = p.smem_w2_off + (uint32_t)p.n_ksteps2 * 4096u;
    // epilogue B through shared memory + TMA (residual and running-sum rows in, output rows out, all as 128-byte wavefronts) instead of
    // row-per-lane global accesses 256 bytes apart (32 wavefronts each): where the staging (32 KB, 64 KB with the running sum) fits beside
    // two x stages and L % 4 == 0 (the grouped maps).  PB_TP_TMAE=0 disables.
    p.smem_rz_off = p.smem_xt_off + 4u * (uint32_t)T16_XT_PLANE;
    const uint32_t rz_bytes = a.accin ? 65536u : 32768u;          // (+ the running-sum rows)
    const bool tmae = pair16_tma_epilogue(a);
    p.smem_x_off = p.smem_rz_off + (tmae ? rz_bytes : 0u);
    const uint32_t left = TP_SMEM_LIMIT - 1024u - p.smem_x_off;
    p.n_x_stages = (int)(left / (uint32_t)p.x_stage_bytes);
    if (p.n_x_stages > TP_MAX_XST) p.n_x_stages = TP_MAX_XST;
    PB_REQUIRE(p.n_x_stages >= 2, PB_ERR_INVALID, "tc_pair16: tile does not fit in shared memory (K=%d D=%d)", a.K, a.D);
    const uint32_t smem_bytes = p.smem_x_off + (uint32_t)p.n_x_stages * (uint32_t)p.x_stage_bytes + 1024;
    p.tiles_t = (int)ceil_div(a.L, p.m_out);
    const int64_t total = (int64_t)a.B * p.tiles_t;
    PB_REQUIRE(total < (1LL << 31), PB_ERR_INVALID, "tc_pair16: too many tiles");
    p.total_tiles = (int)total;
    p.mid_slope = a.mid_slope; p.post_slope = a.post_slope; p.scale = a.scale; p.res_unslope = a.res_unslope;
    static const int ablate = getenv("PB_TP_ABLATE") ? atoi(getenv("PB_TP_ABLATE")) : 0;
    p.ablate = ablate;

    CUtensorMap xmap;
    PB_TRY(tc_make_tmap_act(&xmap, a.x, 16, a.L, a.B, 32, p.x_rows / 2, false));
    CUtensorMap rmap = xmap, amap = xmap, omap = xmap, omap_tail = xmap;      // only read when tmae
    if (tmae) {
        PB_TRY(tc_make_tmap_grouped16(&rmap, a.x, a.L, a.B, 128));
        if (a.accin) PB_TRY(tc_make_tmap_grouped16(&amap, a.accin, a.L, a.B, 128));
        PB_TRY(tc_make_tmap_grouped16(&omap, a.out, a.L, a.B, 32));
        PB_TRY(tc_make_tmap_grouped16(&omap_tail, a.out, a.L, a.B, p.rows_out - 96));
    }

    static int n_sm = 0;
    static bool attr_set = false;
    if (!attr_set) {
        int dev = 0;
        PB_CUDA(cudaGetDevice(&dev));
        PB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        PB_CUDA(cudaFuncSetAttribute(tc_pair16_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_LIMIT));
        PB_CUDA(cudaFuncSetAttribute(tc_pair16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_LIMIT));
        PB_CUDA(cudaFuncSetAttribute(tc_pair16_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_LIMIT));
        PB_CUDA(cudaFuncSetAttribute(tc_pair16_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_LIMIT));
        attr_set = true;
    }
    const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(T16_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("PB_TC_NO_PDL") != nullptr;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    if (a.accin && tmae) PB_CUDA(cudaLaunchKernelEx(&cfg, tc_pair16_kernel<true, true>, xmap, rmap, amap, omap, omap_tail, p));
    else if (a.accin) PB_CUDA(cudaLaunchKernelEx(&cfg, tc_pair16_kernel<true, false>, xmap, rmap, amap, omap, omap_tail, p));
    else if (tmae) PB_CUDA(cudaLaunchKernelEx(&cfg, tc_pair16_kernel<false, true>, xmap, rmap, amap, omap, omap_tail, p));
    else PB_CUDA(cudaLaunchKernelEx(&cfg, tc_pair16_kernel<false, false>, xmap, rmap, amap, omap, omap_tail, p));
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int launch_tc_pair(const TcPairArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.x && a.w1 && a.w2 && a.out && a.bias1 && a.bias2, PB_ERR_INVALID, "tc_pair: null tensor");
    PB_REQUIRE(tc_pair_supported(a.C, a.K, a.D), PB_ERR_INVALID, "tc_pair: C=%d K=%d D=%d unsupported", a.C, a.K, a.D);
    PB_REQUIRE(a.B > 0 && a.L > 0, PB_ERR_INVALID, "tc_pair: bad shape");
    // C = 16: time-grouped conv2 where it pays - k >= 7 (per launch at 8 x 1024 units, same box: k = 11 170 -> 148 us, with the running
    // sum 180 -> 159; k = 7 128 -> 113 / 143 -> 142; k = 3 is not tensor-bound: 103 -> 104).  With the running sum only when its rows
    // can be staged through shared memory too: as row-per-lane global accesses they cost more than the MMAs save (190 -> 201 us).
    // PB_TP_TOEPLITZ=0 / 1 forces it off / on for every 16-channel pair.
    static const int toeplitz_env = getenv("PB_TP_TOEPLITZ") ? atoi(getenv("PB_TP_TOEPLITZ")) : -1;
    if (a.C == 16 && a.K >= 3 && (toeplitz_env < 0 ? (a.K >= 7 && (!a.accin || pair16_tma_epilogue(a))) : toeplitz_env != 0)) return launch_tc_pair16(a, st);
    PairParams p{};
    p.w1 = a.w1; p.w2 = a.w2; p.bias1 = a.bias1; p.bias2 = a.bias2; p.res = a.x; p.accin = a.accin; p.out = a.out;
    p.lens = a.lens; p.len_mul = a.len_mul;
    static const int ablate = getenv("PB_TP_ABLATE") ? atoi(getenv("PB_TP_ABLATE")) : 0;
    p.ablate = ablate;
    p.B = a.B; p.L = a.L; p.C = a.C; p.K = a.K; p.D = a.D;
    p.p1 = (a.K - 1) * a.D / 2; p.p2 = (a.K - 1) / 2; p.m_out = TP_M - (a.K - 1);
    p.ksteps = a.C / 16;
    p.row_bytes = a.C == 16 ? 64 : 128;
    p.lo_off = a.C == 16 ? 32 : 64;
    p.a_layout = a.C == 16 ? 4 : 2;
    p.a_sbo = 8 * p.row_bytes;
    p.x_rows = (int)round_up(TP_M + (a.K - 1) * a.D, 16);
    p.x_stage_bytes = (int)round_up((int64_t)p.x_rows * p.row_bytes, 1024);
    p.mid_bytes = (int)round_up((int64_t)(TP_M + a.K - 1) * p.row_bytes, 1024);
    p.w_bytes = a.K * 2 * p.ksteps * a.C * 32;
    p.smem_bias_off = 1024;
    p.smem_w1_off = 2048;
    p.smem_w2_off = p.smem_w1_off + (uint32_t)round_up(p.w_bytes, 1024);
    p.smem_mid_off = p.smem_w2_off + (uint32_t)round_up(p.w_bytes, 1024);
    // Tiles in flight (= epilogue groups = accumulator / xt stages).  The k <= 7 launches are bound by the per-tile chain conv1 -> xt ->
    // conv2 -> output, not by the tensor pipe: a fourth group buys 4-8 % there (per launch at 8 x 1024 units: k = 3 107 -> 99 us at
    // C = 16, 115 -> 110 at C = 32; k = 7 152 -> 140 / 153 -> 141).  At k = 11 (tensor-bound; fewer x stages) and with the running sum
    // (96 registers per thread: spills) it costs 5-15 us, so those keep three.  PB_TP_GROUPS=3/4 forces it where it fits.
    static const int ng_env = getenv("PB_TP_GROUPS") ? atoi(getenv("PB_TP_GROUPS")) : 0;
    int ng = ng_env == 3 || ng_env == 4 ? ng_env : (!a.accin && a.K <= 7 ? 4 : 3);
    if (ng == 4 && (p.smem_mid_off + 4u * (uint32_t)p.mid_bytes + 2u * (uint32_t)p.x_stage_bytes + 2048u > TP_SMEM_LIMIT || 8 * 2 * a.C > 512)) ng = 3;
    p.smem_x_off = p.smem_mid_off + (uint32_t)ng * (uint32_t)p.mid_bytes;
    const uint32_t left = TP_SMEM_LIMIT - 1024u - p.smem_x_off;
    // TMA-store epilogue: 2 KB of staging per epilogue warp, when at least 3 xt stages remain (PB_TC_TMA_OUT=0 disables).  Measured
    // per launch on B200: -6 us at C = 32 (k = 3), but +5 .. +30 us at C = 16, where the kernel is bound by shared-memory traffic
    // (narrow A-operand reads + the xt tile) and the staging tiles add to it - so 32-channel layers only.
    static const int tma_out_mode = getenv("PB_TC_TMA_OUT") ? atoi(getenv("PB_TC_TMA_OUT")) : 1;
    const uint32_t stage_bytes = (uint32_t)ng * 4u * 2048u;
    p.tma_out = (tma_out_mode > 0 && a.C == 32 && left >= stage_bytes + 3u * (uint32_t)p.x_stage_bytes && p.m_out > 96) ? 1 : 0;
    p.n_x_stages = (int)((left - (p.tma_out ? stage_bytes : 0u)) / (uint32_t)p.x_stage_bytes);
    if (p.n_x_stages > TP_MAX_XST) p.n_x_stages = TP_MAX_XST;
    // the residual from the x tile (held until epilogue B): needs a stage per tile in flight and one to prefetch into.  At C = 32 a lane's
    // residual row is its own 128-byte line (32 L1 wavefronts per global load): k = 3 112 -> 105 us, k = 7 with the running sum 179 -> 171;
    // at C = 16 (two rows per line) nothing: 104.5 -> 105.  PB_TP_RES_SMEM=0 disables, 2 forces it on at C = 16 too.
    static const int res_smem_env = getenv("PB_TP_RES_SMEM") ? atoi(getenv("PB_TP_RES_SMEM")) : 1;
    p.res_smem = ((res_smem_env == 2 || (res_smem_env == 1 && a.C == 32)) && p.n_x_stages >= ng + 1) ? 1 : 0;
    PB_REQUIRE(p.n_x_stages >= 2, PB_ERR_INVALID, "tc_pair: tile does not fit in shared memory (C=%d K=%d D=%d)", a.C, a.K, a.D);
    p.smem_stage_off = p.smem_x_off + (uint32_t)p.n_x_stages * (uint32_t)p.x_stage_bytes;
    const uint32_t smem_bytes = p.smem_stage_off + (p.tma_out ? stage_bytes : 0u) + 1024;
    p.tiles_t = (int)ceil_div(a.L, p.m_out);
    const int64_t total = (int64_t)a.B * p.tiles_t;
    PB_REQUIRE(total < (1LL << 31), PB_ERR_INVALID, "tc_pair: too many tiles");
    p.total_tiles = (int)total;
    p.mid_slope = a.mid_slope; p.post_slope = a.post_slope; p.scale = a.scale; p.res_unslope = a.res_unslope;
    const uint32_t ab_fmt = (1u << 7) | (1u << 10);
    p.idesc = (1u << 4) | ab_fmt | ((uint32_t)(a.C >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
    p.idesc_stack = (1u << 4) | ab_fmt | ((uint32_t)((2 * a.C) >> 3) << 17) | ((uint32_t)(TP_M >> 4) << 24);
    p.acc_cols = 2 * a.C;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * ng * p.acc_cols) p.tmem_cols *= 2;

    // x tile = two boxes of x_rows / 2 rows (a TMA box holds at most 256 rows)
    CUtensorMap xmap;
    PB_TRY(tc_make_tmap_act(&xmap, a.x, a.C, a.L, a.B, a.C == 16 ? 32 : 64, p.x_rows / 2, a.C != 16));

    CUtensorMap omap = xmap, omap_tail = xmap;      // only read when tma_out
    if (p.tma_out) {
        PB_TRY(tc_make_tmap_out(&omap, a.out, a.C, a.L, a.B, 32));
        PB_TRY(tc_make_tmap_out(&omap_tail, a.out, a.C, a.L, a.B, p.m_out - 96));
    }

    static int n_sm = 0;
    typedef void (*PairKernel)(CUtensorMap, CUtensorMap, CUtensorMap, PairParams);
    static const PairKernel kernels[2][2][2] = {      // [accin][ksteps - 1][ng - 3]
        {{tc_pair_kernel<false, 1, 3>, tc_pair_kernel<false, 1, 4>}, {tc_pair_kernel<false, 2, 3>, tc_pair_kernel<false, 2, 4>}},
        {{tc_pair_kernel<true, 1, 3>, tc_pair_kernel<true, 1, 4>}, {tc_pair_kernel<true, 2, 3>, tc_pair_kernel<true, 2, 4>}}};
    static bool attr_set = false;
    if (!attr_set) {
        int dev = 0;
        PB_CUDA(cudaGetDevice(&dev));
        PB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        for (int i = 0; i < 8; ++i)
            PB_CUDA(cudaFuncSetAttribute(kernels[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_LIMIT));
        attr_set = true;
    }
    const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(128u + 128u * (unsigned)ng); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("PB_TC_NO_PDL") != nullptr;      // dev switch, as in tc_conv.cu
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    PB_CUDA(cudaLaunchKernelEx(&cfg, kernels[a.accin ? 1 : 0][p.ksteps - 1][ng - 3], xmap, omap, omap_tail, p));
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
