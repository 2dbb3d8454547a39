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
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pb_tc.cuh"
#include "pb_umma.cuh"

namespace pb {

namespace {

constexpr int TC_MAX_GROUPS = 3;
constexpr int TC_MAX_THREADS = 128 + 128 * TC_MAX_GROUPS;   // 4 service warps + up to 3 epilogue groups of 4 warps (128 registers per thread)
constexpr int TC_M = 128;                 // time rows per tile (UMMA M, cta_group::1)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_ACC = 8;                // accumulator stages in TMEM
constexpr uint32_t TC_SMEM_LIMIT = 232448;  // 227 KB: one CTA per SM
constexpr uint32_t TC_SMEM_HALF = 115712;   // 113 KB: two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)

struct TcParams {
    const uint8_t* w;
    const float* bias;
    const __nv_bfloat16* res;
    const float* res_f32;
    const __nv_bfloat16* accin;
    __nv_bfloat16* out;
    float* out_f32;
    const int32_t* lens;
    unsigned long long* argmax_key;   // fused row argmax (see TcConvArgs::argmax_key) or null
    int argmax_n;
    int len_mul;
    int B, L, Cin, N, NT, n_tiles_n, out_C, U;
    int K, D, pad;
    int nchunks, ksteps, cin_t;
    int a_loads, a_row_bytes, a_tile_bytes, a_stage_bytes, lo_off, a_layout, a_sbo;
    int n_a_stages, n_b_stages, b_block_bytes, b_sub_bytes, kpb, w_total_bytes;
    int res_blocks, total_blocks;          // weight blocks [nt][chunk][tap] below res_blocks stay resident in shared memory
    uint32_t smem_a_off, smem_b_off, smem_ring_off, smem_bias_off, smem_stage_off;
    int tma_out;    // epilogue writes through per-warp staging tiles + TMA stores (U == 1 layers)
    int tiles_t, total_tiles;
    float post_slope, scale, res_unslope;
    int tanh_out;
    uint32_t idesc, idesc_stack;
    int single_pass, stack, acc_cols;
    int cluster;                         // 1: CTA pairs (cta_group::2): M = 256 per pair, each CTA streams half of the weight rows
    int fmt;                             // split storage / MMA operand type: 0 bf16, 1 fp16
    int tmem_cols, n_acc, n_groups;      // n_groups epilogue warpgroups take the tiles of a CTA round-robin (n_acc % n_groups == 0)
    int ablate;   // dev only (PB_TC_ABLATE): 1 = skip the MMAs, 2 = skip the output stores, 4 = skip residual / running-sum loads, 8 = skip the activation TMA loads,
                  // 16 = skip the accumulator reads, 32 = skip the epilogue arithmetic, 64 = skip the streamed weight loads
    int epi_split;  // column-split epilogue groups: every group works on EVERY tile and takes the 16-column items grp, grp + n_groups, ...
                    // - a tile's epilogue latency drops from (NT/16 items) to (NT/16 / n_groups) and n_groups is no longer capped by the
                    // number of accumulators (measured on B200, 8 x 1024 units: 8.08 -> 7.94 ms on the CTA-pair layers, slower on the others)
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------------------
// MMA issuer (warp 1).  The loops run warp-uniformly; one elected lane issues.  KSTEPS (16-channel k-steps per chunk) and MODE
// are compile-time so that one tap is a straight line of tcgen05.mma with immediate descriptor offsets (measured: with runtime
// loop bounds the ~270 cycles of constant-bank loads and branches per tap, not the tensor pipe, bounded the narrow layers).
//   MODE 2 (NT <= 128):  D[0,2NT) += A_hi.[B_hi;B_lo]^T ;  D[0,NT) += A_lo.B_hi^T      (2 MMAs per k-step, epilogue adds the halves)
//   MODE 1 (NT = 256):   D += A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T                   (3 MMAs)
//   MODE 0:              D += A_hi.B_hi^T                                               (plain bf16)
// ------------------------------------------------------------------------------------------------------------------------
struct MmaCtx {
    uint64_t *full_a, *empty_a, *full_b, *empty_b, *acc_full, *acc_empty, *b_res_full, *peer_a, *peer_b;
    uint32_t sm_base, tmem_base;
    int cta0, nctas;      // first tile and tile stride of this CTA (pair)
};

template <int MODE, bool CL>
__device__ __forceinline__ void issue_kstep(uint32_t d_tmem, uint32_t a_hi_word, uint32_t b_hi_word, uint32_t a_k, uint32_t b_k,
                                            uint32_t lo_step, uint32_t lo_rows_step, uint32_t idesc, uint32_t idesc_stack,
                                            uint32_t accumulate) {
    const uint64_t da_hi = ((uint64_t)a_hi_word << 32) | a_k;
    const uint64_t db_hi = ((uint64_t)b_hi_word << 32) | b_k;
    if (CL) {
        // CTA pair, M = 256.  Operand R1 (N = 256 over the pair: rank 0 holds the 128 hi rows, rank 1 the 128 lo rows) gives
        // a_hi.[w_hi; w_lo]; operand R2 (N = 128: rank r holds hi rows 64r..64r+63, K-half stride 1 KB) gives a_lo.w_hi.
        // lo_rows_step carries R2's offset inside the k-step, idesc_stack the N = 256 descriptor.
        umma2_bf16(d_tmem, da_hi, db_hi, idesc_stack, accumulate);
        const uint64_t db_r2 = ((uint64_t)b_hi_word << 32) | (b_k + lo_rows_step);   // lo_rows_step = R2 offset + the LBO change (2 KB -> 1 KB)
        umma2_bf16(d_tmem, ((uint64_t)a_hi_word << 32) | (a_k + lo_step), db_r2, idesc, 1);
    } else if (MODE == 2) {
        umma_bf16(d_tmem, da_hi, db_hi, idesc_stack, accumulate);
        umma_bf16(d_tmem, ((uint64_t)a_hi_word << 32) | (a_k + lo_step), db_hi, idesc, 1);
    } else {
        umma_bf16(d_tmem, da_hi, db_hi, idesc, accumulate);
        if (MODE == 1) {
            umma_bf16(d_tmem, ((uint64_t)a_hi_word << 32) | (a_k + lo_step), db_hi, idesc, 1);
            umma_bf16(d_tmem, da_hi, ((uint64_t)b_hi_word << 32) | (b_k + lo_rows_step), idesc, 1);
        }
    }
}

template <int KSTEPS, int MODE, bool CL = false>
__device__ __forceinline__ void mma_issuer(const TcParams& p, const MmaCtx& mc) {
    auto commit_ = [](uint64_t* bar) { if (CL) umma2_commit(bar); else umma_commit(bar); };
    if (p.res_blocks > 0) { mbar_wait(mc.b_res_full, 0); tc_fence_after(); }
    uint32_t a_s = 0, a_ph = 0, b_s = 0, b_ph = 0, acc_s = 0, acc_ph = 0;
    // everything the inner loops need, read from the parameter block once
    const int K = p.K, nchunks = p.nchunks, n_tiles_n = p.n_tiles_n, res_blocks = p.res_blocks, kpb = p.kpb;
    const int nsub = KSTEPS / kpb;
    const uint32_t n_a_stages = (uint32_t)p.n_a_stages, n_b_stages = (uint32_t)p.n_b_stages, n_acc = (uint32_t)p.n_acc;
    const uint32_t acc_cols = (uint32_t)p.acc_cols, a_stage_bytes = (uint32_t)p.a_stage_bytes, b_sub_bytes = (uint32_t)p.b_sub_bytes;
    const uint32_t idesc = p.idesc, idesc_stack = p.idesc_stack;
    const bool skip_mma = (p.ablate & 1) != 0;
    // descriptor halves that never change: hi word = SBO | version | layout, lo word = start address | LBO
    const uint32_t a_desc_hi = (uint32_t)(smem_desc(0, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0) >> 32);
    const uint32_t a_desc_lo = (uint32_t)smem_desc(0, 16, (uint32_t)p.a_sbo, (uint32_t)p.a_layout, 0);
    // B k-step = [K half 0: hi rows, lo rows][K half 1: hi rows, lo rows]: LBO (K-half stride) = 2*NT*16 bytes, the lo rows
    // start NT*16 bytes after the hi rows, so [hi; lo] is also ONE operand of 2*NT rows
    // (CTA pair: a k-step of this CTA is [R1: 2 K halves x NT rows][R2: 2 K halves x NT/2 rows] = 6 KB at NT = 128; R1's K-half
    //  stride is NT*16 bytes, R2's NT/2*16 - the difference of the two descriptor low words is folded into lo_rows_step)
    const uint32_t b_lbo = CL ? (uint32_t)p.NT * 16u : (uint32_t)p.NT * 32u;
    const uint32_t b_desc_hi = (uint32_t)(smem_desc(0, b_lbo, 128, 0, 0) >> 32);
    const uint32_t b_desc_lo = (uint32_t)smem_desc(0, b_lbo, 128, 0, 0);
    const uint32_t tap_step = (uint32_t)(p.D * p.a_row_bytes) >> 4;   // in 16-byte units
    const uint32_t lo_step = (uint32_t)p.lo_off >> 4;
    const uint32_t lo_rows_step = CL ? ((uint32_t)smem_desc(0, (uint32_t)p.NT * 8u, 128, 0, 0) - b_desc_lo) + (uint32_t)p.NT * 2u
                                     : (uint32_t)p.NT;               // NT*16 bytes (pair: R2 starts NT*32 bytes into the k-step)
    const uint32_t kstep_step = CL ? (uint32_t)p.NT * 3u : (uint32_t)p.NT * 4u;   // k-step size in 16-byte units
    const uint32_t block_step = kstep_step * KSTEPS;
    const uint32_t a_base0 = mc.sm_base + p.smem_a_off, b_res0 = mc.sm_base + p.smem_b_off, b_ring0 = mc.sm_base + p.smem_ring_off;
    int nt = mc.cta0 % n_tiles_n;
    const int nt_adv = mc.nctas % n_tiles_n;
    for (int tile = mc.cta0; tile < p.total_tiles; tile += mc.nctas) {
        mbar_wait(&mc.acc_empty[acc_s], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = mc.tmem_base + acc_s * acc_cols;
        uint32_t accumulate = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&mc.full_a[a_s], a_ph);
            tc_fence_after();
            // The swizzle is a function of the absolute shared-memory address (measured on B200: a start address advanced by
            // whole rows inside a 1024-byte-aligned tile needs base_offset = 0), so the K taps are K views of one staged
            // tile: tap j starts j*D rows further down.
            const uint32_t a_lo0 = a_desc_lo | (((a_base0 + a_s * a_stage_bytes) >> 4) & 0x3fffu);   // (tap 0, k-step 0, hi)
            const int blk0 = (nt * nchunks + c) * K;
            int j0 = res_blocks - blk0;                 // taps below j0 are resident
            j0 = j0 < 0 ? 0 : (j0 > K ? K : j0);
            const bool last_c = c + 1 == nchunks;
            if (j0 > 0) {
                // resident taps: one elected lane, no further synchronisation
                if (elect_one()) {
                    uint32_t b_tap = b_desc_lo | (((b_res0 + (uint32_t)blk0 * (block_step << 4)) >> 4) & 0x3fffu);
                    uint32_t a_tap = a_lo0;
                    uint32_t acc = accumulate;
#pragma unroll 1
                    for (int j = 0; j < (skip_mma ? 0 : j0); ++j) {
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            issue_kstep<MODE, CL>(d_tmem, a_desc_hi, b_desc_hi, a_tap + 2u * ks, b_tap + kstep_step * ks, lo_step, lo_rows_step,
                                              idesc, idesc_stack, acc);
                            acc = 1;
                        }
                        a_tap += tap_step;
                        b_tap += block_step;
                    }
                    if (j0 == K) {
                        commit_(&mc.empty_a[a_s]);
                        if (last_c) commit_(&mc.acc_full[acc_s]);
                    }
                }
                __syncwarp();
                accumulate = 1;
            }
            for (int j = j0; j < K; ++j) {
                for (int kb = 0; kb < nsub; ++kb) {
                    mbar_wait(&mc.full_b[b_s], b_ph);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t b_k = b_desc_lo | (((b_ring0 + b_s * b_sub_bytes) >> 4) & 0x3fffu);
                        uint32_t a_k = a_lo0 + (uint32_t)j * tap_step + (uint32_t)(kb * kpb) * 2u;
                        uint32_t acc = accumulate;
                        if (kpb == KSTEPS) {
#pragma unroll
                            for (int ks = 0; ks < KSTEPS; ++ks) {
                                issue_kstep<MODE, CL>(d_tmem, a_desc_hi, b_desc_hi, a_k + 2u * ks, b_k + kstep_step * ks, lo_step, lo_rows_step,
                                                  idesc, idesc_stack, acc);
                                acc = 1;
                            }
                        } else {
                            for (int ks = 0; ks < kpb; ++ks) {
                                issue_kstep<MODE, CL>(d_tmem, a_desc_hi, b_desc_hi, a_k, b_k, lo_step, lo_rows_step, idesc, idesc_stack, acc);
                                acc = 1;
                                a_k += 2u;
                                b_k += kstep_step;
                            }
                        }
                        commit_(&mc.empty_b[b_s]);
                        if (j + 1 == K && kb + 1 == nsub) {
                            commit_(&mc.empty_a[a_s]);
                            if (last_c) commit_(&mc.acc_full[acc_s]);
                        }
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++b_s == n_b_stages) { b_s = 0; b_ph ^= 1; }
                }
            }
            if (++a_s == n_a_stages) { a_s = 0; a_ph ^= 1; }
        }
        if (++acc_s == n_acc) { acc_s = 0; acc_ph ^= 1; }
        nt += nt_adv;
        if (nt >= n_tiles_n) nt -= n_tiles_n;
    }
}

// ------------------------------------------------------------------------------------------------------------------------
template <bool RES, bool ACC, bool CL>
__global__ void __launch_bounds__(TC_MAX_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap omap,
               const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint64_t* full_a = bars;
    uint64_t* empty_a = bars + TC_MAX_STAGES;
    uint64_t* full_b = bars + 2 * TC_MAX_STAGES;
    uint64_t* empty_b = bars + 3 * TC_MAX_STAGES;
    uint64_t* acc_full = bars + 4 * TC_MAX_STAGES;
    uint64_t* acc_empty = acc_full + TC_MAX_ACC;
    uint64_t* b_res_full = acc_full + 2 * TC_MAX_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2 * TC_MAX_ACC + 2);
    uint64_t* peer_a = acc_full + 2 * TC_MAX_ACC + 4;        // leader CTA of a pair: the peer's A tile / weight half has landed
    uint64_t* peer_b = peer_a + TC_MAX_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sm_base = smem_u32(sm);
    // CTA pair (cta_group::2): rank 0 issues the MMAs for both CTAs; tiles are handed out per pair
    const uint32_t rank = CL ? cluster_ctarank() : 0u;
    const int cta0 = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nctas = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int ROWS_PER_TILE = CL ? 2 * TC_M : TC_M;
    const bool EPI_SPLIT = p.epi_split != 0;        // column-split epilogue groups (see TcParams::epi_split)

    if (threadIdx.x == 0) {
        prefetch_tmap(&xmap);
        if (CL) prefetch_tmap(&wmap);
        if (p.tma_out) prefetch_tmap(&omap);
        for (int s = 0; s < TC_MAX_STAGES; ++s) {
            mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1);
            mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1);
            mbar_init(&peer_a[s], 1); mbar_init(&peer_b[s], 1);
        }
        // accumulator release: 4 epilogue warps per CTA; the leader of a pair also collects the peer's 4
        for (int s = 0; s < TC_MAX_ACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], (CL ? 8 : 4) * (EPI_SPLIT ? p.n_groups : 1)); }
        mbar_init(b_res_full, 1);
        fence_mbar_init();
    }
    {   // bias -> shared memory (zeros when the layer has none)
        float* bias_w = reinterpret_cast<float*>(sm + p.smem_bias_off);
        for (int i = threadIdx.x; i < p.out_C; i += blockDim.x) bias_w[i] = p.bias ? p.bias[i] : 0.f;
    }
    // programmatic dependent launch: the next kernel of the stream may become resident and run its prologue (barriers, TMEM,
    // weights) while this grid drains; it touches activations only after its own griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 2) { if (CL) tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); }
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_b = p.tiles_t * p.n_tiles_n;

    // Producer and MMA warps run their loops warp-uniformly (all 32 lanes), so that addresses, descriptors and barrier
    // phases live in uniform registers; only the asynchronous instructions themselves are issued by one elected lane.
    // (Measured: running the loops inside `if (lane == 0)` costs ~25 dependent R2UR/ELECT instructions per tcgen05.mma
    // and left the tensor pipe 57-80 % idle.)
    if (warp == 0) {
        // =========================== producer ===========================
        if (p.res_blocks > 0) {
            if (elect_one()) {
                if (CL) {
                    // CTA pair with RESIDENT weights: each CTA keeps its half of every block ([rank 0 half][rank 1 half] per block in
                    // the packed operand) for the whole launch; both CTAs' loads complete on the LEADER's barrier (the MMA issuer
                    // lives there and reads B from both CTAs' shared memory at the same offsets)
                    if (rank == 0) mbar_expect_tx(b_res_full, (uint32_t)(2 * p.res_blocks * p.b_sub_bytes));
                    const uint32_t lb = cluster_bar_addr(b_res_full, 0);
                    for (int blk = 0; blk < p.res_blocks; ++blk) {
                        const size_t off = (size_t)blk * p.b_block_bytes + (size_t)rank * p.b_sub_bytes;
                        tma_load_2d_2cta(sm + p.smem_b_off + (size_t)blk * p.b_sub_bytes, &wmap, 0, (int)(off >> 8), lb);
                    }
                } else {
                    const int rbytes = p.res_blocks * p.b_block_bytes;
                    mbar_expect_tx(b_res_full, (uint32_t)rbytes);
                    for (int off = 0; off < rbytes; off += 32768) {
                        const int n = rbytes - off < 32768 ? rbytes - off : 32768;
                        bulk_load_1d(sm + p.smem_b_off + off, p.w + off, (uint32_t)n, b_res_full);
                    }
                }
            }
            __syncwarp();
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");   // activations are written by the previous kernel of the stream
        uint32_t a_s = 0, a_ph = 0, b_s = 0, b_ph = 0;      // ring slot + phase counters (no divisions in the loops)
        const int nsub = p.ksteps / p.kpb;
        int b = cta0 / tiles_per_b, rem = cta0 - b * tiles_per_b;
        for (int tile = cta0; tile < p.total_tiles; tile += nctas) {
            int tt = rem, nt = 0;
            if (p.n_tiles_n != 1) { tt = rem / p.n_tiles_n; nt = rem - tt * p.n_tiles_n; }
            const int row0 = tt * ROWS_PER_TILE + (int)rank * TC_M - p.pad;
            for (int c = 0; c < p.nchunks; ++c) {
                mbar_wait(&empty_a[a_s], a_ph ^ 1);
                if (elect_one()) {
                    if ((p.ablate & 8)) {
                        mbar_arrive(&full_a[a_s]);
                    } else if (CL) {
                        // both CTAs load their own 128 rows; both signal the LEADER's barrier, which expects both halves
                        if (rank == 0) mbar_expect_tx(&full_a[a_s], (uint32_t)(2 * p.a_loads * p.a_tile_bytes));
                        const uint32_t lb = cluster_bar_addr(&full_a[a_s], 0);
                        uint8_t* dst = sm + p.smem_a_off + a_s * p.a_stage_bytes;
                        tma_load_3d_2cta(dst, &xmap, c * p.cin_t, row0, b, lb);
                        if (p.a_loads == 2) tma_load_3d_2cta(dst + p.a_tile_bytes, &xmap, p.Cin + c * p.cin_t, row0, b, lb);
                    } else {
                        mbar_expect_tx(&full_a[a_s], (uint32_t)(p.a_loads * p.a_tile_bytes));
                        uint8_t* dst = sm + p.smem_a_off + a_s * p.a_stage_bytes;
                        tma_load_3d(dst, &xmap, c * p.cin_t, row0, b, &full_a[a_s]);
                        if (p.a_loads == 2) tma_load_3d(dst + p.a_tile_bytes, &xmap, p.Cin + c * p.cin_t, row0, b, &full_a[a_s]);
                    }
                }
                __syncwarp();
                if (++a_s == (uint32_t)p.n_a_stages) { a_s = 0; a_ph ^= 1; }
                const int blk0 = (nt * p.nchunks + c) * p.K;
                int j0 = p.res_blocks - blk0;                 // taps below j0 are resident
                j0 = j0 < 0 ? 0 : (j0 > p.K ? p.K : j0);
                // (CTA pair: a sub-block is stored as [rank 0 half][rank 1 half]; b_sub_bytes is one half)
                const uint8_t* wsrc = p.w + (size_t)(blk0 + j0) * p.b_block_bytes + (CL ? (size_t)rank * p.b_sub_bytes : 0);
                const size_t sub_stride = CL ? 2 * (size_t)p.b_sub_bytes : (size_t)p.b_sub_bytes;
                for (int jb = 0; jb < (p.K - j0) * nsub; ++jb) {
                    mbar_wait(&empty_b[b_s], b_ph ^ 1);
                    if (elect_one()) {
                        if (p.ablate & 64) {
                            if (rank == 0) mbar_arrive(&full_b[b_s]);
                        } else if (CL) {
                            // weights through a 2-D map over the packed bytes ([rows][256 B]) so that the load can signal the leader
                            if (rank == 0) mbar_expect_tx(&full_b[b_s], (uint32_t)(2 * p.b_sub_bytes));
                            const size_t off = (size_t)(wsrc - p.w) + (size_t)jb * sub_stride;
                            tma_load_2d_2cta(sm + p.smem_ring_off + b_s * p.b_sub_bytes, &wmap, 0, (int)(off >> 8), cluster_bar_addr(&full_b[b_s], 0));
                        } else {
                            mbar_expect_tx(&full_b[b_s], (uint32_t)p.b_sub_bytes);
                            bulk_load_1d(sm + p.smem_ring_off + b_s * p.b_sub_bytes, wsrc + (size_t)jb * sub_stride, (uint32_t)p.b_sub_bytes,
                                         &full_b[b_s]);
                        }
                    }
                    __syncwarp();
                    if (++b_s == (uint32_t)p.n_b_stages) { b_s = 0; b_ph ^= 1; }
                }
            }
            rem += nctas;
            while (rem >= tiles_per_b) { rem -= tiles_per_b; ++b; }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        MmaCtx mc{full_a, empty_a, full_b, empty_b, acc_full, acc_empty, b_res_full, peer_a, peer_b, sm_base, tmem_base, cta0, nctas};
        if (CL) {
            if (rank == 0) {
                mma_issuer<4, 1, true>(p, mc);
            }      // the peer CTA's warp 1 has nothing to do: its loads signal the leader's barriers directly
        } else {
            switch (p.ksteps * 4 + (p.single_pass ? 0 : (p.stack ? 2 : 1))) {
                case 1 * 4 + 2: mma_issuer<1, 2>(p, mc); break;
                case 2 * 4 + 2: mma_issuer<2, 2>(p, mc); break;
                case 4 * 4 + 2: mma_issuer<4, 2>(p, mc); break;
                case 4 * 4 + 1: mma_issuer<4, 1>(p, mc); break;
                case 1 * 4 + 0: mma_issuer<1, 0>(p, mc); break;
                case 2 * 4 + 0: mma_issuer<2, 0>(p, mc); break;
                default: mma_issuer<4, 0>(p, mc); break;
            }
        }
    }
    else if (warp >= 4) {
        // =========================== epilogue ===========================
        // Work items are (tile, 16-column chunk) in order; the residual / running-sum rows of item k+1 are requested as soon
        // as item k has consumed its own, so their latency hides behind the rest of item k and the next accumulator read.
        // Everything that costs an integer division is done once per tile; bias comes from shared memory.
        asm volatile("griddepcontrol.wait;" ::: "memory");   // residual / running-sum reads and the output stores depend on the previous kernel
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = (warp - 4) >> 2;              // epilogue group: takes this CTA's tiles grp, grp + n_groups, ...
        const int tile_stride = EPI_SPLIT ? nctas : nctas * p.n_groups;
        const int n0_first = EPI_SPLIT ? 16 * grp : 0;            // (column split: items grp, grp + n_groups, ... of every tile)
        const int n0_step = EPI_SPLIT ? 16 * p.n_groups : 16;
        const int row = q * 32 + lane;
        const int Lout = p.L * p.U;
        const size_t out_pitch = (size_t)2 * p.out_C;
        const float* bias_s = reinterpret_cast<const float*>(sm + p.smem_bias_off);

        // Tile coordinates advance incrementally by gridDim.x tiles (no division per tile); the valid length of a batch
        // item is read from global memory only when the item changes.
        struct TileCtx { int b, rem, nt, orow0, valid_rows; size_t row_el, row_idx; bool ok, row_ok; };
        const int n_tiles_n = p.n_tiles_n;
        auto valid_of = [&](int b) {
            int valid = Lout;
            if (p.lens) { const int v = __ldg(p.lens + b) * p.len_mul; valid = v < valid ? v : valid; }
            return valid;
        };
        auto place = [&](TileCtx& c) {       // (b, rem) -> row coordinates
            int tt = c.rem;
            c.nt = 0;
            if (n_tiles_n != 1) { tt = c.rem / n_tiles_n; c.nt = c.rem - tt * n_tiles_n; }
            const int t = tt * ROWS_PER_TILE + (int)rank * TC_M + row;
            c.row_ok = t < p.L;
            c.orow0 = t * p.U;
            c.row_idx = (size_t)c.b * Lout + c.orow0;
            c.row_el = c.row_idx * out_pitch;
        };
        auto locate_first = [&](int tile, TileCtx& c) {
            c.ok = tile < p.total_tiles;
            if (!c.ok) return;
            c.b = tile / tiles_per_b;
            c.rem = tile - c.b * tiles_per_b;
            c.valid_rows = valid_of(c.b);
            place(c);
        };
        auto locate_next = [&](int tile, TileCtx& c) {      // c holds the previous tile of this CTA
            c.ok = tile < p.total_tiles;
            if (!c.ok) return;
            c.rem += tile_stride;
            if (c.rem >= tiles_per_b) {
                do { c.rem -= tiles_per_b; ++c.b; } while (c.rem >= tiles_per_b);
                c.valid_rows = valid_of(c.b);
            }
            place(c);
        };
        // element offset of the item's 16 hi values, its first output channel, and whether the row is inside the valid length
        auto item = [&](const TileCtx& c, int n0, size_t& off, int& co, bool& live) {
            const int ng = c.nt * p.NT + n0;
            if (p.U == 1) {
                co = ng; off = c.row_el + ng; live = c.orow0 < c.valid_rows;
            } else {
                const int phs = ng / p.out_C;
                co = ng - phs * p.out_C;
                off = c.row_el + (size_t)phs * out_pitch + co;
                live = c.orow0 + phs < c.valid_rows;
            }
        };
        auto request = [&](bool ok, size_t off, uint4 (&r)[4], uint4 (&a)[4]) {
            if (!ok || (p.ablate & 4)) return;
            if (RES) {
                ld_global_nc_v8(p.res + off, r[0], r[1]);
                ld_global_nc_v8(p.res + off + p.out_C, r[2], r[3]);
            }
            if (ACC) {   // may alias `out` of an earlier item of this same thread: plain loads
                ld_global_v8(p.accin + off, a[0], a[1]);
                ld_global_v8(p.accin + off + p.out_C, a[2], a[3]);
            }
        };

        int tile = EPI_SPLIT ? cta0 : cta0 + grp * nctas, n0 = n0_first;
        uint32_t acc_i = EPI_SPLIT ? 0u : (uint32_t)grp;  // local tile counter: this group sees i = grp + k*n_groups
        const uint32_t acc_mask = (uint32_t)p.n_acc - 1u, acc_shift = (uint32_t)__ffs(p.n_acc) - 1u;   // n_acc is a power of two
        TileCtx cur;
        locate_first(tile, cur);
        size_t off = 0; int co = 0; bool live = false;
        if (cur.ok) item(cur, n0_first, off, co, live);
        uint4 R[4], A[4];
        request(cur.ok && cur.row_ok, off, R, A);
        float best_v = -INFINITY;      // fused argmax: this thread's running maximum over its items of the current tile
        int best_c = 0;
        while (cur.ok) {
            // next item
            TileCtx nxt = cur;
            int tile_n = tile, n0n = n0 + n0_step;
            if (n0n >= p.NT) {
                n0n = n0_first; tile_n += tile_stride; locate_next(tile_n, nxt);
            }
            size_t offn = 0; int con = 0; bool liven = false;
            if (nxt.ok) item(nxt, n0n, offn, con, liven);

            if ((RES || ACC) && n0 == n0_first && p.U == 1 && (!EPI_SPLIT || grp == 0)) {
                // residual / running-sum rows of this group's NEXT tile: pull this N tile's columns into L2 now, a whole
                // tile period ahead of their use (the loads then see L2 latency, not DRAM latency under load)
                TileCtx pf = cur;
                locate_next(tile + tile_stride, pf);
                if (pf.ok && pf.row_ok) {
                    const int part_bytes = 2 * p.NT;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        const size_t rb = (pf.row_el + (size_t)h * p.out_C + (size_t)pf.nt * p.NT) * sizeof(__nv_bfloat16);
                        for (int o = 0; o < part_bytes; o += 128) {
                            if (RES) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p.res) + rb + o));
                            if (ACC) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p.accin) + rb + o));
                        }
                    }
                }
            }
            const uint32_t as = acc_i & acc_mask;
            if (n0 == n0_first) {
                mbar_wait(&acc_full[as], (acc_i >> acc_shift) & 1u);
                tc_fence_after();
            }
            float4 bv4[4];                                // bias of this item, fetched while the accumulator read is in flight
            {
                const float4* bp = reinterpret_cast<const float4*>(bias_s + co);
#pragma unroll
                for (int i = 0; i < 4; ++i) bv4[i] = bp[i];
            }
            float v[16];
            {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.acc_cols + (uint32_t)n0;
                if ((p.ablate & 16)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                } else if (p.stack) tmem_ld16x2_sum(taddr, taddr + (uint32_t)p.NT, v);
                else tmem_ld16(taddr, v);
            }
            if (n0n == n0_first) {                     // accumulator fully read: hand it back before the arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                              // one arrival per epilogue warp, at the MMA issuer's CTA
                    if (CL && rank != 0) mbar_arrive_remote(&acc_empty[as], 0); else mbar_arrive(&acc_empty[as]);
                }
            }
            if (cur.row_ok && !(p.ablate & 32)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bv = bv4[i];
                    add2(v[4 * i], v[4 * i + 1], bv.x, bv.y);
                    add2(v[4 * i + 2], v[4 * i + 3], bv.z, bv.w);
                }
                if (RES) {
                    const uint32_t hh[8] = {R[0].x, R[0].y, R[0].z, R[0].w, R[1].x, R[1].y, R[1].z, R[1].w};
                    const uint32_t ll[8] = {R[2].x, R[2].y, R[2].z, R[2].w, R[3].x, R[3].y, R[3].z, R[3].w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 r = join2(hh[i], ll[i], p.fmt);
                        float tx = r.x, ty = r.y;
                        mul2(tx, ty, p.res_unslope, p.res_unslope);
                        add2(v[2 * i], v[2 * i + 1], fminf(r.x, tx), fminf(r.y, ty));      // undo the stored leaky relu (1/slope >= 1)
                    }
                }
                if (ACC) {
                    const uint32_t hh[8] = {A[0].x, A[0].y, A[0].z, A[0].w, A[1].x, A[1].y, A[1].z, A[1].w};
                    const uint32_t ll[8] = {A[2].x, A[2].y, A[2].z, A[2].w, A[3].x, A[3].y, A[3].z, A[3].w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 r = join2(hh[i], ll[i], p.fmt);
                        add2(v[2 * i], v[2 * i + 1], r.x, r.y);
                    }
                }
            }
            // R / A are dead from here: the next item's residual / running-sum rows are requested into the SAME registers (no
            // register-to-register hand-over at the loop edge - that copy made every item wait for its loads right away) and
            // have the rest of this item plus the next accumulator read to arrive
            request(nxt.ok && nxt.row_ok, offn, R, A);
            uint32_t hw[8], lw[8];
            if (cur.row_ok && !(p.ablate & 32)) {
                if (p.res_f32) {      // fp32 residual stream [B][Lout][out_C] (TTE)
                    const float4* rp = reinterpret_cast<const float4*>(p.res_f32 + (off - (size_t)co) / 2 + co);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 rv = __ldg(rp + i);
                        v[4 * i] += rv.x; v[4 * i + 1] += rv.y; v[4 * i + 2] += rv.z; v[4 * i + 3] += rv.w;
                    }
                }
                const float sc = live ? p.scale : 0.f;      // rows past the valid length are written as zeros
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    mul2(v[2 * i], v[2 * i + 1], sc, sc);
                    float tx = v[2 * i], ty = v[2 * i + 1];
                    mul2(tx, ty, p.post_slope, p.post_slope);
                    v[2 * i] = fmaxf(v[2 * i], tx);         // leaky relu for slope in [0,1]; slope 1 = identity
                    v[2 * i + 1] = fmaxf(v[2 * i + 1], ty);
                }
                if (p.tanh_out) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
                }
                if (!RES && !ACC && p.argmax_key) {      // (plain GEMM instantiations only: keeps the residual variants' registers)
                    // argmax over the row as an epilogue (Parrot.infer's torch.argmax, parrot.py:115): the logits never reach HBM.  A
                    // thread keeps the first maximum of its own columns (visited in increasing order) and publishes it once per tile
                    // with a 64-bit atomicMax on (order-preserving bits of the value, ~column): ties resolve to the lowest column.
                    if (n0 == n0_first) { best_v = -INFINITY; best_c = 0; }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x = v[i] + 0.f;                    // -0 -> +0: equal under float comparison, equal keys
                        if (co + i < p.argmax_n && x > best_v) { best_v = x; best_c = co + i; }
                    }
                    if (n0n == n0_first) {
                        const uint32_t bits = __float_as_uint(best_v);
                        const uint32_t ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
                        atomicMax(p.argmax_key + cur.row_idx, ((unsigned long long)ord << 32) | (unsigned long long)(0xffffffffu - (uint32_t)best_c));
                    }
                }
                if (p.out && !(p.ablate & 2)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], p.fmt, hw[i], lw[i]);
                    if (!p.tma_out) {
                        __nv_bfloat16* oh = p.out + off;
                        __nv_bfloat16* ol = oh + p.out_C;
                        st_global_v8(oh, make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]));
                        st_global_v8(ol, make_uint4(lw[0], lw[1], lw[2], lw[3]), make_uint4(lw[4], lw[5], lw[6], lw[7]));
                    }
                }
                if (p.out_f32) {
                    float* of = p.out_f32 + (off - (size_t)co) / 2 + co;     // [B][Lout][out_C]: half the split row pitch
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(of + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            if (p.tma_out && !(p.ablate & 2)) {
                // Row-per-thread 32-byte global stores cost ~50 L1 wavefronts per instruction (every lane its own line); instead the
                // warp parks its [32 rows x 16 ch] hi and lo tiles in shared memory (32B-swizzled: conflict-free 16-byte stores) and
                // one lane hands them to the TMA unit.  Rows past the end of the sequence are clipped by the tensor map.
                uint8_t* sbw = sm + p.smem_stage_off + (uint32_t)(warp - 4) * 2048u;
                if (lane == 0) bulk_wait_group_read<0>();          // the previous item's stores have read this buffer
                __syncwarp();
                const uint32_t sx = ((uint32_t)lane >> 2) & 1u;     // 32B swizzle: 16-byte chunk ^= address bit 7
                uint8_t* rp = sbw + lane * 32;
                *reinterpret_cast<uint4*>(rp + (sx << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(rp + ((sx ^ 1u) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                *reinterpret_cast<uint4*>(rp + 1024 + (sx << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                *reinterpret_cast<uint4*>(rp + 1024 + ((sx ^ 1u) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                fence_proxy_async();
                __syncwarp();
                const int t0 = cur.orow0 - lane;                    // first row of this warp's 32 (U == 1)
                if (lane == 0 && t0 < p.L) {
                    tma_store_3d(&omap, sbw, co, t0, cur.b);
                    tma_store_3d(&omap, sbw + 1024, p.out_C + co, t0, cur.b);
                    bulk_commit_group();
                }
            }
            if (n0n == n0_first) acc_i += EPI_SPLIT ? 1u : (uint32_t)p.n_groups;
            cur = nxt; tile = tile_n; n0 = n0n; off = offn; co = con; live = liven;
        }
    }

    if (p.tma_out && warp >= 4 && lane == 0) bulk_wait_group<0>();     // this lane's TMA stores have completed
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();          // no CTA of the pair exits (or frees TMEM) while the other may still signal it
    if (warp == 2) {
        tc_fence_after();
        if (CL) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------
// N tile: 128 wherever N allows it.  With the stacked [hi; lo] operand a 128-wide tile costs 128 + 64 tensor cycles per
// k-step - exactly half of the 3 x 128 of a 256-wide tile - and twice as many tiles fill the 148 SMs' last wave better
// (320 tiles of 128 x 256 at C = 256, B x L = 8 x 5120, are 2.16 waves; 640 tiles are 4.3).
int tc_n_tile(int N) {
    if (N <= 0 || N % 16) return 0;
    if (N <= 128) return N;
    if (N % 128 == 0) return 128;
    if (N <= 256) return N;
    return 0;
}

static int tc_cin_tile(int Cin) {
    if (Cin >= 64) return Cin % 64 == 0 ? 64 : 0;
    if (Cin == 32 || Cin == 16) return Cin;
    return 0;
}

bool tc_use_cluster(int Cin, int N, int K, int pair64);
// split 16-bit hi + lo = 4 bytes per weight; the CTA-pair layout stores the hi rows a second time (operand R2): 6 bytes
size_t tc_packed_weight_bytes(int Cin, int N, int K, int pair64) { return (size_t)Cin * N * K * (tc_use_cluster(Cin, N, K, pair64) ? 6 : 4); }

// CTA pairs (cta_group::2) for the wide layers whose weights cannot stay resident: M = 256 per pair, each CTA streams (and
// reads) half of the weight rows - halves the L2 -> SM weight traffic and the B-operand shared-memory traffic that bound
// the 128 / 256-channel layers.  A pure function of the layer shape: the packed weight layout depends on it.
// Both CTAs' TMA loads signal the LEADER's mbarriers directly (cp.async.bulk.tensor ... cta_group::2) - a first version that
// relayed "my half has landed" through the peer's idle warp left the leader waiting on that hop 40 % of the time and was
// 25 % slower than the single-CTA path; with direct signalling the pair path is ~8 % faster on the 256 / 128-channel stages.
// PB_TC_CLUSTER=0 selects the single-CTA path (same results).
bool tc_use_cluster(int Cin, int N, int K, int pair64) {
    static const bool on = !(getenv("PB_TC_CLUSTER") != nullptr && atoi(getenv("PB_TC_CLUSTER")) == 0);
    if (!on) return false;
    if (Cin % 64 != 0) return false;
    if (tc_n_tile(N) == 128 && (size_t)Cin * N * K * 4 > 160 * 1024) return true;
    // 64 -> 64 channels with 7 or more taps (vocoder stage 3, k = 7 / 11), when the caller asks for it (pair64): as a pair each CTA
    // keeps its half of the weights RESIDENT (66 / 103 KB; one CTA per SM streams 6 of 11 taps per tile at k = 11) and one issuing
    // thread drives both SMs' tensor cores.  Measured per launch at 8 x 1024 units (profiles/r02_pair64.txt): conv1 k = 7 98 -> 84 us,
    // k = 11 162 -> 123 us; the residual conv2 at k = 11 169 -> 134 us; but the k = 7 residual launches and the ones that also add
    // the running sum are better off as one CTA per SM with the TMA-store epilogue - hence a per-layer choice, not a shape rule.
    static const bool pair64_on = !(getenv("PB_TC_PAIR64") != nullptr && atoi(getenv("PB_TC_PAIR64")) == 0);
    return pair64_on && pair64 != 0 && Cin == 64 && N == 64 && K >= 7;      // (k = 3 as a pair: 66 -> 74 us conv1, 105 -> 127 us conv2)
}

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
static int make_tmap_act(CUtensorMap* map, const __nv_bfloat16* base, int C, int L, int B, int box_c, int rows, bool sw128) {   // 16-bit elements; the copy is type-agnostic
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

int tc_make_tmap_act(CUtensorMap* map, const __nv_bfloat16* base, int C, int L, int B, int box_c, int rows, bool sw128) {
    return make_tmap_act(map, base, C, L, B, box_c, rows, sw128);
}

// 3-D map over a split output [B][L][2C] for the TMA-store epilogue: box (16 ch, rows <= 32, 1), 32B swizzle
int tc_make_tmap_out(CUtensorMap* map, __nv_bfloat16* base, int C, int L, int B, int rows) {
    EncodeTiledFn fn = encode_fn();
    PB_REQUIRE(fn != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    PB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, PB_ERR_INVALID, "tc_conv: output base must be 16-byte aligned");
    PB_REQUIRE(rows >= 1 && rows <= 32, PB_ERR_INVALID, "tc_conv: output box of %d rows", rows);
    cuuint64_t gdim[3] = {(cuuint64_t)2 * C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)4 * C, (cuuint64_t)4 * C * L};
    cuuint32_t box[3] = {16, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed (%d): C=%d L=%d B=%d rows=%d", (int)r, C, L, B, rows);
    return PB_OK;
}

// 4-D map over a split 16-channel tensor [B][L][32] seen as [B][L / 4 groups][4 positions][32] (tc_pair16_kernel's epilogue B: a
// box = one position of `rows` consecutive groups = `rows` 64-byte rows, 64B swizzle; L must be a multiple of 4)
int tc_make_tmap_grouped16(CUtensorMap* map, const __nv_bfloat16* base, int L, int B, int rows) {
    EncodeTiledFn fn = encode_fn();
    PB_REQUIRE(fn != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    PB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && L % 4 == 0, PB_ERR_INVALID, "tc_pair16: base alignment / L=%d not a multiple of 4", L);
    PB_REQUIRE(rows >= 1 && rows <= 256, PB_ERR_INVALID, "tc_pair16: box of %d rows", rows);
    cuuint64_t gdim[4] = {32, 4, (cuuint64_t)(L / 4), (cuuint64_t)B};
    cuuint64_t gstr[3] = {64, 256, (cuuint64_t)64 * L};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled(grouped16) failed (%d): L=%d B=%d rows=%d", (int)r, L, B, rows);
    return PB_OK;
}

int launch_tc_conv(const TcConvArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.x && a.w && (a.out || a.out_f32 || a.argmax_key), PB_ERR_INVALID, "tc_conv: null tensor");
    PB_REQUIRE(!a.argmax_key || (a.U == 1 && a.argmax_n > 0 && a.argmax_n <= a.N && !a.res && !a.accin), PB_ERR_INVALID,
               "tc_conv: fused argmax needs U = 1, 0 < argmax_n <= N and no residual / running sum");
    PB_REQUIRE(a.B > 0 && a.L > 0 && a.K >= 1 && a.D >= 1 && a.U >= 1, PB_ERR_INVALID, "tc_conv: bad shape");
    const int cin_t = tc_cin_tile(a.Cin);
    const int NT = tc_n_tile(a.N);
    PB_REQUIRE(cin_t > 0, PB_ERR_INVALID, "tc_conv: Cin=%d unsupported (16, 32 or a multiple of 64)", a.Cin);
    PB_REQUIRE(NT > 0, PB_ERR_INVALID, "tc_conv: N=%d unsupported", a.N);
    PB_REQUIRE(a.out_C > 0 && a.out_C % 16 == 0 && a.N == a.U * a.out_C, PB_ERR_INVALID, "tc_conv: N=%d != U=%d * out_C=%d", a.N, a.U, a.out_C);
    PB_REQUIRE((a.K - 1) * a.D % 2 == 0, PB_ERR_INVALID, "tc_conv: (K-1)*D must be even ('same' padding)");
    TcParams p{};
    p.w = a.w; p.bias = a.bias; p.res = a.res; p.res_f32 = a.res_f32; p.accin = a.accin; p.out = a.out; p.out_f32 = a.out_f32; p.lens = a.lens;
    p.argmax_key = a.argmax_key; p.argmax_n = a.argmax_n;
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
    p.cluster = tc_use_cluster(a.Cin, a.N, a.K, a.pair64) ? 1 : 0;      // (always the 3-product form: the packed layout is the pair layout)
    p.b_block_bytes = 2 * p.ksteps * NT * 32;           // one (n tile, channel chunk, tap): ksteps x {hi, lo} planes
    p.kpb = 32768 / (2 * NT * 32);                       // k-steps per streamed sub-block (<= 32 KB: 6+ MMAs per barrier round trip)
    if (p.kpb < 1) p.kpb = 1;
    if (p.kpb > p.ksteps) p.kpb = p.ksteps;
    while (p.ksteps % p.kpb) --p.kpb;
    p.b_sub_bytes = p.kpb * 2 * NT * 32;
    if (p.cluster) {      // one block per ring slot; a block is [rank 0: ksteps x (R1 + R2)][rank 1: ...], R1 + R2 = NT*48 bytes
        p.kpb = p.ksteps;
        p.b_sub_bytes = p.ksteps * NT * 48;
        p.b_block_bytes = 2 * p.b_sub_bytes;
    }
    p.total_blocks = p.n_tiles_n * p.nchunks * a.K;
    p.w_total_bytes = p.total_blocks * p.b_block_bytes;
    p.smem_bias_off = 1024;
    p.smem_a_off = 1024 + (uint32_t)round_up((int64_t)a.out_C * 4, 1024);
    // Shared-memory plan.  Weight blocks are [n tile][chunk][tap] in the packed operand; the first `res_blocks` of them
    // stay resident for the whole launch, the rest stream through a ring of sub-blocks once per tile.
    //  - everything resident + >= 3 activation stages in HALF an SM: two CTAs per SM (narrow layers are latency bound per tile)
    //  - everything resident + >= 2 stages in a whole SM: one CTA per SM
    //  - otherwise 2 activation stages, a 3-deep ring, and as many resident blocks as still fit - kept only when they
    //    cover at least a quarter of the operand (else the space buys more pipeline depth instead)
    int ctas_per_sm = 1;
    uint32_t b_res_bytes = 0, ring_bytes = 0;
    const uint32_t w_all = (uint32_t)round_up(p.w_total_bytes, 1024);
    // TMA-store epilogue (U == 1, split output): 2 KB of staging per epilogue warp, up to 3 groups of 4 warps.
    // PB_TC_TMA_OUT: 0 = never, 1 = layers that run one CTA per SM without CTA pairs (default), 2 = every eligible layer (dev switch;
    // measured +1-2 % on the CTA-pair layers: the staging tiles cost a ring slot; restricted to their k <= 3 launches: no change)
    static const int tma_out_mode = getenv("PB_TC_TMA_OUT") ? atoi(getenv("PB_TC_TMA_OUT")) : 1;
    // two CTAs per SM (two producers, two MMA issuers, one epilogue group each) measured best for the 16-channel layers;
    // wider layers run one CTA per SM with up to 3 epilogue groups (PB_TC_2CTA=0/1 forces either, dev switch)
    static const int force_2cta = getenv("PB_TC_2CTA") ? atoi(getenv("PB_TC_2CTA")) : -1;
    const bool want_2cta = force_2cta >= 0 ? (force_2cta != 0 && NT <= 128) : NT <= 16;
    p.tma_out = (a.out != nullptr && a.U == 1 && tma_out_mode > 0 && !want_2cta &&
                 (tma_out_mode == 2 || !p.cluster || (NT == 64 && a.res && a.accin))) ? 1 : 0;
    uint32_t stage_bytes = p.tma_out ? (uint32_t)TC_MAX_GROUPS * 4u * 2048u : 0u;
    auto budget = [&](uint32_t limit) { return limit - 1024u - p.smem_a_off - stage_bytes; };   // barriers + bias + alignment slack
    if (p.tma_out && !p.cluster && budget(TC_SMEM_LIMIT) < w_all + 2u * (uint32_t)p.a_stage_bytes) {
        // the staging tiles would cost weight residency: every resident block counts (measured ~6 us per block on the 64-channel
        // k = 11 layers), so they are kept only where the weights are partly resident anyway AND the epilogue does three global
        // accesses per item (residual + running sum + store)
        const bool full_without = TC_SMEM_LIMIT - 1024u - p.smem_a_off >= w_all + 2u * (uint32_t)p.a_stage_bytes;
        if (full_without || !(a.res && a.accin)) { p.tma_out = 0; stage_bytes = 0; }
    }
    const uint32_t w_half_all = (uint32_t)round_up((int64_t)p.total_blocks * p.b_sub_bytes, 1024);      // (pairs: this CTA's half of every block)
    static const bool cl_resident = !(getenv("PB_TC_CL_RESIDENT") != nullptr && atoi(getenv("PB_TC_CL_RESIDENT")) == 0);
    if (p.cluster && cl_resident && budget(TC_SMEM_LIMIT) >= w_half_all + 2u * (uint32_t)p.a_stage_bytes) {
        // CTA pair, weights split across the pair and RESIDENT: nothing is streamed per tile (the 64-channel k = 7 / 11 layers,
        // the 128-channel k = 3 layers); the rest of the shared memory deepens the activation ring
        p.res_blocks = p.total_blocks; b_res_bytes = w_half_all;
        p.n_a_stages = (int)((budget(TC_SMEM_LIMIT) - w_half_all) / (uint32_t)p.a_stage_bytes);
        p.n_b_stages = 0;
    } else if (p.cluster) {
        const uint32_t bud = budget(TC_SMEM_LIMIT);
        p.res_blocks = 0;
        p.n_a_stages = 2;
        uint32_t left = bud - 2u * (uint32_t)p.a_stage_bytes;
        p.n_b_stages = (int)(left / (uint32_t)p.b_sub_bytes);
        if (p.n_b_stages > TC_MAX_STAGES) p.n_b_stages = TC_MAX_STAGES;
        left -= (uint32_t)p.n_b_stages * (uint32_t)p.b_sub_bytes;
        p.n_a_stages += (int)(left / (uint32_t)p.a_stage_bytes);
        PB_REQUIRE(p.n_b_stages >= 2, PB_ERR_INVALID, "tc_conv: tile does not fit in shared memory (Cin=%d N=%d K=%d D=%d)", a.Cin, a.N, a.K, a.D);
    } else if (want_2cta && budget(TC_SMEM_HALF) >= w_all + 3u * (uint32_t)p.a_stage_bytes) {
        ctas_per_sm = 2;
        p.res_blocks = p.total_blocks; b_res_bytes = w_all;
        p.n_a_stages = (int)((budget(TC_SMEM_HALF) - w_all) / (uint32_t)p.a_stage_bytes);
        p.n_b_stages = 0;
    } else if (budget(TC_SMEM_LIMIT) >= w_all + 2u * (uint32_t)p.a_stage_bytes) {
        p.res_blocks = p.total_blocks; b_res_bytes = w_all;
        p.n_a_stages = (int)((budget(TC_SMEM_LIMIT) - w_all) / (uint32_t)p.a_stage_bytes);
        p.n_b_stages = 0;
    } else {
        const uint32_t bud = budget(TC_SMEM_LIMIT);
        PB_REQUIRE(bud >= 2u * (uint32_t)p.a_stage_bytes + 2u * (uint32_t)p.b_sub_bytes, PB_ERR_INVALID,
                   "tc_conv: tile does not fit in shared memory (Cin=%d N=%d K=%d D=%d)", a.Cin, a.N, a.K, a.D);
        p.n_a_stages = 2;
        uint32_t left = bud - 2u * (uint32_t)p.a_stage_bytes;
        // ring depth vs resident blocks (PB_TC_RING, dev switch): measured on the 64-channel k = 11 layers, 2 / 3 / 4 / 6 slots ->
        // 161 / 161 / 165 / 179 us - the ring hides L2 latency with 2-3 slots already, every further slot costs a resident block
        static const int ring_max = getenv("PB_TC_RING") ? atoi(getenv("PB_TC_RING")) : 3;
        p.n_b_stages = (int)(left / (uint32_t)p.b_sub_bytes);
        if (p.n_b_stages > ring_max) p.n_b_stages = ring_max;
        left -= (uint32_t)p.n_b_stages * (uint32_t)p.b_sub_bytes;
        p.res_blocks = (int)(left / (uint32_t)p.b_block_bytes);
        if (p.res_blocks * 4 < p.total_blocks) {
            p.res_blocks = 0;
            left = bud - 2u * (uint32_t)p.a_stage_bytes;
            p.n_b_stages = (int)(left / (uint32_t)p.b_sub_bytes);
            if (p.n_b_stages > TC_MAX_STAGES) p.n_b_stages = TC_MAX_STAGES;
            left -= (uint32_t)p.n_b_stages * (uint32_t)p.b_sub_bytes;
            p.n_a_stages += (int)(left / (uint32_t)p.a_stage_bytes);
        }
        b_res_bytes = (uint32_t)round_up((int64_t)p.res_blocks * p.b_block_bytes, 1024);
    }
    if (p.n_a_stages > TC_MAX_STAGES) p.n_a_stages = TC_MAX_STAGES;
    ring_bytes = (uint32_t)p.n_b_stages * (uint32_t)p.b_sub_bytes;
    p.smem_b_off = p.smem_a_off + (uint32_t)p.n_a_stages * (uint32_t)p.a_stage_bytes;
    p.smem_ring_off = p.smem_b_off + b_res_bytes;
    p.smem_stage_off = p.smem_ring_off + ring_bytes;
    const uint32_t smem_bytes = p.smem_stage_off + stage_bytes + 1024;
    PB_REQUIRE(smem_bytes <= (ctas_per_sm == 2 ? TC_SMEM_HALF : TC_SMEM_LIMIT), PB_ERR_INVALID, "tc_conv: shared memory plan %u exceeds the limit", smem_bytes);
    p.tiles_t = (int)ceil_div(a.L, p.cluster ? 2 * TC_M : TC_M);
    const int64_t total = (int64_t)a.B * p.tiles_t * p.n_tiles_n;
    PB_REQUIRE(total < (1LL << 31), PB_ERR_INVALID, "tc_conv: too many tiles");
    p.total_tiles = (int)total;
    p.post_slope = a.post_slope; p.scale = a.scale; p.res_unslope = a.res_unslope; p.tanh_out = a.tanh_out;
    // instruction descriptor: D f32 (bit 4), A / B format at bits 7 / 10 (0 = f16, 1 = bf16), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t ab_fmt = a.fmt == 0 ? ((1u << 7) | (1u << 10)) : 0u;
    p.idesc = (1u << 4) | ab_fmt | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    p.idesc_stack = (1u << 4) | ab_fmt | ((uint32_t)((2 * NT) >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    p.fmt = a.fmt;
    p.single_pass = a.single_pass;
    p.stack = ((!a.single_pass && NT <= 128) || p.cluster) ? 1 : 0;
    if (p.cluster) {      // M = 256 over the pair
        p.idesc = (1u << 4) | ab_fmt | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((2 * TC_M) >> 4) << 24);
        p.idesc_stack = (1u << 4) | ab_fmt | ((uint32_t)((2 * NT) >> 3) << 17) | ((uint32_t)((2 * TC_M) >> 4) << 24);
    }
    p.acc_cols = p.stack ? 2 * NT : NT;
    p.n_acc = (512 / ctas_per_sm) / p.acc_cols;
    if (p.n_acc > TC_MAX_ACC) p.n_acc = TC_MAX_ACC;
    while (p.n_acc & (p.n_acc - 1)) --p.n_acc;        // power of two
    // The epilogue of a narrow tile is a short latency-bound chain (accumulator wait, TMEM load, arithmetic, stores):
    // several warpgroups take alternate tiles so that those chains overlap.
    static const int env_groups = getenv("PB_TC_GROUPS") ? atoi(getenv("PB_TC_GROUPS")) : 0;
    p.n_groups = a.n_groups > 0 ? a.n_groups : (env_groups > 0 ? env_groups : (ctas_per_sm == 2 ? 1 : TC_MAX_GROUPS));
    // PB_TC_EPI_SPLIT (dev switch): 0 = off, 1 = the CTA-pair layers (default), 2 = every one-CTA-per-SM layer
    static const int epi_split_env = getenv("PB_TC_EPI_SPLIT") ? atoi(getenv("PB_TC_EPI_SPLIT")) : 1;
    // (pairs with a 64-wide tile keep whole tiles per group: 4 items over 3 groups split 2 / 1 / 1 - measured 118 vs 84 us at k = 7)
    p.epi_split = (ctas_per_sm == 1 && (epi_split_env >= 2 || (epi_split_env == 1 && p.cluster && NT >= 128))) ? 1 : 0;
    if (p.epi_split) {
        if (p.n_groups > NT / 16) p.n_groups = NT / 16;       // every group needs an item in every tile
    } else if (p.n_groups > p.n_acc) p.n_groups = p.n_acc;
    if (p.n_groups > TC_MAX_GROUPS) p.n_groups = TC_MAX_GROUPS;
    const int threads = 128 + 128 * p.n_groups;
    int cols = 32;
    while (cols < p.n_acc * p.acc_cols) cols *= 2;
    p.tmem_cols = cols;

    CUtensorMap xmap;
    PB_TRY(make_tmap_act(&xmap, a.x, a.Cin, a.L, a.B, cin_t == 16 ? 32 : 64, a_rows, cin_t != 16));
    CUtensorMap wmap = xmap;      // only read by the CTA-pair path
    if (p.cluster) {
        // packed weights as a [rows][256 B] byte matrix: one box = one ring slot (this CTA's half of a block)
        EncodeTiledFn fn = encode_fn();
        PB_REQUIRE(fn != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        PB_REQUIRE(p.b_sub_bytes % 256 == 0 && p.b_sub_bytes / 256 <= 256 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0, PB_ERR_INVALID,
                   "tc_conv: weight ring slot of %d bytes cannot be described as a TMA box", p.b_sub_bytes);
        cuuint64_t gdim[2] = {256, (cuuint64_t)p.w_total_bytes / 256};
        cuuint64_t gstr[1] = {256};
        cuuint32_t box[2] = {256, (cuuint32_t)p.b_sub_bytes / 256};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(a.w), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
    }

    CUtensorMap omap = xmap;      // only read when tma_out
    if (p.tma_out) PB_TRY(tc_make_tmap_out(&omap, a.out, a.out_C, a.L, a.B, 32));
    {
        static const int abl = getenv("PB_TC_ABLATE") ? atoi(getenv("PB_TC_ABLATE")) : 0;
        p.ablate = abl;
    }
    static int n_sm = 0;
    static bool attr_set = false;
    if (!attr_set) {
        int dev = 0;
        PB_CUDA(cudaGetDevice(&dev));
        PB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
#define TC_SET_ATTR(R, A, C) \
        PB_CUDA(cudaFuncSetAttribute(tc_conv_kernel<R, A, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_LIMIT))
        TC_SET_ATTR(false, false, false); TC_SET_ATTR(true, false, false); TC_SET_ATTR(true, true, false); TC_SET_ATTR(false, true, false);
        TC_SET_ATTR(false, false, true); TC_SET_ATTR(true, false, true); TC_SET_ATTR(true, true, true); TC_SET_ATTR(false, true, true);
#undef TC_SET_ATTR
        attr_set = true;
    }
    int grid = p.total_tiles < n_sm * ctas_per_sm ? p.total_tiles : n_sm * ctas_per_sm;
    if (p.cluster) grid = 2 * (p.total_tiles < n_sm / 2 ? p.total_tiles : n_sm / 2);      // CTA pairs
    {
        static const bool dbg = getenv("PB_TC_DEBUG") != nullptr;
        if (dbg)
            fprintf(stderr, "[tc_conv] Cin=%d N=%d(NT=%d) K=%d D=%d L=%d B=%d: tiles=%d grid=%d ctas/sm=%d smem=%u a_stages=%dx%d res_blocks=%d/%d (block %d B) ring=%dx%d kpb=%d n_acc=%d groups=%d cluster=%d tma_out=%d\n",
                    a.Cin, a.N, NT, a.K, a.D, a.L, a.B, p.total_tiles, grid, ctas_per_sm, smem_bytes, p.n_a_stages, p.a_stage_bytes, p.res_blocks,
                    p.total_blocks, p.b_block_bytes, p.n_b_stages, p.b_sub_bytes, p.kpb, p.n_acc, p.n_groups, p.cluster, p.tma_out);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    static const bool no_pdl = getenv("PB_TC_NO_PDL") != nullptr;
    if (!no_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.cluster) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
#define TC_LAUNCH(R, A, C) PB_CUDA(cudaLaunchKernelEx(&cfg, tc_conv_kernel<R, A, C>, xmap, wmap, omap, p))
    if (p.cluster) {
        if (a.res && a.accin) TC_LAUNCH(true, true, true);
        else if (a.res) TC_LAUNCH(true, false, true);
        else if (a.accin) TC_LAUNCH(false, true, true);
        else TC_LAUNCH(false, false, true);
    } else {
        if (a.res && a.accin) TC_LAUNCH(true, true, false);
        else if (a.res) TC_LAUNCH(true, false, false);
        else if (a.accin) TC_LAUNCH(false, true, false);
        else TC_LAUNCH(false, false, false);
    }
#undef TC_LAUNCH
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// weight packing: B operand blocks in the no-swizzle K-major canonical layout (8x16-byte core matrices)
//   packed[((nt*nchunks + c)*K + j)] = block of ksteps x [2 K-halves][hi, lo][NT/8 row groups][8 rows][8 bf16]
//   element (n, k) of plane pl of a k-step: byte ((kc*2 + pl)*(NT/8) + n/8)*128 + (n%8)*16 + (k%8)*2, kc = k/8
// ------------------------------------------------------------------------------------------------------------------------
struct PackDev {
    const float* w;
    int64_t s_n, s_ci, s_j;   // strides of the source for (n, ci, tap); convtr handled by the tap table below
    int Cin, N, K, NT, cin_t, ksteps, nchunks;
    // transposed conv: N = U*Cout, 3 taps; source tap index for (phase, tap) or -1
    int convtr, Cout, Kt, U;
    int fmt;
    int cluster;      // CTA-pair layout: block = [rank 0: ksteps x 4 KB][rank 1: ...], rank = (n % NT) / (NT / 2)
    int pair64;
};

__global__ void tc_pack_kernel(const PackDev d, uint8_t* __restrict__ packed) {
    const int64_t total = (int64_t)d.Cin * d.N * d.K;
    uint16_t* out = reinterpret_cast<uint16_t*>(packed);
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
        uint16_t hi, lo;
        split1(v, d.fmt, hi, lo);
        // block base (in 16-bit elements): block index * (2 planes * ksteps * NT * 16); inside a k-step the K half is outermost,
        // then the plane, so that rows [0, NT) = hi and [NT, 2NT) = lo of ONE 2NT-row operand with LBO = 2*NT*16 bytes
        const int64_t blk = ((int64_t)(nt * d.nchunks + c) * d.K + j) * (2LL * d.ksteps * d.NT * 16);
        if (!d.cluster) {
            const int64_t kstep = blk + (int64_t)ks * (2LL * d.NT * 16);
            const int64_t within = ((int64_t)ng * 8 + r) * 8 + e;
            out[kstep + (int64_t)(kc * 2 + 0) * d.NT * 8 + within] = hi;
            out[kstep + (int64_t)(kc * 2 + 1) * d.NT * 8 + within] = lo;
        } else {
            // pair layout, in 16-bit elements: block = [rank][k-step][R1: 2 K halves x NT rows x 8][R2: 2 K halves x NT/2 rows x 8]
            //   rank 0: R1 = hi rows 0..NT-1, R2 = hi rows 0..NT/2-1;   rank 1: R1 = lo rows 0..NT-1, R2 = hi rows NT/2..NT-1
            const int half = d.NT / 2;
            const int nn = ng * 8 + r;
            const int64_t ks_el = (int64_t)d.NT * 24;                        // R1 (NT*16) + R2 (NT*8) elements per k-step
            const int64_t rank_el = (int64_t)d.ksteps * ks_el;
            const int64_t blk2 = ((int64_t)(nt * d.nchunks + c) * d.K + j) * (2 * rank_el);
            const int64_t r1 = (int64_t)ks * ks_el + ((int64_t)kc * d.NT + nn) * 8 + e;
            out[blk2 + r1] = hi;                                             // rank 0, R1
            out[blk2 + rank_el + r1] = lo;                                   // rank 1, R1
            const int64_t r2 = (int64_t)ks * ks_el + (int64_t)d.NT * 16 + ((int64_t)kc * half + nn % half) * 8 + e;
            out[blk2 + (int64_t)(nn / half) * rank_el + r2] = hi;            // R2 of the rank that owns this row
        }
    }
}

static int pack_common(PackDev d, uint8_t* packed, cudaStream_t st) {
    d.cin_t = tc_cin_tile(d.Cin);
    d.NT = tc_n_tile(d.N);
    d.cluster = tc_use_cluster(d.Cin, d.N, d.K, d.pair64) ? 1 : 0;
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

int tc_pack_conv_weight(const float* w, uint8_t* packed, int Cout, int Cin, int K, cudaStream_t st, int fmt, int pair64) {
    PackDev d{};
    d.fmt = fmt;
    d.pair64 = pair64;
    d.w = w; d.s_n = (int64_t)Cin * K; d.s_ci = K; d.s_j = 1; d.Cin = Cin; d.N = Cout; d.K = K;
    return pack_common(d, packed, st);
}

int tc_pack_convtr_weight(const float* w, uint8_t* packed, int Cin, int Cout, int Kt, int U, cudaStream_t st, int fmt) {
    PB_REQUIRE((Kt - U) % 2 == 0 && Kt <= 3 * U, PB_ERR_INVALID, "tc_pack_convtr: kernel %d / stride %d unsupported", Kt, U);
    // every phase must only reach x[q-1], x[q], x[q+1]
    for (int ph = 0; ph < U; ++ph) {
        const int sft = ph + (Kt - U) / 2, a = sft / U, r0 = sft % U;
        for (int m = 0; r0 + m * U < Kt; ++m)
            PB_REQUIRE(a - m >= -1 && a - m <= 1, PB_ERR_INVALID, "tc_pack_convtr: kernel %d / stride %d needs more than 3 input taps", Kt, U);
    }
    PackDev d{};
    d.fmt = fmt;
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

// keys written by the fused argmax epilogue -> ids (the column is stored complemented so that atomicMax prefers the lowest index)
__global__ void argmax_unpack_kernel(const unsigned long long* __restrict__ key, int64_t M, int64_t* __restrict__ ids) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        ids[i] = (int64_t)(0xffffffffu - (uint32_t)(key[i] & 0xffffffffull));
}

int tc_argmax_unpack(const unsigned long long* key, int64_t M, int64_t* ids, cudaStream_t st) {
    if (M <= 0) return PB_OK;
    const int blocks = (int)(ceil_div(M, 256) < 1184 ? ceil_div(M, 256) : 1184);
    argmax_unpack_kernel<<<blocks, 256, 0, st>>>(key, M, ids);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
