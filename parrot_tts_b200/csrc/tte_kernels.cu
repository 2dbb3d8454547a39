// tte_kernels.cu - TTE (text -> HuBERT units) kernels for sm_100a, fp32 SIMT family.
//
// Replaces, for the reference's TTE hot path:
//   nn.Embedding + pe[L] add            modules/parrot.py:94-95, modules/fft.py:17-19   -> embed_pos_kernel
//   nn.LayerNorm                        modules/fft.py:91-92,98-99                      -> layernorm_kernel
//   nn.Linear / Conv1d(k9,k1,k3)        modules/fft.py:48-57,65-80; duration.py:30-34   -> sgemm_kernel (im2col view)
//   baddbmm + softmax + bmm             torch F.multi_head_attention_forward            -> sgemm_kernel (batched) + softmax_kernel
//   proj + masked_fill + exp/round      modules/duration.py:45-48; parrot.py:82-86      -> duration_head_kernel
//   length_regulator + mask             modules/duration.py:6-24; data.py:8-20          -> duration_scan / length_regulate kernels
//   argmax                              modules/parrot.py:115                           -> argmax_kernel
#include <cuda_fp16.h>

#include "pb_tte.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------------
// SGEMM 128x128x16, 256 threads, 8x8 register tile, double-buffered shared memory
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int BM = 128, BN = 128, BK = 16, LDS_ = 132;

struct GemmDev {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    const float* bias;
    const float* res; int64_t ldres;
    int M, N, K, Ka;   // Ka = K rounded up to 4 (A may be read up to there)
    int relu;
    int taps, Cin, L, pad;
    int H;
    int64_t sAb, sAh, sBb, sBh, sCb, sCh;
};
}  // namespace

template <bool B_KN, bool TAPS>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const GemmDev g) {
    __shared__ __align__(16) float As[2][BK][LDS_];
    __shared__ __align__(16) float Bs[2][BK][LDS_];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int zb = blockIdx.z / g.H, zh = blockIdx.z % g.H;
    const float* __restrict__ A = g.A + zb * g.sAb + zh * g.sAh;
    const float* __restrict__ Bm = g.B + zb * g.sBb + zh * g.sBh;
    float* C = g.C + zb * g.sCb + zh * g.sCh;
    const float* res = g.res ? g.res + zb * g.sCb + zh * g.sCh : nullptr;

    // global -> register staging: A tile 128 x 16 (2 float4 per thread), B tile likewise
    float4 ra[2], rb[2];
    const int a_row[2] = {tid >> 2, (tid + 256) >> 2};
    const int a_kq = tid & 3;
    int a_t[2] = {0, 0};
    if (TAPS) {
#pragma unroll
        for (int r = 0; r < 2; ++r) a_t[r] = (m0 + a_row[r]) % g.L;
    }
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = m0 + a_row[r];
            const int kk = k0 + a_kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < g.M && kk < g.Ka) {
                if (TAPS) {
                    const int j = kk / g.Cin, ci = kk - j * g.Cin;
                    const int tt = a_t[r] + j - g.pad;
                    if (tt >= 0 && tt < g.L) v = *reinterpret_cast<const float4*>(A + (int64_t)(m + j - g.pad) * g.lda + ci);
                } else {
                    v = *reinterpret_cast<const float4*>(A + (int64_t)m * g.lda + kk);
                }
            }
            ra[r] = v;
        }
        if (!B_KN) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int n = n0 + a_row[r];
                const int kk = k0 + a_kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < g.N && kk < g.K) v = *reinterpret_cast<const float4*>(Bm + (int64_t)n * g.ldb + kk);
                rb[r] = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int f = tid + 256 * r;
                const int kr = f >> 5, nq = f & 31;
                const int kk = k0 + kr, n = n0 + nq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kk < g.K) {
                    const float* src = Bm + (int64_t)kk * g.ldb + n;
                    if (n + 3 < g.N) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (n < g.N) v.x = src[0];
                        if (n + 1 < g.N) v.y = src[1];
                        if (n + 2 < g.N) v.z = src[2];
                    }
                }
                rb[r] = v;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = a_row[r];
            As[buf][a_kq * 4 + 0][row] = ra[r].x; As[buf][a_kq * 4 + 1][row] = ra[r].y;
            As[buf][a_kq * 4 + 2][row] = ra[r].z; As[buf][a_kq * 4 + 3][row] = ra[r].w;
        }
        if (!B_KN) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = a_row[r];
                Bs[buf][a_kq * 4 + 0][row] = rb[r].x; Bs[buf][a_kq * 4 + 1][row] = rb[r].y;
                Bs[buf][a_kq * 4 + 2][row] = rb[r].z; Bs[buf][a_kq * 4 + 3][row] = rb[r].w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int f = tid + 256 * r;
                *reinterpret_cast<float4*>(&Bs[buf][f >> 5][(f & 31) * 4]) = rb[r];
            }
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (g.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
    const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        (!res || ((g.ldres % 4 == 0) && ((reinterpret_cast<uintptr_t>(res) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= g.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][jh * 4 + j];
            if (g.bias) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] += g.bias[n + j];
            }
            if (g.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (n + 3 < g.N && vec_ok) {
                if (res) {
                    const float4 r = *reinterpret_cast<const float4*>(res + (int64_t)m * g.ldres + n);
                    v[0] = r.x + v[0]; v[1] = r.y + v[1]; v[2] = r.z + v[2]; v[3] = r.w + v[3];
                }
                *reinterpret_cast<float4*>(C + (int64_t)m * g.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.N) {
                        float o = v[j];
                        if (res) o = res[(int64_t)m * g.ldres + n + j] + o;
                        C[(int64_t)m * g.ldc + n + j] = o;
                    }
            }
        }
    }
}

int launch_sgemm(const GemmArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, PB_ERR_INVALID, "sgemm: empty problem %dx%dx%d", a.M, a.N, a.K);
    PB_REQUIRE(a.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0, PB_ERR_INVALID, "sgemm: A must be 16-byte aligned (lda=%lld)", (long long)a.lda);
    PB_REQUIRE(a.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0, PB_ERR_INVALID, "sgemm: B must be 16-byte aligned (ldb=%lld)", (long long)a.ldb);
    PB_REQUIRE(a.sAb % 4 == 0 && a.sAh % 4 == 0 && a.sBb % 4 == 0 && a.sBh % 4 == 0, PB_ERR_INVALID, "sgemm: batch strides must be multiples of 4");
    if (!a.b_kn) PB_REQUIRE(a.K % 4 == 0, PB_ERR_INVALID, "sgemm: K=%d must be a multiple of 4", a.K);
    if (a.taps > 1) PB_REQUIRE(a.Cin % 4 == 0 && a.K == a.taps * a.Cin && a.L > 0 && a.M % a.L == 0, PB_ERR_INVALID, "sgemm: bad conv view");
    GemmDev g{a.A, a.lda, a.B, a.ldb, a.C, a.ldc, a.bias, a.res, a.ldres, a.M, a.N, a.K, (int)round_up(a.K, 4), a.relu,
              a.taps, a.Cin, a.L, a.pad, a.H > 0 ? a.H : 1, a.sAb, a.sAh, a.sBb, a.sBh, a.sCb, a.sCh};
    dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, BM), (unsigned)a.nbatch);
    PB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PB_ERR_INVALID, "sgemm: problem too large for the grid (M=%d)", a.M);
    if (a.b_kn) {
        PB_REQUIRE(a.taps <= 1, PB_ERR_INVALID, "sgemm: conv view with [K,N] weights unsupported");
        sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
    } else if (a.taps > 1) {
        sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
    } else {
        sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
    }
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
// row kernels (one warp per row of D = 256 floats: 2 x float4 per lane)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_pos_kernel(const int64_t* __restrict__ phones, const float* __restrict__ tok_emb, const float* __restrict__ pe_row,
                 const int32_t* __restrict__ lens, int L, int vocab, int M, float* __restrict__ out, int* err_flag) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    long long id = phones[row];
    if (id < 0 || id >= vocab) { if (lane == 0 && err_flag) atomicExch(err_flag, 1); id = 0; }
    const float4* src = reinterpret_cast<const float4*>(tok_emb + (size_t)id * 256);
    // ragged mode (lens != null): pe_row is the TABLE and every item takes the row of its OWN length - what the reference adds
    // when the item is run alone (fft.py:18 indexes the table by the padded length, which is the item's length at batch 1)
    const float4* pe = reinterpret_cast<const float4*>(lens ? pe_row + (size_t)lens[row / L] * 256 : pe_row);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * 256);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 a = src[lane + 32 * h], p = pe[lane + 32 * h];
        dst[lane + 32 * h] = make_float4(p.x + a.x, p.y + a.y, p.z + a.z, p.w + a.w);
    }
}

int launch_embed_pos(const int64_t* phones, const float* tok_emb, const float* pe_row, const int32_t* lens, int L, int vocab, int M,
                     int D, float* out, int* err_flag, cudaStream_t st) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "embed: d_model=%d unsupported (256 expected)", D);
    embed_pos_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(phones, tok_emb, pe_row, lens, L > 0 ? L : 1, vocab, M, out, err_flag);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__device__ __forceinline__ void ln_row_256(const float4 v0, const float4 v1, float& mean, float& rstd) {
    float s = (v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w);
    s = warp_sum(s);
    mean = s * (1.0f / 256.0f);
    float d, q = 0.f;
    d = v0.x - mean; q = fmaf(d, d, q); d = v0.y - mean; q = fmaf(d, d, q);
    d = v0.z - mean; q = fmaf(d, d, q); d = v0.w - mean; q = fmaf(d, d, q);
    d = v1.x - mean; q = fmaf(d, d, q); d = v1.y - mean; q = fmaf(d, d, q);
    d = v1.z - mean; q = fmaf(d, d, q); d = v1.w - mean; q = fmaf(d, d, q);
    q = warp_sum(q);
    rstd = rsqrtf(q * (1.0f / 256.0f) + 1e-5f);
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float* __restrict__ y, int M, const int32_t* __restrict__ lens, int L) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float4* dst = reinterpret_cast<float4*>(y + (size_t)row * 256);
    if (lens && row % L >= lens[row / L]) {      // ragged mode: past the item's end the following convolution must see its zero padding
        dst[lane] = make_float4(0.f, 0.f, 0.f, 0.f); dst[lane + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * 256);
    const float4 v0 = src[lane], v1 = src[lane + 32];
    float mean, rstd;
    ln_row_256(v0, v1, mean, rstd);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    const float4 g0 = g4[lane], g1 = g4[lane + 32], c0 = b4[lane], c1 = b4[lane + 32];
    dst[lane] = make_float4((v0.x - mean) * rstd * g0.x + c0.x, (v0.y - mean) * rstd * g0.y + c0.y,
                            (v0.z - mean) * rstd * g0.z + c0.z, (v0.w - mean) * rstd * g0.w + c0.w);
    dst[lane + 32] = make_float4((v1.x - mean) * rstd * g1.x + c1.x, (v1.y - mean) * rstd * g1.y + c1.y,
                                 (v1.z - mean) * rstd * g1.z + c1.z, (v1.w - mean) * rstd * g1.w + c1.w);
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int M, int D, cudaStream_t st,
                     const int32_t* lens, int L) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "layernorm: width %d unsupported (256 expected)", D);
    PB_REQUIRE(M > 0, PB_ERR_INVALID, "layernorm: empty input");
    layernorm_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(x, gamma, beta, y, M, lens, L > 0 ? L : 1);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ---- tensor-core path: rows in split-fp16 format [M][2C] = [hi(0..C-1) | lo(0..C-1)], value = hi + lo (pb_tc.cuh) --------
__device__ __forceinline__ void split_h4(const float4 v, uint2& hi, uint2& lo) {
    const float a = fminf(fmaxf(v.x, -65504.f), 65504.f), b = fminf(fmaxf(v.y, -65504.f), 65504.f);
    const float c = fminf(fmaxf(v.z, -65504.f), 65504.f), d = fminf(fmaxf(v.w, -65504.f), 65504.f);
    const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(a - f0.x, b - f0.y), l1 = __floats2half2_rn(c - f1.x, d - f1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// LayerNorm(256) with the result written in split-fp16 rows (the A operand of the following tensor-core GEMM)
__global__ void __launch_bounds__(256)
layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       uint16_t* __restrict__ y, int M, const int32_t* __restrict__ lens, int L) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    uint2* dst = reinterpret_cast<uint2*>(y + (size_t)row * 512);     // 4 halves = 8 bytes per lane
    if (lens && row % L >= lens[row / L]) {      // ragged mode: zero rows past the item's end (the k = 9 convolution's zero padding)
        const uint2 z = make_uint2(0u, 0u);
        dst[lane] = z; dst[lane + 32] = z; dst[64 + lane] = z; dst[96 + lane] = z;
        return;
    }
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * 256);
    const float4 v0 = src[lane], v1 = src[lane + 32];
    float mean, rstd;
    ln_row_256(v0, v1, mean, rstd);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    const float4 g0 = g4[lane], g1 = g4[lane + 32], c0 = b4[lane], c1 = b4[lane + 32];
    const float4 o0 = make_float4((v0.x - mean) * rstd * g0.x + c0.x, (v0.y - mean) * rstd * g0.y + c0.y,
                                  (v0.z - mean) * rstd * g0.z + c0.z, (v0.w - mean) * rstd * g0.w + c0.w);
    const float4 o1 = make_float4((v1.x - mean) * rstd * g1.x + c1.x, (v1.y - mean) * rstd * g1.y + c1.y,
                                  (v1.z - mean) * rstd * g1.z + c1.z, (v1.w - mean) * rstd * g1.w + c1.w);
    uint2 h0, l0, h1, l1;
    split_h4(o0, h0, l0);
    split_h4(o1, h1, l1);
    dst[lane] = h0; dst[lane + 32] = h1;                              // hi: channels 4*lane.., 128 + 4*lane..
    dst[64 + lane] = l0; dst[96 + lane] = l1;                         // lo plane starts at element 256
}

int launch_layernorm_split(const float* x, const float* gamma, const float* beta, uint16_t* y, int M, int D, cudaStream_t st,
                           const int32_t* lens, int L) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "layernorm: width %d unsupported (256 expected)", D);
    PB_REQUIRE(M > 0, PB_ERR_INVALID, "layernorm: empty input");
    layernorm_split_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(x, gamma, beta, y, M, lens, L > 0 ? L : 1);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// x f32 [M][C] (row stride ldx) -> split-fp16 rows [M][2C]; C % 4 == 0
__global__ void __launch_bounds__(256)
rows_to_split_kernel(const float* __restrict__ x, int64_t ldx, uint16_t* __restrict__ y, int64_t M, int C) {
    const int per_row = C / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M * per_row; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / per_row;
        const int c4 = (int)(i - row * per_row);
        const float4 v = *reinterpret_cast<const float4*>(x + row * ldx + 4 * c4);
        uint2 h, l;
        split_h4(v, h, l);
        uint2* dst = reinterpret_cast<uint2*>(y + row * 2 * C);
        dst[c4] = h;
        dst[per_row + c4] = l;
    }
}

int launch_rows_to_split(const float* x, int64_t ldx, uint16_t* y, int64_t M, int C, cudaStream_t st) {
    PB_REQUIRE(C % 4 == 0 && ldx % 4 == 0, PB_ERR_INVALID, "rows_to_split: width %d / stride %lld must be multiples of 4", C, (long long)ldx);
    const int64_t n = M * (C / 4);
    const int blocks = (int)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
    rows_to_split_kernel<<<blocks, 256, 0, st>>>(x, ldx, y, M, C);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void __launch_bounds__(256)
add_indexed_row_kernel(float* __restrict__ x, const float* __restrict__ table, const int64_t* __restrict__ idx, int n_rows,
                       int L, int M, int* err_flag, const int32_t* __restrict__ lens) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float4* dst = reinterpret_cast<float4*>(x + (size_t)row * 256);
    if (lens && row % L >= lens[row / L]) {      // ragged mode: the encoder output is zero past the item's end (duration-predictor conv padding)
        dst[lane] = make_float4(0.f, 0.f, 0.f, 0.f); dst[lane + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    if (!table) return;
    long long id = idx[row / L];
    if (id < 0 || id >= n_rows) { if (lane == 0 && err_flag) atomicExch(err_flag, 2); id = 0; }
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)id * 256);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 a = dst[lane + 32 * h];
        const float4 s = src[lane + 32 * h];
        a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
        dst[lane + 32 * h] = a;
    }
}

int launch_add_indexed_row(float* x, const float* table, const int64_t* idx, int n_rows, int B, int L, int D, int* err_flag,
                           cudaStream_t st, const int32_t* lens) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "add_row: width %d unsupported", D);
    add_indexed_row_kernel<<<(unsigned)ceil_div((int64_t)B * L, 8), 256, 0, st>>>(x, table, idx, n_rows, L, B * L, err_flag, lens);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// in-place masked softmax, one warp per row
__global__ void __launch_bounds__(256)
softmax_kernel(float* __restrict__ S, const uint8_t* __restrict__ key_valid, int H, int L, int Lp, float scale, int64_t rows) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int z = (int)(row / L);
    const int b = z / H;
    float* s = S + row * Lp;
    const uint8_t* kv = key_valid ? key_valid + (size_t)b * L : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) {
        const float v = (kv == nullptr || kv[j]) ? s[j] * scale : -INFINITY;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
        const float v = (kv == nullptr || kv[j]) ? expf(s[j] * scale - mx) : 0.f;
        s[j] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Lp; j += 32) s[j] = (j < L) ? s[j] * inv : 0.f;
}

int launch_softmax(float* S, const uint8_t* key_valid, int Z, int H, int L, int Lp, float scale, cudaStream_t st) {
    const int64_t rows = (int64_t)Z * L;
    softmax_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(S, key_valid, H, L, Lp, scale, rows);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void __launch_bounds__(256)
duration_head_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ w, const float* __restrict__ b, const uint8_t* __restrict__ src_mask, int M,
                     float* __restrict__ log_dur, int64_t* __restrict__ dur) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * 256);
    const float4 v0 = src[lane], v1 = src[lane + 32];
    float mean, rstd;
    ln_row_256(v0, v1, mean, rstd);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4 g0 = g4[lane], g1 = g4[lane + 32], c0 = b4[lane], c1 = b4[lane + 32], w0 = w4[lane], w1 = w4[lane + 32];
    float d = 0.f;
    d = fmaf((v0.x - mean) * rstd * g0.x + c0.x, w0.x, d); d = fmaf((v0.y - mean) * rstd * g0.y + c0.y, w0.y, d);
    d = fmaf((v0.z - mean) * rstd * g0.z + c0.z, w0.z, d); d = fmaf((v0.w - mean) * rstd * g0.w + c0.w, w0.w, d);
    d = fmaf((v1.x - mean) * rstd * g1.x + c1.x, w1.x, d); d = fmaf((v1.y - mean) * rstd * g1.y + c1.y, w1.y, d);
    d = fmaf((v1.z - mean) * rstd * g1.z + c1.z, w1.z, d); d = fmaf((v1.w - mean) * rstd * g1.w + c1.w, w1.w, d);
    d = warp_sum(d);
    if (lane == 0) {
        float ld = d + b[0];
        if (src_mask && !src_mask[row]) ld = 0.f;   // masked_fill(~src_mask, 0.0)
        log_dur[row] = ld;
        if (dur) {
            // clamp(round(exp(ld) - 1), min=0): exp in double then rounded to fp32, half-to-even rounding
            const float e = (float)exp((double)ld);
            const float r = rintf(e - 1.0f);
            dur[row] = (int64_t)fmaxf(r, 0.f);
        }
    }
}

int launch_duration_head(const float* x, const float* gamma, const float* beta, const float* w, const float* b,
                         const uint8_t* src_mask, int M, int D, float* log_dur, int64_t* dur, cudaStream_t st) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "duration head: width %d unsupported", D);
    duration_head_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(x, gamma, beta, w, b, src_mask, M, log_dur, dur);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// one block per item: inclusive scan of durations (S <= a few thousand)
__global__ void __launch_bounds__(256)
duration_scan_kernel(const int64_t* __restrict__ dur, int S, int32_t* __restrict__ cum, int64_t* __restrict__ out_lens,
                     int32_t* __restrict__ out_lens32) {
    __shared__ long long part[256];
    const int b = blockIdx.x;
    const int per = (S + 255) / 256;
    const int s0 = threadIdx.x * per;
    long long local = 0;
    for (int i = 0; i < per; ++i) { const int s = s0 + i; if (s < S) local += dur[(size_t)b * S + s]; }
    part[threadIdx.x] = local;
    __syncthreads();
    // exclusive prefix over the 256 partials (serial per thread: 256 adds, negligible)
    long long base = 0;
    for (int i = 0; i < threadIdx.x; ++i) base += part[i];
    long long run = base;
    for (int i = 0; i < per; ++i) {
        const int s = s0 + i;
        if (s < S) { run += dur[(size_t)b * S + s]; cum[(size_t)b * S + s] = (int32_t)(run > 0x7fffffffLL ? 0x7fffffffLL : run); }
    }
    if (threadIdx.x == 255) {
        out_lens[b] = base + local;
        if (out_lens32) out_lens32[b] = (int32_t)(base + local > 0x7fffffffLL ? 0x7fffffffLL : base + local);
    }
}

int launch_duration_scan(const int64_t* dur, int B, int S, int32_t* cum, int64_t* out_lens, cudaStream_t st, int32_t* out_lens32) {
    duration_scan_kernel<<<B, 256, 0, st>>>(dur, S, cum, out_lens, out_lens32);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void __launch_bounds__(256)
length_regulate_kernel(const float* __restrict__ seq, const int32_t* __restrict__ cum, const float* __restrict__ pe_row, int S,
                       int T, int64_t rows, float* __restrict__ out, uint8_t* __restrict__ tgt_mask, int ragged) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = (int)(row / T), t = (int)(row % T);
    const int32_t* c = cum + (size_t)b * S;
    const int len = c[S - 1];
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (t < len) {
        // first s with cum[s] > t
        int lo = 0, hi = S - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid] > t) hi = mid; else lo = mid + 1; }
        const float4* src = reinterpret_cast<const float4*>(seq + ((size_t)b * S + lo) * 256);
        a0 = src[lane]; a1 = src[lane + 32];
    }
    if (pe_row) {
        // ragged mode: pe_row is the TABLE and the item takes the row of its own unit count (what it gets when run alone, fft.py:18)
        const float4* pe = reinterpret_cast<const float4*>(ragged ? pe_row + (size_t)len * 256 : pe_row);
        const float4 p0 = pe[lane], p1 = pe[lane + 32];
        a0 = make_float4(p0.x + a0.x, p0.y + a0.y, p0.z + a0.z, p0.w + a0.w);
        a1 = make_float4(p1.x + a1.x, p1.y + a1.y, p1.z + a1.z, p1.w + a1.w);
    }
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * 256);
    dst[lane] = a0; dst[lane + 32] = a1;
    // `<=`: reference data.py:18 (one extra live position for every item shorter than the batch maximum); run alone an item has
    // exactly `len` positions (arange(len) <= len is all true), which is what the ragged mode reproduces
    if (lane == 0 && tgt_mask) tgt_mask[row] = (ragged ? t < len : t <= len) ? 1 : 0;
}

int launch_length_regulate(const float* seq, const int32_t* cum, const float* pe_row, int B, int S, int D, int T, float* out,
                           uint8_t* tgt_mask, cudaStream_t st, int ragged) {
    PB_REQUIRE(D == 256, PB_ERR_INVALID, "length_regulate: width %d unsupported", D);
    const int64_t rows = (int64_t)B * T;
    if (rows == 0) return PB_OK;
    length_regulate_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(seq, cum, pe_row, S, T, rows, out, tgt_mask, ragged);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void __launch_bounds__(256)
argmax_kernel(const float* __restrict__ logits, int M, int N, int64_t ldl, int64_t* __restrict__ ids) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* r = logits + (size_t)row * ldl;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int j = lane; j < N; j += 32) {
        const float v = r[j];
        if (v > best || (v == best && j < bi)) { best = v; bi = j; }   // NaN never wins, like a strict > scan
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) ids[row] = (bi == 0x7fffffff) ? 0 : bi;
}

int launch_argmax(const float* logits, int M, int N, int64_t ldl, int64_t* ids, cudaStream_t st) {
    argmax_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(logits, M, N, ldl, ids);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void permute_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int N, int Cin, int K) {
    const int64_t total = (int64_t)N * Cin * K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        int64_t r = i / Cin;
        const int j = (int)(r % K);
        const int n = (int)(r / K);
        out[i] = w[((int64_t)n * Cin + ci) * K + j];
    }
}

int launch_permute_conv_weight(const float* w, float* out, int N, int Cin, int K, cudaStream_t st) {
    const int64_t total = (int64_t)N * Cin * K;
    const int blocks = (int)(ceil_div(total, 256) < 2048 ? ceil_div(total, 256) : 2048);
    permute_conv_weight_kernel<<<blocks, 256, 0, st>>>(w, out, N, Cin, K);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
