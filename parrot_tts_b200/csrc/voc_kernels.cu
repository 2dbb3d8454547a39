// voc_kernels.cu - HiFi-GAN code-generator kernels for sm_100a.
//
// Replaces, for the vocoder hot path of the reference (utils/vocoder/models.py):
//   F.conv1d(+leaky_relu, +residual)      ResBlock1.forward           models.py:31-38   -> conv1d_kernel
//   F.conv1d                              conv_pre                    models.py:96      -> conv1d_kernel
//   F.conv_transpose1d                    ups[i]                      models.py:98-99   -> convtr1d_kernel
//   xs += resblock(x); x = xs / 3         resblock mean               models.py:100-106 -> conv1d epilogue
//   leaky_relu(0.01) -> conv_post -> tanh                             models.py:107-109 -> conv_post_kernel
//   dict(code) (+) spkr(spkr) upsampled                               models.py:153-160 -> voc_embed_kernel
//
// Design: activations f32 [B, C, pitch] (time contiguous).  Every conv is an explicit shared-memory
// sliding-window kernel: TMA (cp.async.bulk.tensor.2d) stages a [CI_CHUNK x (128 + halo)] input box per
// warp time-slot, out-of-range columns zero-filled by the TMA unit (= the conv's zero padding), a 1-D bulk
// copy stages the matching [CI_CHUNK][K][CO_T] filter slab; a 3-deep mbarrier ring overlaps the copies
// with the FMAs.  Each thread owns CO_PT output channels x NG*4 consecutive samples in registers and
// slides the K taps over a register window read with 128-bit conflict-free LDS.
#include "pb_voc.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------------
// conv1d ('same' padding, dilation D, kernel K)
// ------------------------------------------------------------------------------------------------
struct ConvDev {
    const float* w;
    const float* bias;
    const float* res;
    const float* accin;
    float* out;
    const int32_t* lens;
    int len_mul;
    int Cin, Cout, L, pitch;
    float pre_slope, post_slope, acc_scale;
};

template <int K_, int D_, int CO_PT_, int CO_GROUPS_, int T_SLOTS_, int NG_>
struct ConvTraits {
    static constexpr int K = K_, D = D_, CO_PT = CO_PT_, CO_GROUPS = CO_GROUPS_, T_SLOTS = T_SLOTS_, NG = NG_;
    static constexpr int CI_CHUNK = 8;
    static constexpr int NSTAGE = 3;
    static constexpr int THREADS = 32 * CO_GROUPS * T_SLOTS;
    static constexpr int CO_T = CO_PT * CO_GROUPS;
    static constexpr int HALO = (K - 1) * D;
    static constexpr int PAD = HALO / 2;
    // TMA needs the innermost box coordinate 16-byte aligned (measured: an unaligned start raises an illegal-
    // instruction fault), so the box starts PAD4 = round_up(PAD, 4) samples early and the taps are shifted by SH.
    static constexpr int PAD4 = (PAD + 3) / 4 * 4;
    static constexpr int SH = PAD4 - PAD;
    static constexpr int HALO4 = (HALO + SH + 3) / 4 * 4;
    static constexpr int BOXW = 128 + HALO4;   // floats per staged row
    static constexpr int WIN4 = 1 + HALO4 / 4; // float4s in a thread's register window
    static constexpr int NBOX = T_SLOTS * NG;
    static constexpr int T_T = 128 * NBOX;     // samples per CTA
    static constexpr int X_FLOATS = NBOX * CI_CHUNK * BOXW;
    static constexpr int W_FLOATS = CI_CHUNK * K * CO_T;
    static constexpr int STAGE_FLOATS = X_FLOATS + W_FLOATS;
    static constexpr int SMEM_BYTES = 256 + NSTAGE * STAGE_FLOATS * 4;
    static_assert(BOXW <= 256, "TMA box dimension limit");
    static_assert(CO_PT % 4 == 0, "CO_PT must be a multiple of 4");
};

template <class Tr>
__global__ void __launch_bounds__(Tr::THREADS, 2)
conv1d_kernel(const __grid_constant__ CUtensorMap xmap, const ConvDev p) {
    constexpr int K = Tr::K, D = Tr::D, CO_PT = Tr::CO_PT, NG = Tr::NG;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // keeps the shared address space (LDS, not LD)
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    float* stage0 = reinterpret_cast<float*>(sm + 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int co_group = warp % Tr::CO_GROUPS, t_slot = warp / Tr::CO_GROUPS;
    const int b = blockIdx.z, co_tile = blockIdx.y, t_tile0 = blockIdx.x * Tr::T_T;
    const int nchunks = p.Cin / Tr::CI_CHUNK;

    if (tid == 0) {
        prefetch_tmap(&xmap);
#pragma unroll
        for (int s = 0; s < Tr::NSTAGE; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int c) {
        const int s = c % Tr::NSTAGE;
        float* xs = stage0 + s * Tr::STAGE_FLOATS;
        float* ws = xs + Tr::X_FLOATS;
        mbar_expect_tx(&full[s], Tr::STAGE_FLOATS * 4);
#pragma unroll
        for (int bx = 0; bx < Tr::NBOX; ++bx)
            tma_load_2d(xs + bx * Tr::CI_CHUNK * Tr::BOXW, &xmap, t_tile0 + bx * 128 - Tr::PAD4,
                        b * p.Cin + c * Tr::CI_CHUNK, &full[s]);
        bulk_load_1d(ws, p.w + (size_t)(co_tile * p.Cin + c * Tr::CI_CHUNK) * K * Tr::CO_T, Tr::W_FLOATS * 4, &full[s]);
    };
    if (tid == 0) {
        const int n0 = nchunks < Tr::NSTAGE ? nchunks : Tr::NSTAGE;
        for (int c = 0; c < n0; ++c) issue(c);
    }

    float acc[CO_PT][NG][4];
#pragma unroll
    for (int c = 0; c < CO_PT; ++c)
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[c][g][i] = 0.f;

    const bool pre_act = p.pre_slope != 1.f;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % Tr::NSTAGE;
        mbar_wait(&full[s], (c / Tr::NSTAGE) & 1);
        float* xs = stage0 + s * Tr::STAGE_FLOATS;
        const float* ws = xs + Tr::X_FLOATS;
        if (pre_act) {
            float4* x4 = reinterpret_cast<float4*>(xs);
            for (int i = tid; i < Tr::X_FLOATS / 4; i += Tr::THREADS) {
                float4 v = x4[i];
                v.x = lrelu(v.x, p.pre_slope); v.y = lrelu(v.y, p.pre_slope);
                v.z = lrelu(v.z, p.pre_slope); v.w = lrelu(v.w, p.pre_slope);
                x4[i] = v;
            }
            __syncthreads();
        }
#pragma unroll 2
        for (int ci = 0; ci < Tr::CI_CHUNK; ++ci) {
            float xw[NG][Tr::WIN4 * 4];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4* xp = reinterpret_cast<const float4*>(
                    xs + ((t_slot * NG + g) * Tr::CI_CHUNK + ci) * Tr::BOXW + lane * 4);
#pragma unroll
                for (int q = 0; q < Tr::WIN4; ++q) {
                    const float4 v = xp[q];
                    xw[g][4 * q + 0] = v.x; xw[g][4 * q + 1] = v.y; xw[g][4 * q + 2] = v.z; xw[g][4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float4* wp = reinterpret_cast<const float4*>(ws + (ci * K + j) * Tr::CO_T + co_group * CO_PT);
                float wv[CO_PT];
#pragma unroll
                for (int q = 0; q < CO_PT / 4; ++q) {
                    const float4 v = wp[q];
                    wv[4 * q + 0] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int cc = 0; cc < CO_PT; ++cc)
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[cc][g][i] = fmaf(wv[cc], xw[g][Tr::SH + j * D + i], acc[cc][g][i]);
            }
        }
        __syncthreads();
        if (tid == 0 && c + Tr::NSTAGE < nchunks) {
            fence_proxy_async();
            issue(c + Tr::NSTAGE);
        }
    }

    // ---- epilogue: bias, residual, resblock running sum / mean, activation, length mask, 128-bit stores
    int valid = p.L;
    if (p.lens) { const int v = p.lens[b] * p.len_mul; valid = v < valid ? v : valid; }
    const float acc_scale = p.acc_scale, post_slope = p.post_slope;
    const float* __restrict__ bias_p = p.bias;
    const float* res_p = p.res;
    const float* accin_p = p.accin;
    float* out_p = p.out;
    const int L = p.L;
    const size_t pitch = p.pitch;
    const size_t row0 = ((size_t)b * p.Cout + co_tile * Tr::CO_T + co_group * CO_PT) * pitch;
#pragma unroll
    for (int cc = 0; cc < CO_PT; ++cc) {
        const int co = co_tile * Tr::CO_T + co_group * CO_PT + cc;
        const float bias = bias_p ? bias_p[co] : 0.f;
        const size_t row = row0 + cc * pitch;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int t = t_tile0 + (t_slot * NG + g) * 128 + lane * 4;
            if (t >= L) continue;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = acc[cc][g][i] + bias;
            if (res_p) {
                const float4 r = *reinterpret_cast<const float4*>(res_p + row + t);
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
            if (accin_p) {
                const float4 r = *reinterpret_cast<const float4*>(accin_p + row + t);
                v[0] = r.x + v[0]; v[1] = r.y + v[1]; v[2] = r.z + v[2]; v[3] = r.w + v[3];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float u = v[i] * acc_scale;          // mean over resblocks: x * (1/n), within 1.5 ulp of x / n
                u = fmaxf(u, u * post_slope);        // leaky relu for slope in [0, 1] (slope 1 = identity)
                v[i] = (t + i < valid) ? u : 0.f;
            }
            *reinterpret_cast<float4*>(out_p + row + t) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

int conv_co_tile(int Cout) {
    if (Cout % 64 == 0) return 64;
    if (Cout % 32 == 0) return 32;
    if (Cout % 16 == 0) return 16;
    return 0;
}

template <class Tr>
static int launch_conv_inst(const ConvArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = conv1d_kernel<Tr>;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tr::SMEM_BYTES));
        attr_set = true;
    }
    CUtensorMap xmap;
    PB_TRY(make_tmap_2d_f32(&xmap, a.x, a.L, (int64_t)a.B * a.Cin, a.pitch, Tr::BOXW, Tr::CI_CHUNK));
    ConvDev p{a.w, a.bias, a.res, a.accin, a.out, a.lens, a.len_mul, a.Cin, a.Cout, a.L, a.pitch,
              a.pre_slope, a.post_slope, 1.0f / a.acc_div};
    dim3 grid((unsigned)ceil_div(a.L, Tr::T_T), (unsigned)(a.Cout / Tr::CO_T), (unsigned)a.B);
    kern<<<grid, Tr::THREADS, Tr::SMEM_BYTES, st>>>(xmap, p);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

template <int K, int D>
static int launch_conv_kd(const ConvArgs& a, cudaStream_t st) {
    constexpr int HALO = (K - 1) * D;
    constexpr int NG = HALO > 24 ? 1 : 2;  // large register windows: one sample group per thread
    switch (a.co_tile) {
        case 64: return launch_conv_inst<ConvTraits<K, D, 8, 8, 1, NG>>(a, st);
        case 32: return launch_conv_inst<ConvTraits<K, D, 8, 4, 2, NG>>(a, st);
        case 16: return launch_conv_inst<ConvTraits<K, D, 8, 2, 4, NG>>(a, st);
    }
    set_error("conv1d: unsupported output-channel tile %d", a.co_tile);
    return PB_ERR_INVALID;
}

int launch_conv1d(const ConvArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.Cin % 8 == 0, PB_ERR_INVALID, "conv1d: Cin=%d must be a multiple of 8", a.Cin);
    PB_REQUIRE(a.co_tile > 0 && a.Cout % a.co_tile == 0, PB_ERR_INVALID, "conv1d: Cout=%d not tiled by %d", a.Cout, a.co_tile);
    PB_REQUIRE(a.pitch % 4 == 0 && a.pitch >= a.L, PB_ERR_INVALID, "conv1d: pitch=%d must be a multiple of 4 >= L=%d", a.pitch, a.L);
    PB_REQUIRE(a.B > 0 && a.L > 0 && a.B <= 65535, PB_ERR_INVALID, "conv1d: bad shape B=%d L=%d", a.B, a.L);
#define PB_KD(k, d) if (a.K == k && a.D == d) return launch_conv_kd<k, d>(a, st);
    PB_KD(3, 1) PB_KD(3, 3) PB_KD(3, 5)
    PB_KD(7, 1) PB_KD(7, 3) PB_KD(7, 5)
    PB_KD(11, 1) PB_KD(11, 3) PB_KD(11, 5)
#undef PB_KD
    set_error("conv1d: kernel size %d / dilation %d has no sm_100a instantiation", a.K, a.D);
    return PB_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// conv_transpose1d, stride U, kernel K, padding (K-U)/2  ->  L_out = U * L_in   (polyphase form)
//   out[U*q + p] = bias + sum_ci sum_m W[ci, co, r + m*U] * x[ci, q + a - m],  s = p + PAD, a = s / U, r = s % U
// ------------------------------------------------------------------------------------------------
struct ConvTrDev {
    const float* w;
    const float* bias;
    float* out;
    const int32_t* lens;
    int len_mul;
    int Cin, Cout, Lin, pitch_out;
    float pre_slope;
};

template <int K_, int U_, int CO_PT_, int CO_GROUPS_, int Q_SLOTS_>
struct ConvTrTraits {
    static constexpr int K = K_, U = U_, CO_PT = CO_PT_, CO_GROUPS = CO_GROUPS_, Q_SLOTS = Q_SLOTS_;
    static constexpr int CI_CHUNK = 8;
    static constexpr int NSTAGE = 3;
    static constexpr int THREADS = 32 * CO_GROUPS * Q_SLOTS;
    static constexpr int CO_T = CO_PT * CO_GROUPS;
    static constexpr int PADT = (K - U) / 2;
    static constexpr int BOXW = 136;  // box starts 4 samples early (16-byte aligned TMA coordinate); 12-float window per thread
    static constexpr int Q_T = 128 * Q_SLOTS;
    static constexpr int X_FLOATS = Q_SLOTS * CI_CHUNK * BOXW;
    static constexpr int W_FLOATS = CI_CHUNK * K * CO_T;
    static constexpr int STAGE_FLOATS = X_FLOATS + W_FLOATS;
    static constexpr int SMEM_BYTES = 256 + NSTAGE * STAGE_FLOATS * 4;
    static_assert((K - U) % 2 == 0, "padding (K-U)/2 must be exact");
    static_assert(CO_PT % 4 == 0, "CO_PT must be a multiple of 4");
};

template <class Tr>
__global__ void __launch_bounds__(Tr::THREADS, 2)
convtr1d_kernel(const __grid_constant__ CUtensorMap xmap, const ConvTrDev p) {
    constexpr int K = Tr::K, U = Tr::U, CO_PT = Tr::CO_PT;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // keeps the shared address space (LDS, not LD)
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    float* stage0 = reinterpret_cast<float*>(sm + 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int co_group = warp % Tr::CO_GROUPS, q_slot = warp / Tr::CO_GROUPS;
    const int b = blockIdx.z, co_tile = blockIdx.y, q_tile0 = blockIdx.x * Tr::Q_T;
    const int nchunks = p.Cin / Tr::CI_CHUNK;

    if (tid == 0) {
        prefetch_tmap(&xmap);
#pragma unroll
        for (int s = 0; s < Tr::NSTAGE; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int c) {
        const int s = c % Tr::NSTAGE;
        float* xs = stage0 + s * Tr::STAGE_FLOATS;
        float* ws = xs + Tr::X_FLOATS;
        mbar_expect_tx(&full[s], Tr::STAGE_FLOATS * 4);
#pragma unroll
        for (int bx = 0; bx < Tr::Q_SLOTS; ++bx)
            tma_load_2d(xs + bx * Tr::CI_CHUNK * Tr::BOXW, &xmap, q_tile0 + bx * 128 - 4,
                        b * p.Cin + c * Tr::CI_CHUNK, &full[s]);
        bulk_load_1d(ws, p.w + (size_t)(co_tile * p.Cin + c * Tr::CI_CHUNK) * K * Tr::CO_T, Tr::W_FLOATS * 4, &full[s]);
    };
    if (tid == 0) {
        const int n0 = nchunks < Tr::NSTAGE ? nchunks : Tr::NSTAGE;
        for (int c = 0; c < n0; ++c) issue(c);
    }

    float acc[CO_PT][4][U];
#pragma unroll
    for (int c = 0; c < CO_PT; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int ph = 0; ph < U; ++ph) acc[c][q][ph] = 0.f;

    const bool pre_act = p.pre_slope != 1.f;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % Tr::NSTAGE;
        mbar_wait(&full[s], (c / Tr::NSTAGE) & 1);
        float* xs = stage0 + s * Tr::STAGE_FLOATS;
        const float* ws = xs + Tr::X_FLOATS;
        if (pre_act) {
            float4* x4 = reinterpret_cast<float4*>(xs);
            for (int i = tid; i < Tr::X_FLOATS / 4; i += Tr::THREADS) {
                float4 v = x4[i];
                v.x = lrelu(v.x, p.pre_slope); v.y = lrelu(v.y, p.pre_slope);
                v.z = lrelu(v.z, p.pre_slope); v.w = lrelu(v.w, p.pre_slope);
                x4[i] = v;
            }
            __syncthreads();
        }
#pragma unroll 2
        for (int ci = 0; ci < Tr::CI_CHUNK; ++ci) {
            // window xw[n] = x[q0 - 4 + n], n = 0..11 (taps use n = 3..8)
            const float4* xp = reinterpret_cast<const float4*>(xs + (q_slot * Tr::CI_CHUNK + ci) * Tr::BOXW + lane * 4);
            const float4 x0 = xp[0], x1 = xp[1], x2 = xp[2];
            const float xw[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
#pragma unroll
            for (int j = 0; j < K; ++j) {
                // tap j contributes to phase ph with (ph + PADT) % U == j % U, using x[q + a - m], m = (j - r) / U
                const float4* wp = reinterpret_cast<const float4*>(ws + (ci * K + j) * Tr::CO_T + co_group * CO_PT);
                float wv[CO_PT];
#pragma unroll
                for (int q = 0; q < CO_PT / 4; ++q) {
                    const float4 v = wp[q];
                    wv[4 * q + 0] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                }
                const int r = j % U;
                const int m = j / U;
                // ph = (r - PADT) mod U ; a = (ph + PADT) / U
                const int ph = ((r - Tr::PADT) % U + U) % U;
                const int a = (ph + Tr::PADT) / U;
                const int off = a - m + 4;  // index into xw for q = 0
#pragma unroll
                for (int cc = 0; cc < CO_PT; ++cc)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[cc][q][ph] = fmaf(wv[cc], xw[q + off], acc[cc][q][ph]);
            }
        }
        __syncthreads();
        if (tid == 0 && c + Tr::NSTAGE < nchunks) {
            fence_proxy_async();
            issue(c + Tr::NSTAGE);
        }
    }

    const int Lout = p.Lin * U;
    int valid = Lout;
    if (p.lens) { const int v = p.lens[b] * p.len_mul; valid = v < valid ? v : valid; }
    const int q0 = q_tile0 + q_slot * 128 + lane * 4;
    if (q0 >= p.Lin) return;
#pragma unroll
    for (int cc = 0; cc < CO_PT; ++cc) {
        const int co = co_tile * Tr::CO_T + co_group * CO_PT + cc;
        const float bias = p.bias ? p.bias[co] : 0.f;
        float* orow = p.out + ((size_t)b * p.Cout + co) * p.pitch_out;
        float v[4 * U];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int ph = 0; ph < U; ++ph) {
                const int o = U * (q0 + q) + ph;
                v[q * U + ph] = (o < valid) ? acc[cc][q][ph] + bias : 0.f;
            }
        const int o0 = U * q0;
#pragma unroll
        for (int e = 0; e < U; ++e) {
            const int o = o0 + 4 * e;
            if (o < Lout) *reinterpret_cast<float4*>(orow + o) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
    }
}

int convtr_co_tile(int Cout, int U) {
    (void)U;
    if (Cout % 32 == 0) return 32;
    if (Cout % 16 == 0) return 16;
    return 0;
}

template <class Tr>
static int launch_convtr_inst(const ConvTrArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = convtr1d_kernel<Tr>;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tr::SMEM_BYTES));
        attr_set = true;
    }
    CUtensorMap xmap;
    PB_TRY(make_tmap_2d_f32(&xmap, a.x, a.Lin, (int64_t)a.B * a.Cin, a.pitch_in, Tr::BOXW, Tr::CI_CHUNK));
    ConvTrDev p{a.w, a.bias, a.out, a.lens, a.len_mul, a.Cin, a.Cout, a.Lin, a.pitch_out, a.pre_slope};
    dim3 grid((unsigned)ceil_div(a.Lin, Tr::Q_T), (unsigned)(a.Cout / Tr::CO_T), (unsigned)a.B);
    kern<<<grid, Tr::THREADS, Tr::SMEM_BYTES, st>>>(xmap, p);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

template <int K, int U>
static int launch_convtr_ku(const ConvTrArgs& a, cudaStream_t st) {
    switch (a.co_tile) {
        case 32: return launch_convtr_inst<ConvTrTraits<K, U, 4, 8, 1>>(a, st);
        case 16: return launch_convtr_inst<ConvTrTraits<K, U, 4, 4, 2>>(a, st);
    }
    set_error("conv_transpose1d: unsupported output-channel tile %d", a.co_tile);
    return PB_ERR_INVALID;
}

int launch_convtr1d(const ConvTrArgs& a, cudaStream_t st) {
    PB_REQUIRE(a.Cin % 8 == 0, PB_ERR_INVALID, "conv_transpose1d: Cin=%d must be a multiple of 8", a.Cin);
    PB_REQUIRE(a.co_tile > 0 && a.Cout % a.co_tile == 0, PB_ERR_INVALID, "conv_transpose1d: Cout=%d not tiled by %d", a.Cout, a.co_tile);
    PB_REQUIRE(a.pitch_in % 4 == 0 && a.pitch_out % 4 == 0 && a.pitch_out >= a.Lin * a.U, PB_ERR_INVALID,
               "conv_transpose1d: bad pitches %d/%d", a.pitch_in, a.pitch_out);
    PB_REQUIRE(a.B > 0 && a.Lin > 0 && a.B <= 65535, PB_ERR_INVALID, "conv_transpose1d: bad shape");
    if (a.K == 11 && a.U == 5) return launch_convtr_ku<11, 5>(a, st);
    if (a.K == 8 && a.U == 4) return launch_convtr_ku<8, 4>(a, st);
    if (a.K == 4 && a.U == 2) return launch_convtr_ku<4, 2>(a, st);
    set_error("conv_transpose1d: kernel %d / stride %d has no sm_100a instantiation", a.K, a.U);
    return PB_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// conv_post (Cin -> 1, K = 7) + tanh.  HBM-bound: each thread produces 4 samples from aligned float4 reads.
// ------------------------------------------------------------------------------------------------
template <int CIN, int K>
__global__ void __launch_bounds__(256)
conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ wav, const int32_t* __restrict__ lens, int len_mul, int L, int pitch_in,
                 int64_t pitch_out) {
    static_assert(K == 7, "window below assumes K = 7 (pad 3)");
    __shared__ float ws[CIN * K];
    for (int i = threadIdx.x; i < CIN * K; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= L) return;
    int valid = L;
    if (lens) { const int v = lens[b] * len_mul; valid = v < valid ? v : valid; }
    const float* xb = x + (size_t)b * CIN * pitch_in;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int ci = 0; ci < CIN; ++ci) {
        const float* xr = xb + (size_t)ci * pitch_in;
        float win[12];  // x[t-4 .. t+7]
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a0 = (t >= 4) ? *reinterpret_cast<const float4*>(xr + t - 4) : z;
        const float4 a1 = *reinterpret_cast<const float4*>(xr + t);
        const float4 a2 = (t + 4 < pitch_in) ? *reinterpret_cast<const float4*>(xr + t + 4) : z;
        win[0] = a0.x; win[1] = a0.y; win[2] = a0.z; win[3] = a0.w;
        win[4] = a1.x; win[5] = a1.y; win[6] = a1.z; win[7] = a1.w;
        win[8] = a2.x; win[9] = a2.y; win[10] = a2.z; win[11] = a2.w;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float wj = ws[ci * K + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(wj, win[i + j + 1], acc[i]);  // x[t + i + j - 3]
        }
    }
    const float bv = bias[0];
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (t + i < valid) ? tanhf(acc[i] + bv) : 0.f;
    float* o = wav + (size_t)b * pitch_out + t;
    if (t + 3 < L && (pitch_out % 4 == 0)) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t + i < L) o[i] = v[i];
    }
}

int launch_conv_post(const float* x, const float* w, const float* bias, float* wav, const int32_t* lens, int len_mul,
                     int B, int Cin, int L, int pitch_in, int64_t pitch_out, int K, cudaStream_t st) {
    PB_REQUIRE(K == 7, PB_ERR_INVALID, "conv_post: kernel size %d unsupported (7 expected)", K);
    PB_REQUIRE(pitch_in % 4 == 0, PB_ERR_INVALID, "conv_post: pitch");
    dim3 grid((unsigned)ceil_div(L, 1024), (unsigned)B);
    if (Cin == 16) conv_post_kernel<16, 7><<<grid, 256, 0, st>>>(x, w, bias, wav, lens, len_mul, L, pitch_in, pitch_out);
    else if (Cin == 32) conv_post_kernel<32, 7><<<grid, 256, 0, st>>>(x, w, bias, wav, lens, len_mul, L, pitch_in, pitch_out);
    else { set_error("conv_post: Cin=%d unsupported", Cin); return PB_ERR_INVALID; }
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
// unit + speaker embedding -> [B, 2E, pitch]  (gather, concat on channels, transpose to time-contiguous)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
voc_embed_kernel(const int64_t* __restrict__ code, const int64_t* __restrict__ spkr, const float* __restrict__ dict_w,
                 const float* __restrict__ spkr_w, const int32_t* __restrict__ lens, int T, int E, int n_codes, int n_spk,
                 int pitch, int C, float* __restrict__ x0, int* __restrict__ err_flag) {
    // block: 32 time steps x all channels; tile transposed through shared memory
    extern __shared__ float tile[];  // [32][E + 1]
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int valid = lens ? min(lens[b], T) : T;
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt;
        long long id = (t < T) ? code[(size_t)b * T + t] : 0;
        if (err_flag && tx == 0 && t < valid && (id < 0 || id >= n_codes)) *err_flag = 1;      // see voc_embed_split_kernel
        id = id < 0 ? 0 : (id >= n_codes ? n_codes - 1 : id);
        for (int c = tx; c < E; c += 32) tile[tt * (E + 1) + c] = (t < valid) ? dict_w[(size_t)id * E + c] : 0.f;
    }
    __syncthreads();
    const int t = t0 + tx;
    float* xb = x0 + (size_t)b * C * pitch;
    if (t < pitch) {
        for (int c = ty; c < E; c += 8) xb[(size_t)c * pitch + t] = (t < T) ? tile[tx * (E + 1) + c] : 0.f;
        if (spkr_w) {
            long long s = spkr[b];
            if (err_flag && threadIdx.x == 0 && blockIdx.x == 0 && valid > 0 && (s < 0 || s >= n_spk)) atomicMax(err_flag, 2);
            s = s < 0 ? 0 : (s >= n_spk ? n_spk - 1 : s);
            for (int c = ty; c < E; c += 8) xb[(size_t)(E + c) * pitch + t] = (t < valid) ? spkr_w[(size_t)s * E + c] : 0.f;
        }
    }
}

int launch_voc_embed(const int64_t* code, const int64_t* spkr, const float* dict_w, const float* spkr_w, const int32_t* lens,
                     int B, int T, int E, int n_codes, int n_spk, int pitch, float* x0, int* err_flag, cudaStream_t st) {
    const int C = spkr_w ? 2 * E : E;
    dim3 grid((unsigned)ceil_div(pitch, 32), (unsigned)B);
    voc_embed_kernel<<<grid, 256, 32 * (E + 1) * sizeof(float), st>>>(code, spkr, dict_w, spkr_w, lens, T, E, n_codes, n_spk, pitch, C, x0, err_flag);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void repitch_kernel(const float* __restrict__ src, int64_t src_pitch, float* __restrict__ dst, int64_t dst_pitch,
                               int64_t rows, int L, const int32_t* __restrict__ lens, int len_mul, int rows_per_item) {
    const int64_t row = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows || t >= dst_pitch) return;
    int valid = L;
    if (lens) { const int v = lens[row / rows_per_item] * len_mul; valid = v < valid ? v : valid; }
    dst[row * dst_pitch + t] = (t < valid) ? src[row * src_pitch + t] : 0.f;
}

int launch_repitch(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int64_t rows, int L,
                   const int32_t* lens, int len_mul, int rows_per_item, cudaStream_t st) {
    PB_REQUIRE(rows <= 65535 * 1024LL, PB_ERR_INVALID, "repitch: too many rows");
    const int64_t width = dst_pitch;
    // grid.y limited to 65535: fold rows
    if (rows <= 65535) {
        dim3 grid((unsigned)ceil_div(width, 256), (unsigned)rows);
        repitch_kernel<<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, rows, L, lens, len_mul, rows_per_item);
        PB_COUNT_LAUNCH();
    } else {
        for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
            const int64_t n = rows - r0 < 65535 ? rows - r0 : 65535;
            PB_REQUIRE(lens == nullptr || (r0 % rows_per_item) == 0, PB_ERR_INVALID, "repitch: row split");
            dim3 grid((unsigned)ceil_div(width, 256), (unsigned)n);
            repitch_kernel<<<grid, 256, 0, st>>>(src + r0 * src_pitch, src_pitch, dst + r0 * dst_pitch, dst_pitch, n, L,
                                                 lens ? lens + r0 / rows_per_item : nullptr, len_mul, rows_per_item);
            PB_COUNT_LAUNCH();
        }
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// w[n, :] = g[n] * v[n, :] / ||v[n, :]||_2   (torch._weight_norm(v, g, dim=0)); one block per n
__global__ void __launch_bounds__(256)
weight_norm_fold_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int inner) {
    const int n = blockIdx.x;
    const float* vr = v + (size_t)n * inner;
    float* wr = w + (size_t)n * inner;
    if (g == nullptr) {
        for (int i = threadIdx.x; i < inner; i += blockDim.x) wr[i] = vr[i];
        return;
    }
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) { const double a = vr[i]; s += a * a; }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float norm = (float)sqrt(red[0]);
    const float scale = g[n] / norm;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) wr[i] = vr[i] * scale;
}

int launch_weight_norm_fold(const float* v, const float* g, float* w, int n0, int inner, cudaStream_t st) {
    weight_norm_fold_kernel<<<n0, 256, 0, st>>>(v, g, w, inner);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ src, int64_t s_co, int64_t s_ci, float* __restrict__ dst,
                                        int Cout, int Cin, int K, int CO_T) {
    const int64_t n = (int64_t)Cout * Cin * K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % CO_T);
        int64_t r = i / CO_T;
        const int j = (int)(r % K); r /= K;
        const int ci = (int)(r % Cin);
        const int ct = (int)(r / Cin);
        const int co = ct * CO_T + c;
        dst[i] = src[co * s_co + ci * s_ci + j];
    }
}

int launch_pack_conv_weight(const float* src, int64_t s_co, int64_t s_ci, float* dst, int Cout, int Cin, int K, int CO_T,
                            cudaStream_t st) {
    const int64_t n = (int64_t)Cout * Cin * K;
    const int blocks = (int)(ceil_div(n, 256) < 1024 ? ceil_div(n, 256) : 1024);
    pack_conv_weight_kernel<<<blocks, 256, 0, st>>>(src, s_co, s_ci, dst, Cout, Cin, K, CO_T);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor-core path helpers: embedding gather into split-bf16 rows, conv_post on split-bf16 rows
// ------------------------------------------------------------------------------------------------
// one warp per (b, t) row; row = [hi(dict E | spkr E) | lo(dict E | spkr E)]
__global__ void __launch_bounds__(256)
voc_embed_split_kernel(const int64_t* __restrict__ code, const int64_t* __restrict__ spkr, const float* __restrict__ dict_w,
                       const float* __restrict__ spkr_w, const int32_t* __restrict__ lens, int T, int E, int n_codes, int n_spk,
                       int64_t rows, __nv_bfloat16* __restrict__ x0, int* __restrict__ err_flag) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = (int)(row / T), t = (int)(row % T);
    const int C = spkr_w ? 2 * E : E;
    const int valid = lens ? min(lens[b], T) : T;
    long long id = code[row];
    long long sp = spkr_w ? spkr[b] : 0;
    // nn.Embedding raises on an id outside the table (models.py:155-156): flag it for pb_voc_check_ids (live positions only - what
    // lies past an item's length is padding and never read); the gather itself clamps so that the launch stays in bounds
    if (err_flag && lane == 0 && t < valid) {
        if (id < 0 || id >= n_codes) *err_flag = 1;
        else if (sp < 0 || sp >= n_spk) *err_flag = 2;
    }
    id = id < 0 ? 0 : (id >= n_codes ? n_codes - 1 : id);
    sp = sp < 0 ? 0 : (sp >= n_spk ? n_spk - 1 : sp);
    __nv_bfloat16* out = x0 + row * 2 * C;
    for (int c = lane; c < C; c += 32) {
        float v = 0.f;
        if (t < valid) v = c < E ? dict_w[(size_t)id * E + c] : spkr_w[(size_t)sp * E + (c - E)];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        out[c] = hi;
        out[C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

int launch_voc_embed_split(const int64_t* code, const int64_t* spkr, const float* dict_w, const float* spkr_w,
                           const int32_t* lens, int B, int T, int E, int n_codes, int n_spk, __nv_bfloat16* x0, int* err_flag,
                           cudaStream_t st) {
    const int64_t rows = (int64_t)B * T;
    voc_embed_split_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(code, spkr, dict_w, spkr_w, lens, T, E, n_codes, n_spk, rows, x0,
                                                                         err_flag);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// conv_post + tanh on split-bf16 rows (reference utils/vocoder/models.py:107-109): a block stages POST_T + K - 1 input rows in shared
// memory as fp32 [channel][position] (hi + lo joined once); a thread owns 4 consecutive samples: 10 inputs (two 128-bit loads and a
// 64-bit one) and the 7 weights of a channel serve 28 FMAs (the first version issued 2 shared-memory loads per FMA: 91 us per
// 8 x 1024 units against a 27 us HBM floor).  The FMA order of one sample is channel outer, tap inner.
// Outputs: the f32 waveform and / or - the generate() tail fused in (utils/vocoder/inference.py:71-73) - int16 PCM
// (x 32768, truncation toward zero, two's-complement wrap: the same expression as wav_to_pcm16_kernel on the same float).
constexpr int POST_T = 512;
__device__ __forceinline__ int16_t pcm16_of(float v) {
    const long long q = (long long)(v * 32768.0f);      // trunc toward zero
    return (int16_t)(uint16_t)(q & 0xffff);
}
template <int CIN, int K>
__global__ void __launch_bounds__(POST_T / 4)
conv_post_split_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ wav, int16_t* __restrict__ pcm, const int32_t* __restrict__ lens, int len_mul, int L) {
    constexpr int ROWS = POST_T + K - 1;
    constexpr int PITCH = (ROWS + 3 + 3) / 4 * 4;     // rows o .. o + 9 of a 4-sample tile stay inside the pitch; multiple of 4 floats
    __shared__ __align__(16) float xs[CIN * PITCH];
    __shared__ float ws[CIN * K];
    for (int i = threadIdx.x; i < CIN * K; i += blockDim.x) ws[i] = w[i];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * POST_T;
    int valid = L;
    if (lens) { const int v = lens[b] * len_mul; valid = v < valid ? v : valid; }
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        const int tt = t0 + r - (K - 1) / 2;
        if (tt >= 0 && tt < L) {
            const uint4* row = reinterpret_cast<const uint4*>(x + ((size_t)b * L + tt) * 2 * CIN);
#pragma unroll
            for (int q = 0; q < CIN / 8; ++q) {
                const uint4 h = __ldg(row + q), l = __ldg(row + CIN / 8 + q);
                const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xs[(q * 8 + 2 * i) * PITCH + r] = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
                    xs[(q * 8 + 2 * i + 1) * PITCH + r] = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xs[c * PITCH + r] = 0.f;
        }
    }
    __syncthreads();
    const float bv = bias[0];
    static_assert(POST_T % 4 == 0 && PITCH % 4 == 0 && K <= 7, "4-sample tiles read xs[o .. o + 9] with aligned vector loads");
    for (int o = 4 * threadIdx.x; o < POST_T; o += 4 * blockDim.x) {
        if (t0 + o >= L) break;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float* xr = xs + c * PITCH + o;
            const float4 a = *reinterpret_cast<const float4*>(xr), bq = *reinterpret_cast<const float4*>(xr + 4);
            const float2 cq = *reinterpret_cast<const float2*>(xr + 8);
            const float in[10] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, cq.x, cq.y};
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float wj = ws[c * K + j];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(wj, in[q + j], acc[q]);
            }
        }
        float y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = t0 + o + q < valid ? tanhf(acc[q] + bv) : 0.f;
        const size_t at = (size_t)b * L + t0 + o;
        if (t0 + o + 3 < L && (L & 3) == 0) {         // whole tile inside the row: one 128-bit (f32) / 64-bit (int16) store
            if (wav) *reinterpret_cast<float4*>(wav + at) = make_float4(y[0], y[1], y[2], y[3]);
            if (pcm) {
                const uint32_t lo = (uint32_t)(uint16_t)pcm16_of(y[0]) | ((uint32_t)(uint16_t)pcm16_of(y[1]) << 16);
                const uint32_t hi = (uint32_t)(uint16_t)pcm16_of(y[2]) | ((uint32_t)(uint16_t)pcm16_of(y[3]) << 16);
                *reinterpret_cast<uint2*>(pcm + at) = make_uint2(lo, hi);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (t0 + o + q < L) {
                    if (wav) wav[at + q] = y[q];
                    if (pcm) pcm[at + q] = pcm16_of(y[q]);
                }
            }
        }
    }
}

int launch_conv_post_split(const __nv_bfloat16* x, const float* w, const float* bias, float* wav, int16_t* pcm, const int32_t* lens,
                           int len_mul, int B, int Cin, int L, int K, cudaStream_t st) {
    PB_REQUIRE(K == 7, PB_ERR_INVALID, "conv_post: kernel size %d unsupported (7 expected)", K);
    PB_REQUIRE(wav || pcm, PB_ERR_INVALID, "conv_post: no output");
    dim3 grid((unsigned)ceil_div(L, POST_T), (unsigned)B);
    if (Cin == 16) conv_post_split_kernel<16, 7><<<grid, POST_T / 4, 0, st>>>(x, w, bias, wav, pcm, lens, len_mul, L);
    else { set_error("conv_post: Cin=%d unsupported on the tensor-core path (16 expected)", Cin); return PB_ERR_INVALID; }
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// generate() tail: (wav * 32768) -> int16 with C-cast semantics (truncate toward zero, wrap modulo 2^16)
__global__ void wav_to_pcm16_kernel(const float* __restrict__ wav, int64_t n, int16_t* __restrict__ pcm) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        pcm[i] = pcm16_of(wav[i]);
    }
}

int launch_wav_to_pcm16(const float* wav, int64_t n, int16_t* pcm, cudaStream_t st) {
    if (n <= 0) return PB_OK;
    const int blocks = (int)(ceil_div(n, 256) < 148 * 8 ? ceil_div(n, 256) : 148 * 8);
    wav_to_pcm16_kernel<<<blocks, 256, 0, st>>>(wav, n, pcm);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// Peak normalisation of generated PCM (librosa.util.normalize defaults: norm = inf over the utterance; an all-zero
// utterance is left as it is).  Pass 1: per-item max |pcm| with one atomicMax per block; pass 2: float(pcm) / float(peak),
// IEEE division (librosa divides in float64 and rounds to float32: identical for 16-bit integers, 53 >= 2*24 + 2).
__global__ void pcm16_peak_kernel(const int16_t* __restrict__ pcm, int64_t n, int32_t* __restrict__ peak) {
    const int16_t* x = pcm + (size_t)blockIdx.y * n;
    int m = 0;
    const int64_t n8 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 8 : 0;     // 16-byte vector body when the item is aligned
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = (int)(int16_t)(w[k] & 0xffffu), b = (int)(int16_t)(w[k] >> 16);
            m = max(m, max(a < 0 ? -a : a, b < 0 ? -b : b));
        }
    }
    for (int64_t i = n8 * 8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)x[i];
        m = max(m, a < 0 ? -a : a);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    __shared__ int sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = max(m, sm[w]);
        if (m > 0) atomicMax(peak + blockIdx.y, m);
    }
}

__global__ void pcm16_scale_kernel(const int16_t* __restrict__ pcm, int64_t n, const int32_t* __restrict__ peak, float* __restrict__ out) {
    const int pk = peak[blockIdx.y];
    const float len = pk > 0 ? (float)pk : 1.0f;
    const int16_t* x = pcm + (size_t)blockIdx.y * n;
    float* o = out + (size_t)blockIdx.y * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = __fdiv_rn((float)x[i], len);
}

int launch_pcm16_peak_normalize(const int16_t* pcm, int B, int64_t n, int32_t* peak, float* out, cudaStream_t st) {
    if (B <= 0 || n <= 0) return PB_OK;
    PB_CUDA(cudaMemsetAsync(peak, 0, sizeof(int32_t) * (size_t)B, st));
    const int64_t want = ceil_div(n, 2048), cap = 148 * 8 / B > 0 ? 148 * 8 / B : 1;     // ~8 blocks per SM over all items
    const int per_item = (int)(want < cap ? want : cap);
    dim3 grid((unsigned)per_item, (unsigned)B);
    pcm16_peak_kernel<<<grid, 256, 0, st>>>(pcm, n, peak);
    PB_COUNT_LAUNCH();
    pcm16_scale_kernel<<<grid, 256, 0, st>>>(pcm, n, peak, out);
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb
