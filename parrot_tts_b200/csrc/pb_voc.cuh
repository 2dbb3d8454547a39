// pb_voc.cuh - internal launch interface of the vocoder kernels (voc_kernels.cu).
#pragma once
#include <cuda_bf16.h>

#include "pb_common.cuh"

namespace pb {

// Activations are f32 [B, C, pitch] (time contiguous, pitch = round_up(L, 4) floats so that every row is
// 16-byte aligned for TMA and 128-bit vector access).  Invariant kept by every epilogue: positions
// t >= valid length (lens[b] * len_mul, or L) hold exact zeros.
struct ConvArgs {
    const float* x = nullptr;      // [B, Cin, pitch_in]
    const float* w = nullptr;      // packed [Cout/CO_T][Cin][K][CO_T]
    int co_tile = 0;               // CO_T the weights were packed with
    const float* bias = nullptr;   // [Cout]
    const float* res = nullptr;    // residual [B, Cout, pitch] or null
    const float* accin = nullptr;  // running sum over resblocks [B, Cout, pitch] or null
    float* out = nullptr;          // [B, Cout, pitch]
    const int32_t* lens = nullptr; // [B] valid length in units, or null
    int len_mul = 1;               // samples per unit at this layer's output
    int B = 0, Cin = 0, Cout = 0, L = 0, pitch = 0;
    int K = 0, D = 1;
    float pre_slope = 1.f;         // leaky-relu on the input (1 = none)
    float post_slope = 1.f;        // leaky-relu on the output (1 = none)
    float acc_div = 1.f;           // (accin + v) is scaled by 1/acc_div (1 = none)
};

// which CO_T the conv / transposed-conv kernels want for a given Cout
int conv_co_tile(int Cout);
int convtr_co_tile(int Cout, int U);

int launch_conv1d(const ConvArgs& a, cudaStream_t st);

struct ConvTrArgs {
    const float* x = nullptr;      // [B, Cin, pitch_in], already activated unless pre_slope != 1
    const float* w = nullptr;      // packed [Cout/CO_T][Cin][K][CO_T]
    int co_tile = 0;
    const float* bias = nullptr;
    float* out = nullptr;          // [B, Cout, pitch_out], L_out = U * L_in
    const int32_t* lens = nullptr;
    int len_mul = 1;               // samples per unit at the OUTPUT
    int B = 0, Cin = 0, Cout = 0, Lin = 0, pitch_in = 0, pitch_out = 0;
    int K = 0, U = 1;
    float pre_slope = 1.f;
};
int launch_convtr1d(const ConvTrArgs& a, cudaStream_t st);

// conv_post: lrelu'd input [B, Cin(16), pitch] -> tanh(conv k7) -> wav [B, 1, Lw] (pitch_out floats per item)
int launch_conv_post(const float* x, const float* w /*[Cin][K]*/, const float* bias, float* wav,
                     const int32_t* lens, int len_mul, int B, int Cin, int L, int pitch_in, int64_t pitch_out,
                     int K, cudaStream_t st);

// unit/speaker embedding gather + concat + transpose: -> x0 [B, 2E (or E), pitch]
int launch_voc_embed(const int64_t* code, const int64_t* spkr, const float* dict_w, const float* spkr_w,
                     const int32_t* lens, int B, int T, int E, int n_codes, int n_spk, int pitch, float* x0, int* err_flag /*or null*/, cudaStream_t st);

// copy [B*C, L] (row stride src_pitch) -> (row stride dst_pitch), zero-filling columns >= valid
int launch_repitch(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int64_t rows, int L,
                   const int32_t* lens, int len_mul, int rows_per_item, cudaStream_t st);

// weight preparation
// w = g * v / ||v|| with the norm over all dims but 0 (v: [n0, inner]); g == null -> plain copy
int launch_weight_norm_fold(const float* v, const float* g, float* w, int n0, int inner, cudaStream_t st);
// src element (co, ci, j) at co*s_co + ci*s_ci + j  ->  dst[((co/CO_T)*Cin + ci)*K + j)*CO_T + co%CO_T]
int launch_pack_conv_weight(const float* src, int64_t s_co, int64_t s_ci, float* dst, int Cout, int Cin, int K, int CO_T,
                            cudaStream_t st);
// ---- tensor-core path (activations split-bf16 channels-last [B][L][2C], see pb_tc.cuh) ----------------------------------
// unit + speaker embedding -> x0 [B][T][2*(E or 2E)] split-bf16 (rows past the valid length are zero)
int launch_voc_embed_split(const int64_t* code, const int64_t* spkr, const float* dict_w, const float* spkr_w,
                           const int32_t* lens, int B, int T, int E, int n_codes, int n_spk, __nv_bfloat16* x0, int* err_flag /*or null*/, cudaStream_t st);
// conv_post on a split-bf16 input that already carries the final leaky-relu: tanh(conv k7 + bias) -> wav f32 [B][L]
int launch_conv_post_split(const __nv_bfloat16* x, const float* w /*[Cin][K]*/, const float* bias, float* wav /*or null*/,
                           int16_t* pcm /*or null*/, const int32_t* lens, int len_mul, int B, int Cin, int L, int K, cudaStream_t st);

int launch_wav_to_pcm16(const float* wav, int64_t n, int16_t* pcm, cudaStream_t st);
int launch_pcm16_peak_normalize(const int16_t* pcm, int B, int64_t n, int32_t* peak, float* out, cudaStream_t st);

}  // namespace pb
