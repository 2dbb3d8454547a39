// pb_tc.cuh - tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM convolution family for sm_100a.
//
// Activation format ("split-bf16, channels-last"): a tensor with C channels and L positions per item is
//   bf16 [B][L][2C], row = [hi(0..C-1) | lo(0..C-1)],  value = float(hi) + float(lo),
// i.e. the two leading bf16 terms of the fp32 value (4 bytes per element, the same HBM footprint as fp32).
// A product x.w is evaluated as x_hi.w_hi + x_lo.w_hi + x_hi.w_lo on the bf16 tensor pipe with fp32
// accumulation in TMEM: relative error ~2^-16 per product instead of 2^-9 (bf16) or 2^-11 (tf32) - what the
// 1e-3 waveform tolerance of the north star needs (hard part 2).
#pragma once
#include <cuda_bf16.h>

#include "pb_common.cuh"

namespace pb {

// One launch = one Conv1d ('same' padding, dilation D, K taps) or one polyphase ConvTranspose1d, as an implicit GEMM
//   acc[t, n] = sum_{tap j} sum_{ci} x[t + j*D - pad, ci] * W_j[ci, n]            (M = 128 time rows per tile)
// followed by the fused epilogue
//   v = acc + bias[co] (+ unact(res)) (+ accin);  v *= scale;  v = lrelu(v, post_slope);  v = 0 past the valid length
// and the store of v as split-bf16 (and/or fp32) at row t*U + n / out_C, channel n % out_C.
struct TcConvArgs {
    const __nv_bfloat16* x = nullptr;      // [B][L][2*Cin] split-bf16 (already activated by its producer)
    const uint8_t* w = nullptr;            // packed by tc_pack_* (see tc_conv.cu)
    const float* bias = nullptr;           // [out_C] or null
    const __nv_bfloat16* res = nullptr;    // [B][U*L][2*out_C] residual or null
    float res_unslope = 1.f;               // the residual tensor is stored leaky-relu'd with slope s: pass 1/s to undo (1 = raw)
    const __nv_bfloat16* accin = nullptr;  // [B][U*L][2*out_C] raw running sum or null
    __nv_bfloat16* out = nullptr;          // [B][U*L][2*out_C] split-bf16 or null
    float* out_f32 = nullptr;              // [B][U*L][out_C] fp32 or null
    const int32_t* lens = nullptr;         // [B] valid units per item or null
    int len_mul = 1;                       // output rows per unit
    int B = 0, L = 0, Cin = 0;
    int N = 0;                             // GEMM N = U * out_C
    int out_C = 0, U = 1;
    int K = 1, D = 1;
    float post_slope = 1.f, scale = 1.f;
    int tanh_out = 0;                      // apply tanh after the activation (conv_post)
    int single_pass = 0;                   // 1: hi.hi only (plain bf16 product, throughput mode)
    int n_groups = 0;                      // epilogue warpgroups per CTA (0 = choose)
    int fmt = 0;                           // split storage + MMA operand type of x / w / res / accin / out: 0 = bf16, 1 = fp16
    int pair64 = 0;                        // as given to tc_pack_conv_weight for this layer
    const float* res_f32 = nullptr;        // [B][U*L][out_C] fp32 residual (added like res, no activation to undo) or null
    // fused row argmax (U == 1): key[b*L + t] = max over columns n < argmax_n of (order-preserving bits of v[t, n]) << 32 | ~n, merged
    // with atomicMax (zero the keys first; tc_argmax_unpack turns them into indices - first maximum, like torch.argmax).  out and
    // out_f32 may then both be null: the GEMM result never reaches HBM.
    unsigned long long* argmax_key = nullptr;
    int argmax_n = 0;
};
int tc_argmax_unpack(const unsigned long long* key, int64_t M, int64_t* ids, cudaStream_t st);

int tc_n_tile(int N);                       // N-tile the kernel uses for a GEMM N (0 = unsupported)
// pair64: the caller's per-layer choice for the 64 -> 64-channel layers with >= 7 taps (see tc_use_cluster); must be the same at pack
// and at launch time.  Every other shape is decided by the shape alone.
size_t tc_packed_weight_bytes(int Cin, int N, int K, int pair64 = 0);
bool tc_use_cluster(int Cin, int N, int K, int pair64 = 0);     // layer runs on CTA pairs (cta_group::2); decides the packed weight layout
// torch Conv1d weight f32 [Cout][Cin][K]  -> packed split-bf16 B operand (N = Cout)
int tc_pack_conv_weight(const float* w, uint8_t* packed, int Cout, int Cin, int K, cudaStream_t st, int fmt = 0, int pair64 = 0);
// torch ConvTranspose1d weight f32 [Cin][Cout][Kt], stride U, padding (Kt-U)/2 -> 3-tap polyphase operand (N = U*Cout)
int tc_pack_convtr_weight(const float* w, uint8_t* packed, int Cin, int Cout, int Kt, int U, cudaStream_t st, int fmt = 0);
int launch_tc_conv(const TcConvArgs& a, cudaStream_t st);

// Fused ResBlock1 step (tc_pair.cu), C = 16 / 32:  out = post( (conv2(lrelu(conv1(x) + b1, mid_slope)) + b2 + unact(x) [+ accin]) * scale )
// with x stored leaky-relu'd (res_unslope undoes it), conv1 dilation D, conv2 dilation 1, both K taps, 'same' padding.
struct TcPairArgs {
    const __nv_bfloat16* x = nullptr;      // [B][L][2C] split-bf16: conv1 input and conv2 residual
    const uint8_t* w1 = nullptr;           // packed by tc_pack_conv_weight (N = Cin = C)
    const uint8_t* w2 = nullptr;
    const float* bias1 = nullptr;
    const float* bias2 = nullptr;
    const __nv_bfloat16* accin = nullptr;  // raw running sum or null (may alias out)
    __nv_bfloat16* out = nullptr;
    const int32_t* lens = nullptr;
    int len_mul = 1;
    int B = 0, L = 0, C = 0, K = 1, D = 1;
    float mid_slope = 0.1f, post_slope = 1.f, scale = 1.f, res_unslope = 10.f;
};
bool tc_pair_supported(int C, int K, int D);
int launch_tc_pair(const TcPairArgs& a, cudaStream_t st);
// 3-D TMA map over split 16-bit activations [B][L][2C] (box: box_c channels x rows positions), shared by tc_conv.cu / tc_pair.cu
int tc_make_tmap_act(CUtensorMap* map, const __nv_bfloat16* base, int C, int L, int B, int box_c, int rows, bool sw128);
int tc_make_tmap_out(CUtensorMap* map, __nv_bfloat16* base, int C, int L, int B, int rows);
int tc_make_tmap_grouped16(CUtensorMap* map, const __nv_bfloat16* base, int L, int B, int rows);

// layout conversion (test entry points, embedding stage): f32 [B][C][L] (row stride pitch) <-> split-bf16 [B][L][2C]
int launch_f32_to_split(const float* x, int64_t pitch, __nv_bfloat16* y, int B, int C, int L, float pre_slope, cudaStream_t st);
int launch_split_to_f32(const __nv_bfloat16* x, float* y, int64_t pitch, int B, int C, int L, cudaStream_t st);

}  // namespace pb
