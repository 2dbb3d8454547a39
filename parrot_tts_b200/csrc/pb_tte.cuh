// pb_tte.cuh - internal launch interface of the TTE kernels (tte_kernels.cu, gemm_tc.cu).
#pragma once
#include "pb_common.cuh"

namespace pb {

// C[M,N] = op(A)[M,K] . op(B) (+ bias[n]) (relu) (+ res[M,N]);  fp32.
//   A: row-major [M,K] with row stride lda.  With taps > 1 the A operand is the im2col view of a
//      channel-last sequence tensor X[B*L, Cin] (row stride lda): K = taps*Cin and
//      A[(b,t), j*Cin + ci] = X[(b, t + j - pad), ci] if 0 <= t + j - pad < L else 0  (Conv1d, 'same' padding).
//   B: b_kn == 0 -> [N,K] row-major (nn.Linear weight), row stride ldb
//      b_kn == 1 -> [K,N] row-major, row stride ldb (A's columns in [K, round_up(K,4)) must then be zero)
//   batching (attention): grid.z = batch*heads; operand z=(zb,zh) is offset by zb*s?b + zh*s?h elements.
struct GemmArgs {
    const float* A = nullptr; int64_t lda = 0;
    const float* B = nullptr; int64_t ldb = 0;
    float* C = nullptr; int64_t ldc = 0;
    const float* bias = nullptr;
    const float* res = nullptr; int64_t ldres = 0;
    int M = 0, N = 0, K = 0;
    int relu = 0;
    int b_kn = 0;
    int taps = 1, Cin = 0, L = 0, pad = 0;
    int nbatch = 1, H = 1;
    int64_t sAb = 0, sAh = 0, sBb = 0, sBh = 0, sCb = 0, sCh = 0;
};
int launch_sgemm(const GemmArgs& g, cudaStream_t st);

// out[b,s,:] = tok_emb[phones[b,s],:] + pe_row[:]      (parrot.py:94-95; fft.py:17-19)
// ragged mode (lens != null, L = padded length): pe_row is the pe TABLE and item b takes row lens[b]
int launch_embed_pos(const int64_t* phones, const float* tok_emb, const float* pe_row, const int32_t* lens, int L, int vocab, int M,
                     int D, float* out, int* err_flag, cudaStream_t st);
// y = LayerNorm(x) * gamma + beta, eps 1e-5, rows of D = 256
// lens / L (optional): rows are [B][L] and row (b, t >= lens[b]) is written as zeros (ragged mode: the next conv's zero padding)
int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, int M, int D, cudaStream_t st,
                     const int32_t* lens = nullptr, int L = 0);
// tensor-core path: LayerNorm / plain rows written as split-fp16 rows [M][2C] = [hi | lo] (pb_tc.cuh, fmt 1)
int launch_layernorm_split(const float* x, const float* gamma, const float* beta, uint16_t* y, int M, int D, cudaStream_t st,
                           const int32_t* lens = nullptr, int L = 0);
int launch_rows_to_split(const float* x, int64_t ldx, uint16_t* y, int64_t M, int C, cudaStream_t st);
// fused attention on tcgen05 (attn_tc.cu): qkv split-fp16 rows [B*L][2*3D] ([q|k|v] hi, then lo), key_valid u8 [B][L] or null ->
// out split-fp16 rows [B*L][2*D] (head h at columns h*128); head width must be 128
int launch_attention_tc(const uint16_t* qkv, const uint8_t* key_valid, uint16_t* out, int B, int L, int H, int D, cudaStream_t st);
// x[b,l,:] += row[idx[b],:]   (speaker embedding, parrot.py:98-99)
// lens (optional): rows (b, t >= lens[b]) are set to zero instead; table may then be null (zeroing only)
int launch_add_indexed_row(float* x, const float* table, const int64_t* idx, int n_rows, int B, int L, int D, int* err_flag, cudaStream_t st,
                           const int32_t* lens = nullptr);
// in-place masked softmax over the last dim of S [Z, L, Lp]: softmax(scale * s + (valid ? 0 : -inf)), pad columns -> 0
int launch_softmax(float* S, const uint8_t* key_valid, int Z, int H, int L, int Lp, float scale, cudaStream_t st);
// duration head: log_dur = mask(LN(x) . w + b); dur = clamp(round_half_even(exp(log_dur) - 1), 0)   (duration.py:36-48; parrot.py:82-86)
int launch_duration_head(const float* x, const float* gamma, const float* beta, const float* w, const float* b,
                         const uint8_t* src_mask, int M, int D, float* log_dur, int64_t* dur, cudaStream_t st);
// per item inclusive prefix sum of durations -> cum i32 [B,S], out_lens i64 [B]
int launch_duration_scan(const int64_t* dur, int B, int S, int32_t* cum, int64_t* out_lens, cudaStream_t st, int32_t* out_lens32 = nullptr);
// length regulator: out[b,t,:] = (t < len_b ? seq[b, s(t), :] : 0) + pe_row ; tgt_mask[b,t] = t <= len_b   (duration.py:6-24; data.py:18)
// ragged != 0: pe_row is the pe TABLE (item b takes row len_b) and tgt_mask[b,t] = t < len_b - every item as if it were run alone
int launch_length_regulate(const float* seq, const int32_t* cum, const float* pe_row, int B, int S, int D, int T,
                           float* out, uint8_t* tgt_mask, cudaStream_t st, int ragged = 0);
// ids[m] = argmax_n logits[m, n] (first maximum)
int launch_argmax(const float* logits, int M, int N, int64_t ldl, int64_t* ids, cudaStream_t st);
// w'[n][j*Cin + ci] = w[n][ci][j]   (Conv1d weight -> GEMM operand)
int launch_permute_conv_weight(const float* w, float* out, int N, int Cin, int K, cudaStream_t st);

}  // namespace pb
