/*
 * parrot_b200.h - C ABI of libparrot_b200.so: the B200 (sm_100a) implementation of the
 * Parrot-TTS inference hot path (TTE: chars -> HuBERT units; HiFi-GAN code generator: units -> wav).
 *
 * The reference has no FFI of its own for this path: its boundary is the Python nn.Module API
 * (SURVEY.md section 8b).  Each entry point below names the reference call it stands in for
 * (paths relative to the reference tree).  The Python mirror of those modules lives in
 * parrot_tts_b200/ and binds this header through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative pb_status; the message is available
 *     through pb_last_error() (thread-local).  No exceptions cross the ABI, no CPU fallback.
 *   - tensor arguments are DEVICE pointers on the current CUDA device unless the name ends in
 *     `_host`; memory is owned by the caller, the library keeps no caller pointer after return
 *     (weights are copied/packed into library-owned storage by *_finalize).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); launches are
 *     asynchronous on it.  Only the functions documented as synchronising wait for the device.
 *   - a handle is bound to the device current at *_create and is not thread-safe.
 *   - dtype codes: 0 = float32, 1 = int64.
 */
#ifndef PARROT_B200_H_
#define PARROT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 3

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_INVALID = -1,   /* bad argument / shape / missing weight */
    PB_ERR_INDEX = -2,     /* length >= max_len (reference modules/fft.py:18 raises IndexError), id out of range */
    PB_ERR_CUDA = -3,      /* CUDA runtime / driver failure */
    PB_ERR_STATE = -4,     /* call order violated (e.g. forward before finalize) */
    PB_ERR_NCCL = -5
} pb_status;

int pb_version(void);
const char* pb_last_error(void);

/* ------------------------------------------------------------------------------------------
 * HiFi-GAN code generator  (reference utils/vocoder/models.py:69-169)
 * ------------------------------------------------------------------------------------------ */
#define PB_VOC_MAX_UPS 8
#define PB_VOC_MAX_RESK 4
#define PB_VOC_MAX_DIL 4

typedef struct pb_voc_cfg {
    int32_t model_in_dim;              /* 256 = embedding_dim (units) + embedding_dim (speaker)   config.json "model_in_dim" */
    int32_t upsample_initial_channel;  /* 512 */
    int32_t num_upsamples;             /* 5 */
    int32_t upsample_rates[PB_VOC_MAX_UPS];         /* 5,4,4,2,2 */
    int32_t upsample_kernel_sizes[PB_VOC_MAX_UPS];  /* 11,8,8,4,4 */
    int32_t num_kernels;               /* 3 */
    int32_t resblock_kernel_sizes[PB_VOC_MAX_RESK]; /* 3,7,11 */
    int32_t num_dilations;             /* 3 */
    int32_t resblock_dilation_sizes[PB_VOC_MAX_RESK][PB_VOC_MAX_DIL]; /* 1,3,5 each */
    int32_t num_embeddings;            /* 1000 (0 = Generator only, no code path) */
    int32_t embedding_dim;             /* 128 */
    int32_t num_speakers;              /* 10  (0 = single speaker: no spkr table) */
    int32_t conv_mode;                 /* 0 = fp32 SIMT sliding-window kernels; 1 = tcgen05 implicit GEMM on split-bf16
                                          operands (3 MMAs per product, fp32-accurate: the default of the Python mirror);
                                          2 = tcgen05 single-pass bf16 (throughput mode, ~2^-9 relative error per product) */
} pb_voc_cfg;

typedef struct pb_voc pb_voc;

/* Generator.__init__ / CodeGenerator.__init__ (models.py:70-93,123-130) */
int pb_voc_create(const pb_voc_cfg* cfg, pb_voc** out);
void pb_voc_destroy(pb_voc* h);

/* load_state_dict (utils/vocoder/inference.py:104-109).  `key` uses the checkpoint names:
 * conv_pre|ups.N|resblocks.N.convs{1,2}.M|conv_post + .weight | .weight_g/.weight_v | .bias ; dict.weight ; spkr.weight.
 * `data` may be a host or device pointer (copied immediately). */
int pb_voc_set_weight(pb_voc* h, const char* key, const void* data, const int64_t* shape, int32_t ndim, int32_t dtype);

/* fold weight norm (w = g*v/||v||, norm over dims != 0 - what torch._weight_norm(v,g,0) computes;
 * models.py:17-28,75-93,113-119 remove_weight_norm) and pack the filters for the kernels. Synchronises. */
int pb_voc_finalize(pb_voc* h, void* stream);

/* CodeGenerator.forward(code=, spkr=) (models.py:153-169) when `code` != NULL, else
 * Generator.forward(x) (models.py:95-111) on a float feature map.
 *   x     f32 [B, model_in_dim, T]  (NULL when code given)
 *   code  i64 [B, T], spkr i64 [B] (or [B,1])
 *   lens  i32 [B] valid units per item, or NULL.  NULL reproduces the reference's padded-rectangle
 *         result; with lens every item equals its own batch-of-1 result (zero beyond its end).
 *   wav   f32 [B, 1, hop*T]  (hop = prod(upsample_rates) = 320)
 * Asynchronous on `stream`. */
int pb_voc_forward(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                   int32_t B, int32_t T, float* wav, void* stream);

/* nn.Embedding raises IndexError on an id outside its table (utils/vocoder/models.py:155-156).  The forward is asynchronous and its
 * embedding kernel clamps such ids (and flags them, live positions only); this call SYNCHRONISES `stream`, returns PB_ERR_INDEX if
 * any forward since the last check saw one, and clears the flag. */
int pb_voc_check_ids(pb_voc* h, void* stream);

/* pb_voc_forward with the generate() tail fused into conv_post's epilogue (utils/vocoder/inference.py:65-74): pcm i16 [B, hop*T] =
 * (int16) trunc(wav * 32768), the same expression pb_wav_to_pcm16 applies to the same float; wav f32 [B,1,hop*T] or NULL (then
 * the waveform never reaches HBM as f32).  conv_mode 0 needs wav (the cast is a separate launch there). */
int pb_voc_forward_pcm16(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                         int32_t B, int32_t T, float* wav, int16_t* pcm, void* stream);

/* generate() tail (utils/vocoder/inference.py:71-73): pcm = (int16) trunc(wav * 32768) (wraps like a C cast). */
int pb_wav_to_pcm16(const float* wav, int64_t n, int16_t* pcm, void* stream);

/* Peak normalisation of the generated PCM before it is written out - `librosa.util.normalize(audio.astype(np.float32))`
 * with its defaults (norm = inf over the utterance; utils/vocoder/inference.py:169, demo.ipynb cell 13):
 *   out[b, i] = float(pcm[b, i]) / max_i |pcm[b, i]|   (an all-zero item is returned unchanged)
 * pcm i16 [B, n]; peak i32 [B] (device, written: the per-item peak); out f32 [B, n].  Asynchronous on `stream`. */
int pb_pcm16_peak_normalize(const int16_t* pcm, int32_t B, int64_t n, int32_t* peak, float* out, void* stream);

/* Number of this library's kernel launches issued by the last pb_voc_forward / pb_tte_* call on the handle. */
int64_t pb_voc_last_launches(const pb_voc* h);
/* Changes whenever the handle's weight arena or workspace has been (re)allocated: device pointers recorded earlier - a captured
 * CUDA graph of pb_voc_forward - are stale once it differs (the workspace grows with the largest forward seen). */
int64_t pb_voc_memory_epoch(const pb_voc* h);

/* Section timing for the roofline report (bench.py): when enabled, pb_voc_forward records CUDA events on the
 * caller's stream at its section boundaries.  Sections: 0 = embedding + conv_pre; 1+2i = ups[i];
 * 2+2i = the 3 ResBlocks of stage i; last = conv_post + tanh.  pb_voc_section_ms waits for the last event. */
int pb_voc_set_timing(pb_voc* h, int32_t enable);
int pb_voc_section_count(const pb_voc* h);
int pb_voc_section_ms(pb_voc* h, float* ms_host, int32_t n);

/* Per-op entry points (unit tests; each is the kernel the model path uses).
 * conv1d: y = post( conv1d(pre(x), w, bias, dilation, 'same') [+ res] ), x/y/res f32 [B, C, L] contiguous.
 *   pre_slope: leaky-relu slope applied to the input (1.0 = none); post_slope likewise on the output;
 *   w f32 [Cout, Cin, K] torch layout.  (F.conv1d / F.leaky_relu in models.py:31-38,96,107-109) */
int pb_conv1d(const float* x, const float* w, const float* bias, const float* res, float* y,
              int32_t B, int32_t Cin, int32_t Cout, int32_t L, int32_t K, int32_t dilation,
              float pre_slope, float post_slope, void* stream);
/* conv_transpose1d(x, w[Cin,Cout,K], bias, stride=u, padding=(K-u)/2), x already activated (models.py:98-99). */
int pb_conv_transpose1d(const float* x, const float* w, const float* bias, float* y,
                        int32_t B, int32_t Cin, int32_t Cout, int32_t L, int32_t K, int32_t stride,
                        float pre_slope, void* stream);

/* Tensor-core variants of the two ops above (tcgen05 implicit GEMM on split-bf16 operands, fp32 accumulation in TMEM;
 * see parrot_tts_b200/csrc/pb_tc.cuh).  Same tensors and semantics; Cin in {16, 32, 64k}, Cout a multiple of 16.
 * `flags` bit 0: single-pass bf16 products (throughput mode) instead of the fp32-accurate 3-pass split. */
int pb_conv1d_tc(const float* x, const float* w, const float* bias, const float* res, float* y,
                 int32_t B, int32_t Cin, int32_t Cout, int32_t L, int32_t K, int32_t dilation,
                 float pre_slope, float post_slope, int32_t flags, void* stream);
int pb_conv_transpose1d_tc(const float* x, const float* w, const float* bias, float* y,
                           int32_t B, int32_t Cin, int32_t Cout, int32_t L, int32_t K, int32_t stride,
                           float pre_slope, int32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * TTE  (reference modules/parrot.py, modules/fft.py, modules/duration.py)
 * ------------------------------------------------------------------------------------------ */
typedef struct pb_tte_cfg {
    int32_t src_vocab_size;
    int32_t src_pad_idx;
    int32_t n_speaker;       /* <= 1: no speaker embedding (parrot.py:28) */
    int32_t d_model;         /* 256 */
    int32_t n_head;          /* 2 */
    int32_t n_enc_layer;     /* 4 */
    int32_t n_dec_layer;     /* 4 */
    int32_t conv_n_filter;   /* 1024 */
    int32_t conv_k1;         /* 9 */
    int32_t conv_k2;         /* 1 */
    int32_t dp_n_filter;     /* 256 */
    int32_t dp_kernel;       /* 3 */
    int32_t n_codes;         /* 1000 */
    int32_t max_len;         /* 3500: rows of pos_emb.pe */
    int32_t fold_projections;/* 1: fold mha.in_proj*qkv and wo*out_proj at finalize (exact in real arithmetic) */
    int32_t gemm_mode;       /* 0 = fp32 SIMT, 1 = tcgen05 on split-fp16 operands, 3 products, fp32 accumulation (fp32-class; default).
                              * There is no single-pass 16-bit mode for the TTE: unit ids are an integer output that must be bit-exact. */
} pb_tte_cfg;

typedef struct pb_tte pb_tte;

/* Parrot.__init__ (parrot.py:13-65) */
int pb_tte_create(const pb_tte_cfg* cfg, pb_tte** out);
void pb_tte_destroy(pb_tte* h);
/* load_state_dict; keys exactly as in Parrot.state_dict(), without the Lightning "parrot." prefix. */
int pb_tte_set_weight(pb_tte* h, const char* key, const void* data, const int64_t* shape, int32_t ndim, int32_t dtype);
int pb_tte_finalize(pb_tte* h, void* stream);

/* Parrot.forward up to and including the duration predictor and rounding (parrot.py:94-100,77-86;
 * duration.py:41-48).  phones i64 [B,S]; src_mask u8 [B,S] (1 = real char); speaker i64 [B] or NULL.
 * dur_in: NULL, or teacher-forced durations i64 [B,S] (batch['duration'], parrot.py:102).
 * src_lens: NULL = the reference's padded-batch semantics (an item's result depends on the padded length: the "positional" row is
 *   pe[S], fft.py:18, padded positions feed the k = 9 convolutions, and the target mask is t <= len, data.py:18), or i32 [B] =
 *   RAGGED mode: item b is computed exactly as the reference computes it ALONE (batch of one, its inference mode, inference.py:34):
 *   its own pe rows pe[src_lens[b]] / pe[out_lens[b]], zeros past its own end wherever a convolution looks, and exactly out_lens[b]
 *   target positions.  src_mask must then be the prefix mask t < src_lens[b].  pb_tte_decode follows the mode of the last encode.
 * Outputs: log_dur f32 [B,S]; dur i64 [B,S]; out_lens i64 [B] (sum of durations per item).
 * *T_host = max_b out_lens (the length-regulated length).  SYNCHRONISES (T is data dependent).
 * Returns PB_ERR_INDEX if S >= max_len or T >= max_len (fft.py:18). */
int pb_tte_encode(pb_tte* h, const int64_t* phones, const uint8_t* src_mask, const int64_t* speaker,
                  const int64_t* dur_in, const int32_t* src_lens, int32_t B, int32_t S,
                  float* log_dur, int64_t* dur, int64_t* out_lens, int32_t* T_host, void* stream);

/* length_regulator + pos_emb + decoder + head (+ argmax of Parrot.infer) (duration.py:6-24; parrot.py:106-108,115)
 * for the batch just encoded.  logits f32 [B,T,n_codes] or NULL; ids i64 [B,T] or NULL;
 * tgt_mask u8 [B,T] out (t <= out_lens[b], data.py:18; ragged mode: t < out_lens[b]).
 * tgt_mask_in: NULL, or the caller's mask u8 [B,T] (batch['tgt_mask'] of the teacher-forced path, parrot.py:104 ->
 * duration.py:22-24: the reference returns a supplied mask unchanged and the decoder attends with it); it then replaces the
 * generated one for the decoder's key padding and is what tgt_mask returns.  Asynchronous. */
int pb_tte_decode(pb_tte* h, int32_t B, int32_t S, int32_t T, const uint8_t* tgt_mask_in, float* logits, int64_t* ids, uint8_t* tgt_mask,
                  void* stream);

int64_t pb_tte_last_launches(const pb_tte* h);

/* Per-op entry points (unit tests and the FFTBlock / DurationPredictor / length_regulator Python mirrors). */
/* nn.LayerNorm(D) eps 1e-5 over rows of x [M, D] (fft.py:91-92). */
int pb_layernorm(const float* x, const float* gamma, const float* beta, float* y, int32_t M, int32_t D, void* stream);
/* FFTBlock.forward (fft.py:94-100) with layer `layer` of the encoder (stack 0) / decoder (stack 1):
 * x,y f32 [B,L,D]; key_valid u8 [B,L] (1 = attend) or NULL. */
int pb_tte_fft_block(pb_tte* h, int32_t stack, int32_t layer, const float* x, const uint8_t* key_valid,
                     int32_t B, int32_t L, float* y, void* stream);
/* DurationPredictor.forward (duration.py:41-48) and the rounding of parrot.py:82-86 on x f32 [B,S,D]; dur may be NULL. */
int pb_tte_duration(pb_tte* h, const float* x, const uint8_t* src_mask, int32_t B, int32_t S, float* log_dur, int64_t* dur, void* stream);
/* length_regulator (duration.py:6-24) + optional pos row add: seq f32 [B,S,D], dur i64 [B,S] ->
 * out f32 [B,T,D] (zero rows past each item's end), tgt_mask u8 [B,T]. pe_row: f32 [D] or NULL. */
int pb_length_regulate(const float* seq, const int64_t* dur, const float* pe_row, int32_t B, int32_t S, int32_t D, int32_t T,
                       float* out, uint8_t* tgt_mask, int64_t* out_lens, void* stream);
/* Y[M,N] = X[M,K] . W[N,K]^T + bias, fp32 (nn.Linear). mode as pb_tte_cfg.gemm_mode. */
int pb_linear(const float* x, const float* w, const float* bias, float* y, int32_t M, int32_t N, int32_t K, int32_t mode, void* stream);

/* mel_spectrogram (utils/vocoder/dataset.py:43-69; spectral_normalize_torch :85-95) of audio on the device - the validation metric
 * beside the sample tolerance: reflect padding by (n_fft - hop) / 2, STFT with the periodic Hann window
 * (center = False; win must equal n_fft, a power of two in 256..2048), sqrt(re^2 + im^2 + 1e-9), mel_basis f32 [n_mels][1 + n_fft/2]
 * (librosa.filters.mel layout), log(clamp(., 1e-5)).  y f32 [B, N] -> out f32 [B, n_mels, *frames].  out == NULL: size query only. */
int pb_mel_spectrogram(const float* y, int32_t B, int64_t N, const float* mel_basis, int32_t n_fft, int32_t hop, int32_t win,
                       int32_t n_mels, float* out, int32_t* frames, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, weights broadcast once at init.
 * `comm` is an ncclComm_t passed as void*.  Must be called after *_finalize on root and after
 * *_create (+ shapes known) elsewhere: non-root handles receive the packed arena and become finalized.
 * ------------------------------------------------------------------------------------------ */
int pb_voc_arena_bytes(pb_voc* h, int64_t* bytes);
void* pb_voc_arena_ptr(pb_voc* h);
int pb_tte_arena_bytes(pb_tte* h, int64_t* bytes);
void* pb_tte_arena_ptr(pb_tte* h);
/* non-root ranks: declare the arena complete after it has been received */
int pb_voc_mark_finalized(pb_voc* h);
int pb_tte_mark_finalized(pb_tte* h);

#ifdef __cplusplus
}
#endif
#endif /* PARROT_B200_H_ */
