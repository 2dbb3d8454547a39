// tte_api.cu - C ABI of the TTE stage: handle, checkpoint ingestion (projection folding, conv weight
// permutation), encode / decode orchestration.  Reference: modules/parrot.py:13-120, modules/fft.py,
// modules/duration.py, modules/data.py:8-20.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "pb_runtime.cuh"
#include "pb_tc.cuh"
#include "pb_tte.cuh"

using namespace pb;

namespace {
struct Slot { size_t off = 0; int64_t numel = 0; std::vector<int64_t> shape; bool loaded = false; bool checkpoint = false; };
}

struct pb_tte {
    pb_tte_cfg cfg;
    int device = 0;
    bool finalized = false;
    std::map<std::string, Slot> slots;    // checkpoint tensors and derived (packed / folded) tensors
    DevBuf arena;
    size_t arena_bytes = 0;
    DevBuf work;                          // per-call activations
    DevBuf keep;                          // encoder output, cum durations, masks kept between encode and decode
    float* enc_out = nullptr; int32_t* cum = nullptr; int64_t* out_lens_dev = nullptr; int* err_flag = nullptr;
    int32_t* tgt_lens32 = nullptr;        // unit count per item as i32 (decoder-side lens of the ragged mode)
    int enc_B = 0, enc_S = 0, enc_T = -1;
    bool enc_ragged = false;              // the last encode ran in ragged mode: every item as if it were run alone
    int64_t last_launches = 0;
    float* P(const std::string& k) {
        auto it = slots.find(k);
        return it == slots.end() ? nullptr : reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(arena.ptr) + it->second.off);
    }
};

static void add_slot(pb_tte* h, const std::string& key, std::vector<int64_t> shape, bool checkpoint) {
    Slot s;
    s.shape = shape;
    s.numel = 1;
    for (auto d : shape) s.numel *= d;
    s.off = h->arena_bytes;
    s.checkpoint = checkpoint;
    h->arena_bytes += ((size_t)s.numel * sizeof(float) + 255) / 256 * 256;
    h->slots[key] = s;
}

static std::string layer_prefix(int stack, int i) {
    return std::string(stack == 0 ? "encoder_layers." : "decoder_layers.") + std::to_string(i);
}

extern "C" int pb_tte_create(const pb_tte_cfg* cfg, pb_tte** out) {
    PB_REQUIRE(cfg && out, PB_ERR_INVALID, "pb_tte_create: null argument");
    PB_REQUIRE(cfg->d_model == 256, PB_ERR_INVALID, "pb_tte_create: d_model=%d unsupported (kernels are built for 256)", cfg->d_model);
    PB_REQUIRE(cfg->n_head > 0 && cfg->d_model % cfg->n_head == 0 && (cfg->d_model / cfg->n_head) % 4 == 0, PB_ERR_INVALID, "pb_tte_create: bad n_head=%d", cfg->n_head);
    PB_REQUIRE(cfg->conv_k1 % 2 == 1 && cfg->conv_k2 % 2 == 1, PB_ERR_INVALID, "pb_tte_create: FFN kernel sizes must be odd");
    PB_REQUIRE(cfg->dp_n_filter == 256 && cfg->dp_kernel == 3, PB_ERR_INVALID, "pb_tte_create: duration predictor must be 256 filters, kernel 3");
    PB_REQUIRE(cfg->conv_n_filter % 4 == 0 && cfg->n_codes > 0 && cfg->max_len > 1 && cfg->src_vocab_size > 0, PB_ERR_INVALID, "pb_tte_create: bad sizes");
    PB_REQUIRE(cfg->gemm_mode == 0 || cfg->gemm_mode == 1, PB_ERR_INVALID, "pb_tte_create: gemm_mode=%d (0 = fp32 SIMT, 1 = tcgen05 split-fp16)", cfg->gemm_mode);
    if (cfg->gemm_mode == 1) {
        PB_REQUIRE(cfg->fold_projections, PB_ERR_INVALID, "pb_tte_create: gemm_mode 1 needs fold_projections = 1 (the unfolded reference mode runs on the fp32 SIMT path)");
        PB_REQUIRE(cfg->d_model / cfg->n_head == 128, PB_ERR_INVALID,
                   "pb_tte_create: gemm_mode 1 needs a head width of 128 (d_model %d / n_head %d = %d): the fused attention kernel is built for it; use gemm_mode 0",
                   cfg->d_model, cfg->n_head, cfg->d_model / cfg->n_head);
        PB_REQUIRE(cfg->conv_n_filter % 256 == 0 && cfg->conv_k2 == 1 && (cfg->conv_k1 - 1) <= 128, PB_ERR_INVALID,
                   "pb_tte_create: gemm_mode 1 needs conv_n_filter %% 256 == 0 and conv_kernel_sizes [k, 1]");
    }
    pb_tte* h = new pb_tte();
    h->cfg = *cfg;
    cudaGetDevice(&h->device);
    const int D = cfg->d_model, F = cfg->conv_n_filter;
    add_slot(h, "pos_emb.pe", {cfg->max_len, D}, true);
    add_slot(h, "tok_emb.weight", {cfg->src_vocab_size, D}, true);
    if (cfg->n_speaker > 1) add_slot(h, "speaker_emb.weight", {cfg->n_speaker, D}, true);
    add_slot(h, "duration_predictor.layers.0.conv.weight", {256, D, 3}, true);
    add_slot(h, "duration_predictor.layers.0.conv.bias", {256}, true);
    add_slot(h, "duration_predictor.layers.2.weight", {256}, true);
    add_slot(h, "duration_predictor.layers.2.bias", {256}, true);
    add_slot(h, "duration_predictor.layers.4.conv.weight", {256, 256, 3}, true);
    add_slot(h, "duration_predictor.layers.4.conv.bias", {256}, true);
    add_slot(h, "duration_predictor.layers.6.weight", {256}, true);
    add_slot(h, "duration_predictor.layers.6.bias", {256}, true);
    add_slot(h, "duration_predictor.proj.weight", {1, 256}, true);
    add_slot(h, "duration_predictor.proj.bias", {1}, true);
    add_slot(h, "duration_predictor.layers.0.conv.packed", {256, 3 * D}, false);
    add_slot(h, "duration_predictor.layers.4.conv.packed", {256, 3 * 256}, false);
    for (int stack = 0; stack < 2; ++stack) {
        const int n = stack == 0 ? cfg->n_enc_layer : cfg->n_dec_layer;
        for (int i = 0; i < n; ++i) {
            const std::string p = layer_prefix(stack, i);
            add_slot(h, p + ".attention.qkv.weight", {3 * D, D}, true);
            add_slot(h, p + ".attention.mha.in_proj_weight", {3 * D, D}, true);
            add_slot(h, p + ".attention.mha.out_proj.weight", {D, D}, true);
            add_slot(h, p + ".attention.wo.weight", {D, D}, true);
            add_slot(h, p + ".convlayer.conv1.weight", {F, D, cfg->conv_k1}, true);
            add_slot(h, p + ".convlayer.conv1.bias", {F}, true);
            add_slot(h, p + ".convlayer.conv2.weight", {D, F, cfg->conv_k2}, true);
            add_slot(h, p + ".convlayer.conv2.bias", {D}, true);
            add_slot(h, p + ".attn_norm.weight", {D}, true);
            add_slot(h, p + ".attn_norm.bias", {D}, true);
            add_slot(h, p + ".conv_norm.weight", {D}, true);
            add_slot(h, p + ".conv_norm.bias", {D}, true);
            add_slot(h, p + ".attention.qkv_eff", {3 * D, D}, false);
            add_slot(h, p + ".attention.wo_eff", {D, D}, false);
            add_slot(h, p + ".convlayer.conv1.packed", {F, cfg->conv_k1 * D}, false);
            add_slot(h, p + ".convlayer.conv2.packed", {D, cfg->conv_k2 * F}, false);
            if (cfg->gemm_mode == 1) {   // tcgen05 operands (split-fp16, packed by tc_pack_conv_weight): same element count in bytes/4
                auto tc_floats = [](int Cin, int N, int K) { return (int64_t)(tc_packed_weight_bytes(Cin, N, K) / sizeof(float)); };
                add_slot(h, p + ".attention.qkv_eff.tc", {tc_floats(D, 3 * D, 1)}, false);
                add_slot(h, p + ".attention.wo_eff.tc", {tc_floats(D, D, 1)}, false);
                add_slot(h, p + ".convlayer.conv1.tc", {tc_floats(D, F, cfg->conv_k1)}, false);
                add_slot(h, p + ".convlayer.conv2.tc", {tc_floats(F, D, cfg->conv_k2)}, false);
            }
        }
    }
    add_slot(h, "head.weight", {cfg->n_codes, D}, true);
    add_slot(h, "head.bias", {cfg->n_codes}, true);
    if (cfg->gemm_mode == 1) {
        const int np = (int)round_up(cfg->n_codes, 256);
        add_slot(h, "head.weight.pad", {np, D}, false);      // zero rows past n_codes
        add_slot(h, "head.weight.tc", {(int64_t)(tc_packed_weight_bytes(D, np, 1) / sizeof(float))}, false);
        add_slot(h, "head.bias.pad", {np}, false);
    }
    int st = h->arena.reserve(h->arena_bytes);
    if (st != PB_OK) { delete h; return st; }
    *out = h;
    return PB_OK;
}

extern "C" void pb_tte_destroy(pb_tte* h) {
    if (!h) return;
    h->arena.release();
    h->work.release();
    h->keep.release();
    delete h;
}

extern "C" int pb_tte_set_weight(pb_tte* h, const char* key, const void* data, const int64_t* shape, int32_t ndim, int32_t dtype) {
    PB_REQUIRE(h && key && data && shape, PB_ERR_INVALID, "pb_tte_set_weight: null argument");
    PB_REQUIRE(dtype == 0, PB_ERR_INVALID, "pb_tte_set_weight(%s): only float32 tensors are accepted", key);
    auto it = h->slots.find(key);
    PB_REQUIRE(it != h->slots.end() && it->second.checkpoint, PB_ERR_INVALID, "pb_tte_set_weight: unexpected key '%s'", key);
    Slot& s = it->second;
    bool same = (int)s.shape.size() == ndim;
    for (int i = 0; same && i < ndim; ++i) same = s.shape[i] == shape[i];
    PB_REQUIRE(same, PB_ERR_INVALID, "pb_tte_set_weight(%s): shape mismatch (expected %d dims, first dim %lld)", key, (int)s.shape.size(), (long long)s.shape[0]);
    PB_CUDA(cudaMemcpy(h->P(key), data, (size_t)s.numel * sizeof(float), cudaMemcpyDefault));
    s.loaded = true;
    h->finalized = false;
    return PB_OK;
}

static bool layer_complete(pb_tte* h, const std::string& p) {
    static const char* names[] = {".attention.qkv.weight", ".attention.mha.in_proj_weight", ".attention.mha.out_proj.weight",
                                  ".attention.wo.weight", ".convlayer.conv1.weight", ".convlayer.conv1.bias",
                                  ".convlayer.conv2.weight", ".convlayer.conv2.bias", ".attn_norm.weight", ".attn_norm.bias",
                                  ".conv_norm.weight", ".conv_norm.bias"};
    for (auto n : names) if (!h->slots[p + n].loaded) return false;
    return true;
}

extern "C" int pb_tte_finalize(pb_tte* h, void* stream) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_tte_finalize: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const pb_tte_cfg& c = h->cfg;
    const int D = c.d_model, F = c.conv_n_filter;
    for (int stack = 0; stack < 2; ++stack) {
        const int n = stack == 0 ? c.n_enc_layer : c.n_dec_layer;
        for (int i = 0; i < n; ++i) {
            const std::string p = layer_prefix(stack, i);
            PB_REQUIRE(layer_complete(h, p), PB_ERR_INVALID, "pb_tte_finalize: layer %s is missing weights", p.c_str());
            // qkv_eff[i] = in_proj[i] . qkv[i]  (x.Wqkv_i^T.Win_i^T = x.(Win_i.Wqkv_i)^T), i = q,k,v
            for (int t = 0; t < 3; ++t) {
                GemmArgs g;
                g.A = h->P(p + ".attention.mha.in_proj_weight") + (size_t)t * D * D; g.lda = D;
                g.B = h->P(p + ".attention.qkv.weight") + (size_t)t * D * D; g.ldb = D; g.b_kn = 1;
                g.C = h->P(p + ".attention.qkv_eff") + (size_t)t * D * D; g.ldc = D;
                g.M = D; g.N = D; g.K = D;
                PB_TRY(launch_sgemm(g, st));
            }
            {   // wo_eff = wo . out_proj
                GemmArgs g;
                g.A = h->P(p + ".attention.wo.weight"); g.lda = D;
                g.B = h->P(p + ".attention.mha.out_proj.weight"); g.ldb = D; g.b_kn = 1;
                g.C = h->P(p + ".attention.wo_eff"); g.ldc = D;
                g.M = D; g.N = D; g.K = D;
                PB_TRY(launch_sgemm(g, st));
            }
            PB_TRY(launch_permute_conv_weight(h->P(p + ".convlayer.conv1.weight"), h->P(p + ".convlayer.conv1.packed"), F, D, c.conv_k1, st));
            PB_TRY(launch_permute_conv_weight(h->P(p + ".convlayer.conv2.weight"), h->P(p + ".convlayer.conv2.packed"), D, F, c.conv_k2, st));
            if (c.gemm_mode == 1) {
                auto TC = [&](const std::string& k) { return reinterpret_cast<uint8_t*>(h->P(k)); };
                PB_TRY(tc_pack_conv_weight(h->P(p + ".attention.qkv_eff"), TC(p + ".attention.qkv_eff.tc"), 3 * D, D, 1, st, 1));
                PB_TRY(tc_pack_conv_weight(h->P(p + ".attention.wo_eff"), TC(p + ".attention.wo_eff.tc"), D, D, 1, st, 1));
                PB_TRY(tc_pack_conv_weight(h->P(p + ".convlayer.conv1.weight"), TC(p + ".convlayer.conv1.tc"), F, D, c.conv_k1, st, 1));
                PB_TRY(tc_pack_conv_weight(h->P(p + ".convlayer.conv2.weight"), TC(p + ".convlayer.conv2.tc"), D, F, c.conv_k2, st, 1));
            }
        }
    }
    if (h->slots["duration_predictor.layers.0.conv.weight"].loaded && h->slots["duration_predictor.layers.4.conv.weight"].loaded) {
        PB_TRY(launch_permute_conv_weight(h->P("duration_predictor.layers.0.conv.weight"), h->P("duration_predictor.layers.0.conv.packed"), 256, D, 3, st));
        PB_TRY(launch_permute_conv_weight(h->P("duration_predictor.layers.4.conv.weight"), h->P("duration_predictor.layers.4.conv.packed"), 256, 256, 3, st));
    }
    if (c.gemm_mode == 1 && h->slots["head.weight"].loaded && h->slots["head.bias"].loaded) {
        const int np = (int)round_up(c.n_codes, 256);
        PB_CUDA(cudaMemsetAsync(h->P("head.weight.pad"), 0, (size_t)np * D * sizeof(float), st));
        PB_CUDA(cudaMemcpyAsync(h->P("head.weight.pad"), h->P("head.weight"), (size_t)c.n_codes * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PB_CUDA(cudaMemsetAsync(h->P("head.bias.pad"), 0, (size_t)np * sizeof(float), st));
        PB_CUDA(cudaMemcpyAsync(h->P("head.bias.pad"), h->P("head.bias"), (size_t)c.n_codes * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PB_TRY(tc_pack_conv_weight(h->P("head.weight.pad"), reinterpret_cast<uint8_t*>(h->P("head.weight.tc")), np, D, 1, st, 1));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    h->finalized = true;
    return PB_OK;
}

extern "C" int pb_tte_arena_bytes(pb_tte* h, int64_t* bytes) {
    PB_REQUIRE(h && bytes, PB_ERR_INVALID, "pb_tte_arena_bytes: null argument");
    *bytes = (int64_t)h->arena_bytes;
    return PB_OK;
}
extern "C" void* pb_tte_arena_ptr(pb_tte* h) { return h ? h->arena.ptr : nullptr; }
extern "C" int pb_tte_mark_finalized(pb_tte* h) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_tte_mark_finalized: null handle");
    for (auto& kv : h->slots) kv.second.loaded = true;
    h->finalized = true;
    return PB_OK;
}
extern "C" int64_t pb_tte_last_launches(const pb_tte* h) { return h ? h->last_launches : 0; }

// ---- one FFT block: y = FFTBlock(x)  (fft.py:94-100) ------------------------------------------------
struct BlockWork {
    float *xn, *qkv, *qkv1, *S, *o, *y1, *hbuf, *f;
    uint16_t *xn_s, *o_s, *f_s, *qkv_s;     // gemm_mode 1: split-fp16 GEMM operands [M][2C]
};
static size_t block_work_plan(Carver& cv, BlockWork& w, int B, int L, int D, int F, int H, bool unfolded, bool tc = false) {
    const size_t M = (size_t)B * L;
    const size_t Lp = (size_t)round_up(L, 4);
    w.xn = cv.take<float>(M * D);
    w.qkv = tc ? nullptr : cv.take<float>(M * 3 * D);
    w.qkv1 = unfolded ? cv.take<float>(M * 3 * D) : nullptr;
    w.S = tc ? nullptr : cv.take<float>((size_t)B * H * L * Lp);      // the tensor-core attention never materialises the scores
    w.o = cv.take<float>(M * D);
    w.qkv_s = tc ? cv.take<uint16_t>(M * 6 * D) : nullptr;
    w.y1 = unfolded ? cv.take<float>(M * D) : nullptr;
    w.hbuf = cv.take<float>(M * D);
    w.f = tc ? nullptr : cv.take<float>(M * F);
    w.xn_s = tc ? cv.take<uint16_t>(M * 2 * D) : nullptr;
    w.o_s = tc ? cv.take<uint16_t>(M * 2 * D) : nullptr;
    w.f_s = tc ? cv.take<uint16_t>(M * 2 * F) : nullptr;
    return cv.off;
}

// One GEMM / Conv1d of the TTE on the tensor-core kernel of the vocoder (tc_conv.cu): x split-fp16 [B][L][2*Cin] ->
// out split-fp16 [B][L][2*N] and/or out_f32 [B][L][N];  acc = conv(x, w) + bias (+ res_f32), then ReLU when relu != 0.
static int tte_tc_gemm(const uint16_t* x, const uint8_t* w, const float* bias, const float* res_f32, uint16_t* out, float* out_f32,
                       int B, int L, int Cin, int N, int K, int relu, cudaStream_t st) {
    TcConvArgs a;
    a.x = reinterpret_cast<const __nv_bfloat16*>(x); a.w = w; a.bias = bias; a.res_f32 = res_f32;
    a.out = reinterpret_cast<__nv_bfloat16*>(out); a.out_f32 = out_f32;
    a.B = B; a.L = L; a.Cin = Cin; a.N = N; a.out_C = N; a.U = 1; a.K = K; a.D = 1;
    a.post_slope = relu ? 0.f : 1.f;       // max(u, 0*u) = ReLU
    a.fmt = 1;
    return launch_tc_conv(a, st);
}

// FFTBlock.forward (fft.py:94-100) with the projections and the FFN convolutions on tcgen05 (split-fp16 operands, fp32
// accumulation in TMEM, 3 passes per product); the residual stream, LayerNorm statistics and softmax stay fp32.
static int run_fft_block_tc(pb_tte* h, const std::string& p, const float* x, const uint8_t* key_valid, int B, int L,
                            const BlockWork& w, float* y, cudaStream_t st, const int32_t* lens) {
    const pb_tte_cfg& c = h->cfg;
    const int D = c.d_model, F = c.conv_n_filter, H = c.n_head;
    const int M = B * L;
    auto TC = [&](const std::string& k) { return reinterpret_cast<const uint8_t*>(h->P(k)); };
    // lens (ragged mode): LayerNorm rows past an item's end are written as zeros - the zero padding its k = 9 convolution sees
    // when the item is run alone; keys past the end are masked by key_valid, so nothing else of a short item sees the padding
    PB_TRY(launch_layernorm_split(x, h->P(p + ".attn_norm.weight"), h->P(p + ".attn_norm.bias"), w.xn_s, M, D, st, lens, L));
    // q | k | v = LN(x) . Wqkv_eff^T   (rows flattened: a 1-tap GEMM has no sequence edges), kept as split-fp16 rows
    PB_TRY(tte_tc_gemm(w.xn_s, TC(p + ".attention.qkv_eff.tc"), nullptr, nullptr, w.qkv_s, nullptr, 1, M, D, 3 * D, 1, 0, st));
    // softmax(q.k^T / sqrt(dh) + mask) . v per (batch, head), fused on tcgen05 (attn_tc.cu): scores stay in tensor memory
    PB_TRY(launch_attention_tc(w.qkv_s, key_valid, w.o_s, B, L, H, D, st));
    // h = x + attn . Wo_eff^T
    PB_TRY(tte_tc_gemm(w.o_s, TC(p + ".attention.wo_eff.tc"), nullptr, x, nullptr, w.hbuf, 1, M, D, D, 1, 0, st));
    PB_TRY(launch_layernorm_split(w.hbuf, h->P(p + ".conv_norm.weight"), h->P(p + ".conv_norm.bias"), w.xn_s, M, D, st, lens, L));
    // f = relu(conv1(LN(h)))  - per item ('same' zero padding at each sequence's own ends, fft.py:65-70)
    PB_TRY(tte_tc_gemm(w.xn_s, TC(p + ".convlayer.conv1.tc"), h->P(p + ".convlayer.conv1.bias"), nullptr, w.f_s, nullptr, B, L, D, F,
                       c.conv_k1, 1, st));
    // y = h + conv2(f)
    PB_TRY(tte_tc_gemm(w.f_s, TC(p + ".convlayer.conv2.tc"), h->P(p + ".convlayer.conv2.bias"), w.hbuf, nullptr, y, 1, M, F, D, 1, 0, st));
    return PB_OK;
}

static int run_fft_block(pb_tte* h, const std::string& p, const float* x, const uint8_t* key_valid, int B, int L,
                         const BlockWork& w, float* y, cudaStream_t st, const int32_t* lens = nullptr) {
    const pb_tte_cfg& c = h->cfg;
    const int D = c.d_model, F = c.conv_n_filter, H = c.n_head, dh = D / H;
    const int M = B * L;
    const int Lp = (int)round_up(L, 4);
    const bool unfolded = !c.fold_projections;
    if (c.gemm_mode == 1) return run_fft_block_tc(h, p, x, key_valid, B, L, w, y, st, lens);
    PB_TRY(launch_layernorm(x, h->P(p + ".attn_norm.weight"), h->P(p + ".attn_norm.bias"), w.xn, M, D, st, lens, L));
    const float* qkv = w.qkv;
    {
        GemmArgs g;
        g.A = w.xn; g.lda = D; g.B = h->P(p + (unfolded ? ".attention.qkv.weight" : ".attention.qkv_eff")); g.ldb = D;
        g.C = w.qkv; g.ldc = 3 * D; g.M = M; g.N = 3 * D; g.K = D;
        PB_TRY(launch_sgemm(g, st));
    }
    if (unfolded) {
        for (int t = 0; t < 3; ++t) {
            GemmArgs g;
            g.A = w.qkv + t * D; g.lda = 3 * D; g.B = h->P(p + ".attention.mha.in_proj_weight") + (size_t)t * D * D; g.ldb = D;
            g.C = w.qkv1 + t * D; g.ldc = 3 * D; g.M = M; g.N = D; g.K = D;
            PB_TRY(launch_sgemm(g, st));
        }
        qkv = w.qkv1;
    }
    {   // S = Q.K^T per (batch, head)
        GemmArgs g;
        g.A = qkv; g.lda = 3 * D; g.sAb = (int64_t)L * 3 * D; g.sAh = dh;
        g.B = qkv + D; g.ldb = 3 * D; g.sBb = (int64_t)L * 3 * D; g.sBh = dh;
        g.C = w.S; g.ldc = Lp; g.sCb = (int64_t)H * L * Lp; g.sCh = (int64_t)L * Lp;
        g.M = L; g.N = L; g.K = dh; g.nbatch = B * H; g.H = H;
        PB_TRY(launch_sgemm(g, st));
    }
    PB_TRY(launch_softmax(w.S, key_valid, B * H, H, L, Lp, sqrtf(1.0f / (float)dh), st));
    {   // O = P.V
        GemmArgs g;
        g.A = w.S; g.lda = Lp; g.sAb = (int64_t)H * L * Lp; g.sAh = (int64_t)L * Lp;
        g.B = qkv + 2 * D; g.ldb = 3 * D; g.b_kn = 1; g.sBb = (int64_t)L * 3 * D; g.sBh = dh;
        g.C = w.o; g.ldc = D; g.sCb = (int64_t)L * D; g.sCh = dh;
        g.M = L; g.N = dh; g.K = L; g.nbatch = B * H; g.H = H;
        PB_TRY(launch_sgemm(g, st));
    }
    if (unfolded) {
        GemmArgs g;
        g.A = w.o; g.lda = D; g.B = h->P(p + ".attention.mha.out_proj.weight"); g.ldb = D; g.C = w.y1; g.ldc = D;
        g.M = M; g.N = D; g.K = D;
        PB_TRY(launch_sgemm(g, st));
    }
    {   // h = x + attn_out . Wo^T
        GemmArgs g;
        g.A = unfolded ? w.y1 : w.o; g.lda = D; g.B = h->P(p + (unfolded ? ".attention.wo.weight" : ".attention.wo_eff")); g.ldb = D;
        g.C = w.hbuf; g.ldc = D; g.res = x; g.ldres = D; g.M = M; g.N = D; g.K = D;
        PB_TRY(launch_sgemm(g, st));
    }
    PB_TRY(launch_layernorm(w.hbuf, h->P(p + ".conv_norm.weight"), h->P(p + ".conv_norm.bias"), w.xn, M, D, st, lens, L));
    {   // f = relu(conv1(hn))
        GemmArgs g;
        g.A = w.xn; g.lda = D; g.B = h->P(p + ".convlayer.conv1.packed"); g.ldb = (int64_t)c.conv_k1 * D;
        g.C = w.f; g.ldc = F; g.bias = h->P(p + ".convlayer.conv1.bias"); g.relu = 1;
        g.M = M; g.N = F; g.K = c.conv_k1 * D; g.taps = c.conv_k1; g.Cin = D; g.L = L; g.pad = (c.conv_k1 - 1) / 2;
        PB_TRY(launch_sgemm(g, st));
    }
    {   // y = h + conv2(f)
        GemmArgs g;
        g.A = w.f; g.lda = F; g.B = h->P(p + ".convlayer.conv2.packed"); g.ldb = (int64_t)c.conv_k2 * F;
        g.C = y; g.ldc = D; g.bias = h->P(p + ".convlayer.conv2.bias"); g.res = w.hbuf; g.ldres = D;
        g.M = M; g.N = D; g.K = c.conv_k2 * F; g.taps = c.conv_k2; g.Cin = F; g.L = L; g.pad = (c.conv_k2 - 1) / 2;
        PB_TRY(launch_sgemm(g, st));
    }
    return PB_OK;
}

static int ensure_keep(pb_tte* h, int B, int S) {
    Carver plan(nullptr);
    plan.take<float>((size_t)B * S * h->cfg.d_model);
    plan.take<int32_t>((size_t)B * S);
    plan.take<int64_t>(B);
    plan.take<int>(4);
    plan.take<int32_t>(B);
    PB_TRY(h->keep.reserve(plan.off));
    Carver cv(h->keep.ptr);
    h->enc_out = cv.take<float>((size_t)B * S * h->cfg.d_model);
    h->cum = cv.take<int32_t>((size_t)B * S);
    h->out_lens_dev = cv.take<int64_t>(B);
    h->err_flag = cv.take<int>(4);
    h->tgt_lens32 = cv.take<int32_t>(B);
    return PB_OK;
}

extern "C" int pb_tte_encode(pb_tte* h, const int64_t* phones, const uint8_t* src_mask, const int64_t* speaker,
                             const int64_t* dur_in, const int32_t* src_lens, int32_t B, int32_t S, float* log_dur, int64_t* dur,
                             int64_t* out_lens, int32_t* T_host, void* stream) {
    PB_REQUIRE(h && phones && src_mask && log_dur && dur && out_lens && T_host, PB_ERR_INVALID, "pb_tte_encode: null argument");
    PB_REQUIRE(h->finalized, PB_ERR_STATE, "pb_tte_encode: weights not finalized");
    PB_REQUIRE(B > 0 && S > 0, PB_ERR_INVALID, "pb_tte_encode: bad shape B=%d S=%d", B, S);
    const pb_tte_cfg& c = h->cfg;
    PB_REQUIRE(S < c.max_len, PB_ERR_INDEX, "index %d is out of bounds for dimension 0 with size %d (pos_emb.pe, source length)", S, c.max_len);
    PB_REQUIRE(c.n_speaker <= 1 || speaker != nullptr, PB_ERR_INVALID, "pb_tte_encode: speaker ids required (multi-speaker model)");
    for (const char* k : {"pos_emb.pe", "tok_emb.weight", "duration_predictor.proj.weight", "duration_predictor.proj.bias",
                          "duration_predictor.layers.0.conv.weight", "duration_predictor.layers.4.conv.weight"})
        PB_REQUIRE(h->slots[k].loaded, PB_ERR_STATE, "pb_tte_encode: weight %s was never loaded", k);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t launches0 = g_launch_count;
    const int D = c.d_model, F = c.conv_n_filter, H = c.n_head;
    const int M = B * S;
    PB_TRY(ensure_keep(h, B, S));
    PB_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), st));

    BlockWork w;
    float *xa, *xb;
    {
        Carver plan(nullptr);
        plan.take<float>((size_t)M * D); plan.take<float>((size_t)M * D);
        block_work_plan(plan, w, B, S, D, F, H, !c.fold_projections, c.gemm_mode == 1);
        PB_TRY(h->work.reserve(plan.off));
        Carver cv(h->work.ptr);
        xa = cv.take<float>((size_t)M * D); xb = cv.take<float>((size_t)M * D);
        block_work_plan(cv, w, B, S, D, F, H, !c.fold_projections, c.gemm_mode == 1);
    }
    // ragged mode (src_lens given): every item is computed as if it were run alone - its own pe row, zero padding past its own end
    // for the convolutions, exactly its own unit count of target positions.  src_mask must be the prefix mask t < src_lens[b].
    const bool ragged = src_lens != nullptr;
    PB_TRY(launch_embed_pos(phones, h->P("tok_emb.weight"), ragged ? h->P("pos_emb.pe") : h->P("pos_emb.pe") + (size_t)S * D, src_lens, S,
                            c.src_vocab_size, M, D, xa, h->err_flag, st));
    float* cur = xa;
    for (int i = 0; i < c.n_enc_layer; ++i) {
        float* dst = (i + 1 == c.n_enc_layer) ? h->enc_out : (cur == xa ? xb : xa);
        PB_TRY(run_fft_block(h, layer_prefix(0, i), cur, src_mask, B, S, w, dst, st, src_lens));
        cur = dst;
    }
    if (c.n_enc_layer == 0) {
        PB_CUDA(cudaMemcpyAsync(h->enc_out, xa, (size_t)M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (c.n_speaker > 1)
        PB_TRY(launch_add_indexed_row(h->enc_out, h->P("speaker_emb.weight"), speaker, c.n_speaker, B, S, D, h->err_flag, st, src_lens));
    else if (ragged)      // zero the encoder output past every item's end (the duration predictor's conv padding)
        PB_TRY(launch_add_indexed_row(h->enc_out, nullptr, nullptr, 0, B, S, D, h->err_flag, st, src_lens));
    // duration predictor (duration.py:41-48): conv3 -> relu -> LN -> conv3 (padding 1) -> relu -> LN -> proj
    {
        GemmArgs g;
        g.A = h->enc_out; g.lda = D; g.B = h->P("duration_predictor.layers.0.conv.packed"); g.ldb = 3 * D;
        g.C = w.o; g.ldc = 256; g.bias = h->P("duration_predictor.layers.0.conv.bias"); g.relu = 1;
        g.M = M; g.N = 256; g.K = 3 * D; g.taps = 3; g.Cin = D; g.L = S; g.pad = 1;
        PB_TRY(launch_sgemm(g, st));
        PB_TRY(launch_layernorm(w.o, h->P("duration_predictor.layers.2.weight"), h->P("duration_predictor.layers.2.bias"), w.xn, M, 256, st,
                                src_lens, S));
        GemmArgs g2;
        g2.A = w.xn; g2.lda = 256; g2.B = h->P("duration_predictor.layers.4.conv.packed"); g2.ldb = 3 * 256;
        g2.C = w.o; g2.ldc = 256; g2.bias = h->P("duration_predictor.layers.4.conv.bias"); g2.relu = 1;
        g2.M = M; g2.N = 256; g2.K = 3 * 256; g2.taps = 3; g2.Cin = 256; g2.L = S; g2.pad = 1;
        PB_TRY(launch_sgemm(g2, st));
        PB_TRY(launch_duration_head(w.o, h->P("duration_predictor.layers.6.weight"), h->P("duration_predictor.layers.6.bias"),
                                    h->P("duration_predictor.proj.weight"), h->P("duration_predictor.proj.bias"), src_mask, M, 256,
                                    log_dur, dur_in ? nullptr : dur, st));
    }
    if (dur_in && dur_in != dur) PB_CUDA(cudaMemcpyAsync(dur, dur_in, (size_t)M * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    PB_TRY(launch_duration_scan(dur, B, S, h->cum, h->out_lens_dev, st, h->tgt_lens32));
    PB_CUDA(cudaMemcpyAsync(out_lens, h->out_lens_dev, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    std::vector<int64_t> lens_host(B);
    int flag = 0;
    PB_CUDA(cudaMemcpyAsync(lens_host.data(), h->out_lens_dev, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(&flag, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    h->last_launches = g_launch_count - launches0;
    PB_REQUIRE(flag == 0, PB_ERR_INDEX, flag == 1 ? "index out of range in self (phones id outside the vocabulary)" : "index out of range in self (speaker id)");
    int64_t T = 0;
    for (int b = 0; b < B; ++b) { PB_REQUIRE(lens_host[b] >= 0, PB_ERR_INVALID, "pb_tte_encode: negative duration"); T = std::max(T, lens_host[b]); }
    h->enc_B = B; h->enc_S = S; h->enc_T = (int)std::min<int64_t>(T, 0x7fffffff);
    h->enc_ragged = ragged;
    *T_host = h->enc_T;
    PB_REQUIRE(T < c.max_len, PB_ERR_INDEX, "index %lld is out of bounds for dimension 0 with size %d (pos_emb.pe, unit length)", (long long)T, c.max_len);
    return PB_OK;
}

extern "C" int pb_tte_decode(pb_tte* h, int32_t B, int32_t S, int32_t T, const uint8_t* tgt_mask_in, float* logits, int64_t* ids,
                             uint8_t* tgt_mask, void* stream) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_tte_decode: null handle");
    PB_REQUIRE(h->finalized, PB_ERR_STATE, "pb_tte_decode: weights not finalized");
    PB_REQUIRE(h->enc_T >= 0 && B == h->enc_B && S == h->enc_S && T == h->enc_T, PB_ERR_STATE,
               "pb_tte_decode: (B,S,T)=(%d,%d,%d) does not match the last pb_tte_encode (%d,%d,%d)", B, S, T, h->enc_B, h->enc_S, h->enc_T);
    PB_REQUIRE(T > 0, PB_ERR_INVALID, "pb_tte_decode: all predicted durations are zero (empty output)");
    const pb_tte_cfg& c = h->cfg;
    PB_REQUIRE(T < c.max_len, PB_ERR_INDEX, "index %d is out of bounds for dimension 0 with size %d (pos_emb.pe)", T, c.max_len);
    PB_REQUIRE(h->slots["head.weight"].loaded && h->slots["head.bias"].loaded, PB_ERR_STATE, "pb_tte_decode: head weights were never loaded");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t launches0 = g_launch_count;
    const int D = c.d_model, F = c.conv_n_filter, H = c.n_head;
    const int64_t M64 = (int64_t)B * T;
    PB_REQUIRE(M64 < (1LL << 30), PB_ERR_INVALID, "pb_tte_decode: batch too large");
    const int M = (int)M64;
    BlockWork w;
    float *xa, *xb, *lg = nullptr, *lgp = nullptr; uint8_t* mask;
    unsigned long long* keys = nullptr;
    // ids only (Parrot.infer): the head GEMM takes the argmax in its epilogue and the [M, 1024] logits (268 MB at 32 x 2048 units)
    // never reach HBM; they are materialised only when the caller asks for them (Parrot.forward)
    const bool fused_argmax = c.gemm_mode == 1 && ids != nullptr && logits == nullptr;
    {
        Carver plan(nullptr);
        plan.take<float>((size_t)M * D); plan.take<float>((size_t)M * D); plan.take<uint8_t>(M);
        if (fused_argmax) plan.take<unsigned long long>(M);
        else if (c.gemm_mode == 1) plan.take<float>((size_t)M * round_up(c.n_codes, 256));
        else if (!logits) plan.take<float>((size_t)M * c.n_codes);
        block_work_plan(plan, w, B, T, D, F, H, !c.fold_projections, c.gemm_mode == 1);
        PB_TRY(h->work.reserve(plan.off));
        Carver cv(h->work.ptr);
        xa = cv.take<float>((size_t)M * D); xb = cv.take<float>((size_t)M * D); mask = cv.take<uint8_t>(M);
        if (fused_argmax) keys = cv.take<unsigned long long>(M);
        else if (c.gemm_mode == 1) { lgp = cv.take<float>((size_t)M * round_up(c.n_codes, 256)); lg = nullptr; }
        else lg = logits ? logits : cv.take<float>((size_t)M * c.n_codes);
        block_work_plan(cv, w, B, T, D, F, H, !c.fold_projections, c.gemm_mode == 1);
    }
    uint8_t* mk = tgt_mask ? tgt_mask : mask;
    const int32_t* tlens = h->enc_ragged ? h->tgt_lens32 : nullptr;
    PB_TRY(launch_length_regulate(h->enc_out, h->cum, h->enc_ragged ? h->P("pos_emb.pe") : h->P("pos_emb.pe") + (size_t)T * D, B, S, D, T, xa,
                                  mk, st, h->enc_ragged ? 1 : 0));
    // a caller-supplied target mask (teacher-forced path) is returned unchanged by the reference's length_regulator and is the
    // decoder's key padding mask (parrot.py:104,107; duration.py:22-24): it replaces the generated t <= len mask
    if (tgt_mask_in && tgt_mask_in != mk) PB_CUDA(cudaMemcpyAsync(mk, tgt_mask_in, (size_t)M, cudaMemcpyDeviceToDevice, st));
    float* cur = xa; float* nxt = xb;
    for (int i = 0; i < c.n_dec_layer; ++i) {
        PB_TRY(run_fft_block(h, layer_prefix(1, i), cur, mk, B, T, w, nxt, st, tlens));
        std::swap(cur, nxt);
    }
    if (c.gemm_mode == 1) {
        // head on tcgen05: N padded to a multiple of 256 (zero rows), logits land in a padded scratch
        const int np = (int)round_up(c.n_codes, 256);
        PB_TRY(launch_rows_to_split(cur, D, w.xn_s, M, D, st));
        if (fused_argmax) {
            PB_CUDA(cudaMemsetAsync(keys, 0, (size_t)M * sizeof(unsigned long long), st));
            TcConvArgs a;
            a.x = reinterpret_cast<const __nv_bfloat16*>(w.xn_s); a.w = reinterpret_cast<const uint8_t*>(h->P("head.weight.tc"));
            a.bias = h->P("head.bias.pad"); a.argmax_key = keys; a.argmax_n = c.n_codes;
            a.B = 1; a.L = M; a.Cin = D; a.N = np; a.out_C = np; a.U = 1; a.K = 1; a.D = 1; a.fmt = 1;
            PB_TRY(launch_tc_conv(a, st));
            PB_TRY(tc_argmax_unpack(keys, M, ids, st));
            h->last_launches = g_launch_count - launches0;
            return PB_OK;
        }
        PB_TRY(tte_tc_gemm(w.xn_s, reinterpret_cast<const uint8_t*>(h->P("head.weight.tc")), h->P("head.bias.pad"), nullptr, nullptr, lgp,
                           1, M, D, np, 1, 0, st));
        if (ids) PB_TRY(launch_argmax(lgp, M, c.n_codes, np, ids, st));
        if (logits) PB_CUDA(cudaMemcpy2DAsync(logits, (size_t)c.n_codes * sizeof(float), lgp, (size_t)np * sizeof(float),
                                              (size_t)c.n_codes * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
    } else {
        GemmArgs g;
        g.A = cur; g.lda = D; g.B = h->P("head.weight"); g.ldb = D; g.C = lg; g.ldc = c.n_codes; g.bias = h->P("head.bias");
        g.M = M; g.N = c.n_codes; g.K = D;
        PB_TRY(launch_sgemm(g, st));
        if (ids) PB_TRY(launch_argmax(lg, M, c.n_codes, c.n_codes, ids, st));
    }
    h->last_launches = g_launch_count - launches0;
    return PB_OK;
}

// ---- per-op entry points ---------------------------------------------------------------------------
extern "C" int pb_layernorm(const float* x, const float* gamma, const float* beta, float* y, int32_t M, int32_t D, void* stream) {
    PB_REQUIRE(x && gamma && beta && y, PB_ERR_INVALID, "pb_layernorm: null argument");
    return launch_layernorm(x, gamma, beta, y, M, D, (cudaStream_t)stream);
}

extern "C" int pb_linear(const float* x, const float* w, const float* bias, float* y, int32_t M, int32_t N, int32_t K, int32_t mode, void* stream) {
    PB_REQUIRE(x && w && y, PB_ERR_INVALID, "pb_linear: null argument");
    PB_REQUIRE(mode == 0 || mode == 1, PB_ERR_INVALID, "pb_linear: gemm mode %d is not built", mode);
    if (mode == 1) {
        // tcgen05 split-fp16 GEMM (unit-test entry: operands are converted / packed on the fly)
        PB_REQUIRE(M > 0 && (K == 16 || K == 32 || K % 64 == 0) && tc_n_tile(N) > 0, PB_ERR_INVALID,
                   "pb_linear(mode 1): K=%d must be 16, 32 or a multiple of 64 and N=%d a multiple of 16 (<= 256, or a multiple of 128)", K, N);
        cudaStream_t st = (cudaStream_t)stream;
        uint16_t* xs = nullptr; uint8_t* wp = nullptr;
        PB_CUDA(cudaMalloc(&xs, (size_t)M * 2 * K * sizeof(uint16_t)));
        if (cudaMalloc(&wp, tc_packed_weight_bytes(K, N, 1)) != cudaSuccess) { cudaFree(xs); set_error("pb_linear: out of memory"); return PB_ERR_CUDA; }
        int rc = launch_rows_to_split(x, K, xs, M, K, st);
        if (rc == PB_OK) rc = tc_pack_conv_weight(w, wp, N, K, 1, st, 1);
        if (rc == PB_OK) rc = tte_tc_gemm(xs, wp, bias, nullptr, nullptr, y, 1, M, K, N, 1, 0, st);
        cudaStreamSynchronize(st);
        cudaFree(xs); cudaFree(wp);
        return rc;
    }
    GemmArgs g;
    g.A = x; g.lda = K; g.B = w; g.ldb = K; g.C = y; g.ldc = N; g.bias = bias; g.M = M; g.N = N; g.K = K;
    return launch_sgemm(g, (cudaStream_t)stream);
}

extern "C" int pb_tte_fft_block(pb_tte* h, int32_t stack, int32_t layer, const float* x, const uint8_t* key_valid, int32_t B, int32_t L,
                                float* y, void* stream) {
    PB_REQUIRE(h && x && y, PB_ERR_INVALID, "pb_tte_fft_block: null argument");
    PB_REQUIRE(h->finalized, PB_ERR_STATE, "pb_tte_fft_block: weights not finalized");
    const pb_tte_cfg& c = h->cfg;
    PB_REQUIRE(stack >= 0 && stack < 2 && layer >= 0 && layer < (stack == 0 ? c.n_enc_layer : c.n_dec_layer), PB_ERR_INVALID, "pb_tte_fft_block: no such layer");
    PB_REQUIRE(B > 0 && L > 0, PB_ERR_INVALID, "pb_tte_fft_block: bad shape");
    BlockWork w;
    Carver plan(nullptr);
    block_work_plan(plan, w, B, L, c.d_model, c.conv_n_filter, c.n_head, !c.fold_projections, c.gemm_mode == 1);
    PB_TRY(h->work.reserve(plan.off));
    Carver cv(h->work.ptr);
    block_work_plan(cv, w, B, L, c.d_model, c.conv_n_filter, c.n_head, !c.fold_projections, c.gemm_mode == 1);
    const int64_t launches0 = g_launch_count;
    int rc = run_fft_block(h, layer_prefix(stack, layer), x, key_valid, B, L, w, y, (cudaStream_t)stream);
    h->last_launches = g_launch_count - launches0;
    return rc;
}

extern "C" int pb_length_regulate(const float* seq, const int64_t* dur, const float* pe_row, int32_t B, int32_t S, int32_t D, int32_t T,
                                  float* out, uint8_t* tgt_mask, int64_t* out_lens, void* stream) {
    PB_REQUIRE(seq && dur && out && out_lens, PB_ERR_INVALID, "pb_length_regulate: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* cum = nullptr;
    PB_CUDA(cudaMalloc(&cum, (size_t)B * S * sizeof(int32_t)));
    int rc = launch_duration_scan(dur, B, S, cum, out_lens, st);
    if (rc == PB_OK && T > 0) rc = launch_length_regulate(seq, cum, pe_row, B, S, D, T, out, tgt_mask, st);
    cudaStreamSynchronize(st);
    cudaFree(cum);
    return rc;
}

// DurationPredictor.forward (duration.py:41-48) + rounding (parrot.py:82-86) on an arbitrary [B,S,256] input.
extern "C" int pb_tte_duration(pb_tte* h, const float* x, const uint8_t* src_mask, int32_t B, int32_t S, float* log_dur, int64_t* dur,
                               void* stream) {
    PB_REQUIRE(h && x && log_dur, PB_ERR_INVALID, "pb_tte_duration: null argument");
    PB_REQUIRE(h->finalized, PB_ERR_STATE, "pb_tte_duration: weights not finalized");
    for (const char* k : {"duration_predictor.proj.weight", "duration_predictor.proj.bias", "duration_predictor.layers.0.conv.weight",
                          "duration_predictor.layers.4.conv.weight", "duration_predictor.layers.2.weight", "duration_predictor.layers.6.weight"})
        PB_REQUIRE(h->slots[k].loaded, PB_ERR_STATE, "pb_tte_duration: weight %s was never loaded", k);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->cfg.d_model, M = B * S;
    Carver plan(nullptr);
    plan.take<float>((size_t)M * 256); plan.take<float>((size_t)M * 256);
    PB_TRY(h->work.reserve(plan.off));
    Carver cv(h->work.ptr);
    float* a = cv.take<float>((size_t)M * 256);
    float* b = cv.take<float>((size_t)M * 256);
    GemmArgs g;
    g.A = x; g.lda = D; g.B = h->P("duration_predictor.layers.0.conv.packed"); g.ldb = 3 * D;
    g.C = a; g.ldc = 256; g.bias = h->P("duration_predictor.layers.0.conv.bias"); g.relu = 1;
    g.M = M; g.N = 256; g.K = 3 * D; g.taps = 3; g.Cin = D; g.L = S; g.pad = 1;
    PB_TRY(launch_sgemm(g, st));
    PB_TRY(launch_layernorm(a, h->P("duration_predictor.layers.2.weight"), h->P("duration_predictor.layers.2.bias"), b, M, 256, st));
    GemmArgs g2;
    g2.A = b; g2.lda = 256; g2.B = h->P("duration_predictor.layers.4.conv.packed"); g2.ldb = 3 * 256;
    g2.C = a; g2.ldc = 256; g2.bias = h->P("duration_predictor.layers.4.conv.bias"); g2.relu = 1;
    g2.M = M; g2.N = 256; g2.K = 3 * 256; g2.taps = 3; g2.Cin = 256; g2.L = S; g2.pad = 1;
    PB_TRY(launch_sgemm(g2, st));
    return launch_duration_head(a, h->P("duration_predictor.layers.6.weight"), h->P("duration_predictor.layers.6.bias"),
                                h->P("duration_predictor.proj.weight"), h->P("duration_predictor.proj.bias"), src_mask, M, 256, log_dur, dur, st);
}
