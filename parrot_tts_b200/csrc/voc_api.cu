// voc_api.cu - C ABI of the HiFi-GAN code generator: handle, checkpoint ingestion (weight-norm fold + pack),
// forward orchestration.  Reference: utils/vocoder/models.py:69-169, utils/vocoder/inference.py:65-74,104-109.
#include <stdlib.h>
#include <string.h>

#include "pb_runtime.cuh"
#include "pb_tc.cuh"
#include "pb_voc.cuh"

using namespace pb;

namespace {

struct ConvSlot {          // one convolution's packed parameters inside the arena
    size_t w_off = 0, b_off = 0;
    int Cin = 0, Cout = 0, K = 0, D = 1, co_tile = 0;
    int pair64 = 0;           // this layer runs on CTA pairs with resident weight halves (tc_use_cluster): fixed at plan time
    bool transposed = false;
    std::string key;
};

}  // namespace

struct pb_voc {
    pb_voc_cfg cfg;
    int device = 0;
    bool finalized = false;
    int hop = 1;
    std::map<std::string, StagedTensor> staged;
    // packed parameters: one arena whose layout is a pure function of cfg (so it can be broadcast)
    DevBuf arena;
    size_t arena_bytes = 0;
    ConvSlot conv_pre, conv_post;
    std::vector<ConvSlot> ups;
    std::vector<ConvSlot> rb1, rb2;   // index ((i * num_kernels + j) * num_dilations + m)
    size_t dict_off = 0, spkr_off = 0;
    DevBuf work;
    int* err_flag = nullptr;      // device: 1 = unit id, 2 = speaker id outside its embedding table (set by the embedding kernels)
    int64_t last_launches = 0;
    // optional section timing (bench.py roofline): events on the caller's stream at section boundaries
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    bool ev_valid = false;
    // chain streams (small forwards): the num_kernels ResBlock chains of a stage are independent until their outputs are summed,
    // so each runs on its own stream and the chains' kernels fill the SMs side by side
    cudaStream_t chain_st[PB_VOC_MAX_RESK] = {};
    cudaEvent_t ev_fork = nullptr, ev_last[PB_VOC_MAX_RESK] = {};
    bool chains_ready = false;
    float* A(size_t off) const { return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(arena.ptr) + off); }
};

static size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

static void plan_slot(pb_voc* h, ConvSlot& s, const std::string& key, int Cin, int Cout, int K, int D, bool transposed,
                      int co_tile, int pair64 = 0) {
    s.key = key; s.Cin = Cin; s.Cout = Cout; s.K = K; s.D = D; s.transposed = transposed; s.co_tile = co_tile; s.pair64 = pair64;
    size_t wbytes = (size_t)Cin * Cout * K * sizeof(float);
    if (h->cfg.conv_mode != 0 && Cout > 1)   // tensor-core operand: a transposed conv becomes U phases x 3 taps (D carries U)
        wbytes = transposed ? tc_packed_weight_bytes(Cin, D * Cout, 3) : tc_packed_weight_bytes(Cin, Cout, K, pair64);
    s.w_off = h->arena_bytes; h->arena_bytes = align_up(h->arena_bytes + wbytes);
    s.b_off = h->arena_bytes; h->arena_bytes = align_up(h->arena_bytes + (size_t)Cout * sizeof(float));
}

extern "C" int pb_voc_create(const pb_voc_cfg* cfg, pb_voc** out) {
    PB_REQUIRE(cfg && out, PB_ERR_INVALID, "pb_voc_create: null argument");
    PB_REQUIRE(cfg->num_upsamples > 0 && cfg->num_upsamples <= PB_VOC_MAX_UPS && cfg->num_kernels > 0 &&
                   cfg->num_kernels <= PB_VOC_MAX_RESK && cfg->num_dilations > 0 && cfg->num_dilations <= PB_VOC_MAX_DIL,
               PB_ERR_INVALID, "pb_voc_create: layer counts out of range");
    PB_REQUIRE(cfg->upsample_initial_channel % (1 << cfg->num_upsamples) == 0, PB_ERR_INVALID,
               "pb_voc_create: upsample_initial_channel=%d not divisible by 2^%d", cfg->upsample_initial_channel, cfg->num_upsamples);
    PB_REQUIRE(cfg->conv_mode >= 0 && cfg->conv_mode <= 2, PB_ERR_INVALID, "pb_voc_create: conv_mode=%d (0, 1 or 2 expected)", cfg->conv_mode);
    if (cfg->conv_mode != 0) {
        // tensor-core path: every layer width must map onto the implicit-GEMM tiles
        bool ok = cfg->model_in_dim % 64 == 0 && tc_n_tile(cfg->upsample_initial_channel) > 0 && cfg->upsample_initial_channel % 64 == 0;
        for (int i = 0; ok && i < cfg->num_upsamples; ++i) {
            const int cout = cfg->upsample_initial_channel >> (i + 1);
            ok = (cout == 16 || cout == 32 || cout % 64 == 0) && tc_n_tile(cfg->upsample_rates[i] * cout) > 0;
        }
        PB_REQUIRE(ok, PB_ERR_INVALID, "pb_voc_create: channel widths unsupported by the tensor-core path (use conv_mode 0)");
    }
    pb_voc* h = new pb_voc();
    h->cfg = *cfg;
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaMalloc(&h->err_flag, 4 * sizeof(int)) != cudaSuccess ||
        cudaMemset(h->err_flag, 0, 4 * sizeof(int)) != cudaSuccess) {
        if (h->err_flag) cudaFree(h->err_flag);
        delete h;
        return cuda_fail(cudaGetLastError(), "pb_voc_create: device / error flag", __FILE__, __LINE__);
    }
    const int c0 = cfg->upsample_initial_channel;
    int co = conv_co_tile(c0);
    if (!co || cfg->model_in_dim % 8) { cudaFree(h->err_flag); delete h; set_error("pb_voc_create: unsupported channel counts"); return PB_ERR_INVALID; }
    plan_slot(h, h->conv_pre, "conv_pre", cfg->model_in_dim, c0, 7, 1, false, co);
    h->hop = 1;
    for (int i = 0; i < cfg->num_upsamples; ++i) {
        const int cin = c0 >> i, cout = c0 >> (i + 1);
        const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
        h->hop *= u;
        const int ct = convtr_co_tile(cout, u);
        if (!ct || (k - u) % 2 || cin % 8) { cudaFree(h->err_flag); delete h; set_error("pb_voc_create: unsupported upsample layer %d", i); return PB_ERR_INVALID; }
        ConvSlot s;
        plan_slot(h, s, "ups." + std::to_string(i), cin, cout, k, u, true, ct);   // D field = stride of the transposed conv
        h->ups.push_back(s);
        const int cr = conv_co_tile(cout);
        if (!cr || cout % 8) { cudaFree(h->err_flag); delete h; set_error("pb_voc_create: unsupported resblock width %d", cout); return PB_ERR_INVALID; }
        for (int j = 0; j < cfg->num_kernels; ++j)
            for (int m = 0; m < cfg->num_dilations; ++m) {
                const std::string p = "resblocks." + std::to_string(i * cfg->num_kernels + j);
                ConvSlot a, b;
                // 64-channel layers with >= 7 taps: conv1 always as a CTA pair; conv2 only at >= 11 taps (measured per launch,
                // see tc_use_cluster).  The shape rule inside tc_use_cluster ignores the flag for every other layer.
                const int kk = cfg->resblock_kernel_sizes[j];
                plan_slot(h, a, p + ".convs1." + std::to_string(m), cout, cout, kk, cfg->resblock_dilation_sizes[j][m], false, cr, 1);
                plan_slot(h, b, p + ".convs2." + std::to_string(m), cout, cout, kk, 1, false, cr, kk >= 11 ? 1 : 0);
                h->rb1.push_back(a);
                h->rb2.push_back(b);
            }
    }
    plan_slot(h, h->conv_post, "conv_post", c0 >> cfg->num_upsamples, 1, 7, 1, false, 1);
    if (cfg->num_embeddings > 0) {
        h->dict_off = h->arena_bytes;
        h->arena_bytes = align_up(h->arena_bytes + (size_t)cfg->num_embeddings * cfg->embedding_dim * sizeof(float));
        if (cfg->num_speakers > 0) {
            h->spkr_off = h->arena_bytes;
            h->arena_bytes = align_up(h->arena_bytes + (size_t)cfg->num_speakers * cfg->embedding_dim * sizeof(float));
        }
    }
    int st = h->arena.reserve(h->arena_bytes);
    if (st != PB_OK) { cudaFree(h->err_flag); delete h; return st; }
    *out = h;
    return PB_OK;
}

extern "C" void pb_voc_destroy(pb_voc* h) {
    if (!h) return;
    for (auto& kv : h->staged) if (kv.second.ptr) cudaFree(kv.second.ptr);
    for (auto e : h->ev) cudaEventDestroy(e);
    if (h->chains_ready) {
        cudaEventDestroy(h->ev_fork);
        for (int j = 0; j < PB_VOC_MAX_RESK; ++j) {
            if (h->ev_last[j]) cudaEventDestroy(h->ev_last[j]);
            if (h->chain_st[j]) cudaStreamDestroy(h->chain_st[j]);
        }
    }
    h->arena.release();
    h->work.release();
    if (h->err_flag) cudaFree(h->err_flag);
    delete h;
}

// nn.Embedding raises IndexError for an id outside its table (reference utils/vocoder/models.py:155-156); the embedding kernels
// clamp (the launch must stay in bounds) and flag.  Synchronises `stream`, reports and clears the flag.
extern "C" int pb_voc_check_ids(pb_voc* h, void* stream) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_voc_check_ids: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    int flag = 0;
    PB_CUDA(cudaMemcpyAsync(&flag, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (flag == 0) return PB_OK;
    PB_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), st));
    set_error(flag == 1 ? "index out of range in self (unit id outside the dictionary)" : "index out of range in self (speaker id)");
    return PB_ERR_INDEX;
}

extern "C" int pb_voc_set_weight(pb_voc* h, const char* key, const void* data, const int64_t* shape, int32_t ndim, int32_t dtype) {
    PB_REQUIRE(h && key && data && shape && ndim > 0 && ndim <= 4, PB_ERR_INVALID, "pb_voc_set_weight: bad argument");
    PB_REQUIRE(dtype == 0, PB_ERR_INVALID, "pb_voc_set_weight(%s): only float32 tensors are accepted", key);
    StagedTensor t;
    t.shape.assign(shape, shape + ndim);
    t.dtype = dtype;
    const size_t nbytes = (size_t)t.numel() * sizeof(float);
    PB_CUDA(cudaMalloc(&t.ptr, nbytes ? nbytes : 4));
    {
        const cudaError_t e = cudaMemcpy(t.ptr, data, nbytes, cudaMemcpyDefault);
        if (e != cudaSuccess) { cudaFree(t.ptr); return cuda_fail(e, "cudaMemcpy(staged weight)", __FILE__, __LINE__); }
    }
    auto it = h->staged.find(key);
    if (it != h->staged.end()) { cudaFree(it->second.ptr); h->staged.erase(it); }
    h->staged[key] = t;
    h->finalized = false;
    return PB_OK;
}

static int finalize_conv(pb_voc* h, const ConvSlot& s, float* scratch, cudaStream_t st) {
    const StagedTensor* w = nullptr; const StagedTensor* g = nullptr;
    auto iw = h->staged.find(s.key + ".weight");
    auto iv = h->staged.find(s.key + ".weight_v");
    auto ig = h->staged.find(s.key + ".weight_g");
    if (iw != h->staged.end()) w = &iw->second;
    else if (iv != h->staged.end() && ig != h->staged.end()) { w = &iv->second; g = &ig->second; }
    PB_REQUIRE(w != nullptr, PB_ERR_INVALID, "pb_voc_finalize: missing %s.weight (or weight_g/weight_v)", s.key.c_str());
    const int n0 = s.transposed ? s.Cin : s.Cout, n1 = s.transposed ? s.Cout : s.Cin;
    PB_REQUIRE(w->shape.size() == 3 && w->shape[0] == n0 && w->shape[1] == n1 && w->shape[2] == s.K, PB_ERR_INVALID,
               "pb_voc_finalize: %s has shape [%lld,%lld,%lld], expected [%d,%d,%d]", s.key.c_str(),
               (long long)(w->shape.size() > 0 ? w->shape[0] : -1), (long long)(w->shape.size() > 1 ? w->shape[1] : -1),
               (long long)(w->shape.size() > 2 ? w->shape[2] : -1), n0, n1, s.K);
    if (g) PB_REQUIRE(g->numel() == n0, PB_ERR_INVALID, "pb_voc_finalize: %s.weight_g has %lld elements, expected %d", s.key.c_str(), (long long)g->numel(), n0);
    PB_TRY(launch_weight_norm_fold((const float*)w->ptr, g ? (const float*)g->ptr : nullptr, scratch, n0, n1 * s.K, st));
    if (s.Cout == 1) {
        // conv_post: [1, Cin, K] is already [ci][j]
        PB_CUDA(cudaMemcpyAsync(h->A(s.w_off), scratch, (size_t)s.Cin * s.K * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (h->cfg.conv_mode != 0) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(h->A(s.w_off));
        if (s.transposed) PB_TRY(tc_pack_convtr_weight(scratch, dst, s.Cin, s.Cout, s.K, s.D, st));
        else PB_TRY(tc_pack_conv_weight(scratch, dst, s.Cout, s.Cin, s.K, st, 0, s.pair64));
    } else if (s.transposed) {
        PB_TRY(launch_pack_conv_weight(scratch, s.K, (int64_t)s.Cout * s.K, h->A(s.w_off), s.Cout, s.Cin, s.K, s.co_tile, st));
    } else {
        PB_TRY(launch_pack_conv_weight(scratch, (int64_t)s.Cin * s.K, s.K, h->A(s.w_off), s.Cout, s.Cin, s.K, s.co_tile, st));
    }
    auto ib = h->staged.find(s.key + ".bias");
    PB_REQUIRE(ib != h->staged.end() && ib->second.numel() == s.Cout, PB_ERR_INVALID, "pb_voc_finalize: missing or mis-shaped %s.bias", s.key.c_str());
    PB_CUDA(cudaMemcpyAsync(h->A(s.b_off), ib->second.ptr, (size_t)s.Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PB_OK;
}

extern "C" int pb_voc_finalize(pb_voc* h, void* stream) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_voc_finalize: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    size_t max_w = 0;
    auto upd = [&](const ConvSlot& s) { max_w = std::max(max_w, (size_t)s.Cin * s.Cout * s.K); };
    upd(h->conv_pre); upd(h->conv_post);
    for (auto& s : h->ups) upd(s);
    for (auto& s : h->rb1) upd(s);
    float* scratch = nullptr;
    PB_CUDA(cudaMalloc(&scratch, max_w * sizeof(float)));
    int rc = PB_OK;
    auto run = [&](const ConvSlot& s) { if (rc == PB_OK) rc = finalize_conv(h, s, scratch, st); };
    run(h->conv_pre);
    for (auto& s : h->ups) run(s);
    for (auto& s : h->rb1) run(s);
    for (auto& s : h->rb2) run(s);
    run(h->conv_post);
    if (rc == PB_OK && h->cfg.num_embeddings > 0) {
        auto id = h->staged.find("dict.weight");
        if (id == h->staged.end() || id->second.numel() != (int64_t)h->cfg.num_embeddings * h->cfg.embedding_dim) {
            set_error("pb_voc_finalize: missing or mis-shaped dict.weight"); rc = PB_ERR_INVALID;
        } else {
            cudaMemcpyAsync(h->A(h->dict_off), id->second.ptr, id->second.numel() * sizeof(float), cudaMemcpyDeviceToDevice, st);
        }
        if (rc == PB_OK && h->cfg.num_speakers > 0) {
            auto is = h->staged.find("spkr.weight");
            if (is == h->staged.end() || is->second.numel() != (int64_t)h->cfg.num_speakers * h->cfg.embedding_dim) {
                set_error("pb_voc_finalize: missing or mis-shaped spkr.weight"); rc = PB_ERR_INVALID;
            } else {
                cudaMemcpyAsync(h->A(h->spkr_off), is->second.ptr, is->second.numel() * sizeof(float), cudaMemcpyDeviceToDevice, st);
            }
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (rc != PB_OK) return rc;
    PB_CUDA(e);
    for (auto& kv : h->staged) if (kv.second.ptr) cudaFree(kv.second.ptr);
    h->staged.clear();
    h->finalized = true;
    return PB_OK;
}

extern "C" int pb_voc_arena_bytes(pb_voc* h, int64_t* bytes) {
    PB_REQUIRE(h && bytes, PB_ERR_INVALID, "pb_voc_arena_bytes: null argument");
    *bytes = (int64_t)h->arena_bytes;
    return PB_OK;
}
extern "C" void* pb_voc_arena_ptr(pb_voc* h) { return h ? h->arena.ptr : nullptr; }
// Non-root ranks call this after receiving the broadcast arena.
extern "C" int pb_voc_mark_finalized(pb_voc* h) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_voc_mark_finalized: null handle");
    h->finalized = true;
    return PB_OK;
}

extern "C" int64_t pb_voc_last_launches(const pb_voc* h) { return h ? h->last_launches : 0; }
extern "C" int64_t pb_voc_memory_epoch(const pb_voc* h) { return h ? h->arena.epoch + h->work.epoch : -1; }

extern "C" int pb_voc_section_count(const pb_voc* h) { return h ? 2 + 2 * h->cfg.num_upsamples : 0; }

extern "C" int pb_voc_set_timing(pb_voc* h, int32_t enable) {
    PB_REQUIRE(h, PB_ERR_INVALID, "pb_voc_set_timing: null handle");
    h->timing = enable != 0;
    h->ev_valid = false;
    if (h->timing && h->ev.empty()) {
        h->ev.resize(pb_voc_section_count(h) + 1);
        for (auto& e : h->ev) PB_CUDA(cudaEventCreate(&e));
    }
    return PB_OK;
}

extern "C" int pb_voc_section_ms(pb_voc* h, float* ms_host, int32_t n) {
    PB_REQUIRE(h && ms_host, PB_ERR_INVALID, "pb_voc_section_ms: null argument");
    PB_REQUIRE(h->timing && h->ev_valid, PB_ERR_STATE, "pb_voc_section_ms: no timed forward has run");
    PB_REQUIRE(n == pb_voc_section_count(h), PB_ERR_INVALID, "pb_voc_section_ms: expected room for %d sections", pb_voc_section_count(h));
    PB_CUDA(cudaEventSynchronize(h->ev.back()));
    for (int i = 0; i < n; ++i) PB_CUDA(cudaEventElapsedTime(&ms_host[i], h->ev[i], h->ev[i + 1]));
    return PB_OK;
}

// ---- tensor-core forward: every activation is split-bf16 channels-last [B][L][2C] (pb_tc.cuh) ------------------------------
// Storage convention: a tensor that feeds a convolution is stored with the leaky-relu that precedes that convolution in
// the reference already applied (slope 0.1, or 0.01 before conv_post); the residual read undoes it (x = a < 0 ? a/slope : a).
static int voc_forward_tc(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                          int32_t B, int32_t T, float* wav, int16_t* pcm, cudaStream_t st) {
    const pb_voc_cfg& c = h->cfg;
    const int64_t launches0 = g_launch_count;
    const int c0 = c.upsample_initial_channel;
    const int single = c.conv_mode == 2;
    typedef __nv_bfloat16 bf;
    size_t slot_elems = 0;
    {
        int64_t L = T;
        for (int i = 0; i < c.num_upsamples; ++i) {
            L *= c.upsample_rates[i];
            PB_REQUIRE(L < (1LL << 30), PB_ERR_INVALID, "pb_voc_forward: T=%d too long", T);
            slot_elems = std::max(slot_elems, (size_t)B * (size_t)L * 2 * (c0 >> (i + 1)));
        }
    }
    // Chain streams: chain j (ResBlock kernel size j) of every stage runs on its own stream with its own XT / XA / XB buffers (the
    // last chain stays on the caller's stream); the running sum keeps its order k3 -> k7 -> k11, so results are bit-identical.
    // Measured on B200: one utterance of 512 units 1.35 -> 1.13 ms (1.29 -> 1.01 ms replayed as a CUDA graph) - at that size the wide
    // stages fill under a third of the SMs and three chains run side by side; 8 x 1024 units 8.30 -> 8.24 ms (nothing to fill).
    // Default: on for small forwards (B x T <= 2048 units); PB_VOC_CHAINS=0 / 1 forces it off / on.
    static const int chains_env = getenv("PB_VOC_CHAINS") ? atoi(getenv("PB_VOC_CHAINS")) : -1;
    const bool chains_want = chains_env >= 0 ? chains_env != 0 : (int64_t)B * T <= 2048;
    const bool chains = chains_want && c.num_kernels > 1 && c.num_kernels <= PB_VOC_MAX_RESK;
    constexpr int MAX_SLOTS = 4 + 3 * PB_VOC_MAX_RESK;
    const int n_slots = chains ? 4 + 3 * c.num_kernels : 7;        // X, XS, Y + the previous stage's Y, and (XT, XA, XB) per chain
    if (chains && !h->chains_ready) {
        PB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        for (int j = 0; j + 1 < c.num_kernels; ++j) {
            PB_CUDA(cudaStreamCreateWithFlags(&h->chain_st[j], cudaStreamNonBlocking));
            PB_CUDA(cudaEventCreateWithFlags(&h->ev_last[j], cudaEventDisableTiming));
        }
        h->chains_ready = true;
    }
    Carver plan(nullptr);
    plan.take<bf>((size_t)B * T * 2 * c.model_in_dim);
    plan.take<bf>((size_t)B * T * 2 * c0);
    for (int s = 0; s < n_slots; ++s) plan.take<bf>(slot_elems);
    PB_TRY(h->work.reserve(plan.off));
    Carver cv(h->work.ptr);
    bf* x0 = cv.take<bf>((size_t)B * T * 2 * c.model_in_dim);
    bf* h0 = cv.take<bf>((size_t)B * T * 2 * c0);
    bf* slot[MAX_SLOTS];
    for (int s = 0; s < n_slots; ++s) slot[s] = cv.take<bf>(slot_elems);

    int ev_i = 0;
    h->ev_valid = false;
#define PB_MARK() do { if (h->timing) PB_CUDA(cudaEventRecord(h->ev[ev_i++], st)); } while (0)
    PB_MARK();
    if (code) {
        PB_TRY(launch_voc_embed_split(code, spkr, h->A(h->dict_off), c.num_speakers > 0 ? h->A(h->spkr_off) : nullptr, lens, B, T,
                                      c.embedding_dim, c.num_embeddings, c.num_speakers, x0, h->err_flag, st));
    } else {
        PB_TRY(launch_f32_to_split(x, T, x0, B, c.model_in_dim, T, 1.f, st));
    }
    auto W = [&](const ConvSlot& s) { return reinterpret_cast<const uint8_t*>(h->A(s.w_off)); };
    {   // conv_pre, stored with the leaky-relu that precedes ups[0]
        TcConvArgs a;
        a.x = x0; a.w = W(h->conv_pre); a.bias = h->A(h->conv_pre.b_off); a.out = h0; a.lens = lens; a.len_mul = 1;
        a.B = B; a.L = T; a.Cin = c.model_in_dim; a.N = c0; a.out_C = c0; a.U = 1; a.K = 7; a.D = 1; a.post_slope = 0.1f;
        a.single_pass = single;
        PB_TRY(launch_tc_conv(a, st));
    }
    PB_MARK();
    const bf* cur_in = h0;
    int in_slot = -1;
    int64_t L = T;
    int mul = 1;
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int C = c0 >> (i + 1);
        const int u = c.upsample_rates[i];
        const int64_t Lo = L * u;
        mul *= u;
        bf* buf[MAX_SLOTS];
        for (int s = 0, n = 0; s < n_slots && n < n_slots - 1; ++s) if (s != in_slot) buf[n++] = slot[s];
        bf *X = buf[0], *XS = buf[4], *Y = buf[5];
        int y_slot = 0;
        for (int s = 0; s < n_slots; ++s) if (slot[s] == Y) y_slot = s;
        {   // ups[i] as a 3-tap polyphase implicit GEMM; X is stored leaky-relu'd (the resblocks' first activation)
            const ConvSlot& s = h->ups[i];
            TcConvArgs a;
            a.x = cur_in; a.w = W(s); a.bias = h->A(s.b_off); a.out = X; a.lens = lens; a.len_mul = mul;
            a.B = B; a.L = (int)L; a.Cin = s.Cin; a.N = u * s.Cout; a.out_C = s.Cout; a.U = u; a.K = 3; a.D = 1;
            a.post_slope = 0.1f; a.single_pass = single;
            PB_TRY(launch_tc_conv(a, st));
        }
        PB_MARK();
        if (chains) {       // fork: the side chains start once ups[i] (X) is complete
            PB_CUDA(cudaEventRecord(h->ev_fork, st));
            for (int j = 0; j + 1 < c.num_kernels; ++j) PB_CUDA(cudaStreamWaitEvent(h->chain_st[j], h->ev_fork, 0));
        }
        for (int j = 0; j < c.num_kernels; ++j) {
            const bf* cur = X;
            bf *XT = buf[1], *XA = buf[2], *XB = buf[3];
            cudaStream_t sj = st;
            if (chains) {
                if (j > 0) { XT = buf[6 + 3 * (j - 1)]; XA = buf[7 + 3 * (j - 1)]; XB = buf[8 + 3 * (j - 1)]; }
                if (j + 1 < c.num_kernels) sj = h->chain_st[j];      // the last chain (it writes Y) stays on the caller's stream
            }
            for (int m = 0; m < c.num_dilations; ++m) {
                const int idx = (i * c.num_kernels + j) * c.num_dilations + m;
                const ConvSlot& s1 = h->rb1[idx];
                const ConvSlot& s2 = h->rb2[idx];
                static const bool no_pair = getenv("PB_TC_NO_PAIR") != nullptr;      // dev switch: two launches per ResBlock step
                if (!no_pair && !single && tc_pair_supported(C, s1.K, s1.D)) {
                    // narrow stages: conv1 -> conv2 fused, the intermediate stays in shared memory (tc_pair.cu)
                    TcPairArgs f;
                    f.x = cur; f.w1 = W(s1); f.w2 = W(s2); f.bias1 = h->A(s1.b_off); f.bias2 = h->A(s2.b_off);
                    f.lens = lens; f.len_mul = mul; f.B = B; f.L = (int)Lo; f.C = C; f.K = s1.K; f.D = s1.D;
                    if (m + 1 < c.num_dilations) {
                        bf* dst = (cur == XA) ? XB : XA;
                        f.out = dst; f.post_slope = 0.1f;
                        PB_TRY(launch_tc_pair(f, sj));
                        cur = dst;
                    } else {
                        if (j > 0) f.accin = XS;
                        if (j + 1 < c.num_kernels) {
                            f.out = XS;
                        } else {
                            f.out = Y;
                            f.scale = 1.0f / (float)c.num_kernels;
                            f.post_slope = (i + 1 < c.num_upsamples) ? 0.1f : 0.01f;
                        }
                        // the chain-closing kernel adds into the running sum the previous chain has written: order k3 -> k7 -> k11 kept
                        if (chains && j > 0) PB_CUDA(cudaStreamWaitEvent(sj, h->ev_last[j - 1], 0));
                        PB_TRY(launch_tc_pair(f, sj));
                        if (chains && j + 1 < c.num_kernels) PB_CUDA(cudaEventRecord(h->ev_last[j], sj));
                    }
                    continue;
                }
                TcConvArgs a;
                a.x = cur; a.w = W(s1); a.bias = h->A(s1.b_off); a.out = XT; a.lens = lens; a.len_mul = mul;
                a.B = B; a.L = (int)Lo; a.Cin = C; a.N = C; a.out_C = C; a.U = 1; a.K = s1.K; a.D = s1.D;
                a.post_slope = 0.1f; a.single_pass = single; a.pair64 = s1.pair64;
                PB_TRY(launch_tc_conv(a, sj));
                TcConvArgs b2;
                b2.x = XT; b2.w = W(s2); b2.bias = h->A(s2.b_off); b2.res = cur; b2.res_unslope = 10.f;
                b2.lens = lens; b2.len_mul = mul; b2.B = B; b2.L = (int)Lo; b2.Cin = C; b2.N = C; b2.out_C = C; b2.U = 1;
                b2.K = s2.K; b2.D = 1; b2.single_pass = single; b2.pair64 = s2.pair64;
                if (m + 1 < c.num_dilations) {
                    bf* dst = (cur == XA) ? XB : XA;
                    b2.out = dst; b2.post_slope = 0.1f;
                    PB_TRY(launch_tc_conv(b2, sj));
                    cur = dst;
                } else {
                    if (j > 0) b2.accin = XS;
                    if (j + 1 < c.num_kernels) {
                        b2.out = XS;                 // raw running sum over the resblocks
                    } else {
                        b2.out = Y;
                        b2.scale = 1.0f / (float)c.num_kernels;
                        b2.post_slope = (i + 1 < c.num_upsamples) ? 0.1f : 0.01f;
                    }
                    if (chains && j > 0) PB_CUDA(cudaStreamWaitEvent(sj, h->ev_last[j - 1], 0));      // (as above)
                    PB_TRY(launch_tc_conv(b2, sj));
                    if (chains && j + 1 < c.num_kernels) PB_CUDA(cudaEventRecord(h->ev_last[j], sj));
                }
            }
        }
        PB_MARK();
        cur_in = Y; in_slot = y_slot; L = Lo;
    }
    PB_TRY(launch_conv_post_split(cur_in, h->A(h->conv_post.w_off), h->A(h->conv_post.b_off), wav, pcm, lens, mul, B,
                                  h->conv_post.Cin, (int)L, 7, st));
    PB_MARK();
#undef PB_MARK
    h->ev_valid = h->timing;
    h->last_launches = g_launch_count - launches0;
    return PB_OK;
}

static int voc_forward_any(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                           int32_t B, int32_t T, float* wav, int16_t* pcm, void* stream);

extern "C" int pb_voc_forward(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                              int32_t B, int32_t T, float* wav, void* stream) {
    PB_REQUIRE(h && wav, PB_ERR_INVALID, "pb_voc_forward: null argument");
    return voc_forward_any(h, x, code, spkr, lens, B, T, wav, nullptr, stream);
}

// forward + generate() tail: int16 PCM written by conv_post's epilogue (utils/vocoder/inference.py:65-74); wav optional
extern "C" int pb_voc_forward_pcm16(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                                    int32_t B, int32_t T, float* wav, int16_t* pcm, void* stream) {
    PB_REQUIRE(h && pcm, PB_ERR_INVALID, "pb_voc_forward_pcm16: null argument");
    PB_REQUIRE(h->cfg.conv_mode != 0 || wav != nullptr, PB_ERR_INVALID,
               "pb_voc_forward_pcm16: conv_mode 0 (fp32 SIMT) casts from the f32 waveform: pass wav as well");
    return voc_forward_any(h, x, code, spkr, lens, B, T, wav, pcm, stream);
}

static int voc_forward_any(pb_voc* h, const float* x, const int64_t* code, const int64_t* spkr, const int32_t* lens,
                           int32_t B, int32_t T, float* wav, int16_t* pcm, void* stream) {
    PB_REQUIRE(h->finalized, PB_ERR_STATE, "pb_voc_forward: weights not finalized");
    PB_REQUIRE((x != nullptr) != (code != nullptr), PB_ERR_INVALID, "pb_voc_forward: give exactly one of x / code");
    PB_REQUIRE(B > 0 && T > 0 && B <= 65535, PB_ERR_INVALID, "pb_voc_forward: bad shape B=%d T=%d", B, T);
    const pb_voc_cfg& c = h->cfg;
    if (code) {
        PB_REQUIRE(c.num_embeddings > 0, PB_ERR_INVALID, "pb_voc_forward: handle was created without a unit dictionary");
        PB_REQUIRE(c.num_speakers == 0 || spkr != nullptr, PB_ERR_INVALID, "pb_voc_forward: spkr ids required (multispkr)");
        const int in_dim = c.embedding_dim * (c.num_speakers > 0 ? 2 : 1);
        PB_REQUIRE(in_dim == c.model_in_dim, PB_ERR_INVALID, "pb_voc_forward: embedding dims (%d) != model_in_dim (%d)", in_dim, c.model_in_dim);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (c.conv_mode != 0) return voc_forward_tc(h, x, code, spkr, lens, B, T, wav, pcm, st);
    const int64_t launches0 = g_launch_count;
    const int c0 = c.upsample_initial_channel;
    const int P0 = (int)round_up(T, 4);

    // ---- workspace plan: x0, h0, and 7 rotating slots sized for the largest stage
    size_t slot_floats = 0;
    {
        int64_t L = T;
        for (int i = 0; i < c.num_upsamples; ++i) {
            L *= c.upsample_rates[i];
            PB_REQUIRE(L < (1LL << 30), PB_ERR_INVALID, "pb_voc_forward: T=%d too long", T);
            slot_floats = std::max(slot_floats, (size_t)B * (c0 >> (i + 1)) * (size_t)round_up(L, 4));
        }
    }
    Carver plan(nullptr);
    plan.take<float>((size_t)B * c.model_in_dim * P0);
    plan.take<float>((size_t)B * c0 * P0);
    for (int s = 0; s < 7; ++s) plan.take<float>(slot_floats);
    PB_TRY(h->work.reserve(plan.off));
    Carver cv(h->work.ptr);
    float* x0 = cv.take<float>((size_t)B * c.model_in_dim * P0);
    float* h0 = cv.take<float>((size_t)B * c0 * P0);
    float* slot[7];
    for (int s = 0; s < 7; ++s) slot[s] = cv.take<float>(slot_floats);

    int ev_i = 0;
    h->ev_valid = false;
#define PB_MARK() do { if (h->timing) PB_CUDA(cudaEventRecord(h->ev[ev_i++], st)); } while (0)
    PB_MARK();
    // ---- input features [B, model_in_dim, P0]
    if (code) {
        PB_TRY(launch_voc_embed(code, spkr, h->A(h->dict_off), c.num_speakers > 0 ? h->A(h->spkr_off) : nullptr, lens, B, T,
                                c.embedding_dim, c.num_embeddings, c.num_speakers, P0, x0, h->err_flag, st));
    } else {
        PB_TRY(launch_repitch(x, T, x0, P0, (int64_t)B * c.model_in_dim, T, lens, 1, c.model_in_dim, st));
    }

    // ---- conv_pre (+ the leaky_relu that precedes ups[0])
    {
        ConvArgs a;
        a.x = x0; a.w = h->A(h->conv_pre.w_off); a.co_tile = h->conv_pre.co_tile; a.bias = h->A(h->conv_pre.b_off);
        a.out = h0; a.lens = lens; a.len_mul = 1; a.B = B; a.Cin = c.model_in_dim; a.Cout = c0; a.L = T; a.pitch = P0;
        a.K = 7; a.D = 1; a.post_slope = 0.1f;
        PB_TRY(launch_conv1d(a, st));
    }

    PB_MARK();
    const float* cur_in = h0;
    int in_slot = -1;
    int64_t L = T;
    int pitch_in = P0;
    int mul = 1;
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int C = c0 >> (i + 1);
        const int u = c.upsample_rates[i];
        const int64_t Lo = L * u;
        const int P = (int)round_up(Lo, 4);
        mul *= u;
        float* buf[6];
        for (int s = 0, n = 0; s < 7 && n < 6; ++s) if (s != in_slot) buf[n++] = slot[s];
        float *X = buf[0], *XT = buf[1], *XA = buf[2], *XB = buf[3], *XS = buf[4], *Y = buf[5];
        int y_slot = 0;
        for (int s = 0; s < 7; ++s) if (slot[s] == Y) y_slot = s;
        {
            const ConvSlot& s = h->ups[i];
            ConvTrArgs a;
            a.x = cur_in; a.w = h->A(s.w_off); a.co_tile = s.co_tile; a.bias = h->A(s.b_off); a.out = X;
            a.lens = lens; a.len_mul = mul; a.B = B; a.Cin = s.Cin; a.Cout = s.Cout; a.Lin = (int)L;
            a.pitch_in = pitch_in; a.pitch_out = P; a.K = s.K; a.U = u;
            PB_TRY(launch_convtr1d(a, st));
        }
        PB_MARK();
        for (int j = 0; j < c.num_kernels; ++j) {
            const float* cur = X;
            for (int m = 0; m < c.num_dilations; ++m) {
                const int idx = (i * c.num_kernels + j) * c.num_dilations + m;
                const ConvSlot& s1 = h->rb1[idx];
                const ConvSlot& s2 = h->rb2[idx];
                ConvArgs a;
                a.x = cur; a.w = h->A(s1.w_off); a.co_tile = s1.co_tile; a.bias = h->A(s1.b_off); a.out = XT;
                a.lens = lens; a.len_mul = mul; a.B = B; a.Cin = C; a.Cout = C; a.L = (int)Lo; a.pitch = P;
                a.K = s1.K; a.D = s1.D; a.pre_slope = 0.1f; a.post_slope = 0.1f;
                PB_TRY(launch_conv1d(a, st));
                ConvArgs b2;
                b2.x = XT; b2.w = h->A(s2.w_off); b2.co_tile = s2.co_tile; b2.bias = h->A(s2.b_off); b2.res = cur;
                b2.lens = lens; b2.len_mul = mul; b2.B = B; b2.Cin = C; b2.Cout = C; b2.L = (int)Lo; b2.pitch = P;
                b2.K = s2.K; b2.D = 1;
                if (m + 1 < c.num_dilations) {
                    float* dst = (cur == XA) ? XB : XA;
                    b2.out = dst;
                    PB_TRY(launch_conv1d(b2, st));
                    cur = dst;
                } else {
                    // xs (+)= resblock output; the last kernel size also takes the mean and the next activation
                    if (j > 0) b2.accin = XS;
                    if (j + 1 < c.num_kernels) {
                        b2.out = XS;
                    } else {
                        b2.out = Y;
                        b2.acc_div = (float)c.num_kernels;
                        b2.post_slope = (i + 1 < c.num_upsamples) ? 0.1f : 0.01f;
                    }
                    PB_TRY(launch_conv1d(b2, st));
                }
            }
        }
        PB_MARK();
        cur_in = Y; in_slot = y_slot; L = Lo; pitch_in = P;
    }
    PB_TRY(launch_conv_post(cur_in, h->A(h->conv_post.w_off), h->A(h->conv_post.b_off), wav, lens, mul, B,
                            h->conv_post.Cin, (int)L, pitch_in, L, 7, st));
    if (pcm) PB_TRY(launch_wav_to_pcm16(wav, (int64_t)B * L, pcm, st));      // fp32 SIMT mode: the cast is its own launch
    PB_MARK();
#undef PB_MARK
    h->ev_valid = h->timing;
    h->last_launches = g_launch_count - launches0;
    return PB_OK;
}

extern "C" int pb_wav_to_pcm16(const float* wav, int64_t n, int16_t* pcm, void* stream) {
    PB_REQUIRE(wav && pcm && n >= 0, PB_ERR_INVALID, "pb_wav_to_pcm16: bad argument");
    return launch_wav_to_pcm16(wav, n, pcm, (cudaStream_t)stream);
}

extern "C" int pb_pcm16_peak_normalize(const int16_t* pcm, int32_t B, int64_t n, int32_t* peak, float* out, void* stream) {
    PB_REQUIRE(pcm && peak && out && B >= 0 && n >= 0, PB_ERR_INVALID, "pb_pcm16_peak_normalize: bad argument");
    return launch_pcm16_peak_normalize(pcm, B, n, peak, out, (cudaStream_t)stream);
}

// ---- per-op entry points (unit tests): torch-layout weights are packed on the fly ------------------
extern "C" int pb_conv1d(const float* x, const float* w, const float* bias, const float* res, float* y, int32_t B, int32_t Cin,
                         int32_t Cout, int32_t L, int32_t K, int32_t dilation, float pre_slope, float post_slope,
                         void* stream) {
    PB_REQUIRE(x && w && y, PB_ERR_INVALID, "pb_conv1d: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int ct = conv_co_tile(Cout);
    PB_REQUIRE(ct > 0, PB_ERR_INVALID, "pb_conv1d: Cout=%d must be a multiple of 16", Cout);
    const int P = (int)round_up(L, 4);
    float *xp = nullptr, *yp = nullptr, *rp = nullptr, *wp = nullptr;
    PB_CUDA(cudaMalloc(&xp, (size_t)B * Cin * P * sizeof(float)));
    PB_CUDA(cudaMalloc(&yp, (size_t)B * Cout * P * sizeof(float)));
    PB_CUDA(cudaMalloc(&wp, (size_t)Cout * Cin * K * sizeof(float)));
    if (res) PB_CUDA(cudaMalloc(&rp, (size_t)B * Cout * P * sizeof(float)));
    int rc = launch_repitch(x, L, xp, P, (int64_t)B * Cin, L, nullptr, 1, Cin, st);
    if (rc == PB_OK && res) rc = launch_repitch(res, L, rp, P, (int64_t)B * Cout, L, nullptr, 1, Cout, st);
    if (rc == PB_OK) rc = launch_pack_conv_weight(w, (int64_t)Cin * K, K, wp, Cout, Cin, K, ct, st);
    if (rc == PB_OK) {
        ConvArgs a;
        a.x = xp; a.w = wp; a.co_tile = ct; a.bias = bias; a.res = rp; a.out = yp; a.B = B; a.Cin = Cin; a.Cout = Cout;
        a.L = L; a.pitch = P; a.K = K; a.D = dilation; a.pre_slope = pre_slope; a.post_slope = post_slope;
        rc = launch_conv1d(a, st);
    }
    if (rc == PB_OK) rc = launch_repitch(yp, P, y, L, (int64_t)B * Cout, L, nullptr, 1, Cout, st);
    cudaStreamSynchronize(st);
    cudaFree(xp); cudaFree(yp); cudaFree(wp); if (rp) cudaFree(rp);
    return rc;
}

extern "C" int pb_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin,
                                   int32_t Cout, int32_t L, int32_t K, int32_t stride, float pre_slope, void* stream) {
    PB_REQUIRE(x && w && y, PB_ERR_INVALID, "pb_conv_transpose1d: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int ct = convtr_co_tile(Cout, stride);
    PB_REQUIRE(ct > 0, PB_ERR_INVALID, "pb_conv_transpose1d: Cout=%d must be a multiple of 16", Cout);
    const int P = (int)round_up(L, 4);
    const int64_t Lo = (int64_t)L * stride;
    const int Po = (int)round_up(Lo, 4);
    float *xp = nullptr, *yp = nullptr, *wp = nullptr;
    PB_CUDA(cudaMalloc(&xp, (size_t)B * Cin * P * sizeof(float)));
    PB_CUDA(cudaMalloc(&yp, (size_t)B * Cout * Po * sizeof(float)));
    PB_CUDA(cudaMalloc(&wp, (size_t)Cout * Cin * K * sizeof(float)));
    int rc = launch_repitch(x, L, xp, P, (int64_t)B * Cin, L, nullptr, 1, Cin, st);
    if (rc == PB_OK) rc = launch_pack_conv_weight(w, K, (int64_t)Cout * K, wp, Cout, Cin, K, ct, st);
    if (rc == PB_OK) {
        ConvTrArgs a;
        a.x = xp; a.w = wp; a.co_tile = ct; a.bias = bias; a.out = yp; a.B = B; a.Cin = Cin; a.Cout = Cout; a.Lin = L;
        a.pitch_in = P; a.pitch_out = Po; a.K = K; a.U = stride; a.pre_slope = pre_slope;
        rc = launch_convtr1d(a, st);
    }
    if (rc == PB_OK) rc = launch_repitch(yp, Po, y, Lo, (int64_t)B * Cout, (int)Lo, nullptr, 1, Cout, st);
    cudaStreamSynchronize(st);
    cudaFree(xp); cudaFree(yp); cudaFree(wp);
    return rc;
}

// ---- tensor-core per-op entry points (unit tests) ---------------------------------------------------------------------

namespace {
struct ScratchGuard {
    std::vector<void*> ptrs;
    ~ScratchGuard() { for (void* q : ptrs) cudaFree(q); }
    template <class T> int get(T** out, size_t n) {
        void* q = nullptr;
        PB_CUDA(cudaMalloc(&q, n * sizeof(T) + 16));
        ptrs.push_back(q);
        *out = reinterpret_cast<T*>(q);
        return PB_OK;
    }
};
}  // namespace

extern "C" int pb_conv1d_tc(const float* x, const float* w, const float* bias, const float* res, float* y, int32_t B, int32_t Cin,
                            int32_t Cout, int32_t L, int32_t K, int32_t dilation, float pre_slope, float post_slope,
                            int32_t debug_flags, void* stream) {
    PB_REQUIRE(x && w && y, PB_ERR_INVALID, "pb_conv1d_tc: null argument");
    PB_REQUIRE(B > 0 && L > 0 && Cin > 0 && Cout > 0 && K > 0, PB_ERR_INVALID, "pb_conv1d_tc: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    ScratchGuard g;
    __nv_bfloat16 *xs = nullptr, *ys = nullptr, *rs = nullptr;
    uint8_t* wp = nullptr;
    PB_TRY(g.get(&xs, (size_t)B * L * 2 * Cin));
    PB_TRY(g.get(&ys, (size_t)B * L * 2 * Cout));
    PB_TRY(g.get(&wp, tc_packed_weight_bytes(Cin, Cout, K, 1)));
    if (res) PB_TRY(g.get(&rs, (size_t)B * L * 2 * Cout));
    PB_TRY(launch_f32_to_split(x, L, xs, B, Cin, L, pre_slope, st));
    if (res) PB_TRY(launch_f32_to_split(res, L, rs, B, Cout, L, 1.f, st));
    PB_TRY(tc_pack_conv_weight(w, wp, Cout, Cin, K, st, 0, 1));
    TcConvArgs a;
    a.x = xs; a.w = wp; a.bias = bias; a.res = rs; a.out = ys; a.B = B; a.L = L; a.Cin = Cin; a.N = Cout; a.out_C = Cout;
    a.U = 1; a.K = K; a.D = dilation; a.post_slope = post_slope; a.single_pass = debug_flags & 1; a.pair64 = 1;
    PB_TRY(launch_tc_conv(a, st));
    PB_TRY(launch_split_to_f32(ys, y, L, B, Cout, L, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_conv_transpose1d_tc(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin,
                                      int32_t Cout, int32_t L, int32_t K, int32_t stride, float pre_slope, int32_t debug_flags,
                                      void* stream) {
    PB_REQUIRE(x && w && y, PB_ERR_INVALID, "pb_conv_transpose1d_tc: null argument");
    PB_REQUIRE(B > 0 && L > 0 && Cin > 0 && Cout > 0 && K > 0 && stride > 0, PB_ERR_INVALID, "pb_conv_transpose1d_tc: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    ScratchGuard g;
    __nv_bfloat16 *xs = nullptr, *ys = nullptr;
    uint8_t* wp = nullptr;
    const int Lo = L * stride;
    PB_TRY(g.get(&xs, (size_t)B * L * 2 * Cin));
    PB_TRY(g.get(&ys, (size_t)B * Lo * 2 * Cout));
    PB_TRY(g.get(&wp, tc_packed_weight_bytes(Cin, stride * Cout, 3)));
    PB_TRY(launch_f32_to_split(x, L, xs, B, Cin, L, pre_slope, st));
    PB_TRY(tc_pack_convtr_weight(w, wp, Cin, Cout, K, stride, st));
    TcConvArgs a;
    a.x = xs; a.w = wp; a.bias = bias; a.out = ys; a.B = B; a.L = L; a.Cin = Cin; a.N = stride * Cout; a.out_C = Cout;
    a.U = stride; a.K = 3; a.D = 1; a.single_pass = debug_flags & 1;
    PB_TRY(launch_tc_conv(a, st));
    PB_TRY(launch_split_to_f32(ys, y, Lo, B, Cout, Lo, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}
