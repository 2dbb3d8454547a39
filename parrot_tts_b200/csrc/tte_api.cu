// tte_api.cu - placeholder entry points (replaced by the real TTE path next).
#include "pb_runtime.cuh"
using namespace pb;
struct pb_tte { int dummy; };
#define NOTYET(name) do { set_error(name ": TTE path not built yet"); return PB_ERR_STATE; } while (0)
extern "C" {
int pb_tte_create(const pb_tte_cfg*, pb_tte**) { NOTYET("pb_tte_create"); }
void pb_tte_destroy(pb_tte*) {}
int pb_tte_set_weight(pb_tte*, const char*, const void*, const int64_t*, int32_t, int32_t) { NOTYET("pb_tte_set_weight"); }
int pb_tte_finalize(pb_tte*, void*) { NOTYET("pb_tte_finalize"); }
int pb_tte_encode(pb_tte*, const int64_t*, const uint8_t*, const int64_t*, const int64_t*, int32_t, int32_t, float*, int64_t*, int64_t*, int32_t*, void*) { NOTYET("pb_tte_encode"); }
int pb_tte_decode(pb_tte*, int32_t, int32_t, int32_t, float*, int64_t*, uint8_t*, void*) { NOTYET("pb_tte_decode"); }
int64_t pb_tte_last_launches(const pb_tte*) { return 0; }
int pb_tte_arena_bytes(pb_tte*, int64_t*) { NOTYET("pb_tte_arena_bytes"); }
void* pb_tte_arena_ptr(pb_tte*) { return nullptr; }
int pb_tte_mark_finalized(pb_tte*) { NOTYET("pb_tte_mark_finalized"); }
int pb_layernorm(const float*, const float*, const float*, float*, int32_t, int32_t, void*) { NOTYET("pb_layernorm"); }
int pb_tte_fft_block(pb_tte*, int32_t, int32_t, const float*, const uint8_t*, int32_t, int32_t, float*, void*) { NOTYET("pb_tte_fft_block"); }
int pb_length_regulate(const float*, const int64_t*, const float*, int32_t, int32_t, int32_t, int32_t, float*, uint8_t*, int64_t*, void*) { NOTYET("pb_length_regulate"); }
int pb_linear(const float*, const float*, const float*, float*, int32_t, int32_t, int32_t, int32_t, void*) { NOTYET("pb_linear"); }
}
