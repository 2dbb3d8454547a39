// pb_runtime.cu - error reporting, TMA descriptor construction, device buffers shared by both stages.
#include <stdarg.h>
#include <string.h>

#include <string>

#include "pb_common.cuh"
#include "pb_runtime.cuh"

namespace pb {

static thread_local char g_err[1024] = "";
thread_local int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return PB_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap_2d_f32(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t pitch, int box_cols,
                     int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    PB_REQUIRE(fn != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    PB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch % 4 == 0, PB_ERR_INVALID,
               "TMA source must be 16-byte aligned with a pitch multiple of 4 floats (pitch=%lld)", (long long)pitch);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_REQUIRE(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) cols=%lld rows=%lld pitch=%lld box=%dx%d",
               (int)r, (long long)cols, (long long)rows, (long long)pitch, box_cols, box_rows);
    return PB_OK;
}

// ---- DevBuf -----------------------------------------------------------------------------------
int DevBuf::reserve(size_t nbytes) {
    if (nbytes <= bytes) return PB_OK;
    if (ptr) { cudaFree(ptr); ptr = nullptr; bytes = 0; }
    ++epoch;
    // grow with slack so that slowly increasing shapes do not reallocate every call
    size_t want = nbytes + nbytes / 8;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { want = nbytes; e = cudaMalloc(&ptr, want); }
    if (e != cudaSuccess) { ptr = nullptr; return cuda_fail(e, "cudaMalloc(workspace)", __FILE__, __LINE__); }
    bytes = want;
    return PB_OK;
}
void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

int upload_any(void* dst_dev, const void* src, size_t nbytes, cudaStream_t st) {
    // src may be host or device memory: cudaMemcpyDefault resolves it through UVA
    PB_CUDA(cudaMemcpyAsync(dst_dev, src, nbytes, cudaMemcpyDefault, st));
    return PB_OK;
}

}  // namespace pb

extern "C" {
int pb_version(void) { return PB_ABI_VERSION; }
const char* pb_last_error(void) { return pb::g_err; }
}
