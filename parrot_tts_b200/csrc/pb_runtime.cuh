// pb_runtime.cuh - host-side runtime pieces shared by the vocoder and TTE handles.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "pb_common.cuh"

namespace pb {

// grow-only device buffer owned by a handle
struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    int64_t epoch = 0;          // bumped whenever the allocation moves: raw pointers captured earlier (CUDA graphs) are stale then
    int reserve(size_t nbytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// bump allocator over a DevBuf (256-byte aligned slices)
struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
    template <class T> T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += (n * sizeof(T) + 255) / 256 * 256;
        return p;
    }
};

// raw checkpoint tensor staged on the device until finalize
struct StagedTensor {
    void* ptr = nullptr;
    std::vector<int64_t> shape;
    int dtype = 0;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

int upload_any(void* dst_dev, const void* src, size_t nbytes, cudaStream_t st);

}  // namespace pb
