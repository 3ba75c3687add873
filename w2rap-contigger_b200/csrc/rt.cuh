// rt.cuh — small host runtime for the step-2 library: error plumbing, device buffers, stream + event timing,
// launch geometry for a 148-SM B200.  Product code only (CUDA required; there is no CPU path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace w2r {

struct Error {
    int code;
    std::string msg;
};

#define W2R_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            char _b[512];                                                                                      \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            throw ::w2r::Error{(_e == cudaErrorMemoryAllocation) ? 4 : 3, _b};                                  \
        }                                                                                                      \
    } while (0)

#define W2R_FAIL(code_, ...)                         \
    do {                                             \
        char _b[512];                                \
        snprintf(_b, sizeof _b, __VA_ARGS__);        \
        throw ::w2r::Error{(code_), _b};             \
    } while (0)

// One stream for the whole pipeline; every launch is counted (the bench reports gpu_launches).
struct Ctx {
    cudaStream_t stream = nullptr;
    int device = 0;
    int sm_count = 148;
    uint32_t launches = 0;
    uint32_t count_launches = 0;
    bool verbose = false;
    void* slab = nullptr;       // DeviceSlab leased to this call (pipeline.cu), or null: large buffers then come from CUDA's pool
};

template <class T>
struct DBuf {   // device buffer
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    explicit DBuf(size_t n_) { alloc(n_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DBuf() { release(); }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) W2R_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
    void zero(cudaStream_t s) { if (n) W2R_CUDA(cudaMemsetAsync(p, 0, bytes(), s)); }
    void fill_ff(cudaStream_t s) { if (n) W2R_CUDA(cudaMemsetAsync(p, 0xff, bytes(), s)); }
};

// Stream-ordered temporary (cudaMallocAsync pool): no driver round trip and no implicit device synchronisation on free.
template <class T>
struct TmpBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s;
    TmpBuf(const Ctx& c, size_t n_) : n(n_), s(c.stream) { if (n) W2R_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), s)); }
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
    ~TmpBuf() { if (p) cudaFreeAsync(p, s); }
};

template <class T>
struct HPinned {   // pinned host buffer (results handed to the caller are plain malloc; this is for staging)
    T* p = nullptr;
    size_t n = 0;
    HPinned() {}
    explicit HPinned(size_t n_) { alloc(n_); }
    HPinned(const HPinned&) = delete;
    HPinned& operator=(const HPinned&) = delete;
    ~HPinned() { if (p) cudaFreeHost(p); }
    void alloc(size_t n_) { if (p) cudaFreeHost(p); p = nullptr; n = n_; if (n) W2R_CUDA(cudaMallocHost((void**)&p, n * sizeof(T))); }
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t s_) : s(s_) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

// grid for an element-wise grid-stride kernel: enough CTAs to fill every SM a few times over, never more than needed.
inline unsigned grid_for(const Ctx& c, uint64_t n, unsigned block, unsigned ctas_per_sm = 16) {
    uint64_t need = (n + block - 1) / block;
    uint64_t cap = (uint64_t)c.sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

#define W2R_LAUNCH(ctx, kernel, grid, block, smem, ...)                 \
    do {                                                                \
        kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__); \
        (ctx).launches++;                                               \
        W2R_CUDA(cudaGetLastError());                                   \
    } while (0)

template <class T>
inline T d2h_scalar(const Ctx& c, const T* dptr) {
    T v;
    W2R_CUDA(cudaMemcpyAsync(&v, dptr, sizeof(T), cudaMemcpyDeviceToHost, c.stream));
    W2R_CUDA(cudaStreamSynchronize(c.stream));
    return v;
}

}  // namespace w2r
