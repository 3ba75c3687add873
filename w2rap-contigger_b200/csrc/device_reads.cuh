// device_reads.cuh — the device-resident read store behind the opaque w2rap_device_reads handle.
#pragma once
#include "kernels.cuh"

namespace w2r {
struct DeviceReads {
    int device = 0;
    uint64_t n = 0, n_bases = 0, bases_bytes = 0, quals_bytes = 0;
    uint32_t max_len = 0;
    uint8_t* bases = nullptr;       // +32 bytes of zero padding after the last read
    uint8_t* quals = nullptr;
    uint64_t* base_off = nullptr;
    uint64_t* qual_off = nullptr;
    uint32_t* len = nullptr;
    bool pooled = false;            // true: buffers come from the stream-ordered pool (freed with cudaFreeAsync on pool_stream)
    cudaStream_t pool_stream = nullptr;
    ReadsView view() const { return ReadsView{n, bases, base_off, len, quals, qual_off}; }
    void release() {
        void* ps[5] = {bases, quals, base_off, qual_off, len};
        for (void* q : ps) { if (!q) continue; if (pooled) cudaFreeAsync(q, pool_stream); else cudaFree(q); }
        bases = quals = nullptr; base_off = qual_off = nullptr; len = nullptr;
    }
};
}  // namespace w2r

struct w2rap_device_reads { w2r::DeviceReads d; };
