// device_reads.cuh — the device-resident read store behind the opaque w2rap_device_reads handle.
#pragma once
#include <vector>

#include "kernels.cuh"

namespace w2r {
void arena_release(int device);   // pipeline.cu
struct DeviceReads {
    int device = 0;
    uint64_t n = 0, n_bases = 0, bases_bytes = 0, quals_bytes = 0;
    uint32_t max_len = 0;
    uint8_t* bases = nullptr;       // +32 bytes of zero padding after the last read
    uint8_t* quals = nullptr;
    uint64_t* base_off = nullptr;
    uint64_t* qual_off = nullptr;
    uint32_t* len = nullptr;
    uint64_t n_inst_upper = 0;      // sum of max(0, len - 59): an upper bound of the k-mer instances, known without the qualities
    uint64_t n_kreads = 0;          // reads with at least one k-mer (len >= 60): each yields at least one super-k-mer record
    // host-buffer entry points upload in batches on a copy stream; batch b = reads [batch_first[b], batch_first[b+1]) is on the
    // device once batch_ready[b] has fired, so extraction of batch b overlaps the transfer of batch b+1
    std::vector<cudaEvent_t> batch_ready;
    std::vector<uint64_t> batch_first;
    bool arena = false;             // buffers belong to the per-device upload arena (pipeline.cu): not freed, only handed back
    bool pooled = false;            // true: buffers come from the stream-ordered pool (freed with cudaFreeAsync on pool_stream)
    cudaStream_t pool_stream = nullptr;
    ReadsView view() const { return ReadsView{n, bases, base_off, len, quals, qual_off}; }
    void clear_batches() { for (cudaEvent_t e : batch_ready) cudaEventDestroy(e); batch_ready.clear(); batch_first.clear(); }
    void release() {
        clear_batches();
        void* ps[5] = {bases, quals, base_off, qual_off, len};
        if (arena) { arena_release(device); arena = false; }
        else for (void* q : ps) { if (!q) continue; if (pooled) cudaFreeAsync(q, pool_stream); else cudaFree(q); }
        bases = quals = nullptr; base_off = qual_off = nullptr; len = nullptr;
    }
};
}  // namespace w2r

struct w2rap_device_reads { w2r::DeviceReads d; };
