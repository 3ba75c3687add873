// slab_freelist.h — the bookkeeping of the device slab (pipeline.cu: DeviceSlab): which pieces of one contiguous range are in use.
// Plain C++ (no CUDA), so that tests/hostcheck can fuzz it on the CPU.
//
// Free ranges are kept disjoint and coalesced in an ordered map (offset -> bytes); requests are served lowest address first, so
// long-lived buffers sink to the bottom and the top of the range is what grows and shrinks with the stages of a call.
#pragma once
#include <cstddef>
#include <iterator>
#include <map>
#include <unordered_map>

namespace w2r {

struct SlabFreeList {
    static constexpr size_t NONE = ~(size_t)0;
    std::map<size_t, size_t> free_;              // offset -> bytes
    std::unordered_map<size_t, size_t> live_;    // offset -> bytes
    size_t used = 0;

    // [off, off + bytes) becomes available (new backing, or a released piece); merges with its neighbours
    void add_free(size_t off, size_t bytes) {
        if (!bytes) return;
        auto nx = free_.lower_bound(off);
        if (nx != free_.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) { off = pv->first; bytes += pv->second; free_.erase(pv); }
        }
        if (nx != free_.end() && off + bytes == nx->first) { bytes += nx->second; free_.erase(nx); }
        free_[off] = bytes;
    }
    // lowest free range that holds `bytes`; NONE if there is none
    size_t take(size_t bytes) {
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second < bytes) continue;
            const size_t off = it->first, len = it->second;
            free_.erase(it);
            if (len > bytes) free_[off + bytes] = len - bytes;
            live_[off] = bytes; used += bytes;
            return off;
        }
        return NONE;
    }
    // returns the size of the piece, 0 if `off` is not a live piece
    size_t give_back(size_t off) {
        auto it = live_.find(off);
        if (it == live_.end()) return 0;
        const size_t bytes = it->second;
        live_.erase(it); used -= bytes;
        add_free(off, bytes);
        return bytes;
    }
    // bytes of the free range that ends exactly at `end` (what a request can reuse when the range is extended there)
    size_t free_tail(size_t end) const {
        if (free_.empty()) return 0;
        auto last = std::prev(free_.end());
        return last->first + last->second == end ? last->second : 0;
    }
};

}  // namespace w2r
