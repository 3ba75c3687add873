// extract.cuh — per-read canonical 60-mer + context extraction (the leaf loop of createDictOMPRecursive,
// paths/long/BuildReadQGraph.cc:1062-1080) as a host/device function over a packed read.
#pragma once
#include "kmer.cuh"

namespace w2r {

// Sequential base reader over a 2-bit packed read (LSB-first in byte); one byte load per four bases.
struct BaseReader {
    const uint8_t* p;
    uint32_t cur, pos;
    W2R_HD explicit BaseReader(const uint8_t* bases, uint32_t start = 0) : p(bases + (start >> 2)), cur(0), pos(start) { cur = (uint32_t)(*p++) >> ((start & 3) * 2); }
    W2R_HD uint32_t next() {
        uint32_t b = cur & 3u;
        cur >>= 2;
        if ((++pos & 3u) == 0) cur = *p++;   // may read one byte past the last base: callers guarantee a padded store
        return b;
    }
};

// Builds the k-mer starting at base `pos` of a packed read.
W2R_HD Kmer kmer_at(const uint8_t* bases, uint32_t pos) {
    uint64_t w0 = 0, w1 = 0;
    for (int i = 0; i < 32; ++i) w0 = (w0 << 2) | packed_base(bases, pos + i);
    for (int i = 32; i < K; ++i) w1 = (w1 << 2) | packed_base(bases, pos + i);
    return Kmer{w0, w1 << 8};
}

// Calls emit(canonical k-mer, context byte) for every k-mer of the quality-floored read prefix [0, good_len).
//   - only reads with good_len > K contribute (BuildReadQGraph.cc:1064)
//   - first k-mer: successor bit only; last: predecessor bit only (:1066-1078)
//   - REV k-mers are stored reverse-complemented with the context bit-reversed (:1069,1074,1078); palindromes as seen.
template <class Emit>
W2R_HD void extract_read_kmers(const uint8_t* bases, uint32_t good_len, Emit& emit) {
    if (good_len <= (uint32_t)K) return;
    Kmer f = kmer_at(bases, 0);
    Kmer r = kmer_rc(f);
    const uint32_t last = good_len - K;       // index of the last k-mer
    uint32_t prev_first = 0;
    for (uint32_t j = 0;; ++j) {
        uint32_t nxt = 0, c = 0;
        if (j < last) { nxt = packed_base(bases, j + K); c |= 1u << nxt; }
        if (j > 0) c |= 16u << prev_first;
        if (kmer_less(r, f)) emit(r, ctx_rc(c)); else emit(f, c);
        if (j == last) break;
        prev_first = kmer_first(f);
        f = kmer_succ(f, nxt);
        r = kmer_pred(r, 3u - nxt);
    }
}

}  // namespace w2r
