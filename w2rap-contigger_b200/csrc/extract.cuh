// extract.cuh — per-read canonical 60-mer + context extraction (the leaf loop of createDictOMPRecursive,
// paths/long/BuildReadQGraph.cc:1062-1080) as a host/device function over a packed read.
#pragma once
#include "kmer.cuh"

namespace w2r {

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
    uint64_t buf = 0;                         // the next bases of the read, 32 at a time (one pair of 8-byte loads per 32 k-mers)
    for (uint32_t j = 0;; ++j) {
        uint32_t nxt = 0, c = 0;
        if ((j & 31u) == 0) buf = bases32_at(bases, (uint64_t)j + K);
        if (j < last) { nxt = (uint32_t)buf & 3u; c |= 1u << nxt; }
        buf >>= 2;
        if (j > 0) c |= 16u << prev_first;
        {   // select first, emit once: the two orientations must not become two divergent copies of the emit body
            const bool rev = kmer_less(r, f);
            const Kmer canon{rev ? r.w0 : f.w0, rev ? r.w1 : f.w1};
            emit(canon, rev ? ctx_rc(c) : c);
        }
        if (j == last) break;
        prev_first = kmer_first(f);
        f = kmer_succ(f, nxt);
        r = kmer_pred(r, 3u - nxt);
    }
}

}  // namespace w2r
