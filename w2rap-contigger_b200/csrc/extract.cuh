// extract.cuh — per-read canonical 60-mer + context extraction (the leaf loop of createDictOMPRecursive,
// paths/long/BuildReadQGraph.cc:1062-1080) as a host/device function over a packed read.
#pragma once
#include "kmer.cuh"

namespace w2r {

// Resumable form of the extraction loop: open() a read, then next() yields one (canonical k-mer, context) at a time.
//   - only reads with good_len > K contribute (BuildReadQGraph.cc:1064)
//   - first k-mer: successor bit only; last: predecessor bit only (:1066-1078)
//   - REV k-mers are stored reverse-complemented with the context bit-reversed (:1069,1074,1078); palindromes as seen.
struct KmerCursor {
    const uint8_t* bases;
    Kmer f, r;
    uint64_t buf;             // the next bases of the read, 32 at a time (one pair of 8-byte loads per 32 k-mers)
    uint32_t j, last, prev_first;
    bool live;
    W2R_HD KmerCursor() : bases(nullptr), f{0, 0}, r{0, 0}, buf(0), j(0), last(0), prev_first(0), live(false) {}
    W2R_HD void open(const uint8_t* b, uint32_t good_len) {
        live = good_len > (uint32_t)K;
        if (!live) return;
        bases = b; f = kmer_at(b, 0); r = kmer_rc(f); last = good_len - K; j = 0; prev_first = 0; buf = 0;
    }
    W2R_HD bool next(Kmer* canon, uint32_t* ctx) {
        if (!live) return false;
        uint32_t nxt = 0, c = 0;
        if ((j & 31u) == 0) buf = bases32_at(bases, (uint64_t)j + K);
        if (j < last) { nxt = (uint32_t)buf & 3u; c |= 1u << nxt; }
        buf >>= 2;
        if (j > 0) c |= 16u << prev_first;
        const bool rev = kmer_less(r, f);       // select first, emit once: no divergent copies of the consumer
        *canon = Kmer{rev ? r.w0 : f.w0, rev ? r.w1 : f.w1};
        *ctx = rev ? ctx_rc(c) : c;
        if (j == last) live = false;
        else { prev_first = kmer_first(f); f = kmer_succ(f, nxt); r = kmer_pred(r, 3u - nxt); ++j; }
        return true;
    }
};

// Calls emit(canonical k-mer, context byte) for every k-mer of the quality-floored read prefix [0, good_len).
template <class Emit>
W2R_HD void extract_read_kmers(const uint8_t* bases, uint32_t good_len, Emit& emit) {
    KmerCursor cur;
    cur.open(bases, good_len);
    Kmer k; uint32_t ctx;
    while (cur.next(&k, &ctx)) emit(k, ctx);
}

// ---------------------------------------------------------------- minimisers (partition key of the single-GPU count)
// The partition of a k-mer is derived from its MINIMISER: the canonical 15-mer with the smallest hash among the 46 inside the
// k-mer.  It is strand-symmetric (a k-mer and its reverse complement contain the same canonical m-mers), so every instance of a
// canonical k-mer lands in the same partition, and consecutive k-mers of a read share it for ~24 positions on average: a read
// appends RUNS of records to a partition (one cursor atomic per run, contiguous stores) instead of scattering single records.
// The counts do not depend on the partition function; this is purely a data-movement choice (the reference's own bucketing,
// MapReduceEngine.h:288-358, hashes whole k-mers).
constexpr int MINI_M = 15;
constexpr int MINI_W = K - MINI_M + 1;     // m-mers per k-mer
// hash of the canonical m-mer starting at base `pos` (a bijection of its 30-bit value, so equal hash <=> equal canonical m-mer)
W2R_HD uint32_t mmer_hash_at(const uint8_t* bases, uint64_t pos) {
    const uint32_t v = bases16_at(bases, pos) & ((1u << (2 * MINI_M)) - 1u);                 // base pos in bits 1:0
    const uint32_t rc = rev2_32(~v) >> (32 - 2 * MINI_M);                                    // reverse complement, same encoding
    uint32_t c = v < rc ? v : rc;
    c ^= c >> 16; c *= 0x85ebca6bu; c ^= c >> 13; c *= 0xc2b2ae35u; c ^= c >> 16;
    return c;
}
// window minima are biased towards 0: re-mix before taking partition bits (top) and pass bits (low 16)
W2R_HD uint32_t mini_mix(uint32_t wmin) {
    uint32_t x = wmin * 0x9e3779b1u;
    x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 13;
    return x;
}
W2R_HD uint32_t mini_part(uint32_t mixed, uint32_t logP) { return logP ? mixed >> (32u - logP) : 0u; }
// reference form for tests: the minimiser hash of the k-mer starting at base j
W2R_HD uint32_t kmer_minimizer_hash(const uint8_t* bases, uint64_t j) {
    uint32_t m = 0xffffffffu;
    for (int t = 0; t < MINI_W; ++t) { const uint32_t h = mmer_hash_at(bases, j + t); if (h < m) m = h; }
    return m;
}

}  // namespace w2r
