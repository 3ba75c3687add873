// pqvec.cuh — streaming decode of the reference's block-compressed quality vectors (PQVec).
//
// Format (feudal/PQVec.cc:87-120 encoder, :122-188 decoder): a stream of blocks
//   u8 nQs (1..255) ; then LSB-first bits: 3 bits nBits, 6 bits minQ, nQs x nBits-bit deltas ; padded to a byte
// terminated by a 0 byte.  quality = minQ + delta.
#pragma once
#include "kmer.cuh"

namespace w2r {

// Sequential reader over one PQVec stream.  next() returns the next quality or -1 at the terminator.
struct PQReader {
    const uint8_t* p;
    uint64_t acc;      // bit accumulator (LSB = next bit)
    uint32_t have;     // valid bits in acc
    uint32_t left;     // quals left in the current block
    uint32_t nbits, minq;
    bool done;

    W2R_HD explicit PQReader(const uint8_t* stream) : p(stream), acc(0), have(0), left(0), nbits(0), minq(0), done(false) {}

    W2R_HD void fill(uint32_t need) { while (have < need) { acc |= (uint64_t)(*p++) << have; have += 8; } }

    W2R_HD int next() {
        if (left == 0) {
            if (done) return -1;
            uint32_t nq = *p++;           // block header byte (always byte aligned: acc is empty here)
            if (nq == 0) { done = true; return -1; }
            acc = 0; have = 0;
            fill(9);
            nbits = (uint32_t)acc & 7u; minq = (uint32_t)(acc >> 3) & 63u;
            acc >>= 9; have -= 9;
            left = nq;
        }
        uint32_t v = 0;
        if (nbits) { fill(nbits); v = (uint32_t)acc & ((1u << nbits) - 1u); acc >>= nbits; have -= nbits; }
        if (--left == 0) { acc = 0; have = 0; }   // the rest of the last byte is padding; p already points past it
        return (int)(minq + v);
    }
};

// paths/long/BuildReadQGraph.cc:962-987 count_good_lengths, evaluated in one forward pass: the reference scans from the END
// and stops at the first position where K consecutive quals >= minQual have been seen; that is the end of the right-most
// maximal run of >= K good quals, which a forward scan finds as "the last position at which the current run is >= K".
// The result is stored in a uint16_t by the reference.  *n_quals receives the number of qualities in the stream.
// 8 bytes from any address as a little-endian word (two aligned loads + a funnel shift; may touch up to 15 bytes past p: the
// quality store carries 32 bytes of padding)
W2R_HD uint64_t pq_load64(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p;
    const uint64_t* q = (const uint64_t*)(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7u) * 8u;
    const uint64_t lo = q[0], hi = q[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}
// `end` = one past the last byte of the stream (qual_off[i+1]): a block header or payload that would cross it (a truncated or
// corrupted .qualp) stops the walk and reports 0xffffffff qualities, which the caller turns into W2RAP_ERR_BAD_ARG.
W2R_HD uint32_t pq_good_length(const uint8_t* stream, const uint8_t* end, uint32_t min_qual, uint32_t* n_quals) {
    const uint8_t* p = stream;
    uint32_t run = 0, good = 0, i = 0;
    for (;;) {
        if (p >= end) { i = 0xffffffffu; break; }               // no terminator inside the stream
        const uint32_t nq = *p++;
        if (!nq) break;
        if (p + 2 > end) { i = 0xffffffffu; break; }
        const uint32_t hdr = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
        const uint32_t nbits = hdr & 7u, minq = (hdr >> 3) & 63u;
        if (p + ((9u + nq * nbits + 7u) >> 3) > end) { i = 0xffffffffu; break; }
        if (nbits == 0) {                       // a constant block is one step: the run grows by nq or is reset
            p += 2;
            i += nq;
            if (minq < min_qual) run = 0;
            else { run += nq; if (run >= (uint32_t)K) good = i; }
            continue;
        }
        if (minq >= min_qual) {                 // every quality of the block is >= minq >= minQual: no need to look at the deltas
            p += (9u + nq * nbits + 7u) >> 3;
            i += nq; run += nq;
            if (run >= (uint32_t)K) good = i;
            continue;
        }
        // Deltas below the floor have to be found: `per` deltas at a time from one 64-bit window, compared with the threshold all at
        // once (even and odd fields separately, so that each field has its neighbour's bits as carry room); only a window that
        // holds a low quality is walked value by value.
        const uint32_t thr = min_qual - minq;                  // quality >= min_qual  <=>  delta >= thr  (thr >= 1 here)
        const uint32_t nbytes = (9u + nq * nbits + 7u) >> 3;
        if (thr >> nbits) { run = 0; i += nq; p += nbytes; continue; }     // no delta of this width reaches the floor
        const uint32_t per = 56u / nbits;                      // fields per window (window start is byte aligned + up to 7 bits)
        const uint64_t field = (1ull << nbits) - 1ull;
        uint64_t ones_even = 0;
        for (uint32_t j = 0; j < per; j += 2) ones_even |= 1ull << (j * nbits);
        const uint64_t add = ((1ull << nbits) - thr) * ones_even, even_fields = field * ones_even;
        for (uint32_t k = 0; k < nq; k += per) {
            const uint32_t n = nq - k < per ? nq - k : per;
            const uint32_t bit = 9u + k * nbits;
            const uint64_t x = pq_load64(p + (bit >> 3)) >> (bit & 7u);
            // bit j*nbits of ge is set iff field j >= thr
            uint64_t ge = (((x & even_fields) + add) >> nbits) & ones_even;
            ge |= (((((x >> nbits) & even_fields) + add) >> nbits) & ones_even) << nbits;
            const uint64_t want = (ones_even | (ones_even << nbits)) & ((1ull << (n * nbits)) - 1ull);      // n * nbits <= 56
            if ((ge & want) == want) { run += n; i += n; if (run >= (uint32_t)K) good = i; continue; }
            for (uint32_t j = 0; j < n; ++j) {
                ++i;
                if (!((ge >> (j * nbits)) & 1ull)) run = 0;
                else if (++run >= (uint32_t)K) good = i;
            }
        }
        p += nbytes;
    }
    if (n_quals) *n_quals = i;
    return good & 0xffffu;
}
W2R_HD uint32_t pq_good_length(const uint8_t* stream, uint32_t min_qual, uint32_t* n_quals) {      // trusted stream (tests)
    return pq_good_length(stream, (const uint8_t*)~(uintptr_t)0, min_qual, n_quals);
}

// One quality by position, without decoding the vector: walk the block headers (a 250-base read has a handful of blocks) and pull
// the delta out of its block by bit offset.  Read pathing only needs qualities at the MISMATCHES of an extension overlap
// (paths/long/ExtendReadPath.cc:15-109), a few per read, so this replaces a full decode into scratch memory.
W2R_HD uint32_t pq_qual_at(const uint8_t* stream, uint32_t pos) {
    const uint8_t* p = stream;
    uint32_t i = 0;
    for (;;) {
        const uint32_t nq = *p;
        if (!nq) return 0;                                     // past the end (callers stay inside the read)
        const uint32_t hdr = (uint32_t)p[1] | ((uint32_t)p[2] << 8);
        const uint32_t nbits = hdr & 7u, minq = (hdr >> 3) & 63u;
        if (pos < i + nq) {
            if (!nbits) return minq;
            const uint32_t bit = 9u + (pos - i) * nbits;       // from the byte after the count
            const uint8_t* b = p + 1 + (bit >> 3);
            return minq + ((((uint32_t)b[0] | ((uint32_t)b[1] << 8)) >> (bit & 7u)) & ((1u << nbits) - 1u));
        }
        p += 1 + ((9u + nq * nbits + 7u) >> 3);
        i += nq;
    }
}

// Decodes up to `cap` quals into out; returns the number of quals in the stream.
W2R_HD uint32_t pq_decode(const uint8_t* stream, uint8_t* out, uint32_t cap) {
    PQReader r(stream);
    uint32_t i = 0;
    for (;;) {
        int q = r.next();
        if (q < 0) break;
        if (i < cap) out[i] = (uint8_t)q;
        ++i;
    }
    return i;
}

}  // namespace w2r
