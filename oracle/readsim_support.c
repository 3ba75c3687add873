/*
 * readsim_support.c — TEST INFRASTRUCTURE: writers for the reference's read-store encodings, used by the
 * synthetic read generator of tests/ and bench.py's cpu_baseline leg.  Not part of the product path.
 *
 * pq_encode*: produce a valid PQVec block stream (format: feudal/PQVec.cc:87-120, block size
 * feudal/PQVec.h "blockSize" = 1 + ceil((9 + nQs*nBits)/8)); any partition into blocks of 1..255 quals decodes to the
 * same qualities, so this encoder does not have to reproduce the reference encoder's block choices.
 *   mode 0: greedy — close the block when the delta width would grow (fast; bench-scale data)
 *   mode 1: minimum-size partition by dynamic programming (what an optimal encoder emits; test data)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static unsigned ceil_lg2(unsigned x) { unsigned b = 0; while ((1u << b) < x) ++b; return b; } /* x in 1..64 */
static unsigned block_size(unsigned nq, unsigned bits) { return 1 + ((9 + nq * bits + 7) >> 3); }

static uint8_t* emit_block(uint8_t* o, const uint8_t* q, unsigned nq, unsigned bits, unsigned minq) {
    *o++ = (uint8_t)nq;
    uint64_t acc = bits | ((uint64_t)minq << 3);
    unsigned have = 9;
    for (unsigned i = 0; i < nq; ++i) {
        acc |= (uint64_t)(q[i] - minq) << have; have += bits;
        while (have >= 8) { *o++ = (uint8_t)acc; acc >>= 8; have -= 8; }
    }
    while (have >= 8) { *o++ = (uint8_t)acc; acc >>= 8; have -= 8; }
    if (have) *o++ = (uint8_t)acc;
    return o;
}

/* returns bytes written, including the 0 terminator.  out must hold at least 2*n + 4 bytes. */
size_t sim_pq_encode(const uint8_t* q, uint32_t n, uint8_t* out, int mode) {
    uint8_t* o = out;
    if (mode == 0) {
        uint32_t i = 0;
        while (i < n) {
            unsigned mn = q[i], mx = q[i], bits = 0, len = 1;
            while (i + len < n && len < 255) {
                unsigned v = q[i + len], nmn = v < mn ? v : mn, nmx = v > mx ? v : mx;
                unsigned nb = ceil_lg2(nmx - nmn + 1);
                if (nb > bits && len >= 8) break;
                mn = nmn; mx = nmx; bits = nb; ++len;
            }
            o = emit_block(o, q + i, len, bits, mn);
            i += len;
        }
    } else {
        uint32_t* cost = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
        uint8_t* blen = (uint8_t*)malloc(n + 1);
        cost[0] = 1;
        for (uint32_t i = 1; i <= n; ++i) {
            unsigned mn = 63, mx = 0; uint32_t best = 0xffffffffu; unsigned bl = 1;
            for (unsigned len = 1; len <= 255 && len <= i; ++len) {
                unsigned v = q[i - len]; if (v < mn) mn = v; if (v > mx) mx = v;
                uint32_t c = cost[i - len] + block_size(len, ceil_lg2(mx - mn + 1));
                if (c < best) { best = c; bl = len; }
            }
            cost[i] = best; blen[i] = (uint8_t)bl;
        }
        uint32_t nb = 0; for (uint32_t i = n; i > 0; i -= blen[i]) ++nb;
        uint32_t* starts = (uint32_t*)malloc(sizeof(uint32_t) * (nb + 1));
        { uint32_t k = nb; for (uint32_t i = n; i > 0; i -= blen[i]) starts[--k] = i - blen[i]; starts[nb] = n; }
        for (uint32_t b = 0; b < nb; ++b) {
            unsigned mn = 63, mx = 0;
            for (uint32_t j = starts[b]; j < starts[b + 1]; ++j) { if (q[j] < mn) mn = q[j]; if (q[j] > mx) mx = q[j]; }
            o = emit_block(o, q + starts[b], starts[b + 1] - starts[b], ceil_lg2(mx - mn + 1), mn);
        }
        free(cost); free(blen); free(starts);
    }
    *o++ = 0;
    return (size_t)(o - out);
}

/* Flattens n reads given as rows of two [n x stride] byte matrices (base codes 0..3, quals 0..63) with per-read lengths
 * into the w2rap_reads layout.  bases_out needs sum(ceil(len/4)); quals_out needs sum(2*len+4).  Returns quals bytes. */
size_t sim_flatten_reads(uint64_t n, uint32_t stride, const uint8_t* codes, const uint8_t* quals, const uint32_t* len,
                         uint8_t* bases_out, uint64_t* base_off, uint8_t* quals_out, uint64_t* qual_off, int pq_mode) {
    uint64_t bo = 0, qo = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* c = codes + r * stride;
        uint32_t L = len[r], nb = (L + 3) / 4;
        base_off[r] = bo; qual_off[r] = qo;
        memset(bases_out + bo, 0, nb);
        for (uint32_t i = 0; i < L; ++i) bases_out[bo + (i >> 2)] |= (uint8_t)(c[i] << ((i & 3) * 2));
        bo += nb;
        qo += sim_pq_encode(quals + r * stride, L, quals_out + qo, pq_mode);
    }
    base_off[n] = bo; qual_off[n] = qo;
    return qo;
}
