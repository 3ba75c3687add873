/*
 * step2_oracle.c — CPU restatement of w2rap-contigger's step 2 (buildReadQGraph + FixPaths).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (w2rap-contigger_b200/) may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * Parity status: PINNED against the reference's own binary (oracle/_ref/w2rap-contigger, built from
 * /root/reference by oracle/Makefile) — see tests/test_oracle_vs_reference.py and tests/golden/.
 * The reference repository itself ships no tests or golden vectors for this path.
 *
 * Plain single-threaded C99, written for clarity, not speed.  Every function cites the reference
 * file:line whose behaviour it restates (paths relative to /root/reference/src).  No reference code is copied.
 *
 * Output order convention (shared with the CUDA path): canonical edges are sorted by sequence
 * (the reference's own edge numbering is a thread race, paths/long/BuildReadQGraph.cc:276-286).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/w2rap_step2.h"

#define KK 60

typedef struct { uint64_t w0, w1; } kmer_t;
typedef struct { uint64_t w0, w1; uint8_t ctx, count; } rec_t;

static void* xmalloc(size_t n) { void* p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); } return p; }
static void* xcalloc(size_t n, size_t s) { void* p = calloc(n ? n : 1, s ? s : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }
static void* xrealloc(void* q, size_t n) { void* p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return p; }

/* ---------------------------------------------------------------- bases and k-mers */

/* feudal/FieldVec.h:765-769 — base i of a packed read: byte i/4, bits 2*(i%4). */
static inline uint8_t packed_base(const uint8_t* p, uint64_t i) { return (p[i >> 2] >> ((i & 3) * 2)) & 3; }

/* kmers/KMer.h:155-162 — base 0 in the top two bits of word 0; bases 32..59 in bits 63..8 of word 1. */
static kmer_t kmer_from_codes(const uint8_t* c) {
    kmer_t k = {0, 0};
    for (int i = 0; i < 32; ++i) k.w0 = (k.w0 << 2) | c[i];
    for (int i = 32; i < KK; ++i) k.w1 = (k.w1 << 2) | c[i];
    k.w1 <<= 8;
    return k;
}
static void kmer_to_codes(kmer_t k, uint8_t* c) {
    for (int i = 0; i < 32; ++i) c[i] = (k.w0 >> (62 - 2 * i)) & 3;
    for (int i = 32; i < KK; ++i) c[i] = (k.w1 >> (62 - 2 * (i - 32))) & 3;
}
static inline int kmer_cmp(kmer_t a, kmer_t b) { /* kmers/KMer.h:312-319 */
    if (a.w0 != b.w0) return a.w0 < b.w0 ? -1 : 1;
    if (a.w1 != b.w1) return a.w1 < b.w1 ? -1 : 1;
    return 0;
}
static inline kmer_t kmer_succ(kmer_t k, uint8_t b) { /* kmers/KMer.h:191-203 toSuccessor */
    kmer_t r; r.w0 = (k.w0 << 2) | (k.w1 >> 62); r.w1 = (k.w1 << 2) | ((uint64_t)b << 8); return r;
}
static inline kmer_t kmer_pred(kmer_t k, uint8_t b) { /* kmers/KMer.h:176-189 toPredecessor */
    kmer_t r; r.w1 = ((k.w1 >> 2) | (k.w0 << 62)) & ~(uint64_t)0xff; r.w0 = (k.w0 >> 2) | ((uint64_t)b << 62); return r;
}
static kmer_t kmer_rc(kmer_t k) { /* kmers/KMer.h:205-227 */
    uint8_t c[KK], r[KK];
    kmer_to_codes(k, c);
    for (int i = 0; i < KK; ++i) r[i] = 3 - c[KK - 1 - i];
    return kmer_from_codes(r);
}
/* dna/CanonicalForm.h:51-63 for even KK: lexicographic compare of the k-mer with its reverse complement.
 * returns 0 FWD, 1 REV, 2 PALINDROME */
static inline int kmer_form(kmer_t k) { int c = kmer_cmp(k, kmer_rc(k)); return c < 0 ? 0 : (c > 0 ? 1 : 2); }

/* kmers/KMerContext.cc:19-37 — reverse-complementing a context reverses its 8 bits. */
static inline uint8_t ctx_rc(uint8_t c) {
    c = (uint8_t)((c >> 4) | (c << 4)); c = (uint8_t)(((c & 0xcc) >> 2) | ((c & 0x33) << 2)); c = (uint8_t)(((c & 0xaa) >> 1) | ((c & 0x55) << 1));
    return c;
}
static inline int pc4(unsigned m) { m &= 15; return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }
static inline int single_bit_code(unsigned m) { m &= 15; return m == 1 ? 0 : m == 2 ? 1 : m == 4 ? 2 : 3; }

/* dna/CanonicalForm.h:34-46 — form of a run-time-length sequence: odd length is decided by the middle base alone. */
static int seq_form(const uint8_t* s, uint64_t len) {
    if (len & 1) return (s[len / 2] & 2) ? 1 : 0;
    for (uint64_t i = 0, j = len; i < j; ++i) {
        --j;
        uint8_t f = s[i], r = s[j] ^ 3;
        if (f < r) return 0;
        if (r < f) return 1;
        if (i + 1 >= j) break;
    }
    return 2;
}
static void seq_rc_inplace(uint8_t* s, uint64_t len) {
    for (uint64_t i = 0, j = len - 1; i < j; ++i, --j) { uint8_t t = 3 - s[i]; s[i] = 3 - s[j]; s[j] = t; }
    if (len & 1) s[len / 2] = 3 - s[len / 2];
}

/* ---------------------------------------------------------------- PQVec */

/* feudal/PQVec.cc:122-188 (decode) / :87-120 (encode) — block stream: u8 nQs; then LSB-first bits: 3 nBits, 6 minQ,
 * nQs*nBits deltas; each block padded to a byte; a 0 byte terminates.  Returns the number of quals. */
static size_t pq_decode(const uint8_t* p, uint8_t* out) {
    size_t n = 0;
    for (;;) {
        unsigned nq = *p++;
        if (!nq) break;
        uint64_t bitpos = 0;
        unsigned hdr = p[0] | ((unsigned)p[1] << 8);
        unsigned nbits = hdr & 7, minq = (hdr >> 3) & 63;
        bitpos = 9;
        for (unsigned i = 0; i < nq; ++i) {
            unsigned v = 0;
            for (unsigned b = 0; b < nbits; ++b, ++bitpos) v |= ((p[bitpos >> 3] >> (bitpos & 7)) & 1u) << b;
            out[n++] = (uint8_t)(minq + v);
        }
        p += (bitpos + 7) >> 3;
    }
    return n;
}

/* paths/long/BuildReadQGraph.cc:962-987 — scan from the END; the first time KK consecutive quals >= minQual have been seen,
 * good_len = index + KK; stored in a uint16_t. */
static unsigned good_length(const uint8_t* q, size_t n, unsigned min_qual) {
    unsigned good = 0;
    for (size_t i = n; i-- > 0;) {
        if (q[i] < min_qual) good = 0;
        else if (++good == KK) return (unsigned)((i + KK) & 0xffff);
    }
    return 0;
}

/* ---------------------------------------------------------------- sort of k-mer records (std::sort in the reference, :1081) */

static void radix_sort_recs(rec_t* a, rec_t* tmp, size_t n) {
    size_t* cnt = (size_t*)xmalloc(sizeof(size_t) * 65537);
    for (int pass = 0; pass < 8; ++pass) {
        int word = pass < 4 ? 1 : 0, shift = (pass & 3) * 16;
        memset(cnt, 0, sizeof(size_t) * 65537);
        for (size_t i = 0; i < n; ++i) { uint64_t w = word ? a[i].w1 : a[i].w0; cnt[((w >> shift) & 0xffff) + 1]++; }
        for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; ++i) { uint64_t w = word ? a[i].w1 : a[i].w0; tmp[cnt[(w >> shift) & 0xffff]++] = a[i]; }
        rec_t* t = a; a = tmp; tmp = t;
    }
    free(cnt); /* 8 passes: result is back in the original `a` */
}

/* ---------------------------------------------------------------- the solid k-mer dictionary (kmers/ReadPather.h:176-349) */

typedef struct {
    size_t n;
    kmer_t* key;      /* sorted canonical k-mers */
    uint8_t* ctx;
    uint32_t* edge;   /* KDef edge id, 0xffffffff = null (kmers/ReadPather.h:104-145) */
    uint32_t* off;
} dict_t;

static long dict_find(const dict_t* d, kmer_t k) {
    size_t lo = 0, hi = d->n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; int c = kmer_cmp(d->key[mid], k); if (c == 0) return (long)mid; if (c < 0) lo = mid + 1; else hi = mid; }
    return -1;
}
/* kmers/ReadPather.h:196-199 findEntry: canonicalise, then look up.  *is_rev = the query was in REV form. */
static long dict_find_any(const dict_t* d, kmer_t k, int* is_rev) {
    kmer_t r = kmer_rc(k);
    if (kmer_cmp(r, k) < 0) { if (is_rev) *is_rev = 1; return dict_find(d, r); }
    if (is_rev) *is_rev = 0;
    return dict_find(d, k);
}
/* BuildReadQGraph.cc:261-273 EdgeBuilder::lookup — entry plus its context seen in the query's orientation. */
static long dict_lookup_oriented(const dict_t* d, kmer_t k, uint8_t* ctx) {
    int rev; long i = dict_find_any(d, k, &rev);
    if (i < 0) { fprintf(stderr, "oracle: lookup of a neighbour k-mer failed (reference ForceAssert, BuildReadQGraph.cc:265)\n"); abort(); }
    *ctx = rev ? ctx_rc(d->ctx[i]) : d->ctx[i];
    return i;
}

/* kmers/ReadPather.h:317-346 AdjProc — only clears bits whose neighbour k-mer is absent. */
static void recompute_adjacencies(dict_t* d) {
    uint8_t* nctx = (uint8_t*)xmalloc(d->n);
    for (size_t i = 0; i < d->n; ++i) {
        uint8_t c = d->ctx[i];
        for (unsigned b = 0; b < 4; ++b)
            if (c & (1u << b)) { if (dict_find_any(d, kmer_succ(d->key[i], (uint8_t)b), NULL) < 0) c &= (uint8_t)~(1u << b); }
        for (unsigned b = 0; b < 4; ++b)
            if (c & (16u << b)) { if (dict_find_any(d, kmer_pred(d->key[i], (uint8_t)b), NULL) < 0) c &= (uint8_t)~(16u << b); }
        nctx[i] = c;
    }
    memcpy(d->ctx, nctx, d->n);
    free(nctx);
}

/* ---------------------------------------------------------------- edges (BuildReadQGraph.cc:99-339) */

typedef struct { uint8_t* seq; uint64_t len; uint32_t* ents; uint64_t n_ents; } edge_t;
typedef struct { edge_t* e; size_t n, cap; } edgelist_t;

static void add_edge(edgelist_t* el, dict_t* d, uint8_t* seq, uint64_t len, uint32_t* ents, uint64_t n_ents) {
    /* BuildReadQGraph.cc:275-306 addEdge: a REV sequence is reverse-complemented, entries reversed; then every k-mer gets (edge, offset) */
    if (seq_form(seq, len) == 1) {
        seq_rc_inplace(seq, len);
        for (uint64_t i = 0, j = n_ents - 1; i < j; ++i, --j) { uint32_t t = ents[i]; ents[i] = ents[j]; ents[j] = t; }
    }
    if (el->n == el->cap) { el->cap = el->cap ? el->cap * 2 : 1024; el->e = (edge_t*)xrealloc(el->e, el->cap * sizeof(edge_t)); }
    edge_t* e = &el->e[el->n];
    e->seq = seq; e->len = len; e->ents = ents; e->n_ents = n_ents;
    for (uint64_t i = 0; i < n_ents; ++i) {
        if (d->edge[ents[i]] != 0xffffffffu) { fprintf(stderr, "oracle: preoccupied k-mer (reference FatalErr, BuildReadQGraph.cc:303)\n"); abort(); }
        d->edge[ents[i]] = (uint32_t)el->n; d->off[ents[i]] = (uint32_t)i;
    }
    el->n++;
}

typedef struct { uint8_t* seq; uint64_t len, cap; uint32_t* ents; uint64_t n_ents, ecap; } walk_t;
static void walk_init(walk_t* w, kmer_t first, uint32_t ent) {
    w->cap = 256; w->seq = (uint8_t*)xmalloc(w->cap); kmer_to_codes(first, w->seq); w->len = KK;
    w->ecap = 64; w->ents = (uint32_t*)xmalloc(w->ecap * sizeof(uint32_t)); w->ents[0] = ent; w->n_ents = 1;
}
static void walk_push(walk_t* w, uint8_t b, uint32_t ent) {
    if (w->len == w->cap) { w->cap *= 2; w->seq = (uint8_t*)xrealloc(w->seq, w->cap); }
    if (w->n_ents == w->ecap) { w->ecap *= 2; w->ents = (uint32_t*)xrealloc(w->ents, w->ecap * sizeof(uint32_t)); }
    w->seq[w->len++] = b; w->ents[w->n_ents++] = ent;
}

/* BuildReadQGraph.cc:192-214 */
static int upstream_possible(const dict_t* d, size_t i) {
    uint8_t c = d->ctx[i];
    if (pc4(c >> 4) != 1) return 0;
    kmer_t p = kmer_pred(d->key[i], (uint8_t)single_bit_code(c >> 4));
    if (kmer_form(p) == 2) return 0;
    uint8_t c2; dict_lookup_oriented(d, p, &c2);
    return pc4(c2) == 1;
}
static int downstream_possible(const dict_t* d, size_t i) {
    uint8_t c = d->ctx[i];
    if (pc4(c) != 1) return 0;
    kmer_t s = kmer_succ(d->key[i], (uint8_t)single_bit_code(c));
    if (kmer_form(s) == 2) return 0;
    uint8_t c2; dict_lookup_oriented(d, s, &c2);
    return pc4(c2 >> 4) == 1;
}
/* BuildReadQGraph.cc:234-259 extend: follow single successors; stop before a palindrome or a k-mer with != 1 predecessors;
 * keep the walk only if the whole sequence is FWD or PALINDROME. */
static void extend_walk(edgelist_t* el, dict_t* d, kmer_t start, uint8_t ctx, uint32_t ent) {
    walk_t w; walk_init(&w, start, ent);
    kmer_t x = start;
    while (pc4(ctx) == 1) {
        uint8_t b = (uint8_t)single_bit_code(ctx);
        kmer_t nx = kmer_succ(x, b);
        if (kmer_form(nx) == 2) break;
        uint8_t c2; long idx = dict_lookup_oriented(d, nx, &c2);
        if (pc4(c2 >> 4) != 1) break;
        walk_push(&w, b, (uint32_t)idx);
        x = nx; ctx = c2;
    }
    int form = seq_form(w.seq, w.len);
    if (form == 1) { free(w.seq); free(w.ents); return; }
    if (form == 2 && w.len != KK) { fprintf(stderr, "oracle: palindromic multi-k-mer edge (reference ForceAssertEq, BuildReadQGraph.cc:249)\n"); abort(); }
    add_edge(el, d, w.seq, w.len, w.ents, w.n_ents);
}
/* BuildReadQGraph.cc:104-115 buildEdge */
static void build_edge(edgelist_t* el, dict_t* d, size_t i) {
    kmer_t k = d->key[i];
    int pal = kmer_form(k) == 2;
    int up = !pal && upstream_possible(d, i);
    int down = !pal && downstream_possible(d, i);
    if (pal || (!up && !down)) { walk_t w; walk_init(&w, k, (uint32_t)i); add_edge(el, d, w.seq, w.len, w.ents, w.n_ents); return; }
    if (up && down) return;
    if (up) extend_walk(el, d, kmer_rc(k), ctx_rc(d->ctx[i]), (uint32_t)i);   /* extendUpstream :222-226 */
    else extend_walk(el, d, k, d->ctx[i], (uint32_t)i);                         /* extendDownstream :228-232 */
}
/* BuildReadQGraph.cc:126-180 simpleCircle + canonicalizeCircle */
static void simple_circle(edgelist_t* el, dict_t* d, size_t i) {
    walk_t w; walk_init(&w, d->key[i], (uint32_t)i);
    kmer_t x = d->key[i]; uint8_t ctx = d->ctx[i];
    for (;;) {
        if (pc4(ctx >> 4) != 1 || pc4(ctx) != 1) { fprintf(stderr, "oracle: circle with a branch (reference ForceAssertEq, BuildReadQGraph.cc:133-134)\n"); abort(); }
        uint8_t b = (uint8_t)single_bit_code(ctx);
        x = kmer_succ(x, b);
        long idx = dict_lookup_oriented(d, x, &ctx);
        if ((size_t)idx == i) break;
        if (d->edge[idx] != 0xffffffffu) { fprintf(stderr, "oracle: failed to close circle (BuildReadQGraph.cc:140-147)\n"); abort(); }
        walk_push(&w, b, (uint32_t)idx);
    }
    /* canonicalizeCircle: rotate to the minimum (canonical) k-mer, flipped so that it reads forward */
    uint64_t best = 0;
    for (uint64_t j = 1; j < w.n_ents; ++j) if (kmer_cmp(d->key[w.ents[j]], d->key[w.ents[best]]) < 0) best = j;
    uint64_t idx = best;
    if (kmer_form(kmer_from_codes(w.seq + idx)) == 1) {
        seq_rc_inplace(w.seq, w.len);
        for (uint64_t a = 0, b = w.n_ents - 1; a < b; ++a, --b) { uint32_t t = w.ents[a]; w.ents[a] = w.ents[b]; w.ents[b] = t; }
        idx = w.len - idx - KK;
    }
    if (idx) {
        uint8_t* s2 = (uint8_t*)xmalloc(w.len);
        uint64_t n = 0;
        for (uint64_t j = idx; j < w.len; ++j) s2[n++] = w.seq[j];
        for (uint64_t j = KK - 1; j < KK + idx - 1; ++j) s2[n++] = w.seq[j];
        uint32_t* e2 = (uint32_t*)xmalloc(w.n_ents * sizeof(uint32_t));
        uint64_t m = 0;
        for (uint64_t j = idx; j < w.n_ents; ++j) e2[m++] = w.ents[j];
        for (uint64_t j = 0; j < idx; ++j) e2[m++] = w.ents[j];
        free(w.seq); free(w.ents); w.seq = s2; w.ents = e2;
    }
    add_edge(el, d, w.seq, w.len, w.ents, w.n_ents);
}

static int edge_seq_cmp(const void* pa, const void* pb) {
    const edge_t* a = (const edge_t*)pa; const edge_t* b = (const edge_t*)pb;
    uint64_t n = a->len < b->len ? a->len : b->len;
    int c = memcmp(a->seq, b->seq, n);
    if (c) return c;
    return a->len < b->len ? -1 : (a->len > b->len ? 1 : 0);
}

/* ---------------------------------------------------------------- HBV (paths/long/HBVFromEdges.cc:76-154) */

typedef struct { uint64_t hash; const uint8_t* seq; uint64_t len; int rc, distal; uint64_t edge; } end_t;
static inline uint8_t end_base(const end_t* e, int i) { /* feudal/BaseVec.h:98-126 SwitchHitterIter */
    uint64_t pos = (e->distal ? e->len - (KK - 1) : 0) + (uint64_t)i;
    return e->rc ? (uint8_t)(3 - e->seq[e->len - 1 - pos]) : e->seq[pos];
}
static int end_cmp(const void* pa, const void* pb) { /* HBVFromEdges.cc:34-38: hash first, then base-wise */
    const end_t* a = (const end_t*)pa; const end_t* b = (const end_t*)pb;
    if (a->hash != b->hash) return a->hash < b->hash ? -1 : 1;
    for (int i = 0; i < KK - 1; ++i) { uint8_t x = end_base(a, i), y = end_base(b, i); if (x != y) return x < y ? -1 : 1; }
    return 0;
}
static int end_cmp_stable(const void* pa, const void* pb) {
    int c = end_cmp(pa, pb);
    if (c) return c;
    const end_t* a = (const end_t*)pa; const end_t* b = (const end_t*)pb; /* equal keys get the same vertex; order is irrelevant */
    return a->edge < b->edge ? -1 : (a->edge > b->edge ? 1 : 0);
}

typedef struct {
    uint64_t n_v, n_e;          /* vertices, hbv edges */
    int32_t* left; int32_t* right; /* per hbv edge: true source / target vertex */
    uint32_t* elen;             /* per hbv edge: bases */
    uint64_t* canon;            /* per hbv edge: canonical edge index */
    uint8_t* is_rc;             /* per hbv edge: 1 if it is the reverse complement of the canonical edge */
    uint64_t* from_off; int32_t* from_v; int32_t* from_e;  /* From(v)/FromEdgeObj(v) in reference order */
    uint64_t* to_off; int32_t* to_v; int32_t* to_e;        /* To(v)/ToEdgeObj(v) */
} hbv_t;

/* paths/HyperBasevector.cc:648-660 HyperBasevector::Involution: sort all hbv edges by sequence, sort their reverse complements by
 * sequence, pair rank with rank: inv[e] = the edge whose reverse complement spells e. */
typedef struct { const uint8_t* seq; uint64_t len; int rc; int32_t id; } invkey_t;
static inline uint8_t invkey_base(const invkey_t* k, uint64_t i) { return k->rc ? (uint8_t)(3 - k->seq[k->len - 1 - i]) : k->seq[i]; }
static int invkey_cmp(const void* pa, const void* pb) {   /* feudal/BaseVec.h operator<: base-wise, then the shorter one first */
    const invkey_t* a = (const invkey_t*)pa; const invkey_t* b = (const invkey_t*)pb;
    uint64_t n = a->len < b->len ? a->len : b->len;
    for (uint64_t i = 0; i < n; ++i) { uint8_t x = invkey_base(a, i), y = invkey_base(b, i); if (x != y) return x < y ? -1 : 1; }
    if (a->len != b->len) return a->len < b->len ? -1 : 1;
    return a->id < b->id ? -1 : (a->id > b->id ? 1 : 0);
}

typedef struct { int32_t a, b, e; } adj_t;
static int adj_cmp(const void* pa, const void* pb) { /* graph/DigraphTemplate.h:1829-1839: upper_bound insertion => (owner, neighbour, edge id) order */
    const adj_t* x = (const adj_t*)pa; const adj_t* y = (const adj_t*)pb;
    if (x->a != y->a) return x->a < y->a ? -1 : 1;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    return x->e < y->e ? -1 : (x->e > y->e ? 1 : 0);
}

static inline uint8_t hbv_edge_base(const hbv_t* h, const edge_t* edges, int32_t he, uint64_t pos) {
    const edge_t* e = &edges[h->canon[he]];
    return h->is_rc[he] ? (uint8_t)(3 - e->seq[e->len - 1 - pos]) : e->seq[pos];
}

/* ---------------------------------------------------------------- read pathing (BuildReadQGraph.cc:403-564, 804-929) */

typedef struct { int gap; uint32_t edge; int rc; uint32_t off, len, elen; } part_t;

static inline uint8_t edge_base_oriented(const edge_t* e, int rc, uint64_t pos) { return rc ? (uint8_t)(3 - e->seq[e->len - 1 - pos]) : e->seq[pos]; }

/* BuildReadQGraph.cc:500-550 BRQ_Pather::path */
static size_t path_read(const dict_t* d, const edge_t* edges, const uint8_t* read, uint32_t rlen, part_t* parts) {
    size_t np = 0;
    if (rlen < KK) { parts[np].gap = 1; parts[np].len = rlen; parts[np].edge = 0; parts[np].rc = 0; parts[np].off = 0; parts[np].elen = 0; return 1; }
    uint32_t nk = rlen - KK + 1, itr = 0;
    while (itr < nk) {
        kmer_t km = kmer_from_codes(read + itr);
        long ent = dict_find_any(d, km, NULL);
        if (ent < 0) {
            uint32_t gap = 1; ++itr;
            while (itr < nk) {
                km = kmer_from_codes(read + itr);
                ent = dict_find_any(d, km, NULL);
                if (ent >= 0) break;
                ++gap; ++itr;
            }
            part_t g = {1, 0, 0, 0, gap, 0}; parts[np++] = g;
        }
        if (ent >= 0) {
            const edge_t* e = &edges[d->edge[ent]];
            uint32_t o = d->off[ent];
            /* dna/CanonicalForm.h:85-92 isRC: the read k-mer differs from the edge k-mer at that offset */
            int rc = memcmp(read + itr, e->seq + o, KK) != 0;
            uint32_t len = 1, off;
            if (!rc) {
                uint64_t rp = itr + KK, ep = o + KK;
                while (rp < rlen && ep < e->len && read[rp] == e->seq[ep]) { ++len; ++rp; ++ep; }
                off = o;
            } else {
                uint64_t ro = e->len - o;      /* position in rc(edge) just past the k-mer */
                uint64_t rp = itr + KK, ep = ro;
                while (rp < rlen && ep < e->len && read[rp] == edge_base_oriented(e, 1, ep)) { ++len; ++rp; ++ep; }
                off = (uint32_t)(ro - KK);
            }
            part_t p = {0, d->edge[ent], rc, off, len, (uint32_t)(e->len - KK + 1)}; parts[np++] = p;
            itr += len;
        }
    }
    return np;
}
/* BuildReadQGraph.cc:552-558 isJoinable: compares the LAST (KK-1)-mer of both oriented edges */
static int is_joinable(const edge_t* edges, const part_t* a, const part_t* b) {
    if (a->edge == b->edge) return 1;
    const edge_t* e1 = &edges[a->edge]; const edge_t* e2 = &edges[b->edge];
    for (int i = 0; i < KK - 1; ++i)
        if (edge_base_oriented(e1, a->rc, e1->len - (KK - 1) + i) != edge_base_oriented(e2, b->rc, e2->len - (KK - 1) + i)) return 0;
    return 1;
}
static inline int same_edge(const part_t* a, const part_t* b) { return a->edge == b->edge && a->rc == b->rc; }

/* paths/long/ExtendReadPath.cc:15-109 — quality-weighted overlap scores.  `penalty -= 0.2*penalty` (unsigned -= double) equals
 * the integer 4*penalty/5 over the whole reachable range (checked for every penalty < 2e6). */
static unsigned score_left(const uint8_t* bases, const uint8_t* quals, size_t start, const hbv_t* h, const edge_t* edges, int32_t he) {
    unsigned qsum = 0, pen = 0;
    long b = (long)start - 1, e = (long)h->elen[he] - KK;       /* bases.rend()-start ; edge.rbegin()+(KK-1) */
    while (b >= 0 && e >= 0) {
        if (bases[b] != hbv_edge_base(h, edges, he, (uint64_t)e)) { unsigned q = quals[b] == 2 ? 20u : quals[b]; pen += q; qsum += pen; }
        else if (pen > 0) pen = 4 * pen / 5;
        --b; --e;
    }
    while (b >= 0) { qsum += 10; --b; }                        /* left-over read bases */
    return qsum;
}
static unsigned score_right(const uint8_t* bases, const uint8_t* quals, uint32_t rlen, size_t start, const hbv_t* h, const edge_t* edges, int32_t he) {
    unsigned qsum = 0, pen = 0;
    uint64_t b = rlen - start, e = KK - 1, elen = h->elen[he];  /* bases.end()-start ; edge.begin()+(KK-1) */
    while (b < rlen && e < elen) {
        if (bases[b] != hbv_edge_base(h, edges, he, e)) { unsigned q = quals[b] == 2 ? 20u : quals[b]; pen += q; qsum += pen; }
        else if (pen > 0) pen = 4 * pen / 5;
        ++b; ++e;
    }
    while (b < rlen) { qsum += 10; ++b; }
    return qsum;
}

typedef struct { int32_t offset; int32_t* e; size_t n, cap; } rpath_t;
static void rpath_push_back(rpath_t* p, int32_t e) { if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 8; p->e = (int32_t*)xrealloc(p->e, p->cap * sizeof(int32_t)); } p->e[p->n++] = e; }
static void rpath_push_front(rpath_t* p, int32_t e) { rpath_push_back(p, 0); memmove(p->e + 1, p->e, (p->n - 1) * sizeof(int32_t)); p->e[0] = e; }

/* Shared classification of ExtendReadPath.cc:150-200 (left) and :262-318 (right).
 * cand_e/cand_v: candidate edges and their far vertices in list order.  Returns the chosen hbv edge or -1. */
static int32_t choose_extension(const hbv_t* h, const edge_t* edges, const int32_t* cand_e, const int32_t* cand_v, size_t nc, size_t last_gap,
                                int leftward, const uint8_t* bases, const uint8_t* quals, uint32_t rlen) {
    int solo = nc == 1;
    size_t nlong = 0, nshort = 0; int32_t short_first = -1; int short_multi = 0;
    uint8_t* hanging = (uint8_t*)xcalloc(nc, 1);
    for (size_t i = 0; i < nc; ++i) {
        int32_t v = cand_v[i];
        uint64_t to_sz = h->to_off[v + 1] - h->to_off[v], from_sz = h->from_off[v + 1] - h->from_off[v];
        if (leftward ? (to_sz == 0 && from_sz == 1) : (from_sz == 0 && to_sz == 1)) hanging[i] = 1;
        int is_long = (size_t)(h->elen[cand_e[i]] - (KK - 1)) >= last_gap;
        if (is_long) ++nlong;
        if (!is_long && !hanging[i]) { if (!nshort) short_first = v; else if (v != short_first) short_multi = 1; ++nshort; }
    }
    if (!solo && nshort > 0) {
        int32_t v = short_first;
        uint64_t to_sz = h->to_off[v + 1] - h->to_off[v], from_sz = h->from_off[v + 1] - h->from_off[v];
        if (nlong > 0 || short_multi || (leftward ? to_sz : from_sz) != 1) { free(hanging); return -1; }
    }
    int32_t least_edge = -1; unsigned least = 0xffffffffu;
    for (size_t i = 0; i < nc; ++i) {
        if (!hanging[i] || solo) {
            unsigned s = leftward ? score_left(bases, quals, last_gap, h, edges, cand_e[i]) : score_right(bases, quals, rlen, last_gap, h, edges, cand_e[i]);
            if (s < least) { least = s; least_edge = cand_e[i]; }
        }
    }
    free(hanging);
    if (least_edge == -1 || least > last_gap * 10) return -1;
    return least_edge;
}
/* ExtendReadPath.cc:124-232 */
static int extend_left(rpath_t* p, const hbv_t* h, const edge_t* edges, const uint8_t* bases, const uint8_t* quals, uint32_t rlen) {
    if (!p->n || p->offset >= 0) return 0;
    size_t last_gap = (size_t)(-(long)p->offset);
    if (last_gap < 10) return 0;
    int32_t v = h->left[p->e[0]];
    size_t nc = h->to_off[v + 1] - h->to_off[v];
    int32_t e = choose_extension(h, edges, h->to_e + h->to_off[v], h->to_v + h->to_off[v], nc, last_gap, 1, bases, quals, rlen);
    if (e < 0) return 0;
    p->offset += (int32_t)(h->elen[e] - KK + 1);
    rpath_push_front(p, e);
    return 1;
}
/* ExtendReadPath.cc:235-348; `to_right` is in fact ToLeft (BuildReadQGraph.cc:836-841) */
static int extend_right(rpath_t* p, const hbv_t* h, const edge_t* edges, const uint8_t* bases, const uint8_t* quals, uint32_t rlen) {
    if (!p->n) return 0;
    int last = (int)rlen + p->offset;
    for (size_t i = 0; i < p->n; ++i) last -= (int)(h->elen[p->e[i]] - KK + 1);
    last -= KK - 1;
    if (last < 10) return 0;
    int32_t v = h->left[p->e[p->n - 1]];          /* the reference quirk: left vertex of the last edge */
    size_t nc = h->from_off[v + 1] - h->from_off[v];
    int32_t e = choose_extension(h, edges, h->from_e + h->from_off[v], h->from_v + h->from_off[v], nc, (size_t)last, 0, bases, quals, rlen);
    if (e < 0) return 0;
    rpath_push_back(p, e);
    return 1;
}

/* ---------------------------------------------------------------- the whole step */

static void pack_bases(const uint8_t* codes, uint64_t n, uint8_t* out) {
    memset(out, 0, (n + 3) / 4);
    for (uint64_t i = 0; i < n; ++i) out[i >> 2] |= (uint8_t)(codes[i] << ((i & 3) * 2));
}

/* a place: a vector of hbv edge ids; order = std::vector<int>::operator< (element by element, a proper prefix first) */
typedef struct { const int32_t* p; size_t n; } place_t;
static int place_cmp(const void* a_, const void* b_) {
    const place_t* a = (const place_t*)a_; const place_t* b = (const place_t*)b_;
    const size_t n = a->n < b->n ? a->n : b->n;
    for (size_t i = 0; i < n; ++i) if (a->p[i] != b->p[i]) return a->p[i] < b->p[i] ? -1 : 1;
    return a->n < b->n ? -1 : (a->n > b->n ? 1 : 0);
}

int oracle_step2_run(const w2rap_reads* in, const w2rap_params* p, w2rap_graph* out) {
    if (!in || !p || !out || p->K != KK) return W2RAP_ERR_BAD_ARG;
    memset(out, 0, sizeof(*out));
    const uint64_t n_reads = in->n_reads;
    uint32_t maxlen = 0;
    uint64_t n_bases = 0;
    for (uint64_t r = 0; r < n_reads; ++r) { if (in->len[r] > maxlen) maxlen = in->len[r]; n_bases += in->len[r]; }
    uint8_t* qbuf = (uint8_t*)xmalloc((size_t)maxlen + 512);
    uint8_t* rbuf = (uint8_t*)xmalloc((size_t)maxlen + 4);

    /* ---- BuildReadQGraph.cc:1052-1083: quality-floored reads -> (canonical k-mer, context, 1) records */
    uint16_t* good = (uint16_t*)xcalloc(n_reads, sizeof(uint16_t));
    uint64_t n_inst = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        size_t nq = pq_decode(in->quals + in->qual_off[r], qbuf);
        unsigned gl = good_length(qbuf, nq, p->min_qual);
        if (gl > in->len[r]) gl = in->len[r];   /* quals and bases have equal length in every valid store */
        good[r] = (uint16_t)gl;
        if (gl > KK) n_inst += gl - KK + 1;
    }
    rec_t* recs = (rec_t*)xmalloc(sizeof(rec_t) * (size_t)n_inst);
    uint64_t ri = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        unsigned gl = good[r];
        if (gl <= KK) continue;                                   /* :1064, strictly greater */
        const uint8_t* pb = in->bases + in->base_off[r];
        for (unsigned i = 0; i < gl; ++i) rbuf[i] = packed_base(pb, i);
        kmer_t f = kmer_from_codes(rbuf), rc = kmer_rc(f);
        for (unsigned j = 0; j + KK <= gl; ++j) {
            if (j) { f = kmer_succ(f, rbuf[j + KK - 1]); rc = kmer_pred(rc, (uint8_t)(3 - rbuf[j + KK - 1])); }
            uint8_t c = 0;
            if (j > 0) c |= (uint8_t)(16u << rbuf[j - 1]);       /* predecessor bit (KMerContext.h:100-105) */
            if (j + KK < gl) c |= (uint8_t)(1u << rbuf[j + KK]);   /* successor bit */
            rec_t* q = &recs[ri++];
            if (kmer_cmp(rc, f) < 0) { q->w0 = rc.w0; q->w1 = rc.w1; q->ctx = ctx_rc(c); }   /* :1069 isRev -> store RC and RC'd context */
            else { q->w0 = f.w0; q->w1 = f.w1; q->ctx = c; }
            q->count = 1;
        }
    }
    free(good);
    /* ---- :1081-1082 sort + collapse (OR contexts, saturating count) */
    rec_t* tmp = (rec_t*)xmalloc(sizeof(rec_t) * (size_t)n_inst);
    radix_sort_recs(recs, tmp, (size_t)n_inst);
    free(tmp);
    uint64_t nd = 0;
    for (uint64_t i = 0; i < n_inst;) {
        uint64_t j = i; unsigned cnt = 0; uint8_t c = 0;
        while (j < n_inst && recs[j].w0 == recs[i].w0 && recs[j].w1 == recs[i].w1) { c |= recs[j].ctx; cnt = cnt + 1 > 255 ? 255 : cnt + 1; ++j; }
        recs[nd].w0 = recs[i].w0; recs[nd].w1 = recs[i].w1; recs[nd].ctx = c; recs[nd].count = (uint8_t)cnt; ++nd;
        i = j;
    }
    /* ---- :1090-1114 histogram, min-frequency filter, dictionary */
    dict_t d; memset(&d, 0, sizeof d);
    for (uint64_t i = 0; i < nd; ++i) { out->hist[recs[i].count > 100 ? 100 : recs[i].count]++; if (recs[i].count >= p->min_freq) d.n++; }
    d.key = (kmer_t*)xmalloc(sizeof(kmer_t) * d.n); d.ctx = (uint8_t*)xmalloc(d.n);
    d.edge = (uint32_t*)xmalloc(sizeof(uint32_t) * d.n); d.off = (uint32_t*)xcalloc(d.n, sizeof(uint32_t));
    { size_t m = 0; for (uint64_t i = 0; i < nd; ++i) if (recs[i].count >= p->min_freq) { d.key[m].w0 = recs[i].w0; d.key[m].w1 = recs[i].w1; d.ctx[m] = recs[i].ctx; d.edge[m] = 0xffffffffu; ++m; } }
    out->n_reads = n_reads; out->n_bases = n_bases; out->n_kmer_instances = n_inst; out->n_distinct = nd; out->n_solid = d.n;
    if (p->dump_kmers == 2) {
        out->n_dump = nd; out->dump = (w2rap_kmer_rec*)xmalloc(sizeof(w2rap_kmer_rec) * nd);
        for (uint64_t i = 0; i < nd; ++i) { w2rap_kmer_rec k = {recs[i].w0, recs[i].w1, recs[i].count, recs[i].ctx, 0xffffffffu, 0}; out->dump[i] = k; }
    }
    free(recs);

    /* ---- :1278 */
    recompute_adjacencies(&d);

    /* ---- :1284 buildEdges: regular edges, then smooth circles */
    edgelist_t el; memset(&el, 0, sizeof el);
    for (size_t i = 0; i < d.n; ++i) if (d.edge[i] == 0xffffffffu) build_edge(&el, &d, i);
    for (size_t i = 0; i < d.n; ++i) if (d.edge[i] == 0xffffffffu) simple_circle(&el, &d, i);
    /* deterministic edge order: by sequence (the reference's order is a race) */
    qsort(el.e, el.n, sizeof(edge_t), edge_seq_cmp);
    for (size_t e = 0; e < el.n; ++e) for (uint64_t j = 0; j < el.e[e].n_ents; ++j) { d.edge[el.e[e].ents[j]] = (uint32_t)e; d.off[el.e[e].ents[j]] = (uint32_t)j; }
    if (p->dump_kmers == 1) {
        out->n_dump = d.n; out->dump = (w2rap_kmer_rec*)xmalloc(sizeof(w2rap_kmer_rec) * d.n);
        for (size_t i = 0; i < d.n; ++i) { w2rap_kmer_rec k = {d.key[i].w0, d.key[i].w1, 0, d.ctx[i], d.edge[i], d.off[i]}; out->dump[i] = k; }
    }

    out->n_edges = el.n;
    out->edge_off = (uint64_t*)xmalloc(sizeof(uint64_t) * (el.n + 1));
    out->edge_len = (uint32_t*)xmalloc(sizeof(uint32_t) * el.n);
    { uint64_t o = 0; for (size_t e = 0; e < el.n; ++e) { out->edge_off[e] = o; out->edge_len[e] = (uint32_t)el.e[e].len; o += (el.e[e].len + 3) / 4; out->n_edge_bases += el.e[e].len; } out->edge_off[el.n] = o;
      out->edge_bases = (uint8_t*)xmalloc(o);
      for (size_t e = 0; e < el.n; ++e) pack_bases(el.e[e].seq, el.e[e].len, out->edge_bases + out->edge_off[e]); }

    /* ---- :1311 buildHBVFromEdges (HBVFromEdges.cc:76-154) */
    hbv_t h; memset(&h, 0, sizeof h);
    end_t* ends = (end_t*)xmalloc(sizeof(end_t) * 4 * el.n);
    size_t n_ends = 0;
    uint8_t* epal = (uint8_t*)xmalloc(el.n);
    for (size_t e = 0; e < el.n; ++e) {
        epal[e] = seq_form(el.e[e].seq, el.e[e].len) == 2;
        for (int rc = 0; rc < (epal[e] ? 1 : 2); ++rc) for (int distal = 0; distal < 2; ++distal) {
            end_t* x = &ends[n_ends++];
            x->seq = el.e[e].seq; x->len = el.e[e].len; x->rc = rc; x->distal = distal; x->edge = e;
            uint64_t hsh = 14695981039346656037ull;                 /* math/Hash.h:26-35 over one byte per base code */
            for (int i = 0; i < KK - 1; ++i) hsh = 1099511628211ull * (hsh ^ end_base(x, i));
            x->hash = hsh;
        }
    }
    qsort(ends, n_ends, sizeof(end_t), end_cmp_stable);
    out->edge_vertices = (int32_t*)xmalloc(sizeof(int32_t) * 4 * el.n);
    for (size_t i = 0; i < 4 * el.n; ++i) out->edge_vertices[i] = -1;
    { int64_t v = 0;
      for (size_t i = 0; i < n_ends; ++i) {
          if (i > 0 && end_cmp(&ends[i - 1], &ends[i]) != 0) ++v;
          out->edge_vertices[4 * ends[i].edge + 2 * ends[i].rc + ends[i].distal] = (int32_t)v;
      }
      h.n_v = n_ends ? (uint64_t)v + 1 : 0; }
    free(ends);
    out->n_vertices = h.n_v;
    out->fwd_xlat = (int32_t*)xmalloc(sizeof(int32_t) * el.n);
    out->rev_xlat = (int32_t*)xmalloc(sizeof(int32_t) * el.n);
    for (size_t e = 0; e < el.n; ++e) { out->fwd_xlat[e] = (int32_t)h.n_e++; out->rev_xlat[e] = epal[e] ? out->fwd_xlat[e] : (int32_t)h.n_e++; }
    out->n_hbv_edges = h.n_e;
    h.left = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e); h.right = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e);
    h.elen = (uint32_t*)xmalloc(sizeof(uint32_t) * h.n_e); h.canon = (uint64_t*)xmalloc(sizeof(uint64_t) * h.n_e); h.is_rc = (uint8_t*)xmalloc(h.n_e);
    for (size_t e = 0; e < el.n; ++e) {
        int32_t f = out->fwd_xlat[e], r = out->rev_xlat[e];
        h.left[f] = out->edge_vertices[4 * e]; h.right[f] = out->edge_vertices[4 * e + 1]; h.elen[f] = (uint32_t)el.e[e].len; h.canon[f] = e; h.is_rc[f] = 0;
        if (r != f) { h.left[r] = out->edge_vertices[4 * e + 2]; h.right[r] = out->edge_vertices[4 * e + 3]; h.elen[r] = (uint32_t)el.e[e].len; h.canon[r] = e; h.is_rc[r] = 1; }
    }
    {   /* step 3's first question (w2rap-contigger.cc:361-371): the involution of the hbv edges */
        invkey_t* x1 = (invkey_t*)xmalloc(sizeof(invkey_t) * h.n_e); invkey_t* x2 = (invkey_t*)xmalloc(sizeof(invkey_t) * h.n_e);
        for (uint64_t i = 0; i < h.n_e; ++i) {
            const edge_t* ce = &el.e[h.canon[i]];
            x1[i].seq = x2[i].seq = ce->seq; x1[i].len = x2[i].len = ce->len; x1[i].id = x2[i].id = (int32_t)i;
            x1[i].rc = h.is_rc[i]; x2[i].rc = !h.is_rc[i];
        }
        qsort(x1, h.n_e, sizeof(invkey_t), invkey_cmp); qsort(x2, h.n_e, sizeof(invkey_t), invkey_cmp);
        out->involution = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e);
        for (uint64_t i = 0; i < h.n_e; ++i) out->involution[x1[i].id] = x2[i].id;
        free(x1); free(x2);
    }
    { adj_t* a = (adj_t*)xmalloc(sizeof(adj_t) * h.n_e);
      h.from_off = (uint64_t*)xcalloc(h.n_v + 1, sizeof(uint64_t)); h.to_off = (uint64_t*)xcalloc(h.n_v + 1, sizeof(uint64_t));
      h.from_v = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e); h.from_e = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e);
      h.to_v = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e); h.to_e = (int32_t*)xmalloc(sizeof(int32_t) * h.n_e);
      for (uint64_t e = 0; e < h.n_e; ++e) { a[e].a = h.left[e]; a[e].b = h.right[e]; a[e].e = (int32_t)e; }
      qsort(a, h.n_e, sizeof(adj_t), adj_cmp);
      for (uint64_t i = 0; i < h.n_e; ++i) { h.from_off[a[i].a + 1]++; h.from_v[i] = a[i].b; h.from_e[i] = a[i].e; }
      for (uint64_t e = 0; e < h.n_e; ++e) { a[e].a = h.right[e]; a[e].b = h.left[e]; a[e].e = (int32_t)e; }
      qsort(a, h.n_e, sizeof(adj_t), adj_cmp);
      for (uint64_t i = 0; i < h.n_e; ++i) { h.to_off[a[i].a + 1]++; h.to_v[i] = a[i].b; h.to_e[i] = a[i].e; }
      for (uint64_t v = 0; v < h.n_v; ++v) { h.from_off[v + 1] += h.from_off[v]; h.to_off[v + 1] += h.to_off[v]; }
      free(a); }

    /* ---- :1316 path_reads_OMP (+ FixPaths, large/GapToyTools.cc:322-335, if asked) */
    if (p->want_paths) {
        out->n_paths = n_reads;
        out->path_offset = (int32_t*)xcalloc(n_reads, sizeof(int32_t));
        out->path_off = (uint64_t*)xcalloc(n_reads + 1, sizeof(uint64_t));
        size_t pe_cap = (size_t)n_reads * 2 + 16; out->path_edges = (int32_t*)xmalloc(sizeof(int32_t) * pe_cap);
        part_t* parts = (part_t*)xmalloc(sizeof(part_t) * ((size_t)maxlen + 4));
        part_t* np = (part_t*)xmalloc(sizeof(part_t) * ((size_t)maxlen + 4));
        rpath_t rp; memset(&rp, 0, sizeof rp);
        for (uint64_t r = 0; r < n_reads; ++r) {
            uint32_t rlen = in->len[r];
            const uint8_t* pb = in->bases + in->base_off[r];
            for (uint32_t i = 0; i < rlen; ++i) rbuf[i] = packed_base(pb, i);
            size_t n = path_read(&d, el.e, rbuf, rlen, parts);
            /* :848-870 hanging-seed rule.  toRight is built with ToLeft (:838), so vright == vleft and the condition
             * ToSize(v)==0 && ToSize(v)>1 can never hold; only the merging of adjacent gaps remains. */
            size_t m = 0;
            for (size_t i = 0; i < n; ++i) {
                part_t pt = parts[i];
                if (!pt.gap) {
                    int32_t he = pt.rc ? out->rev_xlat[pt.edge] : out->fwd_xlat[pt.edge];
                    int32_t vl = h.left[he], vr = h.left[he];
                    uint64_t to_l = h.to_off[vl + 1] - h.to_off[vl], to_r = h.to_off[vr + 1] - h.to_off[vr], from_r = h.from_off[vr + 1] - h.from_off[vr];
                    if (to_l == 0 && to_r > 1 && from_r > 0 && pt.elen <= 100) { part_t g = {1, 0, 0, 0, pt.len, 0}; pt = g; }
                }
                if (pt.gap && m && np[m - 1].gap) np[m - 1].len += pt.len; else np[m++] = pt;
            }
            memcpy(parts, np, m * sizeof(part_t)); n = m;
            /* :875-898 captured-gap consistency */
            if (n >= 3) {
                size_t seeds = parts[0].gap ? 0 : 1;
                for (size_t i = 1; i + 1 < n; ++i) {
                    if (!parts[i].gap) { ++seeds; continue; }
                    const part_t* pv = &parts[i - 1]; const part_t* nx = &parts[i + 1];
                    uint32_t gd = nx->off - (pv->off + pv->len);                  /* :470 unsigned arithmetic */
                    if (!same_edge(pv, nx)) gd += pv->elen;
                    int32_t diff = (int32_t)(parts[i].len - gd);
                    uint32_t ad = (uint32_t)(diff < 0 ? -diff : diff);
                    if (!(ad <= 3u) || !is_joinable(el.e, pv, nx)) {
                        if (seeds > 1) {
                            uint32_t tot = parts[i - 1].len;
                            for (size_t j = i; j < n; ++j) tot += parts[j].len;
                            part_t g = {1, 0, 0, 0, tot, 0}; parts[i - 1] = g; n = i;
                        } else {
                            for (size_t j = i + 1; j < n; ++j) parts[i].len += parts[j].len;
                            n = i + 1;
                        }
                        break;
                    }
                }
            }
            /* :904-918 short trailing seed back-off */
            if (parts[n - 1].gap && n > 1) {
                const part_t* l2 = &parts[n - 2];
                if (l2->off == 0 && l2->len <= 5) { part_t g = parts[n - 1]; g.len += l2->len; parts[n - 2] = g; n -= 1; }
            } else if (!parts[n - 1].gap) {
                if (parts[n - 1].off == 0 && parts[n - 1].len <= 5) { part_t g = {1, 0, 0, 0, parts[n - 1].len, 0}; parts[n - 1] = g; }
            }
            /* :804-827 pathPartsToReadPath */
            rp.n = 0; rp.offset = 0;
            { const part_t* last = NULL;
              for (size_t i = 0; i < n; ++i) {
                  if (parts[i].gap) continue;
                  if (last && same_edge(last, &parts[i])) continue;
                  rpath_push_back(&rp, parts[i].rc ? out->rev_xlat[parts[i].edge] : out->fwd_xlat[parts[i].edge]);
                  last = &parts[i];
              }
              if (rp.n) rp.offset = !parts[0].gap ? (int32_t)parts[0].off : (int32_t)parts[1].off - (int32_t)parts[0].len; }
            /* :922-923 quality-aware extension */
            pq_decode(in->quals + in->qual_off[r], qbuf);
            while (extend_left(&rp, &h, el.e, rbuf, qbuf, rlen)) {}
            while (extend_right(&rp, &h, el.e, rbuf, qbuf, rlen)) {}
            if (p->apply_fixpaths)
                for (size_t i = 0; i + 1 < rp.n; ++i) if (h.right[rp.e[i]] != h.left[rp.e[i + 1]]) { rp.n = i + 1; break; }
            out->path_offset[r] = rp.offset;
            out->path_off[r] = out->n_path_edges;
            if (out->n_path_edges + rp.n > pe_cap) { pe_cap = (out->n_path_edges + rp.n) * 2; out->path_edges = (int32_t*)xrealloc(out->path_edges, sizeof(int32_t) * pe_cap); }
            memcpy(out->path_edges + out->n_path_edges, rp.e, rp.n * sizeof(int32_t));
            out->n_path_edges += rp.n;
            if (rp.n > 0) out->n_pathed++;
            if (rp.n > 2) out->n_multipathed++;
        }
        out->path_off[n_reads] = out->n_path_edges;
        free(parts); free(np); free(rp.e);
    }

    /* ---- step-3 input: RepathInMemory's `places` (paths/long/large/Repath.cc:46-72).  Per path x: nkmers = sum of
     * edges[x[j]].size() - (K-1) (:57-59); dropped if nkmers + (K-1) < K2 (:60); y = reversed x with every edge replaced by its
     * involution (:61-62); the smaller of x and y is the place (:63); then sort + unique (:69-71). */
    if (p->places_K2 && p->want_paths) {
        place_t* pl = (place_t*)xmalloc(sizeof(place_t) * (n_reads + 1));
        int32_t* store = (int32_t*)xmalloc(sizeof(int32_t) * (out->n_path_edges + 1));
        size_t npl = 0, used = 0;
        for (uint64_t r = 0; r < n_reads; ++r) {
            const int32_t* x = out->path_edges + out->path_off[r];
            const size_t n = (size_t)(out->path_off[r + 1] - out->path_off[r]);
            int nkmers = 0;
            for (size_t j = 0; j < n; ++j) nkmers += (int)h.elen[x[j]] - (KK - 1);
            if (nkmers + (KK - 1) < (int)p->places_K2) continue;
            int32_t* dst = store + used;
            int flip = 0;                                         /* is y < x ? */
            for (size_t j = 0; j < n; ++j) { int32_t yj = out->involution[x[n - 1 - j]]; if (yj != x[j]) { flip = yj < x[j]; break; } }
            for (size_t j = 0; j < n; ++j) dst[j] = flip ? out->involution[x[n - 1 - j]] : x[j];
            pl[npl].p = dst; pl[npl].n = n; ++npl; used += n;
        }
        out->n_places_kept = npl;
        qsort(pl, npl, sizeof(place_t), place_cmp);
        size_t nu = 0, ne = 0;
        for (size_t i = 0; i < npl; ++i) if (i == 0 || place_cmp(&pl[i - 1], &pl[i]) != 0) { pl[nu++] = pl[i]; ne += pl[i].n; }
        out->n_places = nu; out->n_place_edges = ne;
        out->place_off = (uint64_t*)xmalloc(sizeof(uint64_t) * (nu + 1));
        out->place_edges = (int32_t*)xmalloc(sizeof(int32_t) * (ne + 1));
        size_t at = 0;
        for (size_t i = 0; i < nu; ++i) { out->place_off[i] = at; memcpy(out->place_edges + at, pl[i].p, pl[i].n * sizeof(int32_t)); at += pl[i].n; }
        out->place_off[nu] = at;
        free(pl); free(store);
    }

    for (size_t e = 0; e < el.n; ++e) { free(el.e[e].seq); free(el.e[e].ents); }
    free(el.e); free(epal);
    free(h.left); free(h.right); free(h.elen); free(h.canon); free(h.is_rc);
    free(h.from_off); free(h.from_v); free(h.from_e); free(h.to_off); free(h.to_v); free(h.to_e);
    free(d.key); free(d.ctx); free(d.edge); free(d.off);
    free(qbuf); free(rbuf);
    return W2RAP_OK;
}

void oracle_step2_free(w2rap_graph* g) {
    if (!g) return;
    free(g->edge_off); free(g->edge_len); free(g->edge_bases); free(g->edge_vertices); free(g->fwd_xlat); free(g->rev_xlat); free(g->involution);
    free(g->path_offset); free(g->path_off); free(g->path_edges); free(g->dump); free(g->place_off); free(g->place_edges);
    memset(g, 0, sizeof(*g));
}

/* ---- stage-level entry points used by kernel parity tests */

/* Decodes one PQVec stream; returns the number of quals (feudal/PQVec.cc:122-188). */
size_t oracle_pq_decode(const uint8_t* stream, uint8_t* out) { return pq_decode(stream, out); }
unsigned oracle_good_length(const uint8_t* q, size_t n, unsigned min_qual) { return good_length(q, n, min_qual); }
