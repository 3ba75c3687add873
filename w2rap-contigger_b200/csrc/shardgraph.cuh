// shardgraph.cuh — the graph stage (adjacency, unipaths, edges) over a dictionary SHARDED across GPUs.
//
// The reference builds one dictionary and walks it (kmers/ReadPather.h:307-346, paths/long/BuildReadQGraph.cc:99-339).  Here the
// solid k-mers stay on the rank that counted them (owner = rank of the k-mer's minimiser partition, as in the counting stage), and
// the stage needs two exchange rounds besides the routing of the counting stage:
//   1. NEIGHBOUR QUERIES.  A k-mer and its <= 8 neighbours share 59 bases, hence mostly the minimiser: ~96 % of the neighbour
//      lookups are local.  The rest are sent to their owners (one all-to-all of 16-byte keys, one of 4-byte replies); neighbours
//      that exist come back as GHOST entries of the local table (key + the slot on the owner), so that adjacency pruning and the
//      successor links run unchanged on local memory.  A second, smaller round fetches the ghosts' pruned contexts.
//   2. CHAIN ENDS.  Local chains (pieces) are ranked with the single-GPU list ranking, a chain ending where the successor is a
//      ghost.  Only one record per piece — head, successor across the cut, length, end k-mers — is all-gathered; every rank ranks
//      the (small) piece list redundantly and derives, for its own k-mers, the edge and the offset in it.
// Edge ids come from a global sort of the head k-mers of the kept strands, edge bases from an all-reduce of disjoint bits.
// Everything per item is host/device (unit-tested on the host with W simulated ranks: tests/hostcheck).
#pragma once
#include "extract.cuh"
#include "shard.cuh"
#include "unipath.cuh"

namespace w2r {

W2R_HD uint32_t kmer_owner(Kmer k, uint32_t logP, uint32_t world) {
    return owner_of_partition(mini_part(mini_mix(kmer_minimizer_hash_words(k)), logP), logP, world);
}

// ---- round 1: which neighbours have to be asked for.  emit(owner rank, canonical neighbour k-mer) for every context bit of an
// owned k-mer whose neighbour lives on another rank (kmers/ReadPather.h:322-342 looks every one of them up).
// The neighbour shares 45 of its 46 m-mers with k: its minimiser is the smaller of (k's minimum without the m-mer that drops out) and
// the one new m-mer, so a k-mer costs 46 + (number of neighbours) m-mer hashes instead of 46 per neighbour.
W2R_HD uint32_t mmer_of_words(uint64_t lo, uint64_t hi, int t) {          // m-mer t of a k-mer in LSB-first words (base i in bits 2i.. of hi:lo)
    const int bit = 2 * t;
    const uint64_t w = bit == 0 ? lo : (bit < 64 ? (lo >> bit) | (hi << (64 - bit)) : hi >> (bit - 64));
    return (uint32_t)w & ((1u << (2 * MINI_M)) - 1u);
}
struct NeighbourScan {
    Kmer k; uint32_t c, logP, world, me, min_wo_first, min_wo_last;
    W2R_HD NeighbourScan(Kmer k_, uint32_t c_, uint32_t logP_, uint32_t world_, uint32_t me_) : k(k_), c(c_ & 0xffu), logP(logP_), world(world_), me(me_),
                                                                                                  min_wo_first(0xffffffffu), min_wo_last(0xffffffffu) {
        if (!c) return;
        const uint64_t lo = rev2(k.w0), hi = rev2(k.w1);
        for (int t = 0; t < MINI_W; ++t) {
            const uint32_t h = mmer_hash_of(mmer_of_words(lo, hi, t));
            if (t > 0 && h < min_wo_first) min_wo_first = h;
            if (t < MINI_W - 1 && h < min_wo_last) min_wo_last = h;
        }
    }
    // context bit b (0-3 successors, 4-7 predecessors): does its neighbour live on another rank?  Then *owner and its canonical form *cn.
    W2R_HD bool remote(uint32_t b, uint32_t* owner, Kmer* cn) const {
        if (!(c & (1u << b))) return false;
        const bool succ = b < 4;
        const Kmer n = succ ? kmer_succ(k, b) : kmer_pred(k, b - 4);
        const uint32_t hn = mmer_hash_of(mmer_of_words(rev2(n.w0), rev2(n.w1), succ ? MINI_W - 1 : 0));
        const uint32_t rest = succ ? min_wo_first : min_wo_last;
        const uint32_t o = owner_of_partition(mini_part(mini_mix(hn < rest ? hn : rest), logP), logP, world);
        if (o == me) return false;
        const Kmer r = kmer_rc(n);
        *owner = o; *cn = kmer_less(r, n) ? r : n;
        return true;
    }
};
template <class Emit>
W2R_HD void neighbour_queries(Kmer k, uint32_t c, uint32_t logP, uint32_t world, uint32_t me, Emit& emit) {
    const NeighbourScan ns(k, c, logP, world, me);
    for (uint32_t b = 0; b < 8; ++b) { uint32_t o; Kmer cn; if (ns.remote(b, &o, &cn)) emit(o, cn); }
}

// ---- round 2: pieces (local chains) and their ranking
constexpr uint64_t GID_NONE = ~0ull;
W2R_HD uint64_t gid_make(uint32_t rank, uint32_t node) { return ((uint64_t)rank << 32) | node; }
struct PieceRec {
    uint64_t head;          // global id (rank << 32 | oriented node) of the first node
    uint64_t succ;          // global id of the node that follows the last node on another rank, or GID_NONE
    uint32_t n;             // k-mers in the piece
    uint32_t flip_local;    // index, among the owner rank's pieces, of the piece made of the flipped nodes (its head is the flip of our tail)
    Kmer head_k, tail_k;    // oriented k-mers of the first and the last node
};
// Is owned node x the head of a local piece?  (no predecessor at all, or a predecessor on another rank)
W2R_HD bool node_is_piece_head(const uint32_t* next0, const uint8_t* ghead, uint32_t x) {
    const uint32_t nx = next0[x];
    if (nx == EMPTY_NODE || nx == GHOST_TAIL) return false;
    return next0[x ^ 1u] == NIL || ghead[x] != 0;
}
// R[x] = (local tail, distance to it | RESOLVED) from the local list ranking
W2R_HD PieceRec piece_of_head(const SolidTable& t, const uint32_t* next0, const RankState* R, uint32_t me, uint32_t h) {
    PieceRec p;
    const uint32_t tail = R[h].x;
    p.head = gid_make(me, h);
    p.n = (R[h].y & ~RANK_RESOLVED) + 1u;
    p.flip_local = tail;                     // the TAIL NODE for now: k_piece_flips turns it into the flipped piece's index
    const uint32_t nt = next0[tail];
    if (nt == NIL) p.succ = GID_NONE;
    else { const SolidSlot& g = t.slots[nt >> 1]; p.succ = gid_make(g.pad - 1u, 2u * g.edge + (nt & 1u)); }     // the ghost knows its owner and its slot there
    p.head_k = node_kmer(t, h);
    p.tail_k = node_kmer(t, tail);
    return p;
}

// Replicated per-piece tables after the gather.  S = ranking state: x = the last piece of the chain, y = k-mers after this piece | RESOLVED.
struct PieceView {
    const PieceRec* rec;
    const uint32_t* flip;       // index of the piece made of the flipped nodes
    const RankState* S;
    uint64_t n;
};
constexpr uint32_t PIECE_DIST_MAX = 0x3fffffffu;        // distances saturate (an edge may not exceed 2^24 k-mers anyway)
W2R_HD RankState piece_rank_init(const PieceRec* rec, const uint32_t* nxt, uint32_t i) {
    return nxt[i] == NIL ? RankState{i, RANK_RESOLVED} : RankState{nxt[i], rec[nxt[i]].n};
}
W2R_HD RankState piece_rank_step(RankState a, RankState b) {     // a: state of i (unresolved), b: state of a.x
    uint64_t d = (uint64_t)(a.y & ~RANK_RESOLVED) + (b.y & ~RANK_RESOLVED);
    if (d > PIECE_DIST_MAX) d = PIECE_DIST_MAX;
    return RankState{b.x, (uint32_t)d | (b.y & RANK_RESOLVED)};
}
// For a chain's LAST piece i: the first piece of the chain and the chain's length in k-mers (through the flipped chain, whose first
// piece is flip[i] and whose last piece is the flip of our first piece).
W2R_HD void chain_of_tail_piece(const PieceView& pv, uint32_t i, uint32_t* head_piece, uint64_t* n_kmers) {
    const uint32_t fi = pv.flip[i];
    *head_piece = pv.flip[pv.S[fi].x];
    *n_kmers = (uint64_t)pv.rec[fi].n + (pv.S[fi].y & ~RANK_RESOLVED);
}
// Even-length edges are oriented by comparing the sequence with its reverse complement outside-in (dna/CanonicalForm.h:34-46): the
// first 60 bases decide, i.e. the head k-mer against the reverse complement of the tail k-mer; both are in the piece records.
// Returns 0/1 = keep flag for an even-length chain, 2 = odd length (decided by the rank that owns the middle k-mer).
W2R_HD uint32_t chain_keep_even(const PieceView& pv, uint32_t tail_piece, uint32_t head_piece, uint64_t n_kmers) {
    if ((n_kmers + K - 1) & 1ull) return 2u;
    const Kmer hk = pv.rec[head_piece].head_k, ok = kmer_rc(pv.rec[tail_piece].tail_k);
    return (kmer_less(hk, ok) || (hk == ok && (pv.rec[head_piece].head & 1ull) == 0)) ? 1u : 0u;
}

// Per piece, once the pieces are ranked: where the piece sits in its chain.
struct PieceInfo {
    uint32_t head_piece;    // first piece of the chain; later replaced by the chain's edge id (NIL: the strand is not kept)
    uint32_t off;           // k-mers of the chain before this piece
    uint32_t after;         // k-mers of the chain after this piece
    uint32_t n;             // k-mers in this piece
};
W2R_HD PieceInfo piece_info(const PieceView& pv, uint32_t i) {
    const uint32_t fi = pv.flip[i];
    return PieceInfo{pv.flip[pv.S[fi].x], pv.S[fi].y & ~RANK_RESOLVED, pv.S[i].y & ~RANK_RESOLVED, pv.rec[i].n};
}
// Odd-length edges are oriented by their middle base alone (dna/CanonicalForm.h:34-46).  The piece that holds the middle k-mer
// walks to it from its head.  Returns -1 if the chain has even length or its middle lies in another piece, else the keep flag.
W2R_HD int piece_keep_odd(const SolidTable& t, const uint32_t* next0, const PieceInfo& pi, uint32_t head_node) {
    const uint64_t n = (uint64_t)pi.off + pi.n + pi.after, L = n + K - 1;
    if (!(L & 1ull)) return -1;
    const uint64_t m = L / 2, ostar = m < n - 1 ? m : n - 1;
    if (ostar < pi.off || ostar >= (uint64_t)pi.off + pi.n) return -1;
    uint32_t x = head_node;
    for (uint64_t s = pi.off; s < ostar; ++s) x = next0[x];
    return (kmer_base(node_kmer(t, x), (int)(m - ostar)) & 2u) ? 0 : 1;
}
// Edge emission + KDef back-fill for one owned node (BuildReadQGraph.cc:287-301).  R[x] = (local tail, distance to it);
// lpiece[tail] = local index of its piece; pinfo[] = this rank's pieces with head_piece already replaced by the edge id.
// put(edge byte offset, base position, base code).
template <class Put>
W2R_HD void emit_node_sharded(const SolidTable& t, const RankState* R, const uint32_t* lpiece, const PieceInfo* pinfo, const uint64_t* edge_off, uint32_t x, Put& put) {
    const RankState a = R[x];
    const PieceInfo pi = pinfo[lpiece[a.x]];
    if (pi.head_piece == NIL) return;
    const uint32_t d = a.y & ~RANK_RESOLVED;
    const uint64_t off = (uint64_t)pi.off + (pi.n - 1u - d);
    SolidSlot* s = t.slots + (x >> 1);
    s->edge = pi.head_piece; s->off = (uint32_t)off;
    const Kmer k = node_kmer(t, x);
    const uint64_t bo = edge_off[pi.head_piece];
    put(bo, off, kmer_first(k));
    if (d == 0 && pi.after == 0)
        for (int i = 1; i < K; ++i) put(bo, off + i, kmer_base(k, i));
}

// 64-bit key -> u32 value open-addressing map (global piece ids -> piece index).  Keys are unique; EMPTY = ~0.
struct GidMap { uint64_t* keys; uint32_t* vals; uint64_t mask; };
W2R_HD uint64_t gid_hash(uint64_t g) { g ^= g >> 31; g *= 0x9e3779b97f4a7c15ull; g ^= g >> 29; g *= 0xbf58476d1ce4e5b9ull; g ^= g >> 32; return g; }
W2R_HD uint32_t gid_find(const GidMap& m, uint64_t g) {
    for (uint64_t h = gid_hash(g) & m.mask;; h = (h + 1) & m.mask) {
        const uint64_t k = m.keys[h];
        if (k == g) return m.vals[h];
        if (k == GID_NONE) return NIL;
    }
}

}  // namespace w2r
