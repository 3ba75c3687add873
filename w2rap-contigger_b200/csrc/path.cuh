// path.cuh — read pathing for one read (path_reads_OMP body, paths/long/BuildReadQGraph.cc:829-929) as a
// host/device function over device-format graph arrays.  One thread paths one read.
//
// The reference materialises a vector<PathPart> and edits it; here the same rules are evaluated in streaming form with
// O(1) state (at most the two most recent seeds can still be dropped by the captured-gap rule and the short-tail rule).
// Every rule cites the reference lines it restates; quirks are kept.
#pragma once
#include "extract.cuh"
#include "kmer.cuh"
#include "pqvec.cuh"

namespace w2r {

struct GraphView {
    PathDict dict;                 // the finished dictionary (k-mer -> edge, offset), sliced over the GPUs, + its negative-lookup filter
    const uint8_t* edge_bases;     // canonical edges, bvec packing, byte aligned per edge
    const uint64_t* edge_off;      // byte offsets
    const uint32_t* edge_len;      // bases
    const int32_t* fwd_xlat;       // canonical edge -> hbv edge id (paths/long/HBVFromEdges.cc:137-151)
    const int32_t* rev_xlat;
    const uint32_t* hcanon;        // hbv edge -> (canonical edge << 1) | is_rc
    const int32_t* hleft;          // hbv edge -> source vertex
    const int32_t* hright;         // hbv edge -> target vertex
    const int32_t* from_e;         // [4*v + i] FromEdgeObj(v) in the reference's list order (graph/DigraphTemplate.h:1829-1839)
    const int32_t* to_e;           // [4*v + i] ToEdgeObj(v)
    const uint8_t* from_n;         // FromSize(v)  (<= 4: the edges leaving a (K-1)-mer start with distinct k-mers)
    const uint8_t* to_n;           // ToSize(v)
};

// staging row of `cap` ints: [0,left_cap) left extensions (stored backwards), [left_cap,cap) seeds + right extensions

struct PathPart { uint32_t edge, rc, off, len, elen; bool after_gap; };

W2R_HD uint32_t edge_base_oriented(const GraphView& g, uint32_t canon, uint32_t rc, uint64_t pos) {
    const uint8_t* p = g.edge_bases + g.edge_off[canon];
    return rc ? 3u - packed_base(p, (uint64_t)g.edge_len[canon] - 1 - pos) : packed_base(p, pos);
}
W2R_HD uint32_t hbv_edge_base(const GraphView& g, int32_t he, uint64_t pos) { uint32_t hc = g.hcanon[he]; return edge_base_oriented(g, hc >> 1, hc & 1u, pos); }
W2R_HD uint32_t hbv_edge_len(const GraphView& g, int32_t he) { return g.edge_len[g.hcanon[he] >> 1]; }
W2R_HD bool part_same_edge(const PathPart& a, const PathPart& b) { return a.edge == b.edge && a.rc == b.rc; }

// n <= 32 bases of an ORIENTED edge starting at oriented position pos, LSB-first (base pos in bits 1:0), bits above 2n clear.
// rc: oriented positions pos..pos+n-1 are canonical positions len-pos-n..len-pos-1, reversed and complemented.
W2R_HD uint64_t oriented_bases(const uint8_t* ep, uint64_t len, uint32_t rc, uint64_t pos, uint32_t n) {
    uint64_t w = rc ? rev2(~bases32_at(ep, len - pos - n)) >> (2u * (32u - n)) : bases32_at(ep, pos);
    if (n < 32u) w &= (1ull << (2u * n)) - 1ull;
    return w;
}
W2R_HD uint64_t read_bases(const uint8_t* bases, uint64_t pos, uint32_t n) {
    uint64_t w = bases32_at(bases, pos);
    if (n < 32u) w &= (1ull << (2u * n)) - 1ull;
    return w;
}

// BuildReadQGraph.cc:552-558 isJoinable: same edge id, or equal LAST (K-1)-mers of the two oriented edges (32 + 27 bases: two words).
W2R_HD bool part_joinable(const GraphView& g, const PathPart& a, const PathPart& b) {
    if (a.edge == b.edge) return true;
    const uint64_t la = g.edge_len[a.edge], lb = g.edge_len[b.edge];
    const uint8_t* pa = g.edge_bases + g.edge_off[a.edge];
    const uint8_t* pb = g.edge_bases + g.edge_off[b.edge];
    if (oriented_bases(pa, la, a.rc, la - (K - 1), 32) != oriented_bases(pb, lb, b.rc, lb - (K - 1), 32)) return false;
    return oriented_bases(pa, la, a.rc, la - (K - 1) + 32, K - 1 - 32) == oriented_bases(pb, lb, b.rc, lb - (K - 1) + 32, K - 1 - 32);
}

// paths/long/ExtendReadPath.cc:15-109.  penalty -= 0.2*penalty (unsigned -= double) == 4*penalty/5 in integers.
// Read and edge are compared 32 bases per step (XOR of packed words); only mismatches cost a quality lookup, and a run of matches
// decays the penalty until it reaches zero.
struct OverlapScore {
    uint32_t qsum = 0, pen = 0;
    W2R_HD void matches(uint32_t m) { while (m-- && pen) pen = 4u * pen / 5u; }
    W2R_HD void mismatch(uint32_t q) { pen += q == 2u ? 20u : q; qsum += pen; }
};
W2R_HD uint32_t score_left_overlap(const GraphView& g, const uint8_t* bases, const uint8_t* qstream, uint32_t start, int32_t he) {
    const uint32_t hc = g.hcanon[he], canon = hc >> 1, rc = hc & 1u;
    const uint64_t elen = g.edge_len[canon];
    const uint8_t* ep = g.edge_bases + g.edge_off[canon];
    OverlapScore sc;
    int64_t b = (int64_t)start - 1, e = (int64_t)elen - K;        // read index downwards from start-1 against edge index downwards from |e|-K
    while (b >= 0 && e >= 0) {
        const uint32_t n = (uint32_t)((b < e ? b : e) + 1 < 32 ? (b < e ? b : e) + 1 : 32);
        // positions b-n+1..b of the read against e-n+1..e of the edge; reversed so that the walk order (downwards) is bit order
        uint64_t x = read_bases(bases, (uint64_t)(b - n + 1), n) ^ oriented_bases(ep, elen, rc, (uint64_t)(e - n + 1), n);
        x = rev2(x) >> (2u * (32u - n));
        uint32_t done = 0;
        while (x) {
            const uint32_t j = (uint32_t)ctz64(x) >> 1;
            sc.matches(j - done);
            sc.mismatch(pq_qual_at(qstream, (uint32_t)(b - j)));
            done = j + 1;
            x &= ~(3ull << (2u * j));
        }
        sc.matches(n - done);
        b -= n; e -= n;
    }
    if (b >= 0) sc.qsum += 10u * (uint32_t)(b + 1);
    return sc.qsum;
}
W2R_HD uint32_t score_right_overlap(const GraphView& g, const uint8_t* bases, const uint8_t* qstream, uint32_t rlen, uint32_t start, int32_t he) {
    const uint32_t hc = g.hcanon[he], canon = hc >> 1, rc = hc & 1u;
    const uint64_t elen = g.edge_len[canon];
    const uint8_t* ep = g.edge_bases + g.edge_off[canon];
    OverlapScore sc;
    uint64_t b = rlen - start, e = K - 1;
    while (b < rlen && e < elen) {
        uint64_t m = rlen - b < elen - e ? rlen - b : elen - e;
        const uint32_t n = m < 32 ? (uint32_t)m : 32u;
        uint64_t x = read_bases(bases, b, n) ^ oriented_bases(ep, elen, rc, e, n);
        uint32_t done = 0;
        while (x) {
            const uint32_t j = (uint32_t)ctz64(x) >> 1;
            sc.matches(j - done);
            sc.mismatch(pq_qual_at(qstream, (uint32_t)(b + j)));
            done = j + 1;
            x &= ~(3ull << (2u * j));
        }
        sc.matches(n - done);
        b += n; e += n;
    }
    if (b < rlen) sc.qsum += 10u * (uint32_t)(rlen - b);
    return sc.qsum;
}

// ExtendReadPath.cc:150-212 (left) / :262-318 (right): classify candidates, then score.  Returns the chosen hbv edge or -1.
W2R_HD int32_t choose_extension(const GraphView& g, int32_t v, bool leftward, uint32_t last_gap, const uint8_t* bases, const uint8_t* quals /* PQVec stream */, uint32_t rlen) {
    const int32_t* cand = (leftward ? g.to_e : g.from_e) + 4 * (int64_t)v;
    const uint32_t nc = leftward ? g.to_n[v] : g.from_n[v];
    const bool solo = nc == 1;
    uint32_t hanging = 0, nlong = 0, nshort = 0;
    int32_t short_v = -1;
    bool short_multi = false;
    for (uint32_t i = 0; i < nc; ++i) {
        int32_t e = cand[i];
        int32_t d = leftward ? g.hleft[e] : g.hright[e];
        bool hang = leftward ? (g.to_n[d] == 0 && g.from_n[d] == 1) : (g.from_n[d] == 0 && g.to_n[d] == 1);
        if (hang) hanging |= 1u << i;
        bool is_long = hbv_edge_len(g, e) - (uint32_t)(K - 1) >= last_gap;
        if (is_long) ++nlong;
        if (!is_long && !hang) { if (!nshort) short_v = d; else if (d != short_v) short_multi = true; ++nshort; }
    }
    if (!solo && nshort > 0) {
        if (nlong > 0 || short_multi) return -1;
        if ((leftward ? g.to_n[short_v] : g.from_n[short_v]) != 1) return -1;
    }
    int32_t least_edge = -1;
    uint32_t least = 0xffffffffu;
    for (uint32_t i = 0; i < nc; ++i) {
        if (!((hanging >> i) & 1u) || solo) {
            uint32_t s = leftward ? score_left_overlap(g, bases, quals, last_gap, cand[i]) : score_right_overlap(g, bases, quals, rlen, last_gap, cand[i]);
            if (s < least) { least = s; least_edge = cand[i]; }
        }
    }
    if (least_edge == -1 || least > last_gap * 10u) return -1;
    return least_edge;
}

struct PathResult { int32_t offset; uint32_t start, len; bool overflow; };

// Pathing of one read as a resumable walker.  scan() runs the seed loop of BRQ_Pather::path (:500-550) until the read is
// exhausted (returns false) or a k-mer misses the dictionary (returns true).  The caller then finds the first position
// p in (itr, nk) whose k-mer IS in the dictionary — serially (path_one_read below, the host check) or with the whole warp
// (k_path_reads: 32 positions per step) — and hands it back with gap_found(p, slot) (p = nk, slot = -1 if there is none).
// finish() applies the tail rules, the quality-aware extensions and FixPaths.
// The walker's state, kept apart from the handle so that a kernel can place it in SHARED memory (one slot per thread): with the
// state in registers the path kernel spilled ~350 bytes per thread, and that local-memory traffic was the largest single source
// of L2/DRAM sectors of the whole kernel (ncu: profiles/r2_ncu_path_reads_c2.txt).
struct PathState {
    const uint8_t* bases;
    uint32_t rlen, nk;
    PathPart pend[2];
    int npend;
    bool last_kept_valid; uint32_t lk_edge, lk_rc;
    uint32_t n_ids, sum_kmers, seeds, nparts;
    bool first_is_gap, have_first_hit;
    uint32_t gap0_len, first_hit_off;
    bool last_is_gap, last2_is_gap;
    bool pending_gap; uint32_t gap_len, gap_index;
    uint32_t itr;
    bool overflow, scan_done;
    const SolidSlot* found_slot;    // dictionary entry of the k-mer at itr, found by the gap screening (nullptr: not known)
};
struct PathWalker {
    PathState& s;
    const GraphView* g;
    int32_t* row;
    uint32_t left_cap, right_cap;
    W2R_HD explicit PathWalker(PathState& st) : s(st), g(nullptr), row(nullptr), left_cap(0), right_cap(0) {}

    W2R_HD void init(const GraphView& g_, const uint8_t* bases_, uint32_t rlen_, int32_t* row_, uint32_t cap, uint32_t left_cap_) {
        g = &g_; s.bases = bases_; row = row_; s.rlen = rlen_; left_cap = left_cap_; right_cap = cap - left_cap_;
        s.nk = s.rlen >= (uint32_t)K ? s.rlen - K + 1 : 0;           // :503-506 a read shorter than K is a single gap part: empty path
        s.npend = 0; s.last_kept_valid = false; s.lk_edge = s.lk_rc = 0;
        s.n_ids = s.sum_kmers = s.seeds = s.nparts = 0;
        s.first_is_gap = s.have_first_hit = false; s.gap0_len = s.first_hit_off = 0;
        s.last_is_gap = s.last2_is_gap = false; s.pending_gap = false; s.gap_len = s.gap_index = 0;
        s.itr = 0; s.overflow = false; s.scan_done = false; s.found_slot = nullptr;
    }
    W2R_HD void commit(const PathPart& h) {                     // :804-815 pathPartsToReadPath, one kept seed
        if (s.last_kept_valid && s.lk_edge == h.edge && s.lk_rc == h.rc) return;
        if (s.n_ids < right_cap) row[left_cap + s.n_ids] = h.rc ? g->rev_xlat[h.edge] : g->fwd_xlat[h.edge]; else s.overflow = true;
        ++s.n_ids; s.sum_kmers += h.elen;
        s.last_kept_valid = true; s.lk_edge = h.edge; s.lk_rc = h.rc;
    }
    // the k-mer at s.itr is dictionary entry `slot`: extend the match along its edge; false = the path ends here (captured-gap rule)
    W2R_HD bool seed(const SolidSlot* slot) {
        const SolidSlot ss = *slot;
        PathPart h;
        h.edge = ss.edge;
        const uint32_t elen = g->edge_len[h.edge];
        uint32_t o = ss.off;
        const uint8_t* ep = g->edge_bases + g->edge_off[h.edge];
        // dna/CanonicalForm.h:85-92 isRC: the read k-mer is the reverse complement of the edge k-mer at that offset
        Kmer ek = kmer_at(ep, o);
        h.rc = !(ek == kmer_at(s.bases, s.itr));
        h.elen = elen - K + 1;
        uint32_t len = 1;
        // matchLen (:341-350), 32 s.bases per step: XOR of two packed words, first differing 2-bit group by count-trailing-zeros
        if (!h.rc) {
            uint64_t rp = (uint64_t)s.itr + K, e2 = (uint64_t)o + K;
            while (rp < s.rlen && e2 < elen) {
                uint64_t n = s.rlen - rp < elen - e2 ? s.rlen - rp : elen - e2;
                if (n > 32) n = 32;
                uint64_t x = bases32_at(s.bases, rp) ^ bases32_at(ep, e2);
                if (n < 32) x &= (1ull << (2 * n)) - 1;
                if (x) { len += (uint32_t)(ctz64(x) >> 1); break; }
                len += (uint32_t)n; rp += n; e2 += n;
            }
            h.off = o;
        } else {
            uint64_t ro = (uint64_t)elen - o;               // position in rc(edge) just past the k-mer
            uint64_t rp = (uint64_t)s.itr + K, e2 = ro;
            while (rp < s.rlen && e2 < elen) {
                uint64_t n = s.rlen - rp < elen - e2 ? s.rlen - rp : elen - e2;
                if (n > 32) n = 32;
                // rc(edge)[e2 .. e2+n) = reverse complement of edge[elen-e2-n .. elen-e2)
                uint64_t c = rev2(~bases32_at(ep, (uint64_t)elen - e2 - n)) >> (2 * (32 - n));
                uint64_t x = bases32_at(s.bases, rp) ^ c;
                if (n < 32) x &= (1ull << (2 * n)) - 1;
                if (x) { len += (uint32_t)(ctz64(x) >> 1); break; }
                len += (uint32_t)n; rp += n; e2 += n;
            }
            h.off = (uint32_t)(ro - K);
        }
        h.len = len;
        h.after_gap = s.pending_gap;
        // :875-898 captured-gap consistency, evaluated when the seed after an interior gap arrives
        if (s.pending_gap && s.gap_index >= 1) {
            const PathPart& pv = s.pend[s.npend - 1];
            uint32_t gd = h.off - (pv.off + pv.len);        // :470 unsigned arithmetic
            if (!part_same_edge(pv, h)) gd += pv.elen;
            int32_t diff = (int32_t)(s.gap_len - gd);
            uint32_t ad = (uint32_t)(diff < 0 ? -diff : diff);
            if (!(ad <= 3u) || !part_joinable(*g, pv, h)) {
                if (s.seeds > 1) { s.last2_is_gap = pv.after_gap; --s.npend; }   // drop the seed before the gap and everything after it
                else { s.last2_is_gap = false; }                             // the gap absorbs everything after it
                s.last_is_gap = true;
                return false;
            }
        }
        s.pending_gap = false;
        if (s.npend == 2) { commit(s.pend[0]); s.pend[0] = s.pend[1]; s.npend = 1; }
        s.pend[s.npend++] = h;
        ++s.seeds;
        if (!s.have_first_hit) { s.have_first_hit = true; s.first_hit_off = h.off; }
        ++s.nparts; s.last2_is_gap = s.last_is_gap; s.last_is_gap = false;
        s.itr += len;
        return true;
    }
    // (seed() is large and is inlined exactly once, here: the kernel's instruction footprint matters — divergent warps roam through
    //  it and instruction-fetch stalls were the second largest stall reason; gap_found() hands its slot over through found_slot)
    W2R_HD bool scan() {
        while (!s.scan_done && s.itr < s.nk) {
            const SolidSlot* slot = s.found_slot;
            s.found_slot = nullptr;
            if (!slot) {
                Kmer f = kmer_at(s.bases, s.itr);
                Kmer r = kmer_rc(f);
                slot = pd_find_filtered(g->dict, kmer_less(r, f) ? r : f);
                if (!slot) return true;
            }
            if (!seed(slot)) s.scan_done = true;
        }
        return false;
    }
    W2R_HD void gap_found(uint32_t p, const SolidSlot* slot) {
        const uint32_t gl = p - s.itr;
        s.itr = p;
        if (s.nparts == 0) { s.first_is_gap = true; s.gap0_len = gl; }
        s.gap_len = gl; s.gap_index = s.nparts; ++s.nparts;
        s.last2_is_gap = s.last_is_gap; s.last_is_gap = true;
        s.pending_gap = true;
        if (!slot) s.scan_done = true; else s.found_slot = slot;            // scan(), which every caller runs next, seeds from it
    }
    W2R_HD PathResult finish(const uint8_t* qstream, bool apply_fixpaths) {
        PathResult res{0, left_cap, 0, false};
        // :904-918 a trailing seed that only reached <= 5 k-mers into an edge from its very start is dropped
        if (s.last_is_gap) {
            if (s.nparts > 1 && !s.last2_is_gap && s.npend > 0) { const PathPart& l2 = s.pend[s.npend - 1]; if (l2.off == 0 && l2.len <= 5) --s.npend; }
        } else if (s.npend > 0) {
            const PathPart& l = s.pend[s.npend - 1];
            if (l.off == 0 && l.len <= 5) --s.npend;
        }
        for (int i = 0; i < s.npend; ++i) commit(s.pend[i]);
        res.overflow = s.overflow;
        if (s.n_ids == 0 || res.overflow) return res;
        int32_t offset = s.first_is_gap ? (int32_t)s.first_hit_off - (int32_t)s.gap0_len : (int32_t)s.first_hit_off;   // :816-826

        // :922-923 quality-aware extension (ExtendReadPath.cc:115-348); all left extensions first, then right
        uint32_t nl = 0;
        int32_t front = row[left_cap], back = row[left_cap + s.n_ids - 1];
        for (int dir = 0; dir < 2; ++dir) {                      // (one loop, one call site: choose_extension is large)
            for (;;) {
                const bool leftward = dir == 0;
                const int32_t lastg = leftward ? -offset : (int32_t)s.rlen + offset - (int32_t)s.sum_kmers - (K - 1);
                if (lastg < 10) break;
                // rightward: the reference hands ToLeft as "to_right" (BuildReadQGraph.cc:836-841): candidates leave the LEFT vertex of the last edge
                const int32_t e = choose_extension(*g, g->hleft[leftward ? front : back], leftward, (uint32_t)lastg, s.bases, qstream, s.rlen);
                if (e < 0) break;
                const uint32_t ek = hbv_edge_len(*g, e) - K + 1;
                s.sum_kmers += ek;
                if (leftward) {
                    offset += (int32_t)ek;
                    if (nl < left_cap) row[left_cap - 1 - nl] = e; else res.overflow = true;
                    ++nl; front = e;
                } else {
                    if (s.n_ids < right_cap) row[left_cap + s.n_ids] = e; else res.overflow = true;
                    ++s.n_ids; back = e;
                }
            }
        }
        if (res.overflow) return res;
        res.offset = offset;
        res.start = left_cap - nl;
        res.len = nl + s.n_ids;
        if (apply_fixpaths) {                                    // large/GapToyTools.cc:322-335
            const int32_t* p = row + res.start;
            for (uint32_t i = 0; i + 1 < res.len; ++i)
                if (g->hright[p[i]] != g->hleft[p[i + 1]]) { res.len = i + 1; break; }
        }
        return res;
    }
};

// Paths one read, serially (the host check; the device kernel drives the walker itself and resolves gaps with the whole warp).
// `row` is this read's staging row of `cap` ints; qualities are looked up in the PQVec stream where an extension needs them.
W2R_HD PathResult path_one_read(const GraphView& g, const uint8_t* bases, uint32_t rlen, const uint8_t* qstream,
                                int32_t* row, uint32_t cap, uint32_t left_cap, bool apply_fixpaths) {
    PathState st;
    PathWalker w(st);
    w.init(g, bases, rlen, row, cap, left_cap);
    while (w.scan()) {
        uint32_t p = st.itr + 1;
        const SolidSlot* slot = nullptr;
        if (p < st.nk) {
            Kmer f = kmer_at(bases, p), r = kmer_rc(f);
            uint64_t nxt = 0;
            for (uint32_t t = 0;; ++t) {
                slot = pd_find_filtered(g.dict, kmer_less(r, f) ? r : f);
                if (slot || p + 1 >= st.nk) { if (!slot) ++p; break; }
                if ((t & 31u) == 0) nxt = bases32_at(bases, (uint64_t)p + K);
                const uint32_t nb = (uint32_t)nxt & 3u;
                nxt >>= 2;
                f = kmer_succ(f, nb); r = kmer_pred(r, 3u - nb);
                ++p;
            }
        }
        w.gap_found(p, slot);
    }
    return w.finish(qstream, apply_fixpaths);
}

}  // namespace w2r
