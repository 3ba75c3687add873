// hostcheck.cpp — TEST INFRASTRUCTURE (never linked into the product library).
//
// The per-item logic of the CUDA kernels lives in host/device headers (w2rap-contigger_b200/csrc/{kmer,pqvec,extract,
// unipath,path}.cuh).  This file compiles those headers with g++ and drives them serially, in the same order the kernels
// run, so that their logic can be checked against the oracle on a machine without a GPU.  It is a unit test of device
// functions, not a CPU implementation of the product: the product library contains none of this and refuses to run
// without a B200.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/w2rap_step2.h"
#include "../../w2rap-contigger_b200/csrc/extract.cuh"
#define W2R_COUNT_PART_HOST_ONLY
#include "../../w2rap-contigger_b200/csrc/shard.cuh"
#include "../../w2rap-contigger_b200/csrc/kmer.cuh"
#include "../../w2rap-contigger_b200/csrc/path.cuh"
#include "../../w2rap-contigger_b200/csrc/pqvec.cuh"
#include "../../w2rap-contigger_b200/csrc/unipath.cuh"

using namespace w2r;

namespace {
struct Rec { uint64_t w0, w1; uint32_t ctx; };
struct Collect {
    std::vector<Rec>* v;
    void operator()(Kmer k, uint32_t ctx) const { v->push_back(Rec{k.w0, k.w1, ctx}); }
};
template <class T> T* dup(const std::vector<T>& v) { T* p = (T*)malloc((v.size() ? v.size() : 1) * sizeof(T)); if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T)); return p; }
}  // namespace

extern "C" {

// k_good_len + k_extract_count + k_count_stats, serially: distinct k-mers with saturated counts and OR-ed contexts.
int hc_count(const w2rap_reads* in, uint32_t min_qual, w2rap_kmer_rec** out, uint64_t* n_out, uint64_t* n_inst) {
    std::vector<Rec> recs;
    Collect emit{&recs};
    for (uint64_t r = 0; r < in->n_reads; ++r) {
        uint32_t nq = 0;
        uint32_t gl = pq_good_length(in->quals + in->qual_off[r], min_qual, &nq);
        if (nq != in->len[r]) return 100;
        if (gl > in->len[r]) gl = in->len[r];
        extract_read_kmers(in->bases + in->base_off[r], gl, emit);
    }
    *n_inst = recs.size();
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.w0 < b.w0 || (a.w0 == b.w0 && a.w1 < b.w1); });
    std::vector<w2rap_kmer_rec> d;
    for (size_t i = 0; i < recs.size();) {
        size_t j = i; uint32_t c = 0, ctx = 0;
        while (j < recs.size() && recs[j].w0 == recs[i].w0 && recs[j].w1 == recs[i].w1) { ctx |= recs[j].ctx; ++c; ++j; }
        d.push_back(w2rap_kmer_rec{recs[i].w0, recs[i].w1, c > 255 ? 255 : c, ctx, 0xffffffffu, 0});
        i = j;
    }
    *n_out = d.size();
    *out = dup(d);
    return 0;
}

void hc_free(void* p) { free(p); }

// ---- pieces of the sharded (multi-GPU) protocol, for the world_size-2 gloo test: the "map" side (k_minimizer_map) and the
// "reduce" side (k_count_smem) as separate calls, with the product's own partition/owner functions.
// recs: n x {w0, w1|ctx}; owner[i] = rank that owns record i's partition (keyed by the k-mer's minimiser, extract.cuh).
int hc_extract_records(const w2rap_reads* in, uint32_t min_qual, uint32_t logP, uint32_t world, uint64_t** recs_out, uint32_t** owner_out, uint64_t* n_out) {
    std::vector<uint64_t> flat;
    std::vector<uint32_t> owner;
    for (uint64_t r = 0; r < in->n_reads; ++r) {
        uint32_t nq = 0;
        uint32_t gl = pq_good_length(in->quals + in->qual_off[r], min_qual, &nq);
        if (nq != in->len[r]) return 100;
        if (gl > in->len[r]) gl = in->len[r];
        const uint8_t* bases = in->bases + in->base_off[r];
        KmerCursor cur;
        cur.open(bases, gl);
        Kmer k; uint32_t ctx;
        for (uint64_t j = 0; cur.next(&k, &ctx); ++j) {
            flat.push_back(k.w0); flat.push_back(k.w1 | ctx);
            owner.push_back(owner_of_partition(mini_part(mini_mix(kmer_minimizer_hash(bases, j)), logP), logP, world));
        }
    }
    *n_out = owner.size(); *recs_out = dup(flat); *owner_out = dup(owner);
    return 0;
}
// Strand symmetry of the minimiser partition key (extract.cuh): the k-mer at position j of a read and the k-mer at position
// len-K-j of its reverse complement are the same canonical k-mer and must get the same key.  Returns the number of mismatches.
uint64_t hc_minimizer_symmetry(const w2rap_reads* in, uint64_t* n_checked) {
    uint64_t bad = 0, n = 0;
    for (uint64_t r = 0; r < in->n_reads; ++r) {
        const uint32_t len = in->len[r];
        if (len < (uint32_t)K) continue;
        const uint8_t* fw = in->bases + in->base_off[r];
        std::vector<uint8_t> rc((len + 3) / 4 + 32, 0);
        for (uint32_t i = 0; i < len; ++i) { uint32_t b = 3u - packed_base(fw, len - 1 - i); rc[i >> 2] |= (uint8_t)(b << ((i & 3) * 2)); }
        for (uint32_t j = 0; j + K <= len; ++j) {
            const uint32_t a = kmer_minimizer_hash(fw, j), b = kmer_minimizer_hash(rc.data(), len - K - j);
            if (a != b || mini_mix(a) != mini_mix(b)) ++bad;
            Kmer f, rcq;                                            // the one-pass (k-mer, reverse complement) pair of the map kernel
            kmer_pair_at(fw, j, &f, &rcq);
            const Kmer f0 = kmer_at(fw, j), r0 = kmer_rc(f0);
            if (!(f == f0) || !(rcq == r0)) ++bad;
            ++n;
        }
    }
    *n_checked = n;
    return bad;
}
int hc_count_records(const uint64_t* recs, uint64_t n, w2rap_kmer_rec** out, uint64_t* n_out) {
    std::vector<Rec> v(n);
    for (uint64_t i = 0; i < n; ++i) v[i] = Rec{recs[2 * i], recs[2 * i + 1] & ~0xffull, (uint32_t)(recs[2 * i + 1] & 0xff)};
    std::sort(v.begin(), v.end(), [](const Rec& a, const Rec& b) { return a.w0 < b.w0 || (a.w0 == b.w0 && a.w1 < b.w1); });
    std::vector<w2rap_kmer_rec> d;
    for (size_t i = 0; i < v.size();) {
        size_t j = i; uint32_t c = 0, ctx = 0;
        while (j < v.size() && v[j].w0 == v[i].w0 && v[j].w1 == v[i].w1) { ctx |= v[j].ctx; ++c; ++j; }
        d.push_back(w2rap_kmer_rec{v[i].w0, v[i].w1, c > 255 ? 255 : c, ctx, 0xffffffffu, 0});
        i = j;
    }
    *n_out = d.size(); *out = dup(d);
    return 0;
}

// Everything after counting, serially, through the same device functions the kernels call.
// `all` = distinct k-mers with counts and raw contexts (dump level 2).
int hc_graph(const w2rap_kmer_rec* all, uint64_t n_all, uint32_t min_freq, const w2rap_reads* in, int want_paths, int apply_fixpaths, uint32_t cap,
             uint32_t left_cap, w2rap_graph* out) {
    memset(out, 0, sizeof(*out));
    // ---- k_insert_solid
    uint64_t n_solid = 0;
    for (uint64_t i = 0; i < n_all; ++i) if (all[i].count >= min_freq) ++n_solid;
    uint32_t lg = 10;
    while ((1ull << lg) < 2 * n_solid) ++lg;
    std::vector<SolidSlot> slots(1ull << lg);
    memset(slots.data(), 0xff, slots.size() * sizeof(SolidSlot));
    SolidTable st{slots.data(), lg};
    const uint64_t T = st.size(), mask = T - 1, nn = 2 * T;
    for (uint64_t i = 0; i < n_all; ++i) {
        if (all[i].count < min_freq) continue;
        Kmer k{all[i].w0, all[i].w1};
        uint64_t h = st.home(k);
        while (slots[h].w0 != EMPTY_W0) h = (h + 1) & mask;
        slots[h].w0 = k.w0; slots[h].w1 = k.w1; slots[h].ctx = all[i].ctx; slots[h].edge = NIL; slots[h].off = 0; slots[h].pad = 0;
    }
    out->n_solid = n_solid; out->n_distinct = n_all;
    // ---- k_adjacency
    for (uint64_t i = 0; i < T; ++i) if (slots[i].w0 != EMPTY_W0) slots[i].ctx = pruned_context(st, Kmer{slots[i].w0, slots[i].w1}, slots[i].ctx & 0xff);
    // ---- k_links, k_rank_*
    std::vector<uint32_t> next0(nn);
    int missing = 0;
    for (uint64_t x = 0; x < nn; ++x) next0[x] = unipath_succ_link(st, (uint32_t)x, &missing);
    if (missing) return 101;
    std::vector<RankState> A(nn, RankState{NIL, 0xffffffffu}), B(nn), D(nn);
    // k_splitter_walk, k_rank_step_inplace, k_splitter_finish
    std::vector<uint32_t> splist;
    for (uint64_t x = 0; x < nn; ++x) if (node_is_splitter(next0.data(), (uint32_t)x)) { splist.push_back((uint32_t)x); splitter_walk(next0.data(), (uint32_t)x, A.data(), B.data()); }
    uint64_t prev_un = ~0ull;
    for (int round = 0; round < 48 && !splist.empty(); ++round) {
        uint64_t un = 0;
        for (uint32_t x : splist) { bool u; B[x] = rank_step_node(B.data(), x, &u); un += u; }
        if (un == 0 || un == prev_un) { prev_un = un; break; }
        prev_un = un;
    }
    prev_un = 0;
    for (uint64_t x = 0; x < nn; ++x) { A[x] = splitter_finish_node(next0.data(), A.data(), B.data(), (uint32_t)x); prev_un += !(A[x].y & RANK_RESOLVED); }
    RankState* cur = A.data(); RankState* oth = B.data();
    out->timings.count_passes = (uint32_t)prev_un;   // reported to the test: nodes that went through the circle path
    if (prev_un) {
        std::vector<uint32_t> list;
        for (uint64_t x = 0; x < nn; ++x) if (!(cur[x].y & RANK_RESOLVED)) list.push_back((uint32_t)x);
        RankState* x0 = oth; RankState* x1 = D.data();
        for (uint32_t x : list) x0[x] = RankState{next0[x], x >> 1};
        for (int round = 0; round < 40; ++round) {
            bool any = false;
            for (uint32_t x : list) { bool ch; x1[x] = cycle_step_node(st, x0, x, &ch); any |= ch; }
            std::swap(x0, x1);
            if (!any) break;
        }
        for (uint32_t x : list) cycle_cut_node(x0, next0.data(), x);
        for (uint32_t x : list) x0[x] = rank_init_node(next0.data(), x);
        for (int round = 0; round < 41; ++round) {
            if (round == 40) return 102;
            uint64_t un = 0;
            for (uint32_t x : list) { bool u; x1[x] = rank_step_node(x0, x, &u); un += u; }
            std::swap(x0, x1);
            if (!un) break;
        }
        for (uint32_t x : list) cur[x] = x0[x];
    }
    const RankState* R = cur;
    // ---- k_strand_decide, k_collect_heads, sort, k_assign_edges, k_emit_edges
    std::vector<uint8_t> keep(nn, 0);
    for (uint64_t x = 0; x < nn; ++x) if (strand_decide_node(st, R, (uint32_t)x, keep.data())) return 103;
    struct Head { Kmer k; uint32_t node, n; };
    std::vector<Head> heads;
    for (uint64_t x = 0; x < nn; ++x) if (head_is_kept(st, R, keep.data(), (uint32_t)x)) heads.push_back(Head{node_kmer(st, (uint32_t)x), (uint32_t)x, (R[x].y & ~RANK_RESOLVED) + 1u});
    std::sort(heads.begin(), heads.end(), [](const Head& a, const Head& b) { return kmer_less(a.k, b.k); });
    const uint64_t E = heads.size();
    std::vector<uint32_t> edge_of_head(nn, NIL), edge_len(E);
    std::vector<uint64_t> edge_off(E + 1, 0);
    for (uint64_t i = 0; i < E; ++i) { edge_of_head[heads[i].node] = (uint32_t)i; edge_len[i] = heads[i].n + K - 1; edge_off[i + 1] = edge_off[i] + (edge_len[i] + 3) / 4; }
    std::vector<uint8_t> edge_bases(edge_off[E] + 32, 0);
    struct Put { uint8_t* b; void operator()(uint64_t bo, uint64_t pos, uint32_t c) const { b[bo + (pos >> 2)] |= (uint8_t)(c << ((pos & 3) * 2)); } } put{edge_bases.data()};
    for (uint64_t x = 0; x < nn; ++x) emit_node(st, R, edge_of_head.data(), edge_off.data(), (uint32_t)x, put);
    // ---- k_edge_ends .. k_adj_sort
    std::vector<EndKey> keys(4 * E);
    std::vector<uint8_t> is_pal(E, 0);
    for (uint64_t idx = 0; idx < 4 * E; ++idx) { bool pal; edge_end_key(edge_bases.data() + edge_off[idx >> 2], edge_len[idx >> 2], (uint32_t)idx & 3, &pal, &keys[idx]); if ((idx & 3) == 0) is_pal[idx >> 2] = pal; }
    std::vector<uint32_t> perm(4 * E);
    for (uint64_t i = 0; i < 4 * E; ++i) perm[i] = (uint32_t)i;
    auto klt = [&](uint32_t a, uint32_t b) { const EndKey &x = keys[a], &y = keys[b]; return x.h != y.h ? x.h < y.h : (x.k0 != y.k0 ? x.k0 < y.k0 : x.k1 < y.k1); };
    std::stable_sort(perm.begin(), perm.end(), klt);
    std::vector<int32_t> edge_vertices(4 * E, -1);
    int64_t vid = -1;
    for (uint64_t i = 0; i < 4 * E; ++i) {
        const EndKey& k = keys[perm[i]];
        if (k.h == ~0ull && k.k0 == ~0ull && k.k1 == ~0ull) continue;
        if (i == 0 || klt(perm[i - 1], perm[i])) ++vid;
        edge_vertices[perm[i]] = (int32_t)vid;
    }
    const uint64_t nv = (uint64_t)(vid + 1);
    std::vector<int32_t> fwd(E), rev(E);
    uint64_t nh = 0;
    for (uint64_t e = 0; e < E; ++e) { fwd[e] = (int32_t)nh++; rev[e] = is_pal[e] ? fwd[e] : (int32_t)nh++; }
    std::vector<uint32_t> hcanon(nh);
    std::vector<int32_t> hleft(nh), hright(nh), from_e(4 * nv, -1), to_e(4 * nv, -1);
    std::vector<uint8_t> from_n(nv, 0), to_n(nv, 0);
    for (uint64_t e = 0; e < E; ++e) {
        hcanon[fwd[e]] = (uint32_t)(e << 1); hleft[fwd[e]] = edge_vertices[4 * e]; hright[fwd[e]] = edge_vertices[4 * e + 1];
        if (rev[e] != fwd[e]) { hcanon[rev[e]] = (uint32_t)(e << 1) | 1u; hleft[rev[e]] = edge_vertices[4 * e + 2]; hright[rev[e]] = edge_vertices[4 * e + 3]; }
    }
    for (uint64_t he = 0; he < nh; ++he) {
        int32_t l = hleft[he], r = hright[he];
        if (from_n[l] >= 4 || to_n[r] >= 4) return 104;
        from_e[4 * l + from_n[l]++] = (int32_t)he; to_e[4 * r + to_n[r]++] = (int32_t)he;
    }
    for (uint64_t v = 0; v < nv; ++v) {
        std::sort(from_e.begin() + 4 * v, from_e.begin() + 4 * v + from_n[v], [&](int32_t a, int32_t b) { return hright[a] != hright[b] ? hright[a] < hright[b] : a < b; });
        std::sort(to_e.begin() + 4 * v, to_e.begin() + 4 * v + to_n[v], [&](int32_t a, int32_t b) { return hleft[a] != hleft[b] ? hleft[a] < hleft[b] : a < b; });
    }
    out->n_edges = E; out->n_vertices = nv; out->n_hbv_edges = nh;
    out->edge_off = dup(edge_off); out->edge_len = dup(edge_len);
    edge_bases.resize(edge_off[E]);
    out->edge_bases = dup(edge_bases);
    edge_bases.resize(edge_off[E] + 32, 0);
    out->edge_vertices = dup(edge_vertices); out->fwd_xlat = dup(fwd); out->rev_xlat = dup(rev);
    for (uint64_t e = 0; e < E; ++e) out->n_edge_bases += edge_len[e];
    {   // dump level 1
        std::vector<w2rap_kmer_rec> d;
        for (uint64_t i = 0; i < T; ++i) if (slots[i].w0 != EMPTY_W0) d.push_back(w2rap_kmer_rec{slots[i].w0, slots[i].w1, 0, slots[i].ctx & 0xff, slots[i].edge, slots[i].off});
        std::sort(d.begin(), d.end(), [](const w2rap_kmer_rec& a, const w2rap_kmer_rec& b) { return a.w0 < b.w0 || (a.w0 == b.w0 && a.w1 < b.w1); });
        out->n_dump = d.size(); out->dump = dup(d);
    }
    if (!want_paths) return 0;
    // ---- k_path_reads (+ overflow retry), k_path_lens, scan, k_path_gather
    // k_bloom_build: the negative-lookup filter of the pathing kernel
    std::vector<uint32_t> bloom_words(std::max<uint64_t>(1024, n_solid / 2), 0u);
    KmerBloom bloom{bloom_words.data(), bloom_words.size()};
    for (uint64_t i = 0; i < T; ++i) if (slots[i].w0 != EMPTY_W0) { uint64_t h = kmer_hash(Kmer{slots[i].w0, slots[i].w1}); bloom_words[bloom_word(bloom, h)] |= bloom_mask(h); }
    GraphView g{st, bloom, edge_bases.data(), edge_off.data(), edge_len.data(), fwd.data(), rev.data(), hcanon.data(), hleft.data(), hright.data(), from_e.data(), to_e.data(), from_n.data(), to_n.data()};
    const uint64_t n = in->n_reads;
    uint32_t maxlen = 0;
    for (uint64_t r = 0; r < n; ++r) maxlen = std::max(maxlen, in->len[r]);
    std::vector<uint8_t> qs(maxlen + 16);
    std::vector<int32_t> row(cap), row2(3 * maxlen + 32), poff(n), pedges;
    std::vector<uint64_t> path_off(n + 1, 0);
    uint64_t n_ovf = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* b = in->bases + in->base_off[r];
        const uint8_t* q = in->quals + in->qual_off[r];
        PathResult pr = path_one_read(g, b, in->len[r], q, qs.data(), row.data(), cap, left_cap, apply_fixpaths != 0);
        const int32_t* src = row.data();
        if (pr.overflow) {
            ++n_ovf;
            pr = path_one_read(g, b, in->len[r], q, qs.data(), row2.data(), 3 * maxlen + 32, maxlen + 16, apply_fixpaths != 0);
            if (pr.overflow) return 105;
            src = row2.data();
        }
        poff[r] = pr.offset;
        path_off[r] = pedges.size();
        pedges.insert(pedges.end(), src + pr.start, src + pr.start + pr.len);
        if (pr.len > 0) out->n_pathed++;
        if (pr.len > 2) out->n_multipathed++;
    }
    path_off[n] = pedges.size();
    out->n_paths = n; out->n_path_edges = pedges.size();
    out->path_offset = dup(poff); out->path_off = dup(path_off); out->path_edges = dup(pedges);
    out->timings.reserved = (uint32_t)n_ovf;   // reported to the test: how many reads took the overflow path
    return 0;
}

void hc_graph_free(w2rap_graph* g) {
    free(g->edge_off); free(g->edge_len); free(g->edge_bases); free(g->edge_vertices); free(g->fwd_xlat); free(g->rev_xlat);
    free(g->path_offset); free(g->path_off); free(g->path_edges); free(g->dump);
    memset(g, 0, sizeof(*g));
}

}  // extern "C"
