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
#include "../../w2rap-contigger_b200/csrc/shardgraph.cuh"
#include "../../w2rap-contigger_b200/csrc/slab_freelist.h"
#include "../../w2rap-contigger_b200/csrc/places.cuh"

using namespace w2r;

namespace {
struct Rec { uint64_t w0, w1; uint32_t ctx; };
struct Collect {
    std::vector<Rec>* v;
    void operator()(Kmer k, uint32_t ctx) const { v->push_back(Rec{k.w0, k.w1, ctx}); }
};
uint64_t table_slots_for(uint64_t n) { return solid_table_slots(n); }      // the product's sizing rule (kmer.cuh)
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

// ---- the map side as the kernel does it (k_minimizer_map): partition of every k-mer by minimiser, runs of equal partition inside a
// 32-position step become one super-k-mer record.  Appends records (4 u64 each) and, per record, its partition.
static int map_read_records(const uint8_t* bases, uint32_t gl, uint32_t logP, uint32_t npass, uint32_t pass, std::vector<uint64_t>* recs, std::vector<uint32_t>* parts) {
    if (gl <= (uint32_t)K) return 0;
    const uint32_t nk = gl - K + 1, last = gl - K;
    std::vector<uint32_t> part(nk);
    for (uint32_t j = 0; j < nk; ++j) {
        const uint32_t mh = mini_mix(kmer_minimizer_hash(bases, j));
        part[j] = (npass > 1 && (mh & 0xffffu) % npass != pass) ? NIL : mini_part(mh, logP);
    }
    for (uint32_t j = 0; j < nk;) {
        if (part[j] == NIL) { ++j; continue; }
        uint32_t e = j + 1;
        while (e < nk && part[e] == part[j] && (e % 32u) != 0) ++e;       // steps are 32 k-mers wide, tiles 192: boundaries at multiples of 32
        const SkmRec r = skm_build(bases, j, e - j, last);
        for (int w = 0; w < 4; ++w) recs->push_back(r.q[w]);
        parts->push_back(part[j]);
        j = e;
    }
    return 0;
}
// Records -> (canonical k-mer, context) per instance, as the reduce expands them (k_count_smem / k_count_region: skm_kmer_at).
static void expand_records(const uint64_t* recs, uint64_t nrec, std::vector<Rec>* out) {
    for (uint64_t i = 0; i < nrec; ++i) {
        const uint64_t* q = recs + 4 * i;
        const uint32_t n = skm_n(q[3]);
        for (uint32_t j = 0; j < n; ++j) { Kmer k; uint32_t ctx; skm_kmer_at(q, j, &k, &ctx); out->push_back(Rec{k.w0, k.w1, ctx}); }
    }
}
// Round trip of the record format: for every read, build + expand must give exactly what extract_read_kmers emits, in order.
// Returns the number of mismatching instances; *n_rec / *n_inst receive the totals.
uint64_t hc_skm_roundtrip(const w2rap_reads* in, uint32_t min_qual, uint32_t logP, uint64_t* n_rec, uint64_t* n_inst) {
    uint64_t bad = 0, nr = 0, ni = 0;
    for (uint64_t r = 0; r < in->n_reads; ++r) {
        uint32_t nq = 0;
        uint32_t gl = pq_good_length(in->quals + in->qual_off[r], min_qual, &nq);
        if (gl > in->len[r]) gl = in->len[r];
        const uint8_t* bases = in->bases + in->base_off[r];
        std::vector<Rec> want, got;
        Collect emit{&want};
        extract_read_kmers(bases, gl, emit);
        std::vector<uint64_t> recs; std::vector<uint32_t> parts;
        map_read_records(bases, gl, logP, 1, 0, &recs, &parts);
        expand_records(recs.data(), parts.size(), &got);
        nr += parts.size(); ni += want.size();
        if (got.size() != want.size()) { bad += want.size() > got.size() ? want.size() - got.size() : got.size() - want.size(); continue; }
        for (size_t i = 0; i < want.size(); ++i) if (want[i].w0 != got[i].w0 || want[i].w1 != got[i].w1 || want[i].ctx != got[i].ctx) ++bad;
    }
    *n_rec = nr; *n_inst = ni;
    return bad;
}

// ---- pieces of the sharded (multi-GPU) protocol, for the world_size-2 gloo test: the "map" side (k_minimizer_map) and the
// "reduce" side (k_count_smem) as separate calls, with the product's own partition/owner functions.
// recs: n x 4 u64 (SkmRec); owner[i] = rank that owns record i's partition (keyed by minimiser, extract.cuh).
int hc_extract_records(const w2rap_reads* in, uint32_t min_qual, uint32_t logP, uint32_t world, uint64_t** recs_out, uint32_t** owner_out, uint64_t* n_out) {
    std::vector<uint64_t> flat;
    std::vector<uint32_t> owner;
    for (uint64_t r = 0; r < in->n_reads; ++r) {
        uint32_t nq = 0;
        uint32_t gl = pq_good_length(in->quals + in->qual_off[r], min_qual, &nq);
        if (nq != in->len[r]) return 100;
        if (gl > in->len[r]) gl = in->len[r];
        map_read_records(in->bases + in->base_off[r], gl, logP, 1, 0, &flat, &owner);
    }
    for (uint32_t& o : owner) o = owner_of_partition(o, logP, world);
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
    std::vector<Rec> v;
    expand_records(recs, n, &v);
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

}  // extern "C"

namespace {
// k_splitter_walk, k_rank_step_inplace, k_splitter_finish, k_cycle_*: list ranking of one (local) table.  R receives the final
// (tail, distance | RESOLVED) of every node; returns the number of nodes that went through the circle path, or -1 on failure.
int64_t rank_local(const SolidTable& st, std::vector<uint32_t>& next0, const uint8_t* ghead, std::vector<RankState>& R) {
    const uint64_t nn = next0.size();
    std::vector<RankState> A(nn, RankState{NIL, 0xffffffffu}), B(nn), D(nn);
    std::vector<uint32_t> splist;
    for (uint64_t x = 0; x < nn; ++x) if (node_is_splitter(next0.data(), ghead, (uint32_t)x)) { splist.push_back((uint32_t)x); splitter_walk(next0.data(), (uint32_t)x, A.data(), B.data()); }
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
            if (round == 40) return -1;
            uint64_t un = 0;
            for (uint32_t x : list) { bool u; x1[x] = rank_step_node(x0, x, &u); un += u; }
            std::swap(x0, x1);
            if (!un) break;
        }
        for (uint32_t x : list) cur[x] = x0[x];
    }
    R.assign(cur, cur + nn);
    return (int64_t)prev_un;
}

struct EdgeSet {
    uint64_t E = 0;
    std::vector<uint32_t> edge_len;
    std::vector<uint64_t> edge_off;
    std::vector<uint8_t> edge_bases;     // + 32 bytes of padding
};
struct PutBase { uint8_t* b; void operator()(uint64_t bo, uint64_t pos, uint32_t c) const { b[bo + (pos >> 2)] |= (uint8_t)(c << ((pos & 3) * 2)); } };

// k_edge_ends .. k_adj_sort, dump level 1, k_path_reads: everything that follows the edges, on the WHOLE dictionary `st`
// (every solid k-mer with its pruned context, edge and offset).
int finish_graph(const std::vector<PathSlice>& slices, EdgeSet& es, uint64_t n_solid, const w2rap_reads* in, int want_paths, int apply_fixpaths, uint32_t cap, uint32_t left_cap, w2rap_graph* out) {
    const uint64_t E = es.E;
    const uint32_t W = (uint32_t)slices.size();
    std::vector<uint32_t>& edge_len = es.edge_len; std::vector<uint64_t>& edge_off = es.edge_off; std::vector<uint8_t>& edge_bases = es.edge_bases;
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
    {   // k_involution
        std::vector<int32_t> inv(nh);
        for (uint64_t e = 0; e < E; ++e) { inv[fwd[e]] = rev[e]; inv[rev[e]] = fwd[e]; }
        out->involution = dup(inv);
    }
    for (uint64_t e = 0; e < E; ++e) out->n_edge_bases += edge_len[e];
    {   // dump level 1
        std::vector<w2rap_kmer_rec> d;
        for (const PathSlice& sl : slices)
            for (uint64_t i = 0; i < sl.nslots; ++i) if (sl.tab[i].w0 != EMPTY_W0) d.push_back(w2rap_kmer_rec{sl.tab[i].w0, sl.tab[i].w1, 0, sl.tab[i].ctx & 0xff, sl.tab[i].edge, sl.tab[i].off});
        std::sort(d.begin(), d.end(), [](const w2rap_kmer_rec& a, const w2rap_kmer_rec& b) { return a.w0 < b.w0 || (a.w0 == b.w0 && a.w1 < b.w1); });
        out->n_dump = d.size(); out->dump = dup(d);
    }
    if (!want_paths) return 0;
    // ---- k_path_reads (+ overflow retry), k_path_lens, scan, k_path_gather
    // k_bloom_build: the negative-lookup filter of the pathing kernel
    // k_bloom_build per slice (sharded: every rank builds the filter slice of its own dictionary slice, the slices are all-gathered)
    const uint32_t slice_words = (uint32_t)std::max<uint64_t>(1024, n_solid / 2 / W);
    std::vector<uint32_t> bloom_words((size_t)slice_words * W, 0u);
    for (const PathSlice& sl : slices)
        for (uint64_t i = 0; i < sl.nslots; ++i) if (sl.tab[i].w0 != EMPTY_W0) { uint32_t h = bloom_hash(Kmer{sl.tab[i].w0, sl.tab[i].w1}); bloom_words[pd_bloom_word(W, slice_words, h)] |= bloom_mask(h); }
    GraphView g{PathDict{slices.data(), W, bloom_words.data(), slice_words}, edge_bases.data(), edge_off.data(), edge_len.data(), fwd.data(), rev.data(), hcanon.data(), hleft.data(), hright.data(), from_e.data(), to_e.data(), from_n.data(), to_n.data()};
    const uint64_t n = in->n_reads;
    uint32_t maxlen = 0;
    for (uint64_t r = 0; r < n; ++r) maxlen = std::max(maxlen, in->len[r]);
    std::vector<int32_t> row(cap), row2(3 * maxlen + 32), poff(n), pedges;
    std::vector<uint64_t> path_off(n + 1, 0);
    uint64_t n_ovf = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* b = in->bases + in->base_off[r];
        const uint8_t* q = in->quals + in->qual_off[r];
        PathResult pr = path_one_read(g, b, in->len[r], q, row.data(), cap, left_cap, apply_fixpaths != 0);
        const int32_t* src = row.data();
        if (pr.overflow) {
            ++n_ovf;
            pr = path_one_read(g, b, in->len[r], q, row2.data(), 3 * maxlen + 32, maxlen + 16, apply_fixpaths != 0);
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


}  // namespace

extern "C" {

// Everything after counting, serially, through the same device functions the kernels call.
// `all` = distinct k-mers with counts and raw contexts (dump level 2).
int hc_graph(const w2rap_kmer_rec* all, uint64_t n_all, uint32_t min_freq, const w2rap_reads* in, int want_paths, int apply_fixpaths, uint32_t cap,
             uint32_t left_cap, w2rap_graph* out) {
    memset(out, 0, sizeof(*out));
    // ---- k_insert_solid
    uint64_t n_solid = 0;
    for (uint64_t i = 0; i < n_all; ++i) if (all[i].count >= min_freq) ++n_solid;
    std::vector<SolidSlot> slots(table_slots_for(n_solid));
    memset(slots.data(), 0xff, slots.size() * sizeof(SolidSlot));
    SolidTable st{slots.data(), slots.size()};
    const uint64_t T = st.size(), nn = 2 * T;
    for (uint64_t i = 0; i < n_all; ++i) {
        if (all[i].count < min_freq) continue;
        Kmer k{all[i].w0, all[i].w1};
        uint64_t h = st.home(k);
        while (slots[h].w0 != EMPTY_W0) h = st.next(h);
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
    std::vector<RankState> Rv;
    const int64_t ncyc = rank_local(st, next0, nullptr, Rv);
    if (ncyc < 0) return 102;
    out->timings.count_passes = (uint32_t)ncyc;   // reported to the test: nodes that went through the circle path
    const RankState* R = Rv.data();
    // ---- k_strand_decide, k_collect_heads, sort, k_assign_edges, k_emit_edges
    std::vector<uint8_t> keep(nn, 0);
    for (uint64_t x = 0; x < nn; ++x) if (strand_decide_node(st, R, (uint32_t)x, keep.data())) return 103;
    struct Head { Kmer k; uint32_t node, n; };
    std::vector<Head> heads;
    for (uint64_t x = 0; x < nn; ++x) if (head_is_kept(st, R, keep.data(), (uint32_t)x)) heads.push_back(Head{node_kmer(st, (uint32_t)x), (uint32_t)x, (R[x].y & ~RANK_RESOLVED) + 1u});
    std::sort(heads.begin(), heads.end(), [](const Head& a, const Head& b) { return kmer_less(a.k, b.k); });
    EdgeSet es;
    es.E = heads.size();
    const uint64_t E = es.E;
    std::vector<uint32_t> edge_of_head(nn, NIL);
    es.edge_len.resize(E); es.edge_off.assign(E + 1, 0);
    for (uint64_t i = 0; i < E; ++i) { edge_of_head[heads[i].node] = (uint32_t)i; es.edge_len[i] = heads[i].n + K - 1; es.edge_off[i + 1] = es.edge_off[i] + (es.edge_len[i] + 3) / 4; }
    es.edge_bases.assign(es.edge_off[E] + 32, 0);
    PutBase put{es.edge_bases.data()};
    for (uint64_t x = 0; x < nn; ++x) emit_node(st, R, edge_of_head.data(), es.edge_off.data(), (uint32_t)x, put);
    return finish_graph(std::vector<PathSlice>{PathSlice{slots.data(), T}}, es, n_solid, in, want_paths, apply_fixpaths, cap, left_cap, out);
}

}  // extern "C"

namespace {
// ---- one rank of the SHARDED graph stage (csrc/shardgraph.cuh), phase by phase, in the order pipeline.cu: graph_stage_sharded() runs
// them.  Between the phases the ranks exchange flat buffers: by plain copies in hc_graph_sharded() below (simulated ranks in one
// process), over torch.distributed/gloo in tests/sharded_gloo_worker.py (one process per rank).
struct CycleNodeH { uint32_t piece, pad; uint64_t w0, w1, gid; };
struct SgRank {
    uint32_t world = 1, me = 0, logP = 0;
    std::vector<w2rap_kmer_rec> owned;
    std::vector<std::vector<Kmer>> q;            // q[d] = canonical k-mers asked of rank d
    std::vector<std::vector<uint32_t>> gslot;    // local slot of every query's ghost
    std::vector<SolidSlot> slots;
    SolidTable st{nullptr, 0};
    std::vector<uint32_t> next0;
    std::vector<uint8_t> ghead;
    std::vector<RankState> R;
    std::vector<uint32_t> lpiece, lhead;
    std::vector<PieceRec> pieces;                // mine
    std::vector<PieceRec> P;                     // everybody's (replicated)
    std::vector<uint64_t> poff;
    std::vector<uint32_t> nxt, flip;
    std::vector<RankState> S;
    std::vector<PieceInfo> pinfo;
    std::vector<uint8_t> is_head;
    std::vector<uint64_t> chain_n;
    std::vector<CycleNodeH> cyc;
    EdgeSet es;
    uint64_t n_ghost = 0;

    // phase 1: neighbour queries (k_neighbour_queries) and the local table (k_insert_solid)
    void start() {
        q.assign(world, {}); gslot.assign(world, {});
        struct Emit { std::vector<std::vector<Kmer>>* q; void operator()(uint32_t o, Kmer k) const { (*q)[o].push_back(k); } } emit{&q};
        uint64_t nq = 0;
        for (const w2rap_kmer_rec& e : owned) neighbour_queries(Kmer{e.w0, e.w1}, e.ctx & 0xffu, logP, world, me, emit);
        for (auto& v : q) nq += v.size();
        slots.resize(table_slots_for(owned.size() + nq));
        memset(slots.data(), 0xff, slots.size() * sizeof(SolidSlot));
        st = SolidTable{slots.data(), slots.size()};
        for (const w2rap_kmer_rec& e : owned) {
            uint64_t h = st.home(Kmer{e.w0, e.w1});
            while (slots[h].w0 != EMPTY_W0) h = st.next(h);
            slots[h] = SolidSlot{e.w0, e.w1, e.ctx & 0xffu, NIL, 0, 0};
        }
    }
    // phase 2 (owner side, before any ghost exists): k_answer_queries
    void answer(const Kmer* keys, uint64_t n, uint32_t* reply) const {
        for (uint64_t i = 0; i < n; ++i) { const int64_t sl = solid_find(st, keys[i]); reply[i] = sl < 0 ? NIL : (uint32_t)sl; }
    }
    // phase 3: k_insert_ghosts for the replies of rank d, then (after all d) k_adjacency
    void insert_ghosts(uint32_t d, const uint32_t* reply) {
        gslot[d].assign(q[d].size(), NIL);
        for (size_t i = 0; i < q[d].size(); ++i) {
            if (reply[i] == NIL) continue;
            const Kmer k = q[d][i];
            uint64_t h = st.home(k);
            while (slots[h].w0 != EMPTY_W0 && !(slots[h].w0 == k.w0 && slots[h].w1 == k.w1)) h = st.next(h);
            if (slots[h].w0 == EMPTY_W0) { slots[h] = SolidSlot{k.w0, k.w1, 0, reply[i], 0, d + 1u}; ++n_ghost; }
            gslot[d][i] = (uint32_t)h;
        }
    }
    void adjacency() {
        for (uint64_t i = 0; i < st.size(); ++i)
            if (slots[i].w0 != EMPTY_W0 && !slot_is_ghost(slots[i])) slots[i].ctx = pruned_context(st, Kmer{slots[i].w0, slots[i].w1}, slots[i].ctx & 0xff);
    }
    // phase 4 (owner side, after adjacency): k_gather_ctx — the pruned contexts of the slots that were asked for
    void ctx_answer(const uint32_t* slot, uint64_t n, uint32_t* ctx) const { for (uint64_t i = 0; i < n; ++i) ctx[i] = slot[i] == NIL ? 0u : (slots[slot[i]].ctx & 0xffu); }
    void apply_ghost_ctx(uint32_t d, const uint32_t* ctx) { for (size_t i = 0; i < gslot[d].size(); ++i) if (gslot[d][i] != NIL) slots[gslot[d][i]].ctx = ctx[i]; }
    // phase 5: k_links_sharded
    int links() {
        const uint64_t nn = 2 * st.size();
        next0.resize(nn); ghead.assign(nn, 0);
        int missing = 0;
        for (uint64_t x = 0; x < nn; ++x) { bool tg; next0[x] = unipath_succ_link(st, (uint32_t)x, &missing, &tg); if (tg) ghead[x ^ 1u] = 1; }
        return missing;
    }
    // phase 6: local list ranking, k_emit_pieces, k_piece_flips
    int rank_and_pieces() {
        if (rank_local(st, next0, ghead.data(), R) < 0) return 102;
        const uint64_t nn = next0.size();
        pieces.clear(); lpiece.assign(nn, NIL); lhead.assign(nn, NIL);
        for (uint64_t x = 0; x < nn; ++x)
            if (node_is_piece_head(next0.data(), ghead.data(), (uint32_t)x)) {
                const PieceRec p = piece_of_head(st, next0.data(), R.data(), me, (uint32_t)x);
                lpiece[p.flip_local] = (uint32_t)pieces.size();          // (flip_local still holds the tail node)
                lhead[x] = (uint32_t)pieces.size();
                pieces.push_back(p);
            }
        for (PieceRec& p : pieces) p.flip_local = lhead[p.flip_local ^ 1u];
        return 0;
    }
    // phase 7 (replicated): all pieces of all ranks -> links, ranks.  Returns the number of unranked pieces (circles), or < 0.
    int64_t set_pieces(const PieceRec* all, uint64_t n_all, const uint64_t* off /* [world + 1] */) {
        P.assign(all, all + n_all); poff.assign(off, off + world + 1);
        const uint64_t np = P.size();
        uint64_t msz = 64; while (msz < 2 * np) msz <<= 1;
        std::vector<uint64_t> mkeys(msz, GID_NONE); std::vector<uint32_t> mvals(msz, NIL);
        GidMap gm{mkeys.data(), mvals.data(), msz - 1};
        for (uint64_t i = 0; i < np; ++i) { uint64_t h = gid_hash(P[i].head) & gm.mask; while (mkeys[h] != GID_NONE) h = (h + 1) & gm.mask; mkeys[h] = P[i].head; mvals[h] = (uint32_t)i; }
        nxt.assign(np, NIL); flip.assign(np, NIL);
        for (uint64_t i = 0; i < np; ++i) {
            if (P[i].succ != GID_NONE) { nxt[i] = gid_find(gm, P[i].succ); if (nxt[i] == NIL) return -107; }
            if (P[i].flip_local == NIL) return -108;
            flip[i] = (uint32_t)poff[P[i].head >> 32] + P[i].flip_local;
        }
        S.resize(np);
        for (uint64_t i = 0; i < np; ++i) S[i] = piece_rank_init(P.data(), nxt.data(), (uint32_t)i);
        uint64_t prev = ~0ull, un = 0;
        for (int round = 0; round < 64; ++round) {
            un = 0;
            for (uint64_t i = 0; i < np; ++i) if (!(S[i].y & RANK_RESOLVED)) { S[i] = piece_rank_step(S[i], S[S[i].x]); un += !(S[i].y & RANK_RESOLVED); }
            if (un == 0 || un == prev) break;
            prev = un;
        }
        return (int64_t)un;
    }
    // phase 7b: circles that span ranks (BuildReadQGraph.cc:126-180): my nodes on unranked pieces (k_collect_cycle_nodes) ...
    void collect_cycle_nodes() {
        cyc.clear();
        for (uint64_t x = 0; x < next0.size(); ++x) {
            if (next0[x] == EMPTY_NODE || next0[x] == GHOST_TAIL) continue;
            const uint32_t pi = (uint32_t)poff[me] + lpiece[R[x].x];
            if (S[pi].y & RANK_RESOLVED) continue;
            const SolidSlot& sl = slots[x >> 1];
            cyc.push_back(CycleNodeH{pi, 0u, sl.w0, sl.w1, gid_make(me, (uint32_t)x)});
        }
    }
    // ... and, from everybody's (the host part of graph_stage_sharded + k_apply_cycle_cuts): cut every circle at its minimum k-mer
    void apply_cuts(const CycleNodeH* all, uint64_t n_all) {
        const uint64_t np = P.size();
        std::vector<uint32_t> lab(np, NIL);
        for (uint64_t i = 0; i < np; ++i) {
            if ((S[i].y & RANK_RESOLVED) || lab[i] != NIL) continue;
            uint32_t mn = (uint32_t)i;
            for (uint32_t j = nxt[i]; j != i; j = nxt[j]) mn = std::min(mn, j);
            lab[i] = mn;
            for (uint32_t j = nxt[i]; j != i; j = nxt[j]) lab[j] = mn;
        }
        std::vector<CycleNodeH> cn(all, all + n_all);
        std::sort(cn.begin(), cn.end(), [&](const CycleNodeH& a, const CycleNodeH& b) {
            const uint32_t la = lab[a.piece], lb = lab[b.piece];
            if (la != lb) return la < lb;
            if (a.w0 != b.w0) return a.w0 < b.w0;
            if (a.w1 != b.w1) return a.w1 < b.w1;
            return a.gid < b.gid;
        });
        std::vector<uint64_t> heads_g, tails_g;            // (kmin,+) becomes a head, (kmin,-) a tail
        for (size_t i = 0; i < cn.size(); ++i) {
            if (i > 0 && lab[cn[i].piece] == lab[cn[i - 1].piece]) continue;
            for (size_t j = i; j < cn.size() && lab[cn[j].piece] == lab[cn[i].piece] && cn[j].w0 == cn[i].w0 && cn[j].w1 == cn[i].w1; ++j)      // (a circle that is its own reverse complement holds both)
                if (cn[j].gid & 1ull) tails_g.push_back(cn[j].gid); else heads_g.push_back(cn[j].gid);
        }
        std::sort(heads_g.begin(), heads_g.end()); std::sort(tails_g.begin(), tails_g.end());
        for (const CycleNodeH& c : cyc) {
            const uint32_t x = (uint32_t)c.gid;
            if (std::binary_search(heads_g.begin(), heads_g.end(), c.gid)) ghead[x] = 0;
            const uint32_t nx = next0[x];
            bool cut = std::binary_search(tails_g.begin(), tails_g.end(), c.gid);
            if (!cut && nx < GHOST_TAIL) {
                const SolidSlot& sl = slots[nx >> 1];
                const uint64_t g = slot_is_ghost(sl) ? gid_make(sl.pad - 1u, 2u * sl.edge + (nx & 1u)) : gid_make(me, nx);
                cut = std::binary_search(heads_g.begin(), heads_g.end(), g);
            }
            if (cut) next0[x] = NIL;
        }
    }
    // phase 8: strands — even lengths from the piece records (k_chain_tails, replicated), odd lengths by the piece that holds the
    // middle k-mer (k_piece_keep_odd); keepp is then max-reduced over the ranks
    int strands(uint8_t* keepp) {
        const uint64_t np = P.size();
        PieceView pv{P.data(), flip.data(), S.data(), np};
        is_head.assign(np, 0); chain_n.assign(np, 0);
        memset(keepp, 0, np);
        for (uint64_t i = 0; i < np; ++i) {
            if (nxt[i] != NIL) continue;
            uint32_t hp; uint64_t n;
            chain_of_tail_piece(pv, (uint32_t)i, &hp, &n);
            if (n > 0x1000000ull) return 103;
            is_head[hp] = 1; chain_n[hp] = n;
            const uint32_t kf = chain_keep_even(pv, (uint32_t)i, hp, n);
            if (kf != 2u) keepp[hp] = (uint8_t)kf;
        }
        pinfo.resize(np);
        for (uint64_t i = 0; i < np; ++i) pinfo[i] = piece_info(pv, (uint32_t)i);
        for (size_t j = 0; j < pieces.size(); ++j) {
            const uint32_t gi = (uint32_t)poff[me] + (uint32_t)j;
            const int kf = piece_keep_odd(st, next0.data(), pinfo[gi], (uint32_t)P[gi].head);
            if (kf >= 0) keepp[pinfo[gi].head_piece] = (uint8_t)kf;
        }
        return 0;
    }
    // phase 9: edge ids from the kept heads (replicated), my k-mers' bases into a zeroed edge array (k_emit_edges_sharded); the arrays are
    // then OR-reduced over the ranks
    void edges(const uint8_t* keepp) {
        const uint64_t np = P.size();
        struct Head { Kmer k; uint32_t piece; };
        std::vector<Head> heads;
        for (uint64_t i = 0; i < np; ++i) if (is_head[i] && keepp[i]) heads.push_back(Head{P[i].head_k, (uint32_t)i});
        std::sort(heads.begin(), heads.end(), [](const Head& a, const Head& b) { return kmer_less(a.k, b.k); });
        es = EdgeSet();
        es.E = heads.size();
        const uint64_t E = es.E;
        std::vector<uint32_t> edge_of_piece(np, NIL);
        es.edge_len.resize(E); es.edge_off.assign(E + 1, 0);
        for (uint64_t i = 0; i < E; ++i) { edge_of_piece[heads[i].piece] = (uint32_t)i; es.edge_len[i] = (uint32_t)chain_n[heads[i].piece] + K - 1; es.edge_off[i + 1] = es.edge_off[i] + (es.edge_len[i] + 3) / 4; }
        es.edge_bases.assign(es.edge_off[E] + 32, 0);
        PutBase put{es.edge_bases.data()};
        for (uint64_t i = 0; i < np; ++i) pinfo[i].head_piece = edge_of_piece[pinfo[i].head_piece];       // k_piece_edges
        for (uint64_t x = 0; x < next0.size(); ++x) {
            if (next0[x] == EMPTY_NODE || next0[x] == GHOST_TAIL) continue;
            emit_node_sharded(st, R.data(), lpiece.data(), pinfo.data() + poff[me], es.edge_off.data(), (uint32_t)x, put);
        }
    }
    // phase 10: my finished entries for dictionary slice d (k_dump_owned_sliced)
    void entries_for(uint32_t d, std::vector<SolidSlot>* out) const {
        for (uint64_t i = 0; i < st.size(); ++i) {
            const SolidSlot& sl = slots[i];
            if (sl.w0 != EMPTY_W0 && !slot_is_ghost(sl) && pd_slice_of(world, bloom_hash(Kmer{sl.w0, sl.w1})) == d) out->push_back(sl);
        }
    }
};
// a dictionary slice from its entries (k_insert_entries)
std::vector<SolidSlot> build_slice(const SolidSlot* e, uint64_t n) {
    std::vector<SolidSlot> t(table_slots_for(n));
    memset(t.data(), 0xff, t.size() * sizeof(SolidSlot));
    SolidTable tt{t.data(), t.size()};
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t h = tt.home(Kmer{e[i].w0, e[i].w1});
        while (t[h].w0 != EMPTY_W0) h = tt.next(h);
        t[h] = e[i];
    }
    return t;
}
}  // namespace

extern "C" {

// `world` simulated ranks in one process: exchanges are plain copies.  Result: the same graph the single-table path builds.
// out->timings.reserved receives the number of ghost entries, out->timings.count_passes the number of pieces (local chains).
int hc_graph_sharded(const w2rap_kmer_rec* all, uint64_t n_all, uint32_t min_freq, uint32_t world, uint32_t logP, const w2rap_reads* in, int want_paths,
                     int apply_fixpaths, uint32_t cap, uint32_t left_cap, w2rap_graph* out) {
    memset(out, 0, sizeof(*out));
    std::vector<SgRank> rk(world);
    uint64_t n_solid = 0;
    for (uint32_t r = 0; r < world; ++r) { rk[r].world = world; rk[r].me = r; rk[r].logP = logP; }
    for (uint64_t i = 0; i < n_all; ++i) {
        if (all[i].count < min_freq) continue;
        ++n_solid;
        rk[kmer_owner(Kmer{all[i].w0, all[i].w1}, logP, world)].owned.push_back(all[i]);
    }
    out->n_solid = n_solid; out->n_distinct = n_all;
    for (SgRank& me : rk) me.start();
    // all-to-all of the keys, answers, all-to-all of the replies
    std::vector<std::vector<std::vector<uint32_t>>> reply(world, std::vector<std::vector<uint32_t>>(world));
    for (uint32_t r = 0; r < world; ++r)
        for (uint32_t d = 0; d < world; ++d) { reply[r][d].resize(rk[r].q[d].size()); rk[d].answer(rk[r].q[d].data(), rk[r].q[d].size(), reply[r][d].data()); }
    for (uint32_t r = 0; r < world; ++r) for (uint32_t d = 0; d < world; ++d) rk[r].insert_ghosts(d, reply[r][d].data());
    for (SgRank& me : rk) me.adjacency();
    // second round: the owners' pruned contexts of the slots they reported
    for (uint32_t r = 0; r < world; ++r)
        for (uint32_t d = 0; d < world; ++d) {
            std::vector<uint32_t> ctx(reply[r][d].size());
            rk[d].ctx_answer(reply[r][d].data(), reply[r][d].size(), ctx.data());
            rk[r].apply_ghost_ctx(d, ctx.data());
        }
    for (SgRank& me : rk) if (me.links()) return 101;
    for (int iteration = 0;; ++iteration) {
        if (iteration > 1) return 106;
        for (SgRank& me : rk) { const int rc = me.rank_and_pieces(); if (rc) return rc; }
        std::vector<PieceRec> P;
        std::vector<uint64_t> off(world + 1, 0);
        for (uint32_t r = 0; r < world; ++r) { off[r] = P.size(); P.insert(P.end(), rk[r].pieces.begin(), rk[r].pieces.end()); }
        off[world] = P.size();
        int64_t un = 0;
        for (SgRank& me : rk) { un = me.set_pieces(P.data(), P.size(), off.data()); if (un < 0) return (int)-un; }
        if (un == 0) break;
        std::vector<CycleNodeH> cn;
        for (SgRank& me : rk) { me.collect_cycle_nodes(); cn.insert(cn.end(), me.cyc.begin(), me.cyc.end()); }
        for (SgRank& me : rk) me.apply_cuts(cn.data(), cn.size());
    }
    const uint64_t np = rk[0].P.size();
    uint64_t n_ghost = 0;
    for (SgRank& me : rk) n_ghost += me.n_ghost;
    // strands (all-reduce max of the keep flags), edges (all-reduce OR of the bases)
    std::vector<uint8_t> keepp(np, 0), mine(np);
    for (SgRank& me : rk) { const int rc = me.strands(mine.data()); if (rc) return rc; for (uint64_t i = 0; i < np; ++i) keepp[i] = std::max(keepp[i], mine[i]); }
    for (SgRank& me : rk) me.edges(keepp.data());
    EdgeSet es = rk[0].es;
    for (uint32_t r = 1; r < world; ++r) for (size_t i = 0; i < es.edge_bases.size(); ++i) es.edge_bases[i] |= rk[r].es.edge_bases[i];
    // the pathing dictionary: finished entries re-sharded by k-mer hash (all-to-all), slice r built on rank r, slices all-gathered
    std::vector<std::vector<SolidSlot>> stab(world);
    for (uint32_t d = 0; d < world; ++d) {
        std::vector<SolidSlot> e;
        for (SgRank& me : rk) me.entries_for(d, &e);
        stab[d] = build_slice(e.data(), e.size());
    }
    std::vector<PathSlice> slices;
    for (uint32_t r = 0; r < world; ++r) slices.push_back(PathSlice{stab[r].data(), stab[r].size()});
    const int rc = finish_graph(slices, es, n_solid, in, want_paths, apply_fixpaths, cap, left_cap, out);
    out->timings.count_passes = (uint32_t)np;
    if (!want_paths) out->timings.reserved = (uint32_t)n_ghost;
    return rc;
}

// ---- the same phases for ONE rank of a real multi-process run (tests/sharded_gloo_worker.py drives them over gloo)
void* sg_new(const w2rap_kmer_rec* owned, uint64_t n, uint32_t world, uint32_t rank, uint32_t logP) {
    SgRank* r = new SgRank();
    r->world = world; r->me = rank; r->logP = logP; r->owned.assign(owned, owned + n);
    r->start();
    return r;
}
void sg_delete(void* h) { delete (SgRank*)h; }
uint32_t sg_owner(uint64_t w0, uint64_t w1, uint32_t logP, uint32_t world) { return kmer_owner(Kmer{w0, w1}, logP, world); }
uint64_t sg_queries(void* h, uint32_t d, const void** keys) { SgRank* r = (SgRank*)h; *keys = r->q[d].data(); return r->q[d].size(); }
void sg_answer(void* h, const void* keys, uint64_t n, uint32_t* reply) { ((SgRank*)h)->answer((const Kmer*)keys, n, reply); }
void sg_insert_ghosts(void* h, uint32_t d, const uint32_t* reply) { ((SgRank*)h)->insert_ghosts(d, reply); }
void sg_adjacency(void* h) { ((SgRank*)h)->adjacency(); }
void sg_ctx_answer(void* h, const uint32_t* slot, uint64_t n, uint32_t* ctx) { ((SgRank*)h)->ctx_answer(slot, n, ctx); }
void sg_apply_ghost_ctx(void* h, uint32_t d, const uint32_t* ctx) { ((SgRank*)h)->apply_ghost_ctx(d, ctx); }
int sg_links(void* h) { return ((SgRank*)h)->links(); }
int sg_rank_and_pieces(void* h, const void** pieces, uint64_t* n) { SgRank* r = (SgRank*)h; const int rc = r->rank_and_pieces(); *pieces = r->pieces.data(); *n = r->pieces.size(); return rc; }
int64_t sg_set_pieces(void* h, const void* all, uint64_t n_all, const uint64_t* off) { return ((SgRank*)h)->set_pieces((const PieceRec*)all, n_all, off); }
uint64_t sg_cycle_nodes(void* h, const void** nodes) { SgRank* r = (SgRank*)h; r->collect_cycle_nodes(); *nodes = r->cyc.data(); return r->cyc.size(); }
void sg_apply_cuts(void* h, const void* all, uint64_t n_all) { ((SgRank*)h)->apply_cuts((const CycleNodeH*)all, n_all); }
int sg_strands(void* h, uint8_t* keepp) { return ((SgRank*)h)->strands(keepp); }
uint64_t sg_edges(void* h, const uint8_t* keepp, void** edge_bases) { SgRank* r = (SgRank*)h; r->edges(keepp); *edge_bases = r->es.edge_bases.data(); return r->es.edge_bases.size(); }
uint64_t sg_entries(void* h, uint32_t d, void** out) {      // caller frees *out with hc_free
    std::vector<SolidSlot> e;
    ((SgRank*)h)->entries_for(d, &e);
    *out = dup(e);
    return e.size();
}
uint32_t sg_sizeof(int what) { return what == 0 ? sizeof(PieceRec) : (what == 1 ? sizeof(CycleNodeH) : sizeof(SolidSlot)); }
// everything after the graph stage, from the (OR-reduced) edge bases of this rank and the entry lists of ALL dictionary slices
int sg_finish(void* h, const void* const* slice_entries, const uint64_t* slice_n, uint64_t n_solid, const w2rap_reads* in, int want_paths, int apply_fixpaths, uint32_t cap,
              uint32_t left_cap, w2rap_graph* out) {
    SgRank* r = (SgRank*)h;
    memset(out, 0, sizeof(*out));
    out->n_solid = n_solid;
    std::vector<std::vector<SolidSlot>> stab(r->world);
    std::vector<PathSlice> slices;
    for (uint32_t d = 0; d < r->world; ++d) { stab[d] = build_slice((const SolidSlot*)slice_entries[d], slice_n[d]); slices.push_back(PathSlice{stab[d].data(), stab[d].size()}); }
    return finish_graph(slices, r->es, n_solid, in, want_paths, apply_fixpaths, cap, left_cap, out);
}

void hc_graph_free(w2rap_graph* g) {
    free(g->edge_off); free(g->edge_len); free(g->edge_bases); free(g->edge_vertices); free(g->fwd_xlat); free(g->rev_xlat); free(g->involution);
    free(g->path_offset); free(g->path_off); free(g->path_edges); free(g->dump);
    memset(g, 0, sizeof(*g));
}

}  // extern "C"


// ---- the device slab's free list (csrc/slab_freelist.h) against a byte-map model: random takes, give-backs and growth steps.
// Checks after every operation: pieces never overlap and lie inside the range, a take returns the LOWEST address that fits,
// free ranges are disjoint and fully coalesced, and `used` matches.  Returns 0, or the number of the first failed check.
extern "C" int hc_slab_freelist_fuzz(uint64_t seed, uint32_t ops) {
    w2r::SlabFreeList fl;
    std::vector<uint8_t> model;                 // 1 = in use, per unit (units keep the model small; sizes are multiples of it)
    const size_t unit = 4096;
    std::vector<std::pair<size_t, size_t>> live;
    uint64_t x = seed * 0x9e3779b97f4a7c15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    auto check = [&]() -> int {
        size_t used = 0, runs = 0;
        for (size_t i = 0; i < model.size(); ++i) { used += model[i]; if (!model[i] && (i == 0 || model[i - 1])) ++runs; }
        if (used * unit != fl.used) return 1;
        if (runs != fl.free_.size()) return 2;                               // not coalesced (or a range lost)
        size_t prev_end = 0; bool first = true;
        for (auto& kv : fl.free_) {
            if (kv.first % unit || kv.second % unit || kv.second == 0) return 3;
            if (!first && kv.first <= prev_end) return 4;                    // overlapping or touching ranges
            for (size_t i = kv.first / unit; i < (kv.first + kv.second) / unit; ++i) if (i >= model.size() || model[i]) return 5;
            prev_end = kv.first + kv.second; first = false;
        }
        return 0;
    };
    for (uint32_t op = 0; op < ops; ++op) {
        const uint64_t r = rnd();
        if (r % 8 == 0 || model.empty()) {                                   // growth: new backing at the top
            const size_t add = (1 + rnd() % 64) * unit;
            fl.add_free(model.size() * unit, add);
            model.resize(model.size() + add / unit, 0);
        } else if (r % 8 < 5) {                                              // take
            const size_t want = (1 + rnd() % 48) * unit;
            size_t expect = w2r::SlabFreeList::NONE, run = 0;
            for (size_t i = 0; i < model.size(); ++i) { run = model[i] ? 0 : run + 1; if (run * unit >= want) { expect = (i + 1 - run) * unit; break; } }
            if (expect != w2r::SlabFreeList::NONE) {                         // (lowest run that fits; its start, not the first fitting position inside it)
                size_t st = expect / unit; while (st > 0 && !model[st - 1]) --st; expect = st * unit;
            }
            const size_t got = fl.take(want);
            if (got != expect) return 10;
            if (got != w2r::SlabFreeList::NONE) { for (size_t i = got / unit; i < (got + want) / unit; ++i) { if (model[i]) return 11; model[i] = 1; } live.push_back({got, want}); }
            else if (fl.free_tail(model.size() * unit) >= want) return 12;
        } else if (!live.empty()) {                                          // give back
            const size_t k = rnd() % live.size();
            if (fl.give_back(live[k].first) != live[k].second) return 20;
            if (fl.give_back(live[k].first) != 0) return 21;                 // twice: refused
            for (size_t i = live[k].first / unit; i < (live[k].first + live[k].second) / unit; ++i) model[i] = 0;
            live[k] = live.back(); live.pop_back();
        }
        if (int c = check()) return 100 + c;
    }
    // everything back: one free range covering the whole slab
    for (auto& pc : live) if (fl.give_back(pc.first) != pc.second) return 30;
    if (!(fl.used == 0 && fl.free_.size() == (model.empty() ? 0u : 1u) && (model.empty() || (fl.free_.begin()->first == 0 && fl.free_.begin()->second == model.size() * unit)))) return 31;
    return 0;
}


// ---- step-3 places (csrc/places.cuh) through the device functions, in the order pipeline.cu: places_stage()/unique_places() runs them:
// measure + fill, hash, sort by hash, neighbour comparison (collisions counted), representatives, stable LSD sort over element
// positions.  `g` is a finished graph with paths (oracle or product); the result goes to malloc'ed arrays the caller frees with hc_free.
extern "C" int hc_places(const w2rap_graph* g, uint32_t K2, uint64_t salt, uint64_t* n_kept, uint64_t* n_places, uint64_t** place_off, int32_t** place_edges,
                         uint64_t* n_collisions) {
    using namespace w2r;
    const uint64_t n = g->n_paths, nh = g->n_hbv_edges;
    std::vector<uint32_t> hcanon(nh ? nh : 1);
    for (uint64_t i = 0; i < g->n_edges; ++i) { hcanon[g->fwd_xlat[i]] = (uint32_t)(i << 1); if (g->rev_xlat[i] != g->fwd_xlat[i]) hcanon[g->rev_xlat[i]] = (uint32_t)(i << 1) | 1u; }   // (k_hbv_edges' encoding)
    // measure + fill
    std::vector<uint64_t> off(1, 0);
    std::vector<int32_t> edges;
    for (uint64_t r = 0; r < n; ++r) {
        const int32_t* x = g->path_edges + g->path_off[r];
        const uint64_t len = g->path_off[r + 1] - g->path_off[r];
        bool flip;
        const uint32_t pl = place_measure(x, len, hcanon.data(), g->edge_len, g->involution, K2, &flip);
        if (!pl) continue;
        for (uint32_t j = 0; j < pl; ++j) edges.push_back(place_element(x, len, g->involution, flip, j));
        off.push_back(edges.size());
    }
    const uint64_t M = off.size() - 1;
    *n_kept = M;
    edges.push_back(0);
    PlacesView s{off.data(), edges.data(), M};
    std::vector<uint64_t> h(M);
    for (uint64_t i = 0; i < M; ++i) h[i] = place_hash(s.edges + s.off[i], s.off[i + 1] - s.off[i], salt);
    std::vector<uint32_t> perm(M);
    for (uint64_t i = 0; i < M; ++i) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return h[a] < h[b]; });
    std::vector<uint32_t> rep;
    uint64_t coll = 0; uint32_t maxlen = 0;
    for (uint64_t i = 0; i < M; ++i) {
        bool first = true;
        if (i) { if (place_equal(s, perm[i], perm[i - 1])) first = false; else if (h[perm[i]] == h[perm[i - 1]]) ++coll; }
        if (first) { rep.push_back(perm[i]); maxlen = std::max<uint32_t>(maxlen, (uint32_t)(s.off[perm[i] + 1] - s.off[perm[i]])); }
    }
    *n_collisions = coll;
    const uint64_t U = rep.size();
    std::vector<uint32_t> order(U);
    for (uint64_t u = 0; u < U; ++u) order[u] = (uint32_t)u;
    for (uint32_t pos = maxlen; pos-- > 0;)
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return place_key(s, rep[a], pos) < place_key(s, rep[b], pos); });
    *n_places = U;
    uint64_t ne = 0;
    for (uint64_t u = 0; u < U; ++u) ne += s.off[rep[u] + 1] - s.off[rep[u]];
    *place_off = (uint64_t*)malloc(sizeof(uint64_t) * (U + 1));
    *place_edges = (int32_t*)malloc(sizeof(int32_t) * (ne + 1));
    uint64_t at = 0;
    for (uint64_t i = 0; i < U; ++i) {
        const uint32_t a = rep[order[i]];
        (*place_off)[i] = at;
        for (uint64_t j = s.off[a]; j < s.off[a + 1]; ++j) (*place_edges)[at++] = s.edges[j];
    }
    (*place_off)[U] = at;
    return 0;
}
