// shardgraph_kernels.cuh — the kernels of the sharded graph stage: thin loops around the host/device functions of shardgraph.cuh
// (which tests/hostcheck drives with simulated ranks).  Orchestration and the NCCL exchanges: pipeline.cu, graph_stage_sharded().
#pragma once
#include "kernels.cuh"
#include "shardgraph.cuh"

namespace w2r {

// ---- round 1: neighbour queries
struct QueryOut {
    ulonglong2* keys;                 // [sum of caps] canonical k-mers, destination d at keys + base[d]
    const uint64_t* base;             // [world]
    unsigned long long* count;        // [world] queries for destination d (keeps counting past cap: the exact need)
    uint64_t cap;                     // per destination
};
// over the solid records {w0, w1 | raw ctx} this rank owns.  All lanes step through the 8 context bits together; per step the
// lanes with a query for the same destination share one cursor atomic (there are only W cursors).
__global__ void __launch_bounds__(256) k_neighbour_queries(const ulonglong2* __restrict__ recs, uint64_t n, uint32_t logP, uint32_t world, uint32_t me, QueryOut q) {
    const uint32_t lane = lane_id();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const ulonglong2 r = i < n ? __ldcs(recs + i) : make_ulonglong2(0, 0);
        const NeighbourScan ns(Kmer{r.x, r.y & ~0xffull}, i < n ? (uint32_t)r.y & 0xffu : 0u, logP, world, me);
        if (!__any_sync(0xffffffffu, ns.c != 0)) continue;
        for (uint32_t b = 0; b < 8; ++b) {
            uint32_t o = 0; Kmer cn{0, 0};
            const bool want = ns.remote(b, &o, &cn);
            if (!__any_sync(0xffffffffu, want)) continue;
            const uint32_t key = want ? o : world + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs((int)peers) - 1;
            unsigned long long pos0 = 0;
            if (want && (int)lane == leader) pos0 = atomicAdd(q.count + o, (unsigned long long)__popc(peers));
            pos0 = __shfl_sync(0xffffffffu, pos0, leader);
            const unsigned long long pos = pos0 + (unsigned)__popc(peers & ((1u << lane) - 1u));
            if (want && pos < q.cap) q.keys[q.base[o] + pos] = make_ulonglong2(cn.w0, cn.w1);
        }
    }
}
// owner side: slot of every asked k-mer, or NIL.  Runs before any ghost is inserted: the table holds owned entries only.
__global__ void k_answer_queries(SolidTable st, const ulonglong2* __restrict__ keys, uint64_t n, uint32_t* __restrict__ reply) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 k = keys[i];
        const int64_t s = solid_find(st, Kmer{k.x, k.y});
        reply[i] = s < 0 ? NIL : (uint32_t)s;
    }
}
// asker side: neighbours that exist on `owner` become ghost entries (find-or-insert: several local k-mers may ask for the same one).
// gslot[i] = local slot of query i's ghost (NIL if the k-mer does not exist).
__global__ void k_insert_ghosts(SolidTable st, const ulonglong2* __restrict__ keys, const uint32_t* __restrict__ reply, uint64_t n, uint32_t owner, uint32_t* __restrict__ gslot) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = reply[i];
        if (r == NIL) { gslot[i] = NIL; continue; }
        const ulonglong2 k = keys[i];
        uint64_t h = st.home(Kmer{k.x, k.y});
        for (;;) {
            SolidSlot* p = st.slots + h;
            const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2*>(p));
            if (cur.x == k.x && cur.y == k.y) break;
            if (cur.x == EMPTY_W0) {
                const U128 old = cas128(p, ~0ull, ~0ull, k.x, k.y);
                if (old.lo == ~0ull && old.hi == ~0ull) { p->ctx = 0; p->edge = r; p->off = 0; p->pad = owner + 1u; break; }
                if (old.lo == k.x && old.hi == k.y) break;
            }
            h = st.next(h);
        }
        gslot[i] = (uint32_t)h;
    }
}
// second round, owner side: the PRUNED context of the entries that were asked for (after k_adjacency)
__global__ void k_gather_ctx(SolidTable st, const uint32_t* __restrict__ slot, uint64_t n, uint32_t* __restrict__ ctx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) ctx[i] = slot[i] == NIL ? 0u : (st.slots[slot[i]].ctx & 0xffu);
}
__global__ void k_apply_ghost_ctx(SolidTable st, const uint32_t* __restrict__ gslot, const uint32_t* __restrict__ ctx, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) if (gslot[i] != NIL) st.slots[gslot[i]].ctx = ctx[i];
}

// ---- links with ghost-predecessor flags
__global__ void k_links_sharded(SolidTable st, uint32_t* __restrict__ next0, uint8_t* __restrict__ ghead, int* __restrict__ missing) {
    const uint64_t nn = 2 * st.size();
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < nn; x += (uint64_t)gridDim.x * blockDim.x) {
        int miss = 0; bool tg = false;
        next0[x] = unipath_succ_link(st, (uint32_t)x, &miss, &tg);
        if (tg) ghead[x ^ 1u] = 1;          // the flip of x has its predecessor on another rank: it heads a local piece
        if (miss) atomicExch(missing, 1);
    }
}

// ---- round 2: one record per local chain
__global__ void k_emit_pieces(SolidTable st, const uint32_t* __restrict__ next0, const uint8_t* __restrict__ ghead, const RankState* __restrict__ R, uint32_t me,
                              PieceRec* __restrict__ out, uint64_t cap, unsigned long long* cursor, uint32_t* __restrict__ lpiece, uint32_t* __restrict__ lhead) {
    const uint64_t nn = 2 * st.size();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < nn; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = base + threadIdx.x;
        const bool want = x < nn && node_is_piece_head(next0, ghead, (uint32_t)x);
        const uint64_t pos = warp_append(cursor, want);
        if (want && pos < cap) {
            const PieceRec p = piece_of_head(st, next0, R, me, (uint32_t)x);
            out[pos] = p;
            lpiece[p.flip_local] = (uint32_t)pos;         // (flip_local still holds the tail node)
            lhead[x] = (uint32_t)pos;
        }
    }
}
// the flipped piece of a piece is the one headed by the flip of its tail: a local look-up (both live on the same rank)
__global__ void k_piece_flips(PieceRec* __restrict__ rec, uint64_t npl, const uint32_t* __restrict__ lhead) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < npl; j += (uint64_t)gridDim.x * blockDim.x) rec[j].flip_local = lhead[rec[j].flip_local ^ 1u];
}
__global__ void k_count_piece_heads(const uint32_t* __restrict__ next0, const uint8_t* __restrict__ ghead, uint64_t nn, unsigned long long* __restrict__ count) {
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < nn; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = base + threadIdx.x;
        const unsigned m = __ballot_sync(__activemask(), x < nn && node_is_piece_head(next0, ghead, (uint32_t)x));
        if (lane_id() == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
    }
}
__global__ void k_gidmap_insert(const PieceRec* __restrict__ rec, uint64_t np, GidMap m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = rec[i].head;
        for (uint64_t h = gid_hash(g) & m.mask;; h = (h + 1) & m.mask) {
            if (atomicCAS((unsigned long long*)m.keys + h, (unsigned long long)GID_NONE, (unsigned long long)g) == (unsigned long long)GID_NONE) { m.vals[h] = (uint32_t)i; break; }
        }
    }
}
// piece0[r] = global index of rank r's first piece
__global__ void k_piece_link(const PieceRec* __restrict__ rec, uint64_t np, GidMap m, const uint64_t* __restrict__ piece0, uint32_t* __restrict__ nxt, uint32_t* __restrict__ flip,
                             int* __restrict__ bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        const PieceRec p = rec[i];
        uint32_t nx = NIL;
        if (p.succ != GID_NONE) { nx = gid_find(m, p.succ); if (nx == NIL) atomicExch(bad, 1); }
        nxt[i] = nx;
        flip[i] = (uint32_t)piece0[p.head >> 32] + p.flip_local;
    }
}
__global__ void k_piece_rank_init(const PieceRec* __restrict__ rec, const uint32_t* __restrict__ nxt, uint64_t np, RankState* __restrict__ S) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) S[i] = piece_rank_init(rec, nxt, (uint32_t)i);
}
// in-place pointer jumping (8-byte states read and written whole: any mix of old and new states composes to a valid state)
__global__ void k_piece_rank_step(unsigned long long* S, uint64_t np, unsigned long long* __restrict__ unresolved) {
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < np; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        bool un = false;
        if (i < np) {
            const unsigned long long a = __ldcg(S + i);
            RankState ra{(uint32_t)a, (uint32_t)(a >> 32)};
            if (!(ra.y & RANK_RESOLVED)) {
                const unsigned long long b = __ldcg(S + ra.x);
                const RankState rn = piece_rank_step(ra, RankState{(uint32_t)b, (uint32_t)(b >> 32)});
                __stcg(S + i, (unsigned long long)rn.x | ((unsigned long long)rn.y << 32));
                un = !(rn.y & RANK_RESOLVED);
            }
        }
        const unsigned m = __ballot_sync(__activemask(), un);
        if (lane_id() == 0 && m) atomicAdd(unresolved, (unsigned long long)__popc(m));
    }
}

// ---- circles that span ranks: the nodes of unresolved pieces, for the host to find the minimum k-mer of every circle
struct CycleNode { uint32_t piece, pad; uint64_t w0, w1, gid; };
__global__ void k_collect_cycle_nodes(SolidTable st, const uint32_t* __restrict__ next0, const RankState* __restrict__ R, const uint32_t* __restrict__ lpiece, uint32_t piece0,
                                      const RankState* __restrict__ S, uint32_t me, CycleNode* __restrict__ out, uint64_t cap, unsigned long long* cursor) {
    const uint64_t nn = 2 * st.size();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < nn; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = base + threadIdx.x;
        bool want = false;
        uint32_t pi = 0;
        if (x < nn && next0[x] != EMPTY_NODE && next0[x] != GHOST_TAIL) { pi = piece0 + lpiece[R[x].x]; want = !(S[pi].y & RANK_RESOLVED); }
        const uint64_t pos = warp_append(cursor, want);
        if (want && pos < cap) { const SolidSlot& sl = st.slots[x >> 1]; out[pos] = CycleNode{pi, 0u, sl.w0, sl.w1, gid_make(me, (uint32_t)x)}; }
    }
}
__device__ __forceinline__ bool gid_in_sorted(const uint64_t* a, uint32_t n, uint64_t g) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < g) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == g;
}
// cut every circle at its minimum k-mer (BuildReadQGraph.cc:156-180): (kmin,-) ends its strand, the predecessor of (kmin,+) ends its strand
__global__ void k_apply_cycle_cuts(SolidTable st, uint32_t* __restrict__ next0, uint8_t* __restrict__ ghead, uint32_t me, const CycleNode* __restrict__ mine, uint64_t n_mine,
                                   const uint64_t* __restrict__ heads_g, uint32_t n_heads, const uint64_t* __restrict__ tails_g, uint32_t n_tails) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mine; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = mine[i].gid;
        const uint32_t x = (uint32_t)g;
        if (gid_in_sorted(heads_g, n_heads, g)) ghead[x] = 0;
        const uint32_t nx = next0[x];
        bool cut = gid_in_sorted(tails_g, n_tails, g);
        if (!cut && nx < GHOST_TAIL) {
            const SolidSlot& sl = st.slots[nx >> 1];
            const uint64_t ng = slot_is_ghost(sl) ? gid_make(sl.pad - 1u, 2u * sl.edge + (nx & 1u)) : gid_make(me, nx);
            cut = gid_in_sorted(heads_g, n_heads, ng);
        }
        if (cut) next0[x] = NIL;
    }
}

// ---- strands and edges
// per chain (at its last piece): first piece, length, even-length orientation
__global__ void k_chain_tails(PieceView pv, const uint32_t* __restrict__ nxt, uint8_t* __restrict__ is_head, uint32_t* __restrict__ chain_n, uint8_t* __restrict__ keepp, int* __restrict__ too_long) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pv.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (nxt[i] != NIL) continue;
        uint32_t hp; uint64_t n;
        chain_of_tail_piece(pv, (uint32_t)i, &hp, &n);
        if (n > 0x1000000ull) { atomicExch(too_long, 1); n = 0x1000000ull; }      // offsets must fit KDef's 24 bits (kmers/ReadPather.h:121-122,144)
        is_head[hp] = 1; chain_n[hp] = (uint32_t)n;
        const uint32_t kf = chain_keep_even(pv, (uint32_t)i, hp, n);
        if (kf != 2u) keepp[hp] = (uint8_t)kf;
    }
}
__global__ void k_piece_info(PieceView pv, PieceInfo* __restrict__ pinfo) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pv.n; i += (uint64_t)gridDim.x * blockDim.x) pinfo[i] = piece_info(pv, (uint32_t)i);
}
// every rank over ITS pieces [piece0, piece0 + npl): the one that holds the middle k-mer of an odd-length chain decides its strand
__global__ void k_piece_keep_odd(SolidTable st, const uint32_t* __restrict__ next0, const PieceRec* __restrict__ rec, const PieceInfo* __restrict__ pinfo, uint32_t piece0, uint64_t npl,
                                 uint8_t* __restrict__ keepp) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < npl; j += (uint64_t)gridDim.x * blockDim.x) {
        const PieceInfo pi = pinfo[piece0 + j];
        const int kf = piece_keep_odd(st, next0, pi, (uint32_t)rec[piece0 + j].head);
        if (kf >= 0) keepp[pi.head_piece] = (uint8_t)kf;
    }
}
__global__ void k_piece_edges(PieceInfo* __restrict__ pinfo, uint64_t np, const uint32_t* __restrict__ edge_of_piece) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) pinfo[i].head_piece = edge_of_piece[pinfo[i].head_piece];
}
__global__ void k_collect_head_pieces(PieceView pv, const uint8_t* __restrict__ is_head, const uint8_t* __restrict__ keepp, const uint32_t* __restrict__ chain_n, uint32_t* __restrict__ h_piece,
                                      uint64_t* __restrict__ h_w0, uint64_t* __restrict__ h_w1, uint32_t* __restrict__ h_n, unsigned long long* cursor, uint64_t cap) {
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < pv.n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const bool want = i < pv.n && is_head[i] && keepp[i];
        const uint64_t pos = warp_append(cursor, want);
        if (want && pos < cap) { h_piece[pos] = (uint32_t)i; h_w0[pos] = pv.rec[i].head_k.w0; h_w1[pos] = pv.rec[i].head_k.w1; h_n[pos] = chain_n[i]; }
    }
}
__global__ void k_emit_edges_sharded(SolidTable st, const uint32_t* __restrict__ next0, const RankState* __restrict__ R, const uint32_t* __restrict__ lpiece,
                                     const PieceInfo* __restrict__ my_pinfo, const uint64_t* __restrict__ edge_off, uint8_t* __restrict__ edge_bases) {
    const uint64_t nn = 2 * st.size();
    OrBase put{edge_bases};
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < nn; x += (uint64_t)gridDim.x * blockDim.x) {
        if (next0[x] == EMPTY_NODE || next0[x] == GHOST_TAIL) continue;
        emit_node_sharded(st, R, lpiece, my_pinfo, edge_off, (uint32_t)x, put);
    }
}
// owned entries (pruned context, edge, offset) out of the local table, bucketed by the dictionary slice (= GPU) that will hold them
// for pathing: slice = pd_slice_of(W, bloom_hash(k-mer)).  count[] keeps counting past cap: the exact need.
__global__ void k_dump_owned_sliced(SolidTable st, uint32_t W, SolidSlot* __restrict__ out, uint64_t cap, unsigned long long* __restrict__ count) {
    const uint64_t T = st.size();
    const uint32_t lane = lane_id();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < T; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        SolidSlot s;
        bool act = false;
        if (i < T) { s = st.slots[i]; act = s.w0 != EMPTY_W0 && !slot_is_ghost(s); }
        // one cursor atomic per destination and warp (W counters only: per-entry atomics would serialise on them)
        const uint32_t d = act ? pd_slice_of(W, bloom_hash(Kmer{s.w0, s.w1})) : W + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs((int)peers) - 1;
        unsigned long long pos0 = 0;
        if (act && (int)lane == leader) pos0 = atomicAdd(count + d, (unsigned long long)__popc(peers));
        pos0 = __shfl_sync(0xffffffffu, pos0, leader);
        const unsigned long long pos = pos0 + (unsigned)__popc(peers & ((1u << lane) - 1u));
        if (act && pos < cap) out[(uint64_t)d * cap + pos] = s;
    }
}

}  // namespace w2r
