// count_part.cuh — k-mer counting as map (partition) + reduce (count per partition).
//
// A global hash table >> L2 costs one random DRAM sector read and one write-back per k-mer INSTANCE (measured: 13 G inserts/s
// on B200, 0.84 TB/s of 32-byte sector traffic).  Instead, the MapReduceEngine shape (MapReduceEngine.h:288-358) on the device:
//   map     k_minimizer_map     : every instance becomes a 16-byte record {w0, w1 | ctx} in the record area of its partition.
//                                 Partitions are keyed by MINIMISER, so a read appends runs of ~24 records; a counting launch
//                                 sizes every partition exactly, a scan lays them out, a second launch stores (no capacity guess).
//   reduce  k_count_smem        : one CTA counts one partition (~20 k records, a few thousand distinct k-mers) in a shared-memory
//                                 hash table, then emits histogram + solid records.
//           k_count_region /    : fallback for partitions that do not fit shared memory, and the legacy path behind the
//           k_scan_region         table_slots test hook: inserts into a 32 MB table region that stays resident in L2
//                                 (128-bit CAS, count/context REDs), then scan + reset.
//   legacy map k_extract_partition : one thread per read, partition = top bits of the k-mer hash, static sub-buffers.
// DRAM traffic: 16 B written + 16 B read per instance — the 34 B/instance of the SURVEY §8d model — and nothing else.
#pragma once
#include "kernels.cuh"
#include "shard.cuh"

namespace w2r {

// ---------------------------------------------------------------- legacy map: hash partitions in static sub-buffers
struct PartParams {
    ulonglong2* recs;        // [P * nsub][cap]
    uint32_t* cursor;        // [P * nsub] * cstride: records appended per sub-buffer, one cursor per L2 line
    uint64_t cap;            // capacity of one sub-buffer (records)
    uint32_t logP;           // partitions = 1 << logP
    uint32_t nsub;           // sub-buffers per partition (power of two): spreads the cursor atomics over nsub x more L2 lines
    uint32_t cstride;        // cursor stride in u32 (32 = one cursor per 128-byte line)
    uint32_t npass, pass;    // outer hash-range passes (when the records of everything would not fit): keep (hash & 0xffff) % npass == pass
    int* overflow;           // set if a sub-buffer overflowed
};

// Appends records in PAIRS: the two cursor atomics are independent, so both are in flight together and the thread waits
// for one round trip per two k-mers (the kernel is bound by the latency of the returning atomic x threads in flight).
struct PartEmit {
    const PartParams& pp;
    ulonglong2 rec;          // the stashed first record of a pair
    uint32_t bucket;
    bool pending;
    uint32_t sub;
    __device__ __forceinline__ PartEmit(const PartParams& p, uint32_t sub_) : pp(p), rec(make_ulonglong2(0, 0)), bucket(0), pending(false), sub(sub_) {}
    __device__ __forceinline__ void put(uint32_t b, uint32_t pos, ulonglong2 r) {
        if (pos < pp.cap) pp.recs[(uint64_t)b * pp.cap + pos] = r;
        else atomicExch(pp.overflow, 1);
    }
    __device__ __forceinline__ void flush() {
        if (pending) { put(bucket, atomicAdd(pp.cursor + (uint64_t)bucket * pp.cstride, 1u), rec); pending = false; }
    }
    __device__ __forceinline__ void operator()(Kmer k, uint32_t ctx) {
        const uint64_t h = kmer_hash(k);
        if (pp.npass > 1 && (uint32_t)(h & 0xffffu) % pp.npass != pp.pass) return;
        const uint32_t b = part_of_hash(h, pp.logP) * pp.nsub + sub;
        const ulonglong2 r = make_ulonglong2(k.w0, k.w1 | ctx);
        if (!pending) { rec = r; bucket = b; pending = true; return; }
        const uint32_t pos0 = atomicAdd(pp.cursor + (uint64_t)bucket * pp.cstride, 1u);
        const uint32_t pos1 = atomicAdd(pp.cursor + (uint64_t)b * pp.cstride, 1u);
        put(bucket, pos0, rec);
        put(b, pos1, r);
        pending = false;
    }
};

// paths/long/BuildReadQGraph.cc:1062-1080 (the "map" step): one thread per read.
__global__ void __launch_bounds__(256, 6) k_extract_partition(ReadsView r, uint64_t first, uint64_t count, const uint16_t* __restrict__ good, PartParams pp) {
    const uint32_t sub = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (pp.nsub - 1);   // per warp
    const uint64_t end = first + count;
    for (uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t gl = good[i];
        if (gl > (uint32_t)K) {
            PartEmit emit(pp, sub);
            extract_read_kmers(r.bases + r.base_off[i], gl, emit);
            emit.flush();
        }
    }
}

// ---------------------------------------------------------------- map keyed by minimiser
// The "map" step: one WARP per read, partitions keyed by minimiser (extract.cuh).  Lane l of step t handles k-mer
// 32t + l: k-mer, context and window minimum are all computed independently per position (no rolling state), lanes whose
// neighbours fall into the same partition form a segment, the segment head reserves the whole segment with one cursor atomic
// and the lanes store their records side by side.  Window minima: the hashes of all m-mers of a tile live in registers (8 per
// lane), five doubling steps of shuffles give the min over 32 consecutive hashes, the 46-wide window is two overlapping 32-wide ones.
// Two launches: COUNT_ONLY sizes every partition exactly (count[p] += segment lengths; no k-mers are formed), an exclusive
// scan turns the counts into partition bases, and the second launch stores.  Exact sizes mean no capacity guess can overflow,
// whatever the multiplicity skew of the read set (a repeat with 10^4 copies just makes its partitions long).
constexpr uint32_t MINI_TILE = 192;                       // k-mers per tile (a 250-base read is one tile)
constexpr uint32_t MINI_NU = 8;                           // hashes per lane: positions 32u + lane, u < 8, cover the 192 + 45 m-mers of a tile
struct MiniParams {
    uint32_t logP, npass, pass;
    uint32_t* count;                 // COUNT_ONLY: records per partition
    const uint64_t* base;            // store: first record of every partition, relative to *batch_off
    uint32_t* cursor;                // store: records appended so far
    ulonglong2* recs;
    const unsigned long long* batch_off;   // store: where this read batch's records start (device scalar: no host round trip)
    uint64_t recs_cap;               // records the buffer holds (it is sized from an upper bound; checked all the same)
    int* overflow;
};
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256, 6) k_minimizer_map(ReadsView r, uint64_t first, uint64_t count, const uint16_t* __restrict__ good, MiniParams mp) {
    __shared__ uint32_t wbuf[8][MINI_TILE];
    uint32_t* wb = wbuf[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned upto = (2u << lane) - 1u;                  // lanes 0..lane
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5, end = first + count;
    const unsigned long long boff = COUNT_ONLY ? 0ull : *mp.batch_off;
    for (uint64_t i = first + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < end; i += nwarps) {
        const uint32_t gl = good[i];
        if (gl <= (uint32_t)K) continue;
        const uint8_t* bases = r.bases + r.base_off[i];
        const uint32_t nk = gl - K + 1, last = gl - K;
        for (uint32_t j0 = 0; j0 < nk; j0 += MINI_TILE) {
            const uint32_t n_k = nk - j0 < MINI_TILE ? nk - j0 : MINI_TILE, n_h = n_k + MINI_W - 1;
            // window minima in registers: h[u] = hash of the m-mer at tile position 32u + lane; a doubling step with shift s takes
            // position p + s from lane (lane + s) & 31 of the same or the next register
            uint32_t h[MINI_NU + 1];
#pragma unroll
            for (uint32_t u = 0; u < MINI_NU; ++u) { const uint32_t t = 32u * u + lane; h[u] = t < n_h ? mmer_hash_at(bases, (uint64_t)j0 + t) : 0xffffffffu; }
            h[MINI_NU] = 0xffffffffu;
#pragma unroll
            for (uint32_t s = 1; s <= 16; s <<= 1) {                      // after these: h = min over 32 consecutive positions
                const uint32_t src = (lane + s) & 31u;
                const bool wrap = lane + s >= 32u;
                uint32_t x = __shfl_sync(0xffffffffu, h[0], src);
#pragma unroll
                for (uint32_t u = 0; u < MINI_NU; ++u) {
                    const uint32_t y = u + 1 < MINI_NU ? __shfl_sync(0xffffffffu, h[u + 1], src) : 0xffffffffu;
                    const uint32_t o = wrap ? y : x;
                    h[u] = h[u] < o ? h[u] : o;
                    x = y;
                }
            }
            {                                                             // 46-wide window = two 32-wide ones, 14 apart
                const uint32_t src = (lane + (MINI_W - 32)) & 31u;
                const bool wrap = lane + (MINI_W - 32) >= 32u;
                uint32_t x = __shfl_sync(0xffffffffu, h[0], src);
                __syncwarp();
#pragma unroll
                for (uint32_t u = 0; u < MINI_TILE / 32; ++u) {
                    const uint32_t y = __shfl_sync(0xffffffffu, h[u + 1], src);
                    const uint32_t o = wrap ? y : x;
                    wb[32u * u + lane] = h[u] < o ? h[u] : o;
                    x = y;
                }
                __syncwarp();
            }
            for (uint32_t t0 = 0; t0 < n_k; t0 += 32) {
                const uint32_t jj = t0 + lane, j = j0 + jj;
                bool active = jj < n_k;
                uint32_t b = 0;
                if (active) {
                    const uint32_t mh = mini_mix(wb[jj]);
                    if (mp.npass > 1 && (mh & 0xffffu) % mp.npass != mp.pass) active = false;
                    b = mini_part(mh, mp.logP);
                }
                const unsigned amask = __ballot_sync(0xffffffffu, active);
                const uint32_t pb = __shfl_up_sync(0xffffffffu, b, 1);
                const bool head = active && (lane == 0 || !((amask >> (lane - 1)) & 1u) || pb != b);
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                unsigned long long pos0 = 0;
                if (head) {
                    const unsigned stop = (heads | ~amask) & ~upto;       // next segment head or first idle lane above this one
                    const uint32_t seglen = (stop ? (uint32_t)__ffs((int)stop) - 1u : 32u) - lane;
                    if (COUNT_ONLY) atomicAdd(mp.count + b, seglen);
                    else pos0 = boff + mp.base[b] + atomicAdd(mp.cursor + b, seglen);
                }
                if (COUNT_ONLY) continue;
                ulonglong2 rec = make_ulonglong2(0, 0);
                if (active) {                                             // (independent of the atomic: overlaps its round trip)
                    Kmer f, rc;
                    kmer_pair_at(bases, j, &f, &rc);
                    uint32_t c = 0;
                    if (j < last) c |= 1u << packed_base(bases, (uint64_t)j + K);
                    if (j > 0) c |= 16u << packed_base(bases, (uint64_t)j - 1);
                    const bool rev = kmer_less(rc, f);
                    rec = make_ulonglong2(rev ? rc.w0 : f.w0, (rev ? rc.w1 : f.w1) | (rev ? ctx_rc(c) : c));
                }
                const uint32_t hl = active ? 31u - (uint32_t)__clz((int)(heads & upto)) : lane;
                const unsigned long long pos = __shfl_sync(0xffffffffu, pos0, hl) + (lane - hl);
                if (active) { if (pos < mp.recs_cap) mp.recs[pos] = rec; else atomicExch(mp.overflow, 1); }
            }
        }
    }
}

// batch_off[1] = batch_off[0] + total[0]: chains the record areas of consecutive read batches on the device
__global__ void k_next_batch_off(unsigned long long* batch_off, const uint64_t* total) { batch_off[1] = batch_off[0] + *total; }

struct RegionParams {
    CountSlot* region;       // 1 << logR slots, resident in L2
    uint32_t logR, logP;
    uint32_t sub_mask, sub_id;   // overflow handling: only records with ((hash >> 3) & sub_mask) == sub_id take part
    int* overflow;
};
constexpr uint32_t REGION_MAX_PROBE = 2048;
constexpr uint32_t COUNT_STOP_REGION = 1u << 24;               // the u32 counter of a region slot stops here (it cannot wrap)

// one 256-bit load of a whole slot (LDG.E.ENL2.256 on sm_100a), L2-coherent
__device__ __forceinline__ void ld_slot(const CountSlot* q, uint64_t& w0, uint64_t& w1, uint64_t& meta) {
    uint64_t pad;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(meta), "=l"(pad) : "l"(q));
}

__device__ __forceinline__ void region_insert(const RegionParams& rp, uint64_t mask, ulonglong2 rec, uint64_t h) {
    const uint64_t kw0 = rec.x, kw1 = rec.y & ~0xffull;
    const uint32_t ctx = (uint32_t)rec.y & 0xffu;
    uint64_t s = region_slot_of_hash(h, rp.logP, rp.logR);
    for (uint32_t probe = 0; probe < REGION_MAX_PROBE; ++probe) {
        CountSlot* q = rp.region + s;
        uint64_t w0, w1, meta;
        ld_slot(q, w0, w1, meta);
        bool hit = false;
        uint32_t have = (uint32_t)(meta >> 32);
        if (w0 == kw0 && w1 == kw1) hit = true;
        else if (w0 == EMPTY_W0) {
            U128 old = cas128(q, ~0ull, ~0ull, kw0, kw1);
            hit = (old.lo == ~0ull && old.hi == ~0ull) || (old.lo == kw0 && old.hi == kw1);
            have = 0;
        }
        if (hit) {
            if ((uint32_t)meta < COUNT_STOP_REGION || w0 == EMPTY_W0) atomicAdd(&q->count, 1u);
            if ((have & ctx) != ctx) atomicOr(&q->ctx, ctx);
            return;
        }
        s = (s + 1) & mask;
    }
    atomicExch(rp.overflow, 1);
}

// Two records at a time: their slot loads are independent, which doubles the L2 requests in flight per thread.
__device__ __forceinline__ void region_count_pair(const RegionParams& rp, uint64_t mask, ulonglong2 ra, ulonglong2 rb, bool two) {
    const uint64_t ha = kmer_hash(Kmer{ra.x, ra.y & ~0xffull});
    const uint64_t hb = kmer_hash(Kmer{rb.x, rb.y & ~0xffull});
    const bool da = (!rp.sub_mask || (((uint32_t)(ha >> 3)) & rp.sub_mask) == rp.sub_id);
    const bool db = two && (!rp.sub_mask || (((uint32_t)(hb >> 3)) & rp.sub_mask) == rp.sub_id);
    CountSlot* qa = rp.region + region_slot_of_hash(ha, rp.logP, rp.logR);
    CountSlot* qb = rp.region + region_slot_of_hash(hb, rp.logP, rp.logR);
    uint64_t a0 = 0, a1 = 0, am = 0, b0 = 0, b1 = 0, bm = 0;
    if (da) ld_slot(qa, a0, a1, am);
    if (db) ld_slot(qb, b0, b1, bm);
    if (da) {
        const uint32_t ctx = (uint32_t)ra.y & 0xffu;
        if (a0 == ra.x && a1 == (ra.y & ~0xffull)) { if ((uint32_t)am < COUNT_STOP_REGION) atomicAdd(&qa->count, 1u); if ((((uint32_t)(am >> 32)) & ctx) != ctx) atomicOr(&qa->ctx, ctx); }
        else region_insert(rp, mask, ra, ha);
    }
    if (db) {
        const uint32_t ctx = (uint32_t)rb.y & 0xffu;
        if (b0 == rb.x && b1 == (rb.y & ~0xffull)) { if ((uint32_t)bm < COUNT_STOP_REGION) atomicAdd(&qb->count, 1u); if ((((uint32_t)(bm >> 32)) & ctx) != ctx) atomicOr(&qb->ctx, ctx); }
        else region_insert(rp, mask, rb, hb);
    }
}

// The "reduce" step (BuildReadQGraph.cc:1081-1082 sort+collapse as a hash count).  blockIdx.y selects the sub-buffer of the group.
// recs/sizes hold one slab per source ([n_src][owned sub-buffers]); blockIdx.y = src * gy + sub-buffer within the group.
// Legacy layout (rv.part_base == nullptr): static sub-buffers of `cap` records, slab src at recs + src * slab_recs.
// Minimiser layout: slab src (a read batch, or a source rank) starts at rv.slab_off[src]; partition b of it is the run
// recs[slab_off[src] + part_base[src * P + b] ...) of sizes[src * slab_cur + b] records.
// plist (optional): the group is partitions plist[b_first .. b_first + gy) instead of the consecutive range starting at b_first.
struct RunView { const uint64_t* part_base; const unsigned long long* slab_off; uint64_t P; const uint32_t* plist; };
__global__ void __launch_bounds__(256) k_count_region(const ulonglong2* __restrict__ recs, const uint32_t* __restrict__ sizes, uint32_t cstride, uint64_t cap,
                                                      uint32_t b_first, uint32_t gy, uint64_t slab_recs, uint64_t slab_cur, RunView rv, RegionParams rp) {
    const uint32_t src = blockIdx.y / gy;
    const uint32_t bi = b_first + (blockIdx.y - src * gy);
    const uint32_t b = rv.plist ? rv.plist[bi] : bi;
    uint64_t n = sizes[(uint64_t)src * slab_cur + (uint64_t)b * cstride];
    if (n > cap) n = cap;
    const uint64_t mask = (1ull << rp.logR) - 1;
    const ulonglong2* base = rv.part_base ? recs + rv.slab_off[src] + rv.part_base[(uint64_t)src * rv.P + b] : recs + (uint64_t)src * slab_recs + (uint64_t)b * cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
        const bool two = i + stride < n;
        const ulonglong2 ra = __ldcs(base + i);
        const ulonglong2 rb = two ? __ldcs(base + i + stride) : make_ulonglong2(0, 0);
        region_count_pair(rp, mask, ra, rb, two);
    }
}

// ---------------------------------------------------------------- reduce in SHARED memory (fine partitions, single GPU)
// With ~2^18 partitions a partition holds ~20 k records and a few thousand distinct k-mers: one CTA counts it in a shared-memory
// table, so the per-record atomics are shared-memory atomics instead of L2 atomics (the L2-resident region count is limited by
// L2 atomic throughput, about one record per 30 ps chip-wide).  A partition is one contiguous run of records per slab (read
// batch on one GPU, source rank when sharded).  Partitions whose distinct k-mers do not fit are listed in `failed`: they are
// retried with a larger table and, failing that, go through the region path.
constexpr uint32_t SMEM_LOG_SLOTS_MAX = 13;                    // 8192 slots: 64 KB w0 + 64 KB w1 + 32 KB count|ctx = 160 KB
constexpr uint32_t SMEM_MAX_PROBE = 512;
constexpr uint32_t COUNT_STOP = 1u << 16;                      // counters stop here (>= 255 is all anyone asks); + one add per racing thread
constexpr uint32_t SMEM_MAX_BATCH = 16;                        // read batches whose runs make up one partition
struct SmemCountParams {
    const ulonglong2* recs;
    const uint32_t* cursor;         // [nbatch][P] records of partition p in slab bi
    const uint64_t* part_base;      // partition p is, per slab bi < nbatch, the contiguous run
    const unsigned long long* batch_off;   //   recs[batch_off[bi] + part_base[bi * P + p] .. + cursor[bi * P + p])
    uint32_t nbatch;
    uint32_t P, logP;
    uint32_t min_freq;
    unsigned long long* hist;       // [104]
    ulonglong2* solid_out; unsigned long long* solid_cursor; uint64_t solid_cap; int* solid_overflow;
    DumpRec* dump_out; unsigned long long* dump_cursor;
    uint32_t* failed; unsigned long long* failed_cursor;       // partitions that did not fit this table size
    uint32_t log_slots;                                        // table size of this launch: 12 (two CTAs of 512 threads per SM) or 13
    const uint32_t* plist; uint32_t nlist;                     // if set: only these partitions (the ones a smaller table could not hold)
};
__global__ void __launch_bounds__(1024, 1) k_count_smem(SmemCountParams sp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t SMEM_SLOTS = 1u << sp.log_slots;
    unsigned long long* w0s = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* w1s = w0s + SMEM_SLOTS;
    uint32_t* ccs = reinterpret_cast<uint32_t*>(w1s + SMEM_SLOTS);          // count (low 24 bits) | ctx << 24
    __shared__ unsigned int sh_hist[104];
    __shared__ int sh_fail;
    __shared__ unsigned long long sh_run[SMEM_MAX_BATCH];
    __shared__ uint32_t sh_pre[SMEM_MAX_BATCH + 1];
    for (int j = threadIdx.x; j < 104; j += blockDim.x) sh_hist[j] = 0;
    const uint32_t n_todo = sp.plist ? sp.nlist : sp.P;
    for (uint32_t pi = blockIdx.x; pi < n_todo; pi += gridDim.x) {
        const uint32_t p = sp.plist ? sp.plist[pi] : pi;
        for (uint32_t j = threadIdx.x; j < SMEM_SLOTS; j += blockDim.x) { w0s[j] = ~0ull; w1s[j] = ~0ull; ccs[j] = 0; }
        if (threadIdx.x == 0) {
            sh_fail = 0;
            uint32_t acc = 0;
            for (uint32_t bi = 0; bi < sp.nbatch; ++bi) {
                sh_run[bi] = sp.batch_off[bi] + sp.part_base[(uint64_t)bi * sp.P + p];
                sh_pre[bi] = acc;
                acc += sp.cursor[(uint64_t)bi * sp.P + p];
            }
            for (uint32_t bi = sp.nbatch; bi <= SMEM_MAX_BATCH; ++bi) sh_pre[bi] = acc;
        }
        __syncthreads();
        const uint32_t n = sh_pre[SMEM_MAX_BATCH];
        auto insert = [&](const ulonglong2 rec) {
            const unsigned long long kw0 = rec.x, kw1 = rec.y & ~0xffull;
            const uint32_t ctx = (uint32_t)rec.y & 0xffu;
            const uint64_t h = kmer_hash(Kmer{kw0, kw1});
            uint32_t s = (uint32_t)((sp.logP ? (h << sp.logP) : h) >> (64 - sp.log_slots));   // the bits below the partition bits
            bool done = false;
            for (uint32_t probe = 0; probe < SMEM_MAX_PROBE && !done; ++probe) {
                unsigned long long k0 = *(volatile unsigned long long*)(w0s + s);
                bool mine = false;
                if (k0 == ~0ull) {
                    const unsigned long long old = atomicCAS(w0s + s, ~0ull, kw0);
                    if (old == ~0ull) { *(volatile unsigned long long*)(w1s + s) = kw1; __threadfence_block(); mine = true; }
                    else k0 = old;
                }
                if (!mine && k0 == kw0) {                     // same first 32 bases: the owner publishes the second word right after its CAS
                    unsigned long long w = *(volatile unsigned long long*)(w1s + s);
                    while (w == ~0ull) w = *(volatile unsigned long long*)(w1s + s);
                    mine = (w == kw1);
                }
                if (mine) {
                    // counts saturate at 255 downstream; the 24-bit field must not run into the context bits however many
                    // instances a k-mer has (a poly-G artefact can have 10^8): stop adding well before that
                    const uint32_t cur = *(volatile uint32_t*)(ccs + s);
                    const uint32_t old = (cur & 0xffffffu) < COUNT_STOP ? atomicAdd(ccs + s, 1u) : cur;
                    if ((((old >> 24) & ctx) != ctx)) atomicOr(ccs + s, ctx << 24);
                    done = true;
                } else s = (s + 1u) & (SMEM_SLOTS - 1u);
            }
            if (!done) sh_fail = 1;
        };
        // four records per thread are in flight before the first is inserted: with one, the stream of records is latency-bound
        // (bytes in flight per SM = 1024 threads x 16 B against ~1 us of DRAM latency)
        constexpr int UNROLL = 4;
        uint32_t bi = 0;                                          // i only grows: the run index is carried along
        for (uint32_t i0 = threadIdx.x; i0 < n; i0 += UNROLL * blockDim.x) {
            ulonglong2 rr[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                if (i < n) {
                    while (i >= sh_pre[bi + 1]) ++bi;
                    rr[u] = __ldcs(sp.recs + sh_run[bi] + (i - sh_pre[bi]));
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i0 + (uint32_t)u * blockDim.x < n) insert(rr[u]);
        }
        __syncthreads();
        const bool failed = sh_fail != 0;
        if (failed) {
            if (threadIdx.x == 0) sp.failed[atomicAdd(sp.failed_cursor, 1ull)] = p;
        } else {
            for (uint32_t base = 0; base < SMEM_SLOTS; base += blockDim.x) {
                const uint32_t j = base + threadIdx.x;
                const unsigned long long k0 = w0s[j];
                const bool occ = k0 != ~0ull;
                const uint32_t cc = ccs[j];
                uint32_t c = cc & 0xffffffu; if (c > 255u) c = 255u;
                const uint32_t ctx = cc >> 24;
                const uint32_t bin = occ ? (c > 100u ? 100u : c) : 103u;
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (occ && (peers & ((1u << lane_id()) - 1u)) == 0) atomicAdd(&sh_hist[bin], (unsigned)__popc(peers));
                const bool solid = occ && c >= sp.min_freq;
                const uint64_t pos = warp_append(sp.solid_cursor, solid);
                if (solid) { if (pos < sp.solid_cap) sp.solid_out[pos] = make_ulonglong2(k0, w1s[j] | ctx); else atomicExch(sp.solid_overflow, 1); }
                if (sp.dump_out) {
                    const uint64_t dp = warp_append(sp.dump_cursor, occ);
                    if (occ) sp.dump_out[dp] = DumpRec{k0, w1s[j], c, ctx, NIL, 0};
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= 100; j += blockDim.x) if (sh_hist[j]) atomicAdd(&sp.hist[j], (unsigned long long)sh_hist[j]);
}

struct ScanParams {
    CountSlot* region;
    uint64_t R;
    uint32_t min_freq;
    unsigned long long* hist;        // [104]: bins 1..100, [101] = solid count
    ulonglong2* solid_out;           // solid records {w0, w1 | ctx}
    unsigned long long* solid_cursor;
    uint64_t solid_cap;
    DumpRec* dump_out;               // test hook (dump level 2) or nullptr
    unsigned long long* dump_cursor;
    const int* overflow;             // if set: the group failed, only reset the region
    int* solid_overflow;
};
// hist[min(100,min(255,count))]++ (BuildReadQGraph.cc:1094-1097), keep count >= minFreq (:1098), reset the region.
__global__ void __launch_bounds__(256) k_scan_region(ScanParams sp) {
    __shared__ unsigned int sh[104];
    for (int j = threadIdx.x; j < 104; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    const bool failed = *sp.overflow != 0;
    uint4* raw = reinterpret_cast<uint4*>(sp.region);
    const uint4 key_empty = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), zero = make_uint4(0, 0, 0, 0);
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < sp.R; base += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = base + threadIdx.x;
        bool occ = false, solid = false;
        uint32_t c = 0, ctx = 0;
        ulonglong2 k = make_ulonglong2(0, 0);
        if (i < sp.R) {
            k = __ldcg(reinterpret_cast<const ulonglong2*>(sp.region + i));
            if (k.x != EMPTY_W0) {
                occ = true;
                uint2 m = __ldcg(reinterpret_cast<const uint2*>(&sp.region[i].count));
                c = m.x > 255u ? 255u : m.x; ctx = m.y & 0xffu;
                raw[2 * i] = key_empty; raw[2 * i + 1] = zero;
            }
        }
        if (failed) continue;
        uint32_t bin = occ ? (c > 100u ? 100u : c) : 103u;
        unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (occ && (peers & ((1u << lane_id()) - 1u)) == 0) atomicAdd(&sh[bin], (unsigned)__popc(peers));
        solid = occ && c >= sp.min_freq;
        uint64_t pos = warp_append(sp.solid_cursor, solid);
        if (solid) { if (pos < sp.solid_cap) sp.solid_out[pos] = make_ulonglong2(k.x, k.y | ctx); else atomicExch(sp.solid_overflow, 1); }
        if (sp.dump_out) {
            uint64_t dp = warp_append(sp.dump_cursor, occ);
            if (occ) sp.dump_out[dp] = DumpRec{k.x, k.y, c, ctx, NIL, 0};
        }
    }
    __syncthreads();
    if (!failed) {
        for (int j = threadIdx.x; j <= 100; j += blockDim.x) if (sh[j]) atomicAdd(&sp.hist[j], (unsigned long long)sh[j]);
    }
}

}  // namespace w2r
