// pipeline.cu — host orchestration of the step-2 path on one B200 and the C ABI over it (include/w2rap_step2.h).
//
// Mirrors buildReadQGraph (paths/long/BuildReadQGraph.cc:1253-1327) stage by stage:
//   createDictOMPRecursive  -> count_stage()      (k_good_len, k_minimizer_map x2, [NCCL exchange], k_count_smem, k_count_region + k_scan_region as
//                                                  fallback / legacy path, k_insert_solid)
//   recomputeAdjacencies    -> k_adjacency
//   buildEdges              -> unipath_stage()    (k_links, pointer-jumping list ranking, circles, edge emission)
//   buildHBVFromEdges       -> hbv_stage()        (end keys, radix sort, vertex ids, incidence)
//   path_reads_OMP(+FixPaths)-> path_stage()
// No CPU fallback: without a usable sm_100 device every compute entry point fails with W2RAP_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <mutex>
#include <thread>

#include "../../include/w2rap_step2.h"
#include "device_reads.cuh"
#include "count_part.cuh"
#include "kernels.cuh"
#include "nccl_dl.h"
#include "prims.cuh"

namespace w2r {

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Output arrays live in pinned host memory.  cudaMallocHost is slow (it maps and locks pages), so blocks are kept in a
// process-wide pool and reused by later calls: steady-state steps pay no allocation.
struct PinnedPool {
    struct Block { void* p; size_t bytes; bool used; };
    std::mutex mu;
    std::vector<Block> blocks;
    void* acquire(size_t bytes) {
        if (bytes < 64) bytes = 64;
        std::lock_guard<std::mutex> g(mu);
        int best = -1;
        for (size_t i = 0; i < blocks.size(); ++i)
            if (!blocks[i].used && blocks[i].bytes >= bytes && blocks[i].bytes <= 2 * bytes + 4096 && (best < 0 || blocks[i].bytes < blocks[best].bytes)) best = (int)i;
        if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
        void* p = nullptr;
        size_t cap = bytes + bytes / 8;       // a little slack so that slightly larger results of the next step still fit
        cudaError_t e = cudaMallocHost(&p, cap);
        if (e != cudaSuccess) {               // drop idle blocks and retry once
            cudaGetLastError();
            for (auto& b : blocks) if (!b.used && b.p) { cudaFreeHost(b.p); b.p = nullptr; b.bytes = 0; }
            W2R_CUDA(cudaMallocHost(&p, cap));
        }
        blocks.push_back(Block{p, cap, true});
        return p;
    }
    void release(void* p) {
        std::lock_guard<std::mutex> g(mu);
        for (auto& b : blocks) if (b.p == p) { b.used = false; return; }
    }
    static PinnedPool& get() { static PinnedPool* pool = new PinnedPool(); return *pool; }   // leaked on purpose: outlives the CUDA context teardown order
};
struct GraphOwner {
    std::vector<void*> pinned;
    ~GraphOwner() { for (void* p : pinned) PinnedPool::get().release(p); }
};
template <class T>
static T* out_alloc(GraphOwner* o, size_t n) {
    void* p = PinnedPool::get().acquire((n ? n : 1) * sizeof(T));
    o->pinned.push_back(p);
    return (T*)p;
}

static void check_device(int device, Ctx& c) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); W2R_FAIL(W2RAP_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path"); }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) W2R_FAIL(W2RAP_ERR_NO_DEVICE, "device %d does not exist (%d visible)", device, n);
    // cudaGetDeviceProperties is slow (it queries everything); three attributes, cached per device, are all that is needed
    static int cached_major[16], cached_sms[16];
    static bool cached[16];
    const int slot = device & 15;
    if (!cached[slot]) {
        int major = 0, minor = 0, sms = 0;
        W2R_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        W2R_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
        W2R_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        if (major < 10) W2R_FAIL(W2RAP_ERR_NO_DEVICE, "device %d (sm_%d%d) is not sm_100; this library only carries sm_100a code", device, major, minor);
        cached_major[slot] = major; cached_sms[slot] = sms; cached[slot] = true;
    }
    W2R_CUDA(cudaSetDevice(device));
    c.device = device;
    c.sm_count = cached_sms[slot];
    static std::once_flag pool_once[16];
    std::call_once(pool_once[device & 15], [&] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { uint64_t thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    });
}

// stream-ordered device buffer (memory comes back from the pool on the next call, so steady-state steps do not hit the driver)
template <class T>
struct SBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    SBuf() {}
    SBuf(Ctx& c, size_t n_) { alloc(c, n_); }
    SBuf(const SBuf&) = delete;
    SBuf& operator=(const SBuf&) = delete;
    SBuf(SBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    SBuf& operator=(SBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; } return *this; }
    ~SBuf() { release(); }
    void alloc(Ctx& c, size_t n_) {
        release();
        s = c.stream; n = n_;
        if (n) W2R_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), s));
    }
    void release() { if (p) cudaFreeAsync(p, s); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
    void zero() { if (n) W2R_CUDA(cudaMemsetAsync(p, 0, bytes(), s)); }
    void fill_ff() { if (n) W2R_CUDA(cudaMemsetAsync(p, 0xff, bytes(), s)); }
};

static size_t device_budget(const Ctx& c) {
    size_t fr = 0, tot = 0;
    W2R_CUDA(cudaMemGetInfo(&fr, &tot));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess) {
        uint64_t reserved = 0, used = 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        if (reserved > used) fr += (size_t)(reserved - used);
    }
    return fr;
}

struct StageTimer {
    Ctx& c;
    cudaEvent_t ev[2];
    explicit StageTimer(Ctx& c_) : c(c_) { cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]); }
    ~StageTimer() { cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); }
    void start() { cudaEventRecord(ev[0], c.stream); }
    float stop() { cudaEventRecord(ev[1], c.stream); cudaEventSynchronize(ev[1]); float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[1]); return ms; }
};

static void say(const Ctx& c, const char* fmt, ...) {
    if (!c.verbose) return;
    va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); fflush(stdout);
}

// ---------------------------------------------------------------- the pipeline
struct Pipeline {
    Ctx c;
    const DeviceReads& dr;
    const w2rap_params& prm;
    w2rap_graph* out;
    GraphOwner* owner;
    std::vector<DumpRec> dump_host;

    // persistent device state between stages
    SBuf<uint16_t> good;
    SBuf<SolidSlot> solid_slots;
    SolidTable st{nullptr, 0};
    uint64_t E = 0, nv = 0, nh = 0;
    SBuf<uint8_t> edge_bases; SBuf<uint64_t> edge_off; SBuf<uint32_t> edge_len;
    SBuf<int32_t> edge_vertices, fwd_xlat, rev_xlat, hleft, hright, from_e, to_e;
    SBuf<uint32_t> hcanon;
    SBuf<uint8_t> from_n, to_n;
    uint64_t edge_bytes = 0;
    int world = 1, rank = 0;            // sharded run: one process per GPU, reads sharded by index, k-mers routed to owners
    ncclComm_t comm = nullptr;

    Pipeline(const DeviceReads& dr_, const w2rap_params& p_, w2rap_graph* out_, GraphOwner* ow) : dr(dr_), prm(p_), out(out_), owner(ow) {}

    unsigned grid(uint64_t n, unsigned block, unsigned per_sm = 16) const { return grid_for(c, n, block, per_sm); }

    // ---- createDictOMPRecursive (BuildReadQGraph.cc:1015-1117): partition (map) + L2-resident hash count (reduce), count_part.cuh
    void set_l2_window(void* base, size_t bytes) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof attr);
        attr.accessPolicyWindow.base_ptr = base;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = bytes ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
        if (cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    }

    // ---- multi-GPU plumbing (NCCL over NVLink/NVSwitch); world == 1 needs none of it
    void nccl_check(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) W2R_FAIL(W2RAP_ERR_CUDA, "NCCL %s failed: %s", what, NcclApi::get().GetErrorString ? NcclApi::get().GetErrorString(r) : "?");
    }
    // element-wise all-reduce of a small host vector of u64 (through a device staging buffer)
    void allreduce_u64(std::vector<unsigned long long>& v, ncclRedOp_t op) {
        if (world == 1) return;
        SBuf<unsigned long long> d(c, v.size());
        W2R_CUDA(cudaMemcpyAsync(d.p, v.data(), v.size() * 8, cudaMemcpyHostToDevice, c.stream));
        nccl_check(NcclApi::get().AllReduce(d.p, d.p, v.size(), ncclUint64, op, comm, c.stream), "all-reduce");
        W2R_CUDA(cudaMemcpyAsync(v.data(), d.p, v.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        W2R_CUDA(cudaStreamSynchronize(c.stream));
    }
    // all-to-all of equal slabs: slab d of `send` goes to rank d, slab s of `recv` comes from rank s (MapReduceEngine's "swizzle",
    // MapReduceEngine.h:337-358, as grouped ncclSend/ncclRecv over NVLink)
    void alltoall_slabs(const void* send, void* recv, size_t slab_bytes) {
        NcclApi& n = NcclApi::get();
        nccl_check(n.GroupStart(), "group start");
        for (int peer = 0; peer < world; ++peer) {
            if (peer == rank) continue;
            nccl_check(n.Send((const char*)send + (size_t)peer * slab_bytes, slab_bytes, ncclUint8, peer, comm, c.stream), "send");
            nccl_check(n.Recv((char*)recv + (size_t)peer * slab_bytes, slab_bytes, ncclUint8, peer, comm, c.stream), "recv");
        }
        nccl_check(n.GroupEnd(), "group end");
        W2R_CUDA(cudaMemcpyAsync((char*)recv + (size_t)rank * slab_bytes, (const char*)send + (size_t)rank * slab_bytes, slab_bytes, cudaMemcpyDeviceToDevice, c.stream));
    }

    // byte-granular all-to-all with per-peer sizes (the minimiser layout: every rank sends each peer exactly the runs of the
    // partitions that peer owns)
    void alltoall_v(const void* send, const std::vector<size_t>& soff, const std::vector<size_t>& scnt, void* recv, const std::vector<size_t>& roff,
                    const std::vector<size_t>& rcnt) {
        NcclApi& n = NcclApi::get();
        nccl_check(n.GroupStart(), "group start");
        for (int peer = 0; peer < world; ++peer) {
            if (peer == rank) continue;
            if (scnt[peer]) nccl_check(n.Send((const char*)send + soff[peer], scnt[peer], ncclUint8, peer, comm, c.stream), "send");
            if (rcnt[peer]) nccl_check(n.Recv((char*)recv + roff[peer], rcnt[peer], ncclUint8, peer, comm, c.stream), "recv");
        }
        nccl_check(n.GroupEnd(), "group end");
        if (scnt[rank]) W2R_CUDA(cudaMemcpyAsync((char*)recv + roff[rank], (const char*)send + soff[rank], scnt[rank], cudaMemcpyDeviceToDevice, c.stream));
    }

    // ---- createDictOMPRecursive (BuildReadQGraph.cc:1000-1108): quality floor, k-mer counting, min-frequency filter, dictionary
    void count_stage() {
        const ReadsView rv = dr.view();
        good.alloc(c, dr.n);
        SBuf<unsigned long long> scal(c, 8);       // [0] k-mer instances, [1] solid cursor, [2] dump cursor, [4], [5] failed-partition cursors
        scal.zero();
        SBuf<int> flags(c, 4);                     // [0] malformed quality vector, [1] record buffer overflow, [2] solid staging overflow
        flags.zero();
        // Everything is sized from an upper bound of the instance count that needs no qualities (sum of len-59), so that the
        // quality floor + map of the first read batch can start while later batches are still crossing PCIe.
        const unsigned long long n_inst_local = dr.n_inst_upper;
        std::vector<unsigned long long> agg = {n_inst_local};
        allreduce_u64(agg, ncclSum);
        std::vector<unsigned long long> mx = {n_inst_local};
        allreduce_u64(mx, ncclMax);
        const unsigned long long n_inst = agg[0], n_inst_max = mx[0];      // whole job / largest shard (upper bounds)
        // read batches: [first, first+count) with an event to wait for (nullptr = already resident)
        struct Batch { uint64_t first, count; cudaEvent_t ready; };
        std::vector<Batch> batches;
        if (!dr.batch_ready.empty()) for (size_t b = 0; b < dr.batch_ready.size(); ++b) batches.push_back(Batch{dr.batch_first[b], dr.batch_first[b + 1] - dr.batch_first[b], dr.batch_ready[b]});
        else batches.push_back(Batch{0, dr.n, nullptr});
        bool good_done = false;

        // region: the L2-resident counting table.  1 Mi slots x 32 B = 32 MB of the 126 MB L2 (smaller for tiny inputs / the test hook).
        uint32_t logR = 20;
        while (logR > 8 && (1ull << (logR - 1)) >= 2 * n_inst + 64) --logR;
        if (prm.table_slots) { logR = 6; while ((2ull << logR) <= prm.table_slots && logR < 24) ++logR; }
        const uint64_t R = 1ull << logR;
        // Default: FINE partitions keyed by minimiser (count_part.cuh: k_minimizer_map), each small enough for one CTA to count in
        // shared memory (k_count_smem); the k-mer hash bits are then all free for the slot inside the table.  Rank r owns the
        // contiguous partition range [r*P/world, (r+1)*P/world).
        // Legacy (the table_slots test hook, W2RAP_STATIC_PARTITIONS): hash partitions in static sub-buffers, counted through the
        // L2 region; few enough records each that even an all-distinct partition fits the region.
        const bool mini = !prm.table_slots && !getenv("W2RAP_STATIC_PARTITIONS");
        // table size of the first k_count_smem launch: 8192 slots and partitions of ~24 k records (measured, config 2: 73 ms; 4096
        // slots with two 512-thread CTAs per SM and ~12 k records: 79 ms)
        static const uint32_t smem_log = getenv("W2RAP_SMEM_LOG") ? (uint32_t)atoi(getenv("W2RAP_SMEM_LOG")) : 13u;
        static const double fine_recs = getenv("W2RAP_FINE_RECS") ? atof(getenv("W2RAP_FINE_RECS")) : (smem_log <= 12 ? 12000.0 : 24000.0);
        uint32_t logP = 0;
        if (mini) { while (((double)n_inst / (double)(1ull << logP) > fine_recs || (1ull << logP) < (uint64_t)world) && logP < 22) ++logP; }
        else
        // (n_inst is an upper bound, and real read sets are far from all-distinct: 0.9 R records per partition; a partition that
        //  does not fit is handled by the hash sub-range fallback)
        while (((double)n_inst / (double)(1ull << logP) > 0.9 * (double)R || (1ull << logP) < (uint64_t)world) && logP < 24) ++logP;
        size_t budget = (size_t)(device_budget(c) * 0.80);
        if (prm.dump_kmers == 2 && world > 1) W2R_FAIL(W2RAP_ERR_BAD_ARG, "dump level 2 is a single-GPU test hook");
        const size_t fixed_bytes = R * sizeof(CountSlot) + (prm.dump_kmers == 2 ? n_inst * sizeof(DumpRec) : 0) +
                                   (size_t)((double)n_inst / world / std::max<uint32_t>(1, prm.min_freq) * 1.3) * sizeof(ulonglong2);
        if (fixed_bytes + (64u << 20) > budget) W2R_FAIL(W2RAP_ERR_OOM, "not enough device memory for the solid k-mer staging buffer");
        SBuf<ulonglong2> solid;                       // solid records of the partitions this rank owns
        uint64_t solid_cap = 0;
        // Two regions on two streams: groups alternate between them, so the scan/reset of one group overlaps the inserts of the next.
        SBuf<CountSlot> region(c, 2 * R);
        SBuf<DumpRec> dump_dev(c, prm.dump_kmers == 2 ? n_inst : 0);
        SBuf<unsigned long long> hist(c, 104); hist.zero();
        W2R_LAUNCH(c, k_init_count_table, grid(4 * R, 256), 256, 0, region.p, 2 * R);
        cudaStream_t s2 = nullptr;
        W2R_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        struct StreamGuard { cudaStream_t s; ~StreamGuard() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } } s2_guard{s2};
        cudaEvent_t ev_fork, ev_join;
        cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
        struct EventGuard { cudaEvent_t a, b; ~EventGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev_guard{ev_fork, ev_join};
        uint32_t group_parity = 0;
        float part_ms = 0, region_ms = 0, xchg_ms = 0;
        uint32_t npass = 1;
        double slack = 1.06;
        uint64_t n_distinct_seen = 0;
        uint32_t n_groups = 0;
        EventTimer kt(c.stream);
        for (int attempt = 0;; ++attempt) {
            if (attempt > 8) W2R_FAIL(W2RAP_ERR_INTERNAL, "k-mer partitioning did not converge");
            const uint64_t P = 1ull << logP, Pown = P / world;
            // legacy layout: 8 sub-buffers per partition with cursors on separate L2 lines; minimiser layout: one cursor per partition
            const uint32_t nsub = (!mini && n_inst_max / P >= 65536 && !prm.table_slots) ? 8u : 1u;
            const uint32_t cstride = mini ? 1 : 32;
            const uint64_t NB = P * nsub, NBown = Pown * nsub;
            const uint64_t per_part = (uint64_t)((double)n_inst_max / (double)NB / (double)npass);
            const uint64_t cap = mini ? (1ull << 40) : (uint64_t)((double)per_part * slack) + 1024;
            // minimiser layout: the local record area is sized from the upper bound (exact sizes come from the counting launch, but
            // only on the device); with several passes the k-mer space is split by minimiser hash, so allow for uneven passes.
            // Sharded: plus the records this rank owns (allocated exactly once the counts have been exchanged).
            const size_t local_recs = mini ? (size_t)((double)n_inst_local / npass * (npass > 1 ? slack * 1.2 : 1.0)) + 64 : NB * cap;
            const size_t rec_bytes = mini ? (local_recs + (world > 1 ? (size_t)((double)n_inst / world / npass * 1.25) : 0)) * sizeof(ulonglong2)
                                          : NB * cap * sizeof(ulonglong2) * (world > 1 ? 2 : 1);   // + the receive slabs
            // Legacy layout: scattered single-record appends over tens of GB run into TLB misses (measured: 2x slower per record on
            // an 82 GB buffer than on a 41 GB one), so that buffer is capped and the k-mer space split into hash-range passes.
            static const double rec_cap_gb = getenv("W2RAP_REC_BUDGET_GB") ? atof(getenv("W2RAP_REC_BUDGET_GB")) : 48.0;
            std::vector<unsigned long long> too_big = {(rec_bytes + fixed_bytes > budget || (!mini && (double)(NB * cap * sizeof(ulonglong2)) > rec_cap_gb * 1e9)) ? 1ull : 0ull};
            allreduce_u64(too_big, ncclMax);                     // every rank must run the same number of passes
            if (too_big[0] && npass < 4096) { npass *= 2; continue; }
            const double t_alloc0 = now_ms();
            SBuf<ulonglong2> recs(c, local_recs), xrecs_buf(c, (!mini && world > 1) ? NB * cap : 0);
            if (getenv("W2RAP_TRACE")) fprintf(stderr, "[w2rap] count: record buffer %.2f GB allocated in %.1f ms (host)\n", recs.bytes() / 1e9, now_ms() - t_alloc0);
            // Minimiser layout, one GPU: every read batch gets its own exactly sized record area, so that a batch is mapped as soon
            // as it has arrived; a partition is then one run per batch.  Sharded: one local area, laid out by partition = by owner,
            // and after the exchange a partition is one run per source rank.  Either way the reduce sees `nslab` runs per partition.
            const uint32_t nmap = mini ? (world > 1 ? 1u : (uint32_t)batches.size()) : 0u;
            const uint32_t nslab = mini ? (world > 1 ? (uint32_t)world : nmap) : (uint32_t)world;
            if (mini && nslab > SMEM_MAX_BATCH) W2R_FAIL(W2RAP_ERR_INTERNAL, "more record slabs per partition than the reduce kernel handles");
            SBuf<uint32_t> part_count(c, nmap * P);
            SBuf<uint64_t> part_base(c, nmap * P), batch_total(c, nmap);
            SBuf<unsigned long long> batch_off(c, mini ? nmap + 1 : 0);
            SBuf<uint32_t> cursor(c, mini ? nmap * P : NB * cstride), xcur_buf(c, world > 1 ? (mini ? world * Pown : NB * cstride) : 0);
            SBuf<uint64_t> xbase(c, (mini && world > 1) ? world * Pown : 0), xtot(c, (mini && world > 1) ? world : 0);
            SBuf<unsigned long long> xoff(c, (mini && world > 1) ? world + 1 : 0);
            const uint64_t slab_recs = NBown * cap, slab_cur = NBown * cstride;
            // what this rank reduces: records, per-slab partition sizes, and (minimiser layout) where the runs start
            const ulonglong2* xrecs = world > 1 ? xrecs_buf.p : recs.p;
            const uint32_t* xcur = world > 1 ? xcur_buf.p : cursor.p;            // [nslab][NBown][cstride]
            RunView runs{nullptr, nullptr, Pown, nullptr};
            if (mini) runs = world > 1 ? RunView{xbase.p, xoff.p, Pown, nullptr} : RunView{part_base.p, batch_off.p, Pown, nullptr};
            std::vector<uint32_t> sizes(Pown), raw_sizes((size_t)nslab * slab_cur);
            std::vector<uint64_t> totals(Pown);
            bool retry = false;
            W2R_CUDA(cudaMemsetAsync(scal.p + 1, 0, 16, c.stream));       // solid cursor, dump cursor
            hist.zero();
            uint64_t solid_used_before = 0;
            for (uint32_t pass = 0; pass < npass && !retry; ++pass) {
                cursor.zero();
                if (mini) {
                    // launch 1 sizes the partitions exactly (per read batch, under the upload), the scan lays them out, launch 2 stores
                    static const unsigned map_ctas = getenv("W2RAP_MAP_CTAS") ? (unsigned)atoi(getenv("W2RAP_MAP_CTAS")) : 8u;
                    part_count.zero();
                    W2R_CUDA(cudaMemsetAsync(batch_off.p, 0, sizeof(unsigned long long), c.stream));
                    std::vector<Batch> mb = batches;
                    if (world > 1) {        // sharded: one map batch (the layout must be by owner rank), after the whole shard has arrived
                        for (const Batch& bt : batches) {
                            if (bt.ready && !good_done) W2R_CUDA(cudaStreamWaitEvent(c.stream, bt.ready, 0));
                            if (bt.count && !good_done) W2R_LAUNCH(c, k_good_len, grid(bt.count, 128), 128, 0, rv, bt.first, bt.count, prm.min_qual, good.p, scal.p, flags.p);
                        }
                        good_done = true;
                        mb.assign(1, Batch{0, dr.n, nullptr});
                    }
                    std::vector<cudaEvent_t> ev(2 * nmap, nullptr);                 // the store launches are timed without stalling the queue
                    struct EvGuard { std::vector<cudaEvent_t>& v; ~EvGuard() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); } } evg{ev};
                    for (uint32_t bi = 0; bi < nmap; ++bi) {
                        const Batch& bt = mb[bi];
                        MiniParams mp{logP, npass, pass, part_count.p + bi * P, part_base.p + bi * P, cursor.p + bi * P, recs.p, batch_off.p + bi, recs.n, flags.p + 1};
                        if (bt.ready && !good_done) W2R_CUDA(cudaStreamWaitEvent(c.stream, bt.ready, 0));
                        if (bt.count && !good_done) W2R_LAUNCH(c, k_good_len, grid(bt.count, 128), 128, 0, rv, bt.first, bt.count, prm.min_qual, good.p, scal.p, flags.p);
                        if (bt.count && n_inst_local) W2R_LAUNCH(c, k_minimizer_map<true>, grid(bt.count * 32, 256, map_ctas), 256, 0, rv, bt.first, bt.count, good.p, mp);
                        exclusive_scan<uint32_t, uint64_t>(c, part_count.p + bi * P, P, part_base.p + bi * P, batch_total.p + bi);
                        W2R_LAUNCH(c, k_next_batch_off, 1, 1, 0, batch_off.p + bi, batch_total.p + bi);
                        W2R_CUDA(cudaEventCreate(&ev[2 * bi])); W2R_CUDA(cudaEventCreate(&ev[2 * bi + 1]));
                        W2R_CUDA(cudaEventRecord(ev[2 * bi], c.stream));
                        if (bt.count && n_inst_local) { W2R_LAUNCH(c, k_minimizer_map<false>, grid(bt.count * 32, 256, map_ctas), 256, 0, rv, bt.first, bt.count, good.p, mp); c.count_launches++; }
                        W2R_CUDA(cudaEventRecord(ev[2 * bi + 1], c.stream));
                    }
                    good_done = true;
                    W2R_CUDA(cudaStreamSynchronize(c.stream));
                    for (uint32_t bi = 0; bi < nmap; ++bi) { float ms = 0; cudaEventElapsedTime(&ms, ev[2 * bi], ev[2 * bi + 1]); part_ms += ms; }
                } else {
                    PartParams pp{recs.p, cursor.p, cap, logP, nsub, cstride, npass, pass, flags.p + 1};
                    kt.start();
                    for (const Batch& bt : batches) {
                        if (!bt.count) continue;
                        if (bt.ready && !good_done) W2R_CUDA(cudaStreamWaitEvent(c.stream, bt.ready, 0));
                        if (!good_done) W2R_LAUNCH(c, k_good_len, grid(bt.count, 128), 128, 0, rv, bt.first, bt.count, prm.min_qual, good.p, scal.p, flags.p);
                        if (n_inst_local) { W2R_LAUNCH(c, k_extract_partition, grid(bt.count, 256, 8), 256, 0, rv, bt.first, bt.count, good.p, pp); c.count_launches++; }
                    }
                    good_done = true;
                    part_ms += kt.stop();
                }
                if (d2h_scalar(c, flags.p)) W2R_FAIL(W2RAP_ERR_BAD_ARG, "a read's quality vector does not have one quality per base");
                std::vector<unsigned long long> of = {(unsigned long long)d2h_scalar(c, flags.p + 1)};
                allreduce_u64(of, ncclSum);
                if (of[0]) {        // a record buffer overflowed somewhere (legacy: skewed k-mer multiplicities; minimiser: an uneven pass): more slack, on every rank
                    W2R_CUDA(cudaMemsetAsync(flags.p + 1, 0, sizeof(int), c.stream));
                    slack *= 1.5; retry = true; break;
                }
                if (world > 1 && !mini) {    // route every record to the rank that owns its partition: fixed-size slabs
                    kt.start();
                    alltoall_slabs(recs.p, xrecs_buf.p, slab_recs * sizeof(ulonglong2));
                    alltoall_slabs(cursor.p, xcur_buf.p, slab_cur * sizeof(uint32_t));
                    xchg_ms += kt.stop();
                }
                if (world > 1 && mini) {     // ... exactly sized runs: counts first, then the records
                    kt.start();
                    alltoall_slabs(part_count.p, xcur_buf.p, Pown * sizeof(uint32_t));          // xcur[s][q] = records of my partition q held by rank s
                    W2R_CUDA(cudaMemsetAsync(xoff.p, 0, sizeof(unsigned long long), c.stream));
                    for (int sidx = 0; sidx < world; ++sidx) {
                        exclusive_scan<uint32_t, uint64_t>(c, xcur_buf.p + (uint64_t)sidx * Pown, Pown, xbase.p + (uint64_t)sidx * Pown, xtot.p + sidx);
                        W2R_LAUNCH(c, k_next_batch_off, 1, 1, 0, xoff.p + sidx, xtot.p + sidx);
                    }
                    std::vector<uint64_t> sbeg(world + 1, 0), rtot(world, 0);
                    for (int d = 0; d < world; ++d) W2R_CUDA(cudaMemcpyAsync(&sbeg[d], part_base.p + (uint64_t)d * Pown, 8, cudaMemcpyDeviceToHost, c.stream));
                    W2R_CUDA(cudaMemcpyAsync(&sbeg[world], batch_total.p, 8, cudaMemcpyDeviceToHost, c.stream));
                    W2R_CUDA(cudaMemcpyAsync(rtot.data(), xtot.p, world * 8, cudaMemcpyDeviceToHost, c.stream));
                    W2R_CUDA(cudaStreamSynchronize(c.stream));
                    std::vector<size_t> soff(world), scnt(world), roff(world), rcnt(world);
                    size_t racc = 0;
                    for (int d = 0; d < world; ++d) {
                        soff[d] = sbeg[d] * sizeof(ulonglong2); scnt[d] = (sbeg[d + 1] - sbeg[d]) * sizeof(ulonglong2);
                        roff[d] = racc * sizeof(ulonglong2); rcnt[d] = rtot[d] * sizeof(ulonglong2); racc += rtot[d];
                    }
                    xrecs_buf.alloc(c, racc + 1);
                    xrecs = xrecs_buf.p;
                    alltoall_v(recs.p, soff, scnt, xrecs_buf.p, roff, rcnt);
                    xchg_ms += kt.stop();
                }
                W2R_CUDA(cudaMemcpyAsync(raw_sizes.data(), xcur, raw_sizes.size() * 4, cudaMemcpyDeviceToHost, c.stream));
                W2R_CUDA(cudaStreamSynchronize(c.stream));
                uint64_t owned_records = 0;
                for (uint64_t q = 0; q < Pown; ++q) {
                    uint32_t mxs = 0;
                    uint64_t tq = 0;
                    for (uint32_t sidx = 0; sidx < nslab; ++sidx)
                        for (uint32_t u = 0; u < nsub; ++u) { uint32_t v = raw_sizes[sidx * slab_cur + (q * nsub + u) * cstride]; mxs = std::max(mxs, v); tq += v; }
                    sizes[q] = mxs;      // largest run / sub-buffer of the partition (sizes the grid)
                    totals[q] = tq;
                    owned_records += tq;
                }
                {   // staging for this pass's solid records (each needs >= min_freq instances)
                    uint64_t need = solid_used_before + owned_records / std::max<uint32_t>(1, prm.min_freq) + 1024;
                    if (need > solid_cap) {
                        SBuf<ulonglong2> bigger(c, need + need / 4);
                        if (solid_used_before) W2R_CUDA(cudaMemcpyAsync(bigger.p, solid.p, solid_used_before * sizeof(ulonglong2), cudaMemcpyDeviceToDevice, c.stream));
                        solid = std::move(bigger);
                        solid_cap = solid.n;
                    }
                }
                // ---- reduce
                kt.start();
                // the persisting carve-out is taken from the normal L2, which the map needs for write combining: hold it only while reducing
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(2 * R * sizeof(CountSlot), 80u << 20)) != cudaSuccess) cudaGetLastError();
                set_l2_window(region.p, 2 * R * sizeof(CountSlot));
                {   // the second stream gets the same L2 window and starts after everything queued so far
                    cudaStreamAttrValue attr; memset(&attr, 0, sizeof attr);
                    attr.accessPolicyWindow.base_ptr = region.p; attr.accessPolicyWindow.num_bytes = 2 * R * sizeof(CountSlot);
                    attr.accessPolicyWindow.hitRatio = 1.0f; attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    if (cudaStreamSetAttribute(s2, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
                }
                SBuf<int> gflag(c, Pown + 1); gflag.zero();
                auto fork = [&] { W2R_CUDA(cudaEventRecord(ev_fork, c.stream)); W2R_CUDA(cudaStreamWaitEvent(s2, ev_fork, 0)); };
                auto join = [&] { W2R_CUDA(cudaEventRecord(ev_join, s2)); W2R_CUDA(cudaStreamWaitEvent(c.stream, ev_join, 0)); };
                fork();
                // a group of owned partitions goes through one counting region: the consecutive range [p0, p0+g), or entries
                // [p0, p0+g) of a partition list (hlist on the host, dlist on the device)
                auto run_group = [&](uint32_t p0, uint32_t g, uint32_t sub_mask, uint32_t sub_id, int* flag, const uint32_t* dlist = nullptr, const uint32_t* hlist = nullptr) {
                    ++n_groups;
                    uint32_t mxg = 0;
                    for (uint32_t q = p0; q < p0 + g; ++q) mxg = std::max(mxg, sizes[hlist ? hlist[q] : q]);
                    cudaStream_t gs = (group_parity & 1u) ? s2 : c.stream;
                    CountSlot* greg = region.p + ((group_parity & 1u) ? R : 0);
                    ++group_parity;
                    if (mxg) {
                        RegionParams rp{greg, logR, mini ? 0u : logP, sub_mask, sub_id, flag};
                        const uint32_t gy = g * nsub;
                        dim3 gr(std::max(1u, std::min<unsigned>((mxg + 511) / 512, (unsigned)(c.sm_count * 8 / std::max(1u, std::min(gy * nslab, 8u))))), gy * nslab);
                        RunView rv2 = runs;
                        rv2.plist = dlist;
                        k_count_region<<<gr, 256, 0, gs>>>(xrecs, xcur, cstride, cap, p0 * nsub, gy, slab_recs, slab_cur, rv2, rp); c.launches++;
                        W2R_CUDA(cudaGetLastError());
                    }
                    ScanParams sp{greg, R, prm.min_freq, hist.p, solid.p, scal.p + 1, solid_cap, prm.dump_kmers == 2 ? dump_dev.p : nullptr, scal.p + 2, flag, flags.p + 2};
                    k_scan_region<<<grid(R, 256, 4), 256, 0, gs>>>(sp); c.launches++;
                    W2R_CUDA(cudaGetLastError());
                };
                std::vector<std::pair<uint32_t, uint32_t>> groups;   // (first owned partition, count)
                std::vector<int> gf(Pown + 1, 0);
                if (mini) {
                    // every partition is counted by one CTA in shared memory; the few that do not fit are retried with the largest
                    // table, and what still fails is redone through the region below
                    SBuf<uint32_t> failed(c, Pown), failed2(c, Pown);
                    W2R_CUDA(cudaMemsetAsync(scal.p + 4, 0, 16, c.stream));
                    // (a per-device attribute: set on every call — ranks driven from threads of one process each have their own device)
                    W2R_CUDA(cudaFuncSetAttribute(k_count_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((20u << SMEM_LOG_SLOTS_MAX))));
                    const uint32_t log1 = std::min(std::max(smem_log, 10u), SMEM_LOG_SLOTS_MAX);
                    SmemCountParams sc{xrecs, xcur, runs.part_base, runs.slab_off, nslab, (uint32_t)Pown, 0u, prm.min_freq, hist.p, solid.p, scal.p + 1, solid_cap, flags.p + 2,
                                       prm.dump_kmers == 2 ? dump_dev.p : nullptr, scal.p + 2, failed.p, scal.p + 4, log1, nullptr, 0};
                    const bool two_per_sm = log1 <= 12;
                    k_count_smem<<<(unsigned)std::min<uint64_t>(Pown, (uint64_t)c.sm_count * (two_per_sm ? 2 : 1)), two_per_sm ? 512 : 1024, (size_t)20u << log1, c.stream>>>(sc); c.launches++; ++n_groups;
                    W2R_CUDA(cudaGetLastError());
                    const uint64_t nfail1 = log1 < SMEM_LOG_SLOTS_MAX ? d2h_scalar(c, scal.p + 4) : 0;
                    if (nfail1) {
                        SmemCountParams sc2 = sc;
                        sc2.failed = failed2.p; sc2.failed_cursor = scal.p + 5; sc2.log_slots = SMEM_LOG_SLOTS_MAX; sc2.plist = failed.p; sc2.nlist = (uint32_t)nfail1;
                        k_count_smem<<<(unsigned)std::min<uint64_t>(nfail1, (uint64_t)c.sm_count), 1024, (size_t)20u << SMEM_LOG_SLOTS_MAX, c.stream>>>(sc2); c.launches++; ++n_groups;
                        W2R_CUDA(cudaGetLastError());
                    }
                    const SBuf<uint32_t>& failed_final = nfail1 ? failed2 : failed;
                    const uint64_t nfail = d2h_scalar(c, nfail1 ? scal.p + 5 : scal.p + 4);
                    std::vector<uint32_t> fl(nfail);
                    if (nfail) { W2R_CUDA(cudaMemcpyAsync(fl.data(), failed_final.p, nfail * 4, cudaMemcpyDeviceToHost, c.stream)); W2R_CUDA(cudaStreamSynchronize(c.stream)); }
                    if (nfail) {
                        say(c, "%llu of %llu k-mer partitions did not fit shared memory; counting them through the L2 region", (unsigned long long)nfail, (unsigned long long)Pown);
                        // in bulk: as many listed partitions per region pass as fit it even if every record were distinct (load <= 0.6),
                        // alternating regions/streams, no host round trip per group; a group that overflows all the same is redone
                        // partition by partition below
                        std::vector<std::pair<uint32_t, uint32_t>> lgroups;        // (offset into fl, count)
                        const uint32_t gmax = std::max(1u, 4000u / nslab);          // gridDim.y = g * nslab <= 65535
                        for (size_t i0 = 0; i0 < fl.size();) {
                            uint64_t acc = 0; size_t j = i0;
                            while (j < fl.size() && j - i0 < gmax && (j == i0 || (double)(acc + totals[fl[j]]) <= 0.6 * (double)R)) { acc += totals[fl[j]]; ++j; }
                            lgroups.push_back({(uint32_t)i0, (uint32_t)(j - i0)});
                            i0 = j;
                        }
                        fork();
                        for (size_t gi = 0; gi < lgroups.size(); ++gi) run_group(lgroups[gi].first, lgroups[gi].second, 0, 0, gflag.p + gi, failed_final.p, fl.data());
                        join();
                        std::vector<int> lgf(lgroups.size());
                        W2R_CUDA(cudaMemcpyAsync(lgf.data(), gflag.p, lgroups.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
                        W2R_CUDA(cudaStreamSynchronize(c.stream));
                        for (size_t gi = 0; gi < lgroups.size(); ++gi)
                            if (lgf[gi]) for (uint32_t k = lgroups[gi].first; k < lgroups[gi].first + lgroups[gi].second; ++k) { groups.push_back({fl[k], 1u}); gf[fl[k]] = 1; }
                    }
                } else {
                    // groups of consecutive owned partitions share the region; the group size comes from the first partition's distinct count
                    run_group(0, 1, 0, 0, gflag.p + 0);
                    groups.push_back({0u, 1u});
                    if (Pown > 1) {
                        std::vector<unsigned long long> hh(104);
                        W2R_CUDA(cudaMemcpyAsync(hh.data(), hist.p, 104 * 8, cudaMemcpyDeviceToHost, c.stream));
                        W2R_CUDA(cudaStreamSynchronize(c.stream));
                        unsigned long long d0 = 0;
                        for (int i = 1; i <= 100; ++i) d0 += hh[i];
                        d0 -= std::min<unsigned long long>(d0, n_distinct_seen);
                        uint32_t g = (uint32_t)std::max<double>(1.0, std::min<double>(4096.0, 0.5 * (double)R / ((double)d0 * 1.15 + 1.0)));
                        g = std::max<uint32_t>(1u, std::min<uint32_t>(g, 32768u / (nsub * world)));
                        for (uint64_t p0 = 1; p0 < Pown; p0 += g) {
                            uint32_t gg = (uint32_t)std::min<uint64_t>(g, Pown - p0);
                            run_group((uint32_t)p0, gg, 0, 0, gflag.p + p0);
                            groups.push_back({(uint32_t)p0, gg});
                        }
                    }
                    join();
                    W2R_CUDA(cudaMemcpyAsync(gf.data(), gflag.p, (Pown + 1) * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
                    W2R_CUDA(cudaStreamSynchronize(c.stream));
                }
                for (auto& gr : groups) {
                    if (!gf[gr.first]) continue;
                    // the group did not fit together: its partitions one by one, and a partition that still fails in hash sub-ranges
                    for (uint32_t q = gr.first; q < gr.first + gr.second; ++q) {
                        unsigned long long snapshot[2];
                        std::vector<unsigned long long> hist_snapshot(104);
                        W2R_CUDA(cudaMemcpyAsync(snapshot, scal.p + 1, 16, cudaMemcpyDeviceToHost, c.stream));
                        W2R_CUDA(cudaMemcpyAsync(hist_snapshot.data(), hist.p, 104 * 8, cudaMemcpyDeviceToHost, c.stream));
                        W2R_CUDA(cudaMemsetAsync(gflag.p + Pown, 0, sizeof(int), c.stream));
                        fork(); run_group(q, 1, 0, 0, gflag.p + Pown); join();
                        if (!d2h_scalar(c, gflag.p + Pown)) continue;
                        for (uint32_t S = 2;; S *= 2) {
                            if (S > 4096) W2R_FAIL(W2RAP_ERR_INTERNAL, "a k-mer partition does not fit the counting region even in 4096 hash sub-ranges");
                            bool ok = true;
                            for (uint32_t sid = 0; sid < S && ok; ++sid) {
                                W2R_CUDA(cudaMemsetAsync(gflag.p + Pown, 0, sizeof(int), c.stream));
                                fork(); run_group(q, 1, S - 1, sid, gflag.p + Pown); join();
                                if (d2h_scalar(c, gflag.p + Pown)) ok = false;
                            }
                            if (ok) break;
                            // roll back what the successful sub-ranges of this split emitted, then split finer
                            W2R_CUDA(cudaMemcpyAsync(scal.p + 1, snapshot, 16, cudaMemcpyHostToDevice, c.stream));
                            W2R_CUDA(cudaMemcpyAsync(hist.p, hist_snapshot.data(), 104 * 8, cudaMemcpyHostToDevice, c.stream));
                            W2R_CUDA(cudaStreamSynchronize(c.stream));
                        }
                    }
                }
                set_l2_window(nullptr, 0);
                region_ms += kt.stop();
                if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
                {
                    std::vector<unsigned long long> hh(104);
                    W2R_CUDA(cudaMemcpyAsync(hh.data(), hist.p, 104 * 8, cudaMemcpyDeviceToHost, c.stream));
                    W2R_CUDA(cudaStreamSynchronize(c.stream));
                    n_distinct_seen = 0;
                    for (int i = 1; i <= 100; ++i) n_distinct_seen += hh[i];
                    solid_used_before = d2h_scalar(c, scal.p + 1);
                }
                if (mini && world > 1) { xrecs_buf.release(); xrecs = nullptr; }
            }
            if (retry) { n_distinct_seen = 0; n_groups = 0; continue; }
            if (d2h_scalar(c, flags.p + 2)) W2R_FAIL(W2RAP_ERR_INTERNAL, "solid staging buffer overflow");
            break;
        }
        if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
        {   // the exact instance count of the whole job (the bound above only sized buffers)
            std::vector<unsigned long long> exact = {(unsigned long long)d2h_scalar(c, scal.p)};
            allreduce_u64(exact, ncclSum);
            out->n_kmer_instances = exact[0];
            say(c, "%llu k-mer instances in quality-floored reads", exact[0]);
        }
        out->timings.count_kernel_ms = part_ms;
        out->timings.region_ms = region_ms;
        out->timings.exchange_ms = xchg_ms;
        out->timings.count_passes = n_groups;
        std::vector<unsigned long long> hh(104);
        W2R_CUDA(cudaMemcpyAsync(hh.data(), hist.p, 104 * 8, cudaMemcpyDeviceToHost, c.stream));
        unsigned long long cursors[2];
        W2R_CUDA(cudaMemcpyAsync(cursors, scal.p + 1, 16, cudaMemcpyDeviceToHost, c.stream));
        W2R_CUDA(cudaStreamSynchronize(c.stream));
        allreduce_u64(hh, ncclSum);                          // histogram of the whole job
        uint64_t n_distinct = 0;
        for (int i = 1; i <= 100; ++i) { out->hist[i] = hh[i]; n_distinct += hh[i]; }
        const uint64_t n_solid_local = cursors[0];
        if (prm.dump_kmers == 2 && cursors[1]) {
            dump_host.resize(cursors[1]);
            W2R_CUDA(cudaMemcpyAsync(dump_host.data(), dump_dev.p, cursors[1] * sizeof(DumpRec), cudaMemcpyDeviceToHost, c.stream));
            W2R_CUDA(cudaStreamSynchronize(c.stream));
        }
        region.release(); dump_dev.release();
        // every rank needs the whole dictionary for adjacency, unipaths and pathing: all-gather the solid records
        std::vector<unsigned long long> per_rank(world, 0ull);
        per_rank[rank] = n_solid_local;
        allreduce_u64(per_rank, ncclSum);
        uint64_t n_solid = 0;
        for (auto v : per_rank) n_solid += v;
        SBuf<ulonglong2> solid_all;
        const ulonglong2* solid_src = solid.p;
        if (world > 1) {
            kt.start();
            solid_all.alloc(c, n_solid);
            uint64_t off = 0;
            for (int sidx = 0; sidx < world; ++sidx) {
                if (per_rank[sidx]) nccl_check(NcclApi::get().Broadcast(sidx == rank ? (const void*)solid.p : (const void*)(solid_all.p + off), solid_all.p + off,
                                                                         per_rank[sidx] * sizeof(ulonglong2), ncclUint8, sidx, comm, c.stream), "broadcast");
                off += per_rank[sidx];
            }
            out->timings.exchange_ms += kt.stop();
            solid_src = solid_all.p;
            solid.release();
        }
        out->n_distinct = n_distinct; out->n_solid = n_solid;
        say(c, "%llu kmers counted, filtering...", (unsigned long long)n_distinct);
        say(c, "%llu / %llu kmers with Freq >= %u", (unsigned long long)n_solid, (unsigned long long)n_distinct, prm.min_freq);

        // dictionary (kmers/ReadPather.h:176-349) as an open-addressing table at load <= 0.5
        uint32_t lg = 10;
        while ((1ull << lg) < 2 * n_solid) ++lg;
        if (lg > 31) W2R_FAIL(W2RAP_ERR_OOM, "more than 2^30 solid k-mers on one device");
        solid_slots.alloc(c, 1ull << lg);
        solid_slots.fill_ff();
        st = SolidTable{solid_slots.p, lg};
        if (n_solid) W2R_LAUNCH(c, k_insert_solid, grid(n_solid, 256), 256, 0, solid_src, n_solid, st);
        W2R_CUDA(cudaStreamSynchronize(c.stream));   // solid / solid_all are released when this function returns
    }

    // ---- buildEdges (BuildReadQGraph.cc:314-339)
    void unipath_stage() {
        const uint64_t T = st.size(), nn = 2 * T;
        SBuf<uint32_t> next0(c, nn);
        SBuf<int> flags(c, 4); flags.zero();
        SBuf<unsigned long long> scal(c, 4); scal.zero();
        W2R_LAUNCH(c, k_links, grid(nn, 256), 256, 0, st, next0.p, flags.p);
        // list ranking: splitters walk to the next splitter, the splitters alone are ranked by pointer jumping (unipath.cuh)
        SBuf<RankState> A(c, nn), B(c, nn);          // A: label, then the final (tail, distance) of every node; B: splitter states
        SBuf<uint32_t> splist(c, nn / 16 + out->n_solid / 2 + 1024);
        W2R_CUDA(cudaMemsetAsync(A.p, 0xff, A.bytes(), c.stream));     // label = {NIL, ...}
        W2R_CUDA(cudaMemsetAsync(scal.p + 3, 0, 8, c.stream));
        W2R_LAUNCH(c, k_splitter_walk, grid(nn, 256), 256, 0, next0.p, nn, A.p, B.p, splist.p, scal.p + 3);
        if (d2h_scalar(c, flags.p)) W2R_FAIL(W2RAP_ERR_INTERNAL, "a neighbour k-mer promised by a pruned context is missing (reference: ForceAssert in EdgeBuilder::lookup)");
        const uint64_t nsp = d2h_scalar(c, scal.p + 3);
        if (nsp > splist.n) W2R_FAIL(W2RAP_ERR_INTERNAL, "splitter list overflow");
        unsigned long long prev_un = ~0ull;
        if (nsp) {
            for (int round = 0; round < 48; ++round) {
                W2R_CUDA(cudaMemsetAsync(scal.p, 0, 8, c.stream));
                W2R_LAUNCH(c, k_rank_step_inplace, grid(nsp, 256), 256, 0, (const uint32_t*)splist.p, nsp, (unsigned long long*)B.p, scal.p);
                unsigned long long un = d2h_scalar(c, scal.p);
                if (un == 0 || un == prev_un) { prev_un = un; break; }
                prev_un = un;
            }
        }
        W2R_CUDA(cudaMemsetAsync(scal.p, 0, 8, c.stream));
        W2R_LAUNCH(c, k_splitter_finish, grid(nn, 256), 256, 0, next0.p, nn, A.p, (const RankState*)B.p, scal.p);
        prev_un = d2h_scalar(c, scal.p);
        RankState* cur = A.p; RankState* oth = B.p;
        if (prev_un) {   // smooth circles (:332-335)
            uint64_t ncyc = prev_un;
            SBuf<uint32_t> list(c, ncyc);
            splist.release();
            W2R_CUDA(cudaMemsetAsync(scal.p + 1, 0, 8, c.stream));
            W2R_LAUNCH(c, k_collect_unresolved, grid(nn, 256), 256, 0, cur, nn, list.p, scal.p + 1);
            if (d2h_scalar(c, scal.p + 1) != ncyc) W2R_FAIL(W2RAP_ERR_INTERNAL, "unresolved node count changed between kernels");
            // minimum canonical k-mer per cycle by pointer doubling.  `cur` holds the final ranks of the path nodes and is
            // preserved; the cycle work ping-pongs between `oth` and D, touching only cycle entries.
            SBuf<RankState> D(c, nn);   // second buffer for cycle work (only cycle entries are touched)
            RankState* x0 = oth; RankState* x1 = D.p;
            W2R_LAUNCH(c, k_cycle_init, grid(ncyc, 256), 256, 0, list.p, ncyc, next0.p, x0);
            for (int round = 0; round < 40; ++round) {
                W2R_CUDA(cudaMemsetAsync(flags.p + 1, 0, sizeof(int), c.stream));
                W2R_LAUNCH(c, k_cycle_step, grid(ncyc, 256), 256, 0, list.p, ncyc, st, x0, x1, flags.p + 1);
                std::swap(x0, x1);
                if (!d2h_scalar(c, flags.p + 1)) break;
            }
            W2R_LAUNCH(c, k_cycle_cut, grid(ncyc, 256), 256, 0, list.p, ncyc, x0, next0.p);
            // the cycles are now paths: rank them
            W2R_LAUNCH(c, k_rank_init_list, grid(ncyc, 256), 256, 0, list.p, ncyc, next0.p, x0);
            for (int round = 0; round < 40; ++round) {
                W2R_CUDA(cudaMemsetAsync(scal.p, 0, 8, c.stream));
                W2R_LAUNCH(c, k_rank_step, grid(ncyc, 256), 256, 0, list.p, ncyc, x0, x1, scal.p);
                std::swap(x0, x1);
                unsigned long long un = d2h_scalar(c, scal.p);
                if (un == 0) break;
                if (round == 39) W2R_FAIL(W2RAP_ERR_INTERNAL, "circle ranking did not converge");
            }
            W2R_LAUNCH(c, k_copy_list, grid(ncyc, 256), 256, 0, list.p, ncyc, x0, cur);
            W2R_CUDA(cudaStreamSynchronize(c.stream));
        }
        const RankState* R = cur;
        SBuf<uint8_t> keep(c, nn); keep.zero();
        W2R_LAUNCH(c, k_strand_decide, grid(nn, 256), 256, 0, st, R, keep.p, flags.p + 2);
        if (d2h_scalar(c, flags.p + 2)) W2R_FAIL(W2RAP_ERR_EDGE_TOO_LONG, "an edge is longer than 2^24 k-mers (reference: KDef offset is 24 bits)");
        // heads of kept strands = edges.  Upper bound: one per solid k-mer.
        uint64_t cap = out->n_solid;
        SBuf<uint32_t> h_node(c, cap), h_n(c, cap);
        SBuf<uint64_t> h_w0(c, cap), h_w1(c, cap);
        W2R_CUDA(cudaMemsetAsync(scal.p + 2, 0, 8, c.stream));
        W2R_LAUNCH(c, k_collect_heads, grid(nn, 256), 256, 0, st, R, keep.p, h_node.p, h_w0.p, h_w1.p, h_n.p, scal.p + 2, cap);
        E = d2h_scalar(c, scal.p + 2);
        if (E > cap) W2R_FAIL(W2RAP_ERR_INTERNAL, "more edges than solid k-mers");
        // deterministic edge order: sorted by sequence == sorted by the first 60 bases (each oriented k-mer heads at most one edge)
        SBuf<uint32_t> perm(c, E), tmp(c, E);
        if (E) {
            W2R_LAUNCH(c, k_rs_iota, grid(E, 256), 256, 0, perm.p, (uint32_t)E);
            SortWord words[2] = {{h_w1.p, 8, 64}, {h_w0.p, 0, 64}};
            radix_sort_perm(c, perm.p, tmp.p, (uint32_t)E, words, 2);
        }
        SBuf<uint32_t> edge_of_head(c, nn); edge_of_head.fill_ff();
        edge_len.alloc(c, E);
        SBuf<uint32_t> nbytes(c, E);
        if (E) W2R_LAUNCH(c, k_assign_edges, grid(E, 256), 256, 0, perm.p, E, h_node.p, h_n.p, edge_of_head.p, edge_len.p, nbytes.p);
        edge_off.alloc(c, E + 1);
        SBuf<unsigned long long> tot(c, 1);
        exclusive_scan<uint32_t, unsigned long long>(c, nbytes.p, E, (unsigned long long*)edge_off.p, tot.p);
        edge_bytes = E ? d2h_scalar(c, tot.p) : 0;
        W2R_CUDA(cudaMemcpyAsync(edge_off.p + E, &edge_bytes, 8, cudaMemcpyHostToDevice, c.stream));
        W2R_CUDA(cudaStreamSynchronize(c.stream));
        edge_bases.alloc(c, (edge_bytes + 3 + 32) & ~3ull);
        edge_bases.zero();
        W2R_LAUNCH(c, k_emit_edges, grid(nn, 256), 256, 0, st, R, edge_of_head.p, edge_off.p, edge_bases.p);
        W2R_CUDA(cudaStreamSynchronize(c.stream));
    }

    // ---- buildHBVFromEdges (paths/long/HBVFromEdges.cc:76-154)
    void hbv_stage() {
        edge_vertices.alloc(c, 4 * E); fwd_xlat.alloc(c, E); rev_xlat.alloc(c, E);
        if (!E) { nv = nh = 0; return; }
        const uint64_t n4 = 4 * E;
        if (n4 >= (1ull << 32)) W2R_FAIL(W2RAP_ERR_OOM, "too many edges for 32-bit end indices");
        SBuf<uint64_t> kh(c, n4), k0(c, n4), k1(c, n4);
        SBuf<uint8_t> is_pal(c, E);
        W2R_LAUNCH(c, k_edge_ends, grid(n4, 128), 128, 0, edge_bases.p, edge_off.p, edge_len.p, E, kh.p, k0.p, k1.p, is_pal.p);
        SBuf<uint32_t> perm(c, n4), tmp(c, n4);
        W2R_LAUNCH(c, k_rs_iota, grid(n4, 256), 256, 0, perm.p, (uint32_t)n4);
        SortWord words[3] = {{k1.p, 10, 64}, {k0.p, 0, 64}, {kh.p, 0, 64}};
        radix_sort_perm(c, perm.p, tmp.p, (uint32_t)n4, words, 3);
        SBuf<uint32_t> flag(c, n4), excl(c, n4), tot(c, 1);
        W2R_LAUNCH(c, k_vertex_flags, grid(n4, 256), 256, 0, perm.p, n4, kh.p, k0.p, k1.p, flag.p);
        exclusive_scan<uint32_t, uint32_t>(c, flag.p, n4, excl.p, tot.p);
        nv = d2h_scalar(c, tot.p);
        W2R_LAUNCH(c, k_scatter_vids, grid(n4, 256), 256, 0, perm.p, n4, flag.p, excl.p, kh.p, k0.p, k1.p, edge_vertices.p);
        SBuf<uint32_t> width(c, E), xl(c, E);
        W2R_LAUNCH(c, k_pal_widths, grid(E, 256), 256, 0, is_pal.p, E, width.p);
        exclusive_scan<uint32_t, uint32_t>(c, width.p, E, xl.p, tot.p);
        nh = d2h_scalar(c, tot.p);
        hcanon.alloc(c, nh); hleft.alloc(c, nh); hright.alloc(c, nh);
        W2R_LAUNCH(c, k_hbv_edges, grid(E, 256), 256, 0, E, is_pal.p, xl.p, edge_vertices.p, fwd_xlat.p, rev_xlat.p, hcanon.p, hleft.p, hright.p);
        from_e.alloc(c, 4 * nv); to_e.alloc(c, 4 * nv); from_n.alloc(c, nv); to_n.alloc(c, nv);
        SBuf<uint32_t> fc(c, nv), tc(c, nv); fc.zero(); tc.zero();
        SBuf<int> bad(c, 1); bad.zero();
        W2R_LAUNCH(c, k_adj_fill, grid(nh, 256), 256, 0, nh, hleft.p, hright.p, from_e.p, to_e.p, fc.p, tc.p, bad.p);
        if (d2h_scalar(c, bad.p)) W2R_FAIL(W2RAP_ERR_INTERNAL, "a vertex has more than four edges on one side");
        W2R_LAUNCH(c, k_adj_sort, grid(nv, 128), 128, 0, nv, hleft.p, hright.p, from_e.p, to_e.p, fc.p, tc.p, from_n.p, to_n.p);
        W2R_CUDA(cudaStreamSynchronize(c.stream));
    }

    // ---- path_reads_OMP (+FixPaths)  (BuildReadQGraph.cc:829-929; large/GapToyTools.cc:322-335)
    void path_stage(SBuf<int32_t>& d_offset, SBuf<uint64_t>& d_path_off, SBuf<int32_t>& d_path_edges, uint64_t* n_path_edges, unsigned long long* pathed,
                    unsigned long long* multipathed) {
        const uint64_t n = dr.n;
        d_offset.alloc(c, n); d_path_off.alloc(c, n + 1);
        *n_path_edges = 0; *pathed = 0; *multipathed = 0;
        if (!n) { W2R_CUDA(cudaMemsetAsync(d_path_off.p, 0, 8, c.stream)); return; }
        // negative-lookup filter (kmer.cuh: KmerBloom), with an L2 persistence window for the duration of the pathing kernel.  Its
        // false-positive rate matters more than full L2 residency: every false positive is a random DRAM fetch in the dictionary.
        SBuf<uint32_t> bloom_words;
        KmerBloom bloom{nullptr, 0};
        if (out->n_solid >= 4096 && !getenv("W2RAP_NO_BLOOM")) {
            static const uint64_t bloom_mb = getenv("W2RAP_BLOOM_MB") ? (uint64_t)atoi(getenv("W2RAP_BLOOM_MB")) : 192;   // measured (config 2): 40 MB 111 ms, 96 MB 100 ms, 192 MB 97.5 ms
            uint64_t bytes = std::min<uint64_t>(bloom_mb << 20, std::max<uint64_t>(4096, out->n_solid * 2));
            if (bytes * 8 >= 2 * out->n_solid) {          // below ~2 bits per key the filter passes most queries: not worth its L2
                bloom_words.alloc(c, bytes / 4); bloom_words.zero();
                bloom = KmerBloom{bloom_words.p, bytes / 4};
                W2R_LAUNCH(c, k_bloom_build, grid(st.size(), 256), 256, 0, st, bloom);
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) cudaGetLastError();
                set_l2_window(bloom_words.p, bytes);
            }
        }
        GraphView g{st, bloom, edge_bases.p, edge_off.p, edge_len.p, fwd_xlat.p, rev_xlat.p, hcanon.p, hleft.p, hright.p, from_e.p, to_e.p, from_n.p, to_n.p};
        const ReadsView rv = dr.view();
        const unsigned block = 128;
        static const int path_occ_grid = getenv("W2RAP_PATH_OCC") ? std::max(12, atoi(getenv("W2RAP_PATH_OCC"))) : 12;
        const unsigned gr = grid(n, block, path_occ_grid);
        const uint32_t qstride = (dr.max_len + 15) & ~15u;
        SBuf<uint8_t> qscratch(c, (size_t)gr * block * std::max<uint32_t>(qstride, 16));
        const uint32_t cap = 24, left_cap = 8;
        SBuf<int32_t> stage(c, n * cap), row_off(c, n);
        SBuf<PathMeta> meta(c, n);
        SBuf<uint32_t> lens(c, n);
        SBuf<unsigned long long> counters(c, 4); counters.zero();
        // resident CTAs per SM the compiler must allow for (register cap): the kernel waits on dependent DRAM fetches, so warps in flight matter
        // measured (config 2, path stage): 8 CTAs/SM (64 registers) 96 ms, 10: 105 ms, 12 (40 registers, more spills) 81.5 ms, 14/16: 86 ms
        static const int path_occ = getenv("W2RAP_PATH_OCC") ? atoi(getenv("W2RAP_PATH_OCC")) : 12;
        auto launch_path = [&](unsigned grd, const uint32_t* list, uint64_t rows, int32_t* stg, uint32_t cp, uint32_t lcp, int32_t* roff, PathMeta* mt) {
            if (path_occ >= 16) W2R_LAUNCH(c, k_path_reads<16>, grd, block, 0, rv, g, list, rows, qscratch.p, qstride, stg, cp, lcp, roff, mt, prm.apply_fixpaths);
            else if (path_occ >= 14) W2R_LAUNCH(c, k_path_reads<14>, grd, block, 0, rv, g, list, rows, qscratch.p, qstride, stg, cp, lcp, roff, mt, prm.apply_fixpaths);
            else if (path_occ >= 12) W2R_LAUNCH(c, k_path_reads<12>, grd, block, 0, rv, g, list, rows, qscratch.p, qstride, stg, cp, lcp, roff, mt, prm.apply_fixpaths);
            else if (path_occ >= 10) W2R_LAUNCH(c, k_path_reads<10>, grd, block, 0, rv, g, list, rows, qscratch.p, qstride, stg, cp, lcp, roff, mt, prm.apply_fixpaths);
            else W2R_LAUNCH(c, k_path_reads<8>, grd, block, 0, rv, g, list, rows, qscratch.p, qstride, stg, cp, lcp, roff, mt, prm.apply_fixpaths);
        };
        launch_path(gr, nullptr, n, stage.p, cap, left_cap, row_off.p, meta.p);
        W2R_LAUNCH(c, k_path_lens, grid(n, 256), 256, 0, meta.p, (const uint32_t*)nullptr, n, lens.p, counters.p);
        unsigned long long cnt[3];
        W2R_CUDA(cudaMemcpyAsync(cnt, counters.p, 24, cudaMemcpyDeviceToHost, c.stream));
        W2R_CUDA(cudaStreamSynchronize(c.stream));
        // rows that overflowed the small staging row are pathed again with a row that always suffices
        uint64_t n_ovf = cnt[2];
        SBuf<uint32_t> olist(c, n_ovf);
        const uint32_t cap2 = 3 * dr.max_len + 32, left2 = dr.max_len + 16;
        SBuf<int32_t> stage2(c, n_ovf * cap2), row_off2(c, n_ovf);
        SBuf<PathMeta> meta2(c, n_ovf);
        if (n_ovf) {
            W2R_CUDA(cudaMemsetAsync(counters.p + 3, 0, 8, c.stream));
            W2R_LAUNCH(c, k_collect_overflow, grid(n, 256), 256, 0, meta.p, n, olist.p, counters.p + 3);
            launch_path(grid(n_ovf, block, 12), olist.p, n_ovf, stage2.p, cap2, left2, row_off2.p, meta2.p);
            W2R_CUDA(cudaMemsetAsync(counters.p + 2, 0, 8, c.stream));
            W2R_LAUNCH(c, k_path_lens, grid(n_ovf, 256), 256, 0, meta2.p, (const uint32_t*)olist.p, n_ovf, lens.p, counters.p);
            W2R_CUDA(cudaMemcpyAsync(cnt, counters.p, 24, cudaMemcpyDeviceToHost, c.stream));
            W2R_CUDA(cudaStreamSynchronize(c.stream));
            if (cnt[2]) W2R_FAIL(W2RAP_ERR_INTERNAL, "a read path overflowed the worst-case staging row");
        }
        *pathed = cnt[0]; *multipathed = cnt[1];
        SBuf<unsigned long long> tot(c, 1);
        exclusive_scan<uint32_t, unsigned long long>(c, lens.p, n, (unsigned long long*)d_path_off.p, tot.p);
        *n_path_edges = d2h_scalar(c, tot.p);
        W2R_CUDA(cudaMemcpyAsync(d_path_off.p + n, n_path_edges, 8, cudaMemcpyHostToDevice, c.stream));
        W2R_CUDA(cudaStreamSynchronize(c.stream));
        d_path_edges.alloc(c, *n_path_edges);
        W2R_LAUNCH(c, k_path_gather, grid(n, 256), 256, 0, stage.p, cap, meta.p, (const uint32_t*)nullptr, row_off.p, d_path_off.p, n, d_path_edges.p, d_offset.p);
        if (bloom.words) {
            set_l2_window(nullptr, 0);
            W2R_CUDA(cudaStreamSynchronize(c.stream));
            if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
        }
        if (n_ovf) W2R_LAUNCH(c, k_path_gather, grid(n_ovf, 256), 256, 0, stage2.p, cap2, meta2.p, (const uint32_t*)olist.p, row_off2.p, d_path_off.p, n_ovf, d_path_edges.p, d_offset.p);
        W2R_CUDA(cudaStreamSynchronize(c.stream));
    }

    template <class T>
    T* to_host(const T* dptr, size_t n) {
        T* h = out_alloc<T>(owner, n);
        if (n) W2R_CUDA(cudaMemcpyAsync(h, dptr, n * sizeof(T), cudaMemcpyDeviceToHost, c.stream));
        return h;
    }

    void run() {
        W2R_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.verbose = prm.verbose != 0;
        StageTimer total(c), st_t(c);
        total.start();
        out->n_reads = dr.n; out->n_bases = dr.n_bases;
        say(c, "creating kmers from reads...");
        st_t.start(); count_stage();
        out->timings.count_ms = st_t.stop();
        good.release();
        say(c, "updating adjacencies");
        st_t.start();
        W2R_LAUNCH(c, k_adjacency, grid(st.size(), 256), 256, 0, st);
        out->timings.adjacency_ms = st_t.stop();
        say(c, "finding edges (unique paths)");
        st_t.start(); unipath_stage(); out->timings.unipath_ms = st_t.stop();
        say(c, "building graph...");
        st_t.start(); hbv_stage(); out->timings.hbv_ms = st_t.stop();
        SBuf<int32_t> d_offset, d_path_edges; SBuf<uint64_t> d_path_off;
        uint64_t npe = 0; unsigned long long pathed = 0, multi = 0;
        if (prm.want_paths) {
            say(c, "pathing reads into graph...");
            st_t.start(); path_stage(d_offset, d_path_off, d_path_edges, &npe, &pathed, &multi); out->timings.path_ms = st_t.stop();
            say(c, "%llu / %llu reads pathed, %llu spanning junctions", pathed, (unsigned long long)dr.n, multi);
        }
        // ---- results to the host
        st_t.start();
        out->n_edges = E; out->n_vertices = nv; out->n_hbv_edges = nh;
        out->edge_off = to_host<uint64_t>(edge_off.p, E + 1);
        out->edge_len = to_host<uint32_t>(edge_len.p, E);
        out->edge_bases = to_host<uint8_t>(edge_bases.p, edge_bytes);
        out->edge_vertices = to_host<int32_t>(edge_vertices.p, 4 * E);
        out->fwd_xlat = to_host<int32_t>(fwd_xlat.p, E);
        out->rev_xlat = to_host<int32_t>(rev_xlat.p, E);
        if (prm.want_paths) {
            out->n_paths = dr.n; out->n_path_edges = npe; out->n_pathed = pathed; out->n_multipathed = multi;
            out->path_offset = to_host<int32_t>(d_offset.p, dr.n);
            out->path_off = to_host<uint64_t>(d_path_off.p, dr.n + 1);
            out->path_edges = to_host<int32_t>(d_path_edges.p, npe);
        }
        if (prm.dump_kmers == 1 && out->n_solid) {
            SBuf<DumpRec> dd(c, out->n_solid);
            SBuf<unsigned long long> cur(c, 1); cur.zero();
            W2R_LAUNCH(c, k_dump_solid, grid(st.size(), 256), 256, 0, st, dd.p, cur.p);
            dump_host.resize(out->n_solid);
            W2R_CUDA(cudaMemcpyAsync(dump_host.data(), dd.p, out->n_solid * sizeof(DumpRec), cudaMemcpyDeviceToHost, c.stream));
            W2R_CUDA(cudaStreamSynchronize(c.stream));
        }
        W2R_CUDA(cudaStreamSynchronize(c.stream));
        out->timings.d2h_ms = st_t.stop();
        if (E == 0 && out->edge_off) out->edge_off[0] = 0;
        uint64_t neb = 0;
        for (uint64_t i = 0; i < E; ++i) neb += out->edge_len[i];
        out->n_edge_bases = neb;
        if (prm.dump_kmers && !dump_host.empty()) {
            std::sort(dump_host.begin(), dump_host.end(), [](const DumpRec& a, const DumpRec& b) { return a.w0 < b.w0 || (a.w0 == b.w0 && a.w1 < b.w1); });
            out->n_dump = dump_host.size();
            out->dump = out_alloc<w2rap_kmer_rec>(owner, dump_host.size());
            static_assert(sizeof(DumpRec) == sizeof(w2rap_kmer_rec), "dump record layout");
            memcpy(out->dump, dump_host.data(), dump_host.size() * sizeof(DumpRec));
        }
        out->timings.total_ms = total.stop();
        out->timings.kernel_launches = c.launches;
        out->timings.count_launches = c.count_launches;
        say(c, "%llu edges of total length %llu; %llu vertices", (unsigned long long)E, (unsigned long long)neb, (unsigned long long)nv);
    }

    ~Pipeline() {
        // free stream-ordered buffers before the stream goes away
        good.release(); solid_slots.release(); edge_bases.release(); edge_off.release(); edge_len.release();
        edge_vertices.release(); fwd_xlat.release(); rev_xlat.release(); hleft.release(); hright.release(); from_e.release(); to_e.release();
        hcanon.release(); from_n.release(); to_n.release();
        if (c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); }
    }
};

static void validate_reads(const w2rap_reads* in) {
    if (!in) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null read store");
    if (in->n_reads && (!in->bases || !in->base_off || !in->len || !in->quals || !in->qual_off)) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null pointer in read store");
    if (in->n_reads >= (1ull << 32)) W2R_FAIL(W2RAP_ERR_BAD_ARG, "more than 2^32-1 reads on one device");
}

// Device buffers for the read stores of the host-buffer entry points.  They are kept between calls (grow-only, one set per
// device): taking them from the stream-ordered pool instead fragments it, and the 80+ GB record pool of the counting stage then
// has to be re-created by the driver on every call (measured: +120 ms per step).
struct UploadArena {
    std::mutex mu;
    bool in_use = false;
    void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0};
    static UploadArena& get(int device) { static UploadArena* a = new UploadArena[16]; return a[device & 15]; }
};
void arena_release(int device) { UploadArena& a = UploadArena::get(device); std::lock_guard<std::mutex> g(a.mu); a.in_use = false; }
static void dev_alloc(DeviceReads* d, void** p, size_t bytes, cudaStream_t s, int which) {
    if (d->arena) {
        UploadArena& a = UploadArena::get(d->device);
        if (a.cap[which] < bytes) {
            if (a.buf[which]) cudaFree(a.buf[which]);
            a.buf[which] = nullptr; a.cap[which] = 0;
            size_t want = bytes + bytes / 16;
            W2R_CUDA(cudaMalloc(&a.buf[which], want));
            a.cap[which] = want;
        }
        *p = a.buf[which];
        return;
    }
    if (d->pooled) W2R_CUDA(cudaMallocAsync(p, bytes, s)); else W2R_CUDA(cudaMalloc(p, bytes));
}
// Validates the flattened store and starts its transfer.  With `batched`, the copy is issued as up to 8 read batches with an
// event each and NOT waited for: the pipeline consumes batch b while batch b+1 is in flight.  Host buffers must stay valid until
// the stream has drained (the entry points synchronise before returning).
static void upload(const w2rap_reads* in, int device, DeviceReads* d, cudaStream_t s, bool batched) {
    d->device = device;
    d->n = in->n_reads;
    const uint64_t n = d->n;
    d->bases_bytes = n ? in->base_off[n] : 0;
    d->quals_bytes = n ? in->qual_off[n] : 0;
    const double t0 = now_ms();
    {   // one pass over the offsets/lengths, on a few host threads
        const unsigned nt = n > (1u << 20) ? 8u : 1u;
        std::vector<uint64_t> t_nb(nt, 0), t_inst(nt, 0), t_bad(nt, ~0ull);
        std::vector<uint32_t> t_mx(nt, 0), t_kind(nt, 0);
        auto work = [&](unsigned t) {
            uint64_t lo = n * t / nt, hi = n * (t + 1) / nt, nb = 0, inst = 0; uint32_t mx = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                uint32_t L = in->len[i];
                nb += L; if (L > mx) mx = L;
                if (L > 59) inst += L - 59;
                if (in->base_off[i + 1] < in->base_off[i] || in->base_off[i + 1] - in->base_off[i] < (uint64_t)(L + 3) / 4) { t_bad[t] = i; t_kind[t] = 1; break; }
                if (in->qual_off[i + 1] <= in->qual_off[i]) { t_bad[t] = i; t_kind[t] = 2; break; }
            }
            t_nb[t] = nb; t_inst[t] = inst; t_mx[t] = mx;
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        uint64_t nb = 0, inst = 0; uint32_t mx = 0;
        for (unsigned t = 0; t < nt; ++t) {
            if (t_kind[t] == 1) W2R_FAIL(W2RAP_ERR_BAD_ARG, "read %llu: base offsets do not hold its bases", (unsigned long long)t_bad[t]);
            if (t_kind[t] == 2) W2R_FAIL(W2RAP_ERR_BAD_ARG, "read %llu: empty quality stream (at least the terminator byte is required)", (unsigned long long)t_bad[t]);
            nb += t_nb[t]; inst += t_inst[t]; mx = std::max(mx, t_mx[t]);
        }
        if (mx > 65535u) W2R_FAIL(W2RAP_ERR_BAD_ARG, "reads longer than 65535 bases are not supported (the reference stores good lengths in uint16_t)");
        d->n_bases = nb; d->max_len = mx; d->n_inst_upper = inst;
    }
    const double t1 = now_ms();
    // +32 bytes of padding: packed bases are read with aligned 8-byte loads that may touch a few bytes past a read
    if (d->pooled) {        // host-buffer entry point: try to take the per-device arena
        UploadArena& a = UploadArena::get(device);
        std::lock_guard<std::mutex> g(a.mu);
        if (!a.in_use) { a.in_use = true; d->arena = true; }
    }
    dev_alloc(d, (void**)&d->bases, d->bases_bytes + 32, s, 0);
    dev_alloc(d, (void**)&d->quals, d->quals_bytes + 32, s, 1);
    dev_alloc(d, (void**)&d->base_off, (n + 1) * 8, s, 2);
    dev_alloc(d, (void**)&d->qual_off, (n + 1) * 8, s, 3);
    dev_alloc(d, (void**)&d->len, (n + 1) * 4, s, 4);
    W2R_CUDA(cudaMemsetAsync(d->bases + d->bases_bytes, 0, 32, s));
    W2R_CUDA(cudaMemsetAsync(d->quals + d->quals_bytes, 0, 32, s));
    const double t2 = now_ms();
    if (n) {
        const unsigned nbatch = (batched && n >= (1u << 20)) ? 8u : 1u;
        for (unsigned b = 0; b < nbatch; ++b) {
            const uint64_t r0 = n * b / nbatch, r1 = n * (b + 1) / nbatch;
            const uint64_t b0 = in->base_off[r0], b1 = in->base_off[r1], q0 = in->qual_off[r0], q1 = in->qual_off[r1];
            W2R_CUDA(cudaMemcpyAsync(d->len + r0, in->len + r0, (r1 - r0) * 4, cudaMemcpyHostToDevice, s));
            W2R_CUDA(cudaMemcpyAsync(d->base_off + r0, in->base_off + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, s));
            W2R_CUDA(cudaMemcpyAsync(d->qual_off + r0, in->qual_off + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, s));
            W2R_CUDA(cudaMemcpyAsync(d->bases + b0, in->bases + b0, b1 - b0, cudaMemcpyHostToDevice, s));
            W2R_CUDA(cudaMemcpyAsync(d->quals + q0, in->quals + q0, q1 - q0, cudaMemcpyHostToDevice, s));
            if (batched) {
                cudaEvent_t e;
                W2R_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                W2R_CUDA(cudaEventRecord(e, s));
                d->batch_ready.push_back(e);
                d->batch_first.push_back(r0);
            }
        }
        if (batched) d->batch_first.push_back(n);
    } else {
        uint64_t z = 0;
        W2R_CUDA(cudaMemcpyAsync(d->base_off, &z, 8, cudaMemcpyHostToDevice, s));
        W2R_CUDA(cudaMemcpyAsync(d->qual_off, &z, 8, cudaMemcpyHostToDevice, s));
    }
    if (!batched) W2R_CUDA(cudaStreamSynchronize(s));
    if (getenv("W2RAP_TRACE")) fprintf(stderr, "[w2rap] upload: validate %.1f ms, alloc %.1f ms, enqueue %.1f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
}

static void check_params(const w2rap_params* p) {
    if (!p) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null params");
    if (p->abi_version != W2RAP_STEP2_ABI_VERSION) W2R_FAIL(W2RAP_ERR_BAD_ARG, "ABI version %u, library has %d", p->abi_version, W2RAP_STEP2_ABI_VERSION);
    if (p->K != W2RAP_K) W2R_FAIL(W2RAP_ERR_BAD_ARG, "K=%u: only K=60 is built (the reference hard-wires it, BuildReadQGraph.cc:51)", p->K);
    if (p->min_freq == 0 || p->min_freq > 255) W2R_FAIL(W2RAP_ERR_BAD_ARG, "min_freq must be in 1..255 (counts saturate at 255)");
}

}  // namespace w2r

struct w2rap_comm { int world, rank, device; w2r::ncclComm_t comm; };

namespace w2r {
static void run_on_device(DeviceReads& dr, const w2rap_params* p, w2rap_graph* out, float h2d_ms, const w2rap_comm* cm = nullptr, double t_entry = 0) {
    GraphOwner* owner = new GraphOwner();
    memset(out, 0, sizeof(*out));
    out->_owner = owner;
    try {
        Pipeline pl(dr, *p, out, owner);
        if (cm) { pl.world = cm->world; pl.rank = cm->rank; pl.comm = cm->comm; }
        check_device(dr.device, pl.c);
        const double t_run = now_ms();
        pl.run();
        out->timings.h2d_ms = h2d_ms;
        out->timings.total_ms += h2d_ms;
        out->timings.host_pre_ms = t_entry ? (float)(t_run - t_entry) : 0.f;
        out->timings.wall_ms = (float)(now_ms() - (t_entry ? t_entry : t_run));
    } catch (...) {
        delete owner;
        memset(out, 0, sizeof(*out));
        throw;
    }
}

}  // namespace w2r

// ================================================================ C ABI
using namespace w2r;

static int fail(const Error& e, char* err, size_t errlen) {
    if (err && errlen) { snprintf(err, errlen, "%s", e.msg.c_str()); }
    return e.code;
}
#define W2R_API_BEGIN try {
#define W2R_API_END                                                                                                  \
    }                                                                                                                \
    catch (const Error& e) { return fail(e, err, errlen); }                                                          \
    catch (const std::exception& e) { return fail(Error{W2RAP_ERR_INTERNAL, e.what()}, err, errlen); }               \
    return W2RAP_OK;

extern "C" {

int w2rap_step2_abi_version(void) { return W2RAP_STEP2_ABI_VERSION; }
const char* w2rap_step2_build_info(void) { return "w2rap step2 B200 (sm_100a), K=60, built " __DATE__ " " __TIME__; }

int w2rap_step2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major >= 10) ++ok; }
    return ok;
}

int w2rap_step2_upload(const w2rap_reads* in, int device, w2rap_device_reads** handle, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!handle) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null handle");
    validate_reads(in);
    Ctx c; check_device(device, c);
    w2rap_device_reads* h = new w2rap_device_reads();
    try { upload(in, c.device, &h->d, 0, false); } catch (...) { h->d.release(); delete h; throw; }
    *handle = h;
    W2R_API_END
}

void w2rap_step2_release(w2rap_device_reads* handle) {
    if (!handle) return;
    cudaSetDevice(handle->d.device);
    handle->d.release();
    delete handle;
}

int w2rap_step2_run_resident(w2rap_device_reads* handle, const w2rap_params* p, w2rap_graph* out, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!handle || !out) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null argument");
    check_params(p);
    run_on_device(handle->d, p, out, 0.f);
    if (p->workdir && p->workdir[0]) { std::string f = std::string(p->workdir) + "/small_K.freqs"; int rc = w2rap_write_freqs(f.c_str(), out, err, errlen); if (rc) return rc; }
    W2R_API_END
}

int w2rap_step2_run(const w2rap_reads* in, const w2rap_params* p, w2rap_graph* out, char* err, size_t errlen) {
    W2R_API_BEGIN
    const double t_entry = now_ms();
    if (!out) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null output");
    check_params(p);
    validate_reads(in);
    Ctx c; check_device(p->device, c);
    DeviceReads d;
    d.pooled = true;
    cudaStream_t us = nullptr;
    W2R_CUDA(cudaStreamCreateWithFlags(&us, cudaStreamNonBlocking));
    d.pool_stream = us;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float h2d = 0;
    try {
        cudaEventRecord(e0, us);
        upload(in, c.device, &d, us, true);
        cudaEventRecord(e1, us);
        run_on_device(d, p, out, 0.f, nullptr, t_entry);          // consumes the batches as they land
        cudaEventSynchronize(e1); cudaEventElapsedTime(&h2d, e0, e1);
        out->timings.h2d_ms = h2d;              // overlapped with the quality floor + extraction, already inside total_ms
    } catch (...) { d.release(); cudaStreamSynchronize(us); cudaStreamDestroy(us); cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
    const double t_post = now_ms();
    d.release(); cudaStreamSynchronize(us); cudaStreamDestroy(us); cudaEventDestroy(e0); cudaEventDestroy(e1);
    out->timings.host_post_ms = (float)(now_ms() - t_post);
    out->timings.wall_ms = (float)(now_ms() - t_entry);
    if (p->workdir && p->workdir[0]) { std::string f = std::string(p->workdir) + "/small_K.freqs"; int rc = w2rap_write_freqs(f.c_str(), out, err, errlen); if (rc) return rc; }
    W2R_API_END
}

int w2rap_step2_comm_unique_id(uint8_t* id128, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!id128) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null id buffer");
    NcclApi& n = NcclApi::get();
    if (n.error) W2R_FAIL(W2RAP_ERR_NO_DEVICE, "%s", n.error);
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) W2R_FAIL(W2RAP_ERR_CUDA, "ncclGetUniqueId failed");
    memcpy(id128, id.internal, 128);
    W2R_API_END
}

int w2rap_step2_comm_init(const uint8_t* id128, int world, int rank, int device, w2rap_comm** comm, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) W2R_FAIL(W2RAP_ERR_BAD_ARG, "bad communicator arguments");
    if (world & (world - 1)) W2R_FAIL(W2RAP_ERR_BAD_ARG, "world size must be a power of two (partition ranges are split evenly)");
    NcclApi& n = NcclApi::get();
    if (n.error) W2R_FAIL(W2RAP_ERR_NO_DEVICE, "%s", n.error);
    Ctx c; check_device(device, c);
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t nc = nullptr;
    ncclResult_t r = n.CommInitRank(&nc, world, id, rank);
    if (r != ncclSuccess) W2R_FAIL(W2RAP_ERR_CUDA, "ncclCommInitRank failed: %s", n.GetErrorString(r));
    *comm = new w2rap_comm{world, rank, c.device, nc};
    W2R_API_END
}

void w2rap_step2_comm_destroy(w2rap_comm* comm) {
    if (!comm) return;
    cudaSetDevice(comm->device);
    if (comm->comm) NcclApi::get().CommDestroy(comm->comm);
    delete comm;
}

int w2rap_step2_run_sharded_resident(w2rap_device_reads* handle, const w2rap_params* p, w2rap_comm* comm, w2rap_graph* out, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!handle || !out || !comm) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null argument");
    check_params(p);
    if (handle->d.device != comm->device) W2R_FAIL(W2RAP_ERR_BAD_ARG, "read shard lives on device %d, communicator on %d", handle->d.device, comm->device);
    run_on_device(handle->d, p, out, 0.f, comm);
    if (comm->rank == 0 && p->workdir && p->workdir[0]) { std::string f = std::string(p->workdir) + "/small_K.freqs"; int rc = w2rap_write_freqs(f.c_str(), out, err, errlen); if (rc) return rc; }
    W2R_API_END
}

int w2rap_step2_run_sharded(const w2rap_reads* shard, const w2rap_params* p, w2rap_comm* comm, w2rap_graph* out, char* err, size_t errlen) {
    W2R_API_BEGIN
    if (!out || !comm) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null argument");
    check_params(p);
    validate_reads(shard);
    Ctx c; check_device(comm->device, c);
    DeviceReads d;
    d.pooled = true;
    cudaStream_t us = nullptr;
    W2R_CUDA(cudaStreamCreateWithFlags(&us, cudaStreamNonBlocking));
    d.pool_stream = us;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float h2d = 0;
    try {
        cudaEventRecord(e0, us);
        upload(shard, c.device, &d, us, true);
        cudaEventRecord(e1, us);
        run_on_device(d, p, out, 0.f, comm);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&h2d, e0, e1);
        out->timings.h2d_ms = h2d;
    } catch (...) { d.release(); cudaStreamSynchronize(us); cudaStreamDestroy(us); cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
    d.release(); cudaStreamSynchronize(us); cudaStreamDestroy(us); cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (comm->rank == 0 && p->workdir && p->workdir[0]) { std::string f = std::string(p->workdir) + "/small_K.freqs"; int rc = w2rap_write_freqs(f.c_str(), out, err, errlen); if (rc) return rc; }
    W2R_API_END
}

void w2rap_step2_free(w2rap_graph* out) {
    if (!out) return;
    delete (GraphOwner*)out->_owner;
    memset(out, 0, sizeof(*out));
}

}  // extern "C"
