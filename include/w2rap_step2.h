/*
 * w2rap_step2.h — C ABI of the B200-native step-2 (K=60 read de Bruijn graph) path.
 *
 * This is the drop-in boundary for w2rap-contigger's
 *     void buildReadQGraph(vecbvec const& reads, VecPQVec const& quals, bool doFillGaps, bool doJoinOverlaps,
 *                          unsigned minQual, unsigned minFreq, double minFreq2Fract, unsigned maxGapSize,
 *                          HyperBasevector* pHBV, ReadPathVec* pPaths, int _K, std::string workdir,
 *                          std::string tmpdir, unsigned char disk_batches)
 *     (reference: src/paths/long/BuildReadQGraph.h:24-29, called once at src/modules/w2rap-contigger.cc:338,
 *      followed by FixPaths at :340 = src/paths/long/large/GapToyTools.cc:322-335).
 * The reference has no FFI or plugin registry for this path; a maintainer replaces the translation unit
 * src/paths/long/BuildReadQGraph.cc by a thin C++ wrapper that flattens its inputs into `w2rap_reads`,
 * calls w2rap_step2_run() and rebuilds HyperBasevector / ReadPathVec from `w2rap_graph`
 * (see INTEGRATION.md and w2rap-contigger_b200/host/BuildReadQGraph_b200.cc).
 *
 * Plain C, plain pointers and sizes.  All arithmetic on the path is integer; results are bit-exact with
 * the reference modulo the reference's own racy edge numbering (edges are emitted here in a deterministic
 * order: sorted by sequence).
 *
 * There is NO CPU fallback: every entry point that computes returns W2RAP_ERR_NO_DEVICE when no sm_100
 * device is usable.
 */
#ifndef W2RAP_STEP2_H_
#define W2RAP_STEP2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2RAP_STEP2_ABI_VERSION 2
#define W2RAP_K 60

/* status codes (0 = ok).  The reference aborts (FatalErr / ForceAssert / CRD::exit(1)) where these are
 * returned; the C++ wrapper turns a non-zero status into the same abort. */
enum {
    W2RAP_OK = 0,
    W2RAP_ERR_BAD_ARG = 1,       /* K != 60, null pointers, inconsistent offsets, read longer than 65535 */
    W2RAP_ERR_NO_DEVICE = 2,     /* no CUDA device / not sm_100 — there is no CPU path */
    W2RAP_ERR_CUDA = 3,          /* a CUDA runtime call failed (message in err) */
    W2RAP_ERR_OOM = 4,           /* device or pinned-host allocation failed */
    W2RAP_ERR_EDGE_TOO_LONG = 5, /* an edge exceeds 2^24-1 k-mers (reference: kmers/ReadPather.h:121-122,144) */
    W2RAP_ERR_INTERNAL = 6,      /* invariant violated (the reference would ForceAssert) */
    W2RAP_ERR_IO = 7
};

/*
 * Input reads: the flattened form of the reference's two in-memory read stores.
 *
 *  bases    = every read's 2-bit packed bases, LSB-first within a byte (base i of a read -> byte i/4,
 *             bits 2*(i%4)..+1; A=0 C=1 G=2 T=3), each read starting on a byte boundary
 *             (reference: feudal/FieldVec.h:765-769,788-794; this is byte-for-byte the variable-data
 *             section of a .fastb feudal file, feudal/FeudalFileWriter.cc:24-38).
 *  base_off = n_reads+1 byte offsets into `bases`; read i occupies [base_off[i], base_off[i+1]),
 *             which must hold at least ceil(len[i]/4) bytes.
 *  len      = n_reads read lengths in bases (the .fastb fixed-data section, u32 each).
 *  quals    = every read's PQVec block stream (reference: feudal/PQVec.cc:87-188: per block
 *             {u8 nQs, 3-bit nBits, 6-bit minQ, nQs x nBits-bit deltas LSB-first, padded to a byte},
 *             terminated by a 0 byte); byte-for-byte the variable-data section of a .qualp file.
 *  qual_off = n_reads+1 byte offsets into `quals`.
 *
 * All pointers are HOST pointers for w2rap_step2_run(); nothing is modified (the reference passes both
 * stores as const& and reuses them in steps 4-6).
 */
typedef struct w2rap_reads {
    uint64_t n_reads;
    const uint8_t* bases;
    const uint64_t* base_off;
    const uint32_t* len;
    const uint8_t* quals;
    const uint64_t* qual_off;
} w2rap_reads;

/* Parameters = the arguments of buildReadQGraph that the driver actually varies
 * (src/modules/w2rap-contigger.cc:336-338: doFillGaps=doJoinOverlaps=false, minFreq2Fract=.75, maxGapSize=0). */
typedef struct w2rap_params {
    uint32_t abi_version;    /* W2RAP_STEP2_ABI_VERSION */
    uint32_t K;              /* must be 60 (reference hard-wires K, BuildReadQGraph.cc:51) */
    uint32_t min_qual;       /* --min_qual, default 7 */
    uint32_t min_freq;       /* --min_freq, default 4 */
    uint32_t want_paths;     /* 0: graph only (pPaths == nullptr, BuildReadQGraph.cc:1300-1307) */
    uint32_t apply_fixpaths; /* 1: also apply FixPaths (file-level drop-in); 0: paths as buildReadQGraph returns them */
    uint32_t dump_kmers;     /* test hook: 0 none, 1 solid k-mers (pruned ctx, edge, offset), 2 all distinct k-mers (count, raw ctx) */
    int32_t device;          /* CUDA device ordinal, -1 = current */
    const char* workdir;     /* if non-null and non-empty: write <workdir>/small_K.freqs (BuildReadQGraph.cc:1108-1112) */
    uint64_t table_slots;    /* 0 = auto; otherwise force the size of the counting tables (test hook: tiny tables push partitions
                                through every fallback of the reduce) */
    uint32_t verbose;        /* 1: reference-style progress lines on stdout */
    uint32_t force_passes;   /* test hook: 0 = auto; n > 0 = count in exactly n hash-range passes (the replacement of the
                                reference's disk batches, BuildReadQGraph.cc:1120-1250; normally chosen from free device memory) */
    uint32_t graph_on_root_only; /* sharded runs: 0 = every rank receives the graph arrays; 1 = only rank 0 does (the process that
                                writes the .hbv): the other ranks get the counters, edge_len, digests and their own paths, with
                                edge_off/edge_bases/edge_vertices/fwd_xlat/rev_xlat/involution null.  Eight identical copies of a
                                1 Gbp graph (1.3 GB each) otherwise share the host's PCIe uplinks with the path arrays. */
    uint32_t places_K2;      /* 0 = off; otherwise also build step 3's "places" for large K = places_K2 (w2rap-contigger -K, default
                                200) from the finished paths: RepathInMemory's first block (paths/long/large/Repath.cc:46-72).  Needs
                                want_paths and apply_fixpaths (step 3 reads the paths as main() left them, after FixPaths). */
} w2rap_params;

/* One record of the optional k-mer dump (sorted by k-mer). */
typedef struct w2rap_kmer_rec {
    uint64_t w0, w1;      /* reference KMer<60> words: base 0 in bits 63-62 of w0, bases 32-59 in bits 63-8 of w1 (kmers/KMer.h:155-162) */
    uint32_t count;       /* min(255, occurrences) (dump level 2), 0 in level 1 */
    uint32_t ctx;         /* KMerContext byte: pred mask << 4 | succ mask (kmers/KMerContext.h:23-121) */
    uint32_t edge;        /* canonical edge index (level 1), 0xffffffff otherwise */
    uint32_t offset;      /* k-mer offset in that edge (level 1) */
} w2rap_kmer_rec;

/* indices into w2rap_timings.kernel_ms */
enum {
    W2RAP_KT_GOOD_LEN = 0, W2RAP_KT_MAP = 1, W2RAP_KT_SCATTER = 2, W2RAP_KT_REDUCE = 3, W2RAP_KT_INSERT_SOLID = 4,
    W2RAP_KT_ADJACENCY = 5, W2RAP_KT_LINKS = 6, W2RAP_KT_SPLITTER_WALK = 7, W2RAP_KT_SPLITTER_FINISH = 8, W2RAP_KT_EMIT_EDGES = 9,
    W2RAP_KT_UNUSED_10 = 10 /* (was the separate filter build; the filter bits are now set by k_insert_solid) */, W2RAP_KT_PATH_READS = 11,
    /* sharded graph stage, phases (kernels + the exchanges between them) */
    W2RAP_KT_SG_QUERIES = 12,   /* neighbour queries, answers, ghost entries, ghost contexts (without k_adjacency) */
    W2RAP_KT_SG_PIECES = 13,    /* chain-end records: emit, all-gather, link, rank */
    W2RAP_KT_SG_EDGES = 14,     /* strands, edge ids, all-reduce of the edge bases (without k_emit_edges) */
    W2RAP_KT_SG_DICT = 15,      /* finished entries gathered into the pathing dictionary (without k_insert_solid) */
    W2RAP_KT_COUNT
};

/* Stage timings, device milliseconds from CUDA events on the stream the kernels run on. */
typedef struct w2rap_timings {
    float h2d_ms, count_ms, solid_ms, adjacency_ms, unipath_ms, hbv_ms, path_ms, d2h_ms, total_ms;
    float count_kernel_ms;        /* the map: k_minimizer_map + scan + k_scatter_records, all read batches */
    float region_ms;              /* the reduce: k_count_smem + all k_count_region / k_scan_region launches */
    float exchange_ms;            /* multi-GPU only: NCCL all-to-all of k-mer records + all-gather of the solid records */
    float host_pre_ms;            /* host wall time from entry to the first pipeline launch (validation, allocation, copy enqueue) */
    float host_post_ms;           /* host wall time after the pipeline finished (buffer release, stream teardown) */
    float wall_ms;                /* host wall time of the whole call */
    uint32_t count_launches;      /* launches of the map kernel (one per read batch and counting pass) */
    uint32_t kernel_launches;     /* all kernels launched by this call */
    uint32_t count_passes;        /* reduce launches: k_count_smem launches + partition groups that went through the counting region */
    uint32_t reserved;
    float dict_ms;                /* dictionary build (k_insert_solid; sharded: + gather of the records it is built from) */
    float graph_exchange_ms;      /* multi-GPU only: the exchanges of the sharded graph stage (neighbour queries, chain ends, edges) */
    uint64_t exchange_bytes;      /* multi-GPU only: bytes this rank sent over NVLink in the whole step */
    uint64_t n_records;           /* super-k-mer records this rank's map produced (32 bytes each) */
    /* CUDA-event time of individual kernels (all launches of the kernel in this call), index = W2RAP_KT_*; 0 where not run */
    float kernel_ms[16];
    float alloc_host_ms;          /* host time spent inside the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync) during this call:
                                     ~0 when the pool serves every request from cached memory, large when it has to grow, trim or re-map */
    uint32_t reserved2;
    uint64_t count_exchange_bytes; /* of exchange_bytes, the super-k-mer records of the counting stage (what exchange_ms moved) */
    float places_ms;              /* params.places_K2: construction + unique sort of the places */
    uint32_t reserved3;
} w2rap_timings;

/*
 * Output graph + paths.  All arrays are allocated by the library and released by w2rap_step2_free().
 *
 *  edges: the reference's `vecbvec edges` (BuildReadQGraph.cc:1282-1284) — each sequence in FWD/PALINDROME
 *         orientation — sorted by sequence; edge i occupies edge_bases[edge_off[i] .. edge_off[i+1]) packed
 *         like a bvec (2-bit, LSB-first in byte), edge_len[i] bases.
 *  edge_vertices[4*i..]: fw_v1, fw_v2, rc_v1, rc_v2 of buildHBVFromEdges (paths/long/HBVFromEdges.cc:61-63,
 *         99-122); rc_* = -1 for a palindromic edge.
 *  fwd_xlat/rev_xlat: HBV edge ids (HBVFromEdges.cc:129-151): edge i -> fwd id, then rc id (same id if palindrome).
 *  paths: ReadPath per read (paths/long/ReadPath.h:25-58): path_offset[r], edge ids
 *         path_edges[path_off[r] .. path_off[r+1]).
 */
typedef struct w2rap_graph {
    /* counters the reference prints (BuildReadQGraph.cc:1091,1106,325,1323) */
    uint64_t n_reads;
    uint64_t n_bases;           /* sum of read lengths */
    uint64_t n_kmer_instances;  /* k-mers extracted from quality-floored reads */
    uint64_t n_distinct;
    uint64_t n_solid;
    uint64_t hist[101];         /* hist[min(100,count)], bins 1..100 are the lines of small_K.freqs */

    uint64_t n_edges;
    uint64_t n_edge_bases;      /* sum of edge_len */
    uint64_t* edge_off;         /* n_edges+1 */
    uint32_t* edge_len;         /* n_edges */
    uint8_t* edge_bases;

    uint64_t n_vertices;
    uint64_t n_hbv_edges;
    int32_t* edge_vertices;     /* 4*n_edges */
    int32_t* fwd_xlat;          /* n_edges */
    int32_t* rev_xlat;          /* n_edges */
    int32_t* involution;        /* n_hbv_edges: hbv edge id -> id of its reverse complement (itself for a palindrome).  What step 3 asks
                                   the reference for right after step 2 (hbv.Involution, paths/HyperBasevector.cc:648-660: two sorts of
                                   all edges); here it is known from the fwd/rc pairing */

    uint64_t n_paths;           /* n_reads if want_paths else 0 */
    uint64_t n_path_edges;
    int32_t* path_offset;       /* n_paths */
    uint64_t* path_off;         /* n_paths+1 */
    int32_t* path_edges;
    uint64_t n_pathed;          /* paths with >0 edges */
    uint64_t n_multipathed;     /* paths with >2 edges */

    /* Order-sensitive 64-bit digests computed on the device (sum over elements of a mix of index and value), for comparing
     * whole results across runs and GPU counts without shipping them: digest_graph covers hist, edge lengths, edge bases,
     * vertex ids and xlat tables; digest_paths covers (read sequence, path offset, path edges) of every read and is summed over
     * all ranks in a sharded run, so both are independent of the number of GPUs. */
    uint64_t digest_graph;
    uint64_t digest_paths;

    uint64_t n_dump;
    w2rap_kmer_rec* dump;       /* sorted by (w0,w1); null unless params.dump_kmers */

    /* Step-3 input (params.places_K2): every read path that implies at least K2 bases, replaced by its inverse (reversed, each edge
     * by its involution) if that is smaller, then sorted lexicographically and made unique — `places` of RepathInMemory
     * (paths/long/large/Repath.cc:46-72; std::vector<int> order: element by element, a proper prefix first).  In a sharded run the
     * places of all ranks are merged: every rank receives the same list. */
    uint64_t n_places_kept;     /* paths that passed the K2 test ("sorting N places" in the reference's log) */
    uint64_t n_places;          /* unique places ("N unique places") */
    uint64_t n_place_edges;
    uint64_t* place_off;        /* n_places+1 */
    int32_t* place_edges;       /* hbv edge ids */

    w2rap_timings timings;
    void* _owner;               /* library-private */
} w2rap_graph;

/* Library / device introspection (never touches the GPU). */
int w2rap_step2_abi_version(void);
const char* w2rap_step2_build_info(void);

/* Number of usable sm_100 devices (0 if none; never fails). */
int w2rap_step2_device_count(void);

/*
 * The whole of step 2 with HOST buffers: H2D of the read stores, count, adjacency, unipaths, HBV vertices,
 * read pathing, D2H of the graph and paths.  Replaces buildReadQGraph (+ FixPaths if apply_fixpaths).
 * Returns a status code; on failure `err` (if non-null) receives a NUL-terminated message.
 */
int w2rap_step2_run(const w2rap_reads* in, const w2rap_params* p, w2rap_graph* out, char* err, size_t errlen);

/*
 * Device-resident variant used to time the path without host<->device copies:
 * upload once, run many times.  `w2rap_device_reads` is opaque.
 */
typedef struct w2rap_device_reads w2rap_device_reads;
int w2rap_step2_upload(const w2rap_reads* in, int device, w2rap_device_reads** handle, char* err, size_t errlen);
int w2rap_step2_run_resident(w2rap_device_reads* handle, const w2rap_params* p, w2rap_graph* out,
                             char* err, size_t errlen);
void w2rap_step2_release(w2rap_device_reads* handle);

/*
 * Multi-GPU (one process or thread per GPU, NCCL over NVLink/NVSwitch).  Reads are sharded by index by the
 * caller; every super-k-mer record is routed to the rank that owns its minimiser partition with one all-to-all (the reference's
 * MapReduceEngine "swizzle", MapReduceEngine.h:337-358); owners count.  The dictionary, adjacency pruning and the unipath walk stay
 * sharded by the same owners (neighbour queries and one record per local chain are exchanged); only the finished
 * dictionary is replicated, for pathing; each rank paths its own shard.  On return every rank holds the WHOLE graph (identical on
 * all ranks: edges, vertices, xlat, histogram, n_kmer_instances / n_distinct / n_solid and the digests of the whole job) — unless
 * graph_on_root_only — and the paths of ITS shard; n_reads, n_bases, n_paths, n_pathed and n_multipathed count the shard.
 * world must be a power of two.  Rank 0 creates the id and hands the 128 bytes to the others (any transport).
 */
typedef struct w2rap_comm w2rap_comm;
int w2rap_step2_comm_unique_id(uint8_t* id128, char* err, size_t errlen);
int w2rap_step2_comm_init(const uint8_t* id128, int world, int rank, int device, w2rap_comm** comm, char* err, size_t errlen);
void w2rap_step2_comm_destroy(w2rap_comm* comm);
int w2rap_step2_run_sharded(const w2rap_reads* shard, const w2rap_params* p, w2rap_comm* comm, w2rap_graph* out, char* err, size_t errlen);
int w2rap_step2_run_sharded_resident(w2rap_device_reads* shard, const w2rap_params* p, w2rap_comm* comm, w2rap_graph* out,
                                     char* err, size_t errlen);

/* Frees everything w2rap_step2_run* put into `out` and zeroes it. */
void w2rap_step2_free(w2rap_graph* out);

/*
 * Pinned (page-locked) host memory for the flattened read stores.  w2rap_step2_run copies pinned buffers at PCIe speed while the first
 * read batches are already being processed; pageable buffers go through the driver's staging at a fraction of that.  Blocks come from
 * a process-wide pool and return to it (the C++ drop-in flattens vecbvec / VecPQVec straight into them).  NULL on failure.
 */
void* w2rap_step2_host_alloc(size_t bytes);
void w2rap_step2_host_free(void* p);

/*
 * File-level drop-in (src/modules/w2rap-contigger.cc:326-327,345-346):
 *   read  <dir>/frag_reads_orig.fastb + .qualp  (feudal files),
 *   write <dir>/<prefix>.small_K.hbv (BINWRITE stream of HyperBasevector, paths/HyperBasevector.cc:121-125),
 *         <dir>/<prefix>.small_K.paths (paths/long/ReadPath.cc:6-20), <dir>/small_K.freqs.
 * Paths are post-FixPaths, so `w2rap-contigger --from_step 3` continues from these files.
 */
int w2rap_step2_run_files(const char* dir, const char* prefix, const w2rap_params* p, w2rap_graph* out_or_null,
                          char* err, size_t errlen);

/* Host-side format helpers (no GPU): used by the file-level drop-in and by tests. */
int w2rap_write_hbv(const char* path, const w2rap_graph* g, char* err, size_t errlen);
int w2rap_write_paths(const char* path, const w2rap_graph* g, char* err, size_t errlen);
int w2rap_write_freqs(const char* path, const w2rap_graph* g, char* err, size_t errlen);

/* Synthetic read generator running on the device (bench input at the BASELINE.json sizes):
 * fills a device-resident read set without touching the host. */
typedef struct w2rap_synth_params {
    uint64_t genome_len;     /* haploid genome length in bases */
    uint32_t read_len;       /* 250 or 150 */
    uint32_t coverage;       /* e.g. 60 */
    uint64_t seed;
    uint32_t het_per_10k;    /* SNP rate of the second haplotype per 10,000 bases (0 = haploid) */
    uint32_t reserved;
    uint64_t n_reads;        /* 0 = derive from coverage; otherwise exact (the number of reads THIS call generates) */
    uint64_t first_read;     /* index of the first read to generate: a shard of the global read set (pairs stay together if even) */
} w2rap_synth_params;
int w2rap_step2_synth(const w2rap_synth_params* sp, int device, w2rap_device_reads** handle,
                      char* err, size_t errlen);
/* Copies a device-resident read set to freshly allocated pinned host buffers (for the end-to-end timing and
 * for writing .fastb/.qualp for the CPU reference).  Free with w2rap_step2_free_host_reads. */
int w2rap_step2_download_reads(w2rap_device_reads* handle, w2rap_reads* out, char* err, size_t errlen);
void w2rap_step2_free_host_reads(w2rap_reads* r);
/* Host-side feudal writers for a flattened read set (reference formats: feudal/FeudalControlBlock.h:156-163). */
int w2rap_write_fastb(const char* path, const w2rap_reads* r, char* err, size_t errlen);
int w2rap_write_qualp(const char* path, const w2rap_reads* r, char* err, size_t errlen);
/* Host-side feudal readers: allocate a flattened read set from .fastb/.qualp (free with w2rap_step2_free_host_reads). */
int w2rap_read_fastb_qualp(const char* fastb, const char* qualp, w2rap_reads* out, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* W2RAP_STEP2_H_ */
