// BuildReadQGraph_b200.cc — link-level drop-in for the reference's src/paths/long/BuildReadQGraph.cc.
//
// Keeps the exact signature of buildReadQGraph (paths/long/BuildReadQGraph.h:24-29), flattens the two read stores,
// calls the B200 library through its C ABI (include/w2rap_step2.h) and rebuilds HyperBasevector / ReadPathVec exactly the
// way buildHBVFromEdges (paths/long/HBVFromEdges.cc:124-151) and path_reads_OMP (BuildReadQGraph.cc:925) fill them.
// It is compiled against the reference's own headers and replaces the object of the original
// translation unit (CMakeLists.txt:378); nothing else in the reference changes.
//
// Inputs are const& and are not modified.  Errors: the reference aborts (FatalErr / CRD::exit); so does this wrapper.
#include "paths/long/BuildReadQGraph.h"

#include <omp.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "Basevector.h"
#include "Qualvector.h"
#include "feudal/PQVec.h"
#include "paths/HyperBasevector.h"
#include "paths/long/ReadPath.h"
#include "system/System.h"
#include "w2rap_step2.h"

namespace {

// FieldVec keeps its packed bytes behind a protected accessor (feudal/FieldVec.h:616-631); a derived view exposes it.
struct BvecBytes : public bvec {
    unsigned char const* bytes() const { return data(); }
    unsigned char* bytes() { return data(); }
};
// PQVec is one word: the low 48 bits are the pointer to its block stream (feudal/PQVec.h:185-201).
static_assert(sizeof(PQVec) == sizeof(size_t), "PQVec layout changed");
inline unsigned char const* pqvec_bytes(PQVec const& q) {
    size_t word;
    std::memcpy(&word, &q, sizeof word);
    return reinterpret_cast<unsigned char const*>(word & 0xffffffffffffULL);
}

}  // namespace

void buildReadQGraph(vecbvec const& reads, VecPQVec const& quals, bool doFillGaps, bool doJoinOverlaps, unsigned minQual, unsigned minFreq,
                     double /*minFreq2Fract*/, unsigned /*maxGapSize*/, HyperBasevector* pHBV, ReadPathVec* pPaths, int _K, std::string workdir,
                     std::string /*tmpdir*/, unsigned char /*disk_batches*/) {
    if (doFillGaps || doJoinOverlaps) FatalErr("buildReadQGraph (B200): fillGaps/joinOverlaps are not built; the driver passes false (w2rap-contigger.cc:336-338)");
    std::cout << Date() << ": creating kmers from reads... (B200)" << std::endl;
    const size_t n = reads.size();
    ForceAssertEq(n, quals.size());

    // ---- flatten: per-read packed bases and PQVec streams into two contiguous PINNED buffers (the library then copies them at PCIe
    // speed under its first kernels); offsets by a serial scan, bytes by all OpenMP threads
    const double t_flat0 = WallClockTime();
    std::vector<uint64_t> base_off(n + 1), qual_off(n + 1);
    std::vector<uint32_t> len(n ? n : 1);
    base_off[0] = qual_off[0] = 0;
    for (size_t i = 0; i < n; ++i) {
        len[i] = reads[i].size();
        base_off[i + 1] = base_off[i] + (len[i] + 3) / 4;
        size_t qs = quals[i].size();
        qual_off[i + 1] = qual_off[i] + (qs ? qs : 1);     // an empty PQVec has no buffer: emit the terminator byte
    }
    unsigned char* bases = static_cast<unsigned char*>(w2rap_step2_host_alloc(base_off[n] + 32));
    unsigned char* qbuf = static_cast<unsigned char*>(w2rap_step2_host_alloc(qual_off[n] + 32));
    if (!bases || !qbuf) FatalErr("buildReadQGraph (B200): cannot allocate pinned host memory for the read stores");
    std::memset(bases + base_off[n], 0, 32); std::memset(qbuf + qual_off[n], 0, 32);
#pragma omp parallel for schedule(static, 4096)
    for (size_t i = 0; i < n; ++i) {
        if (len[i]) std::memcpy(bases + base_off[i], static_cast<BvecBytes const&>(reads[i]).bytes(), (len[i] + 3) / 4);
        size_t qs = quals[i].size();
        if (qs) std::memcpy(qbuf + qual_off[i], pqvec_bytes(quals[i]), qs);
        else qbuf[qual_off[i]] = 0;
    }
    const double t_flat = WallClockTime() - t_flat0;

    w2rap_reads in;
    in.n_reads = n; in.bases = bases; in.base_off = base_off.data(); in.len = len.data(); in.quals = qbuf; in.qual_off = qual_off.data();
    w2rap_params p;
    std::memset(&p, 0, sizeof p);
    p.abi_version = W2RAP_STEP2_ABI_VERSION; p.K = (uint32_t)_K; p.min_qual = minQual; p.min_freq = minFreq;
    p.want_paths = pPaths ? 1 : 0; p.apply_fixpaths = 0;   // main() calls FixPaths itself (w2rap-contigger.cc:340)
    p.device = -1; p.workdir = workdir.empty() ? nullptr : workdir.c_str(); p.verbose = 1;
    w2rap_graph g;
    char err[1024];
    const double t_run0 = WallClockTime();
    int rc = w2rap_step2_run(&in, &p, &g, err, sizeof err);
    const double t_run = WallClockTime() - t_run0;
    w2rap_step2_host_free(bases); w2rap_step2_host_free(qbuf);
    if (rc != W2RAP_OK) FatalErr("buildReadQGraph (B200) failed with status " << rc << ": " << err);
    const double t_build0 = WallClockTime();

    // ---- HyperBasevector, as buildHBVFromEdges does it (HBVFromEdges.cc:79,124-151)
    std::cout << Date() << ": building graph..." << std::endl;
    pHBV->Clear();
    if (g.n_edges) {
        pHBV->SetK(_K);
        pHBV->AddVertices(g.n_vertices);
        pHBV->EdgesMutable().reserve(2 * g.n_edges);
        for (uint64_t i = 0; i < g.n_edges; ++i) {
            bvec edge(g.edge_len[i]);
            std::memcpy(static_cast<BvecBytes&>(edge).bytes(), g.edge_bases + g.edge_off[i], (g.edge_len[i] + 3) / 4);
            int fw = pHBV->AddEdge(g.edge_vertices[4 * i], g.edge_vertices[4 * i + 1], edge);
            ForceAssertEq(fw, g.fwd_xlat[i]);
            if (g.rev_xlat[i] != g.fwd_xlat[i]) {
                int bw = pHBV->AddEdge(g.edge_vertices[4 * i + 2], g.edge_vertices[4 * i + 3], edge);
                pHBV->EdgeObjectMutable(bw).ReverseComplement();
                ForceAssertEq(bw, g.rev_xlat[i]);
            }
        }
    }
    std::cout << Date() << ": graph built" << std::endl;
    if (pPaths) {
        pPaths->clear();
        pPaths->resize(n);
#pragma omp parallel for schedule(static, 4096)
        for (size_t r = 0; r < n; ++r) {
            ReadPath& rp = (*pPaths)[r];
            rp.setOffset(g.path_offset[r]);
            rp.assign(g.path_edges + g.path_off[r], g.path_edges + g.path_off[r + 1]);
        }
        std::cout << Date() << ": " << g.n_pathed << " / " << n << " reads pathed, " << g.n_multipathed << " spanning junctions" << std::endl;
    }
    const double device_s = g.timings.total_ms * 1e-3;
    w2rap_step2_free(&g);
    std::cout << Date() << ": B200 step 2: flatten " << t_flat << " s, w2rap_step2_run " << t_run << " s (device " << device_s << " s), rebuild "
              << (WallClockTime() - t_build0) << " s" << std::endl;
}
