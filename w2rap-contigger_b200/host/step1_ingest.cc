// step1_ingest.cc — paired FASTQ -> flattened read stores (include/w2rap_step1.h).  Host code, no CUDA.
//
// What it replaces in the reference: ExtractReads' paired-FASTQ branch (paths/long/large/ExtractReads.cc:372-479) and the quality
// compressor PQVecEncoder (feudal/PQVec.cc:18-120).  The reference reads the two files line by line on one thread, builds a
// basevector and a qvec per read and compresses the qualities in batches; here both files are mapped (or, for .gz, piped through
// zcat into memory), cut into per-thread record ranges by counting newlines, and every thread packs bases and compresses
// qualities of its records straight into flat buffers; a last parallel pass interleaves the mates into the result arrays.
//
// The PQVec stream of a read must be byte-identical to the reference's (it is part of the step file .qualp), so the encoder
// follows PQVecEncoder::init/encode exactly — including the effect of the reference's lookup table for ceil(log2(range))
// (math/PowerOf2.h:33-43), whose entries for range >= 2 are 63, 62, 62, 61, ... (the leading-zero count, not the logarithm): with
// those "bit widths" a block that mixes two different qualities costs at least 7 bytes per quality against 3 bytes for a block of
// equal qualities, so the dynamic programme never chooses one and every block the reference emits is a run of EQUAL qualities of
// 1..255 elements: {nQs, (minQ << 3) & 0xff, minQ >> 5}.  pq_encode_ref() below is that programme restricted to what can win.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/w2rap_step1.h"

namespace {

struct Fail { int code; std::string msg; };
[[noreturn]] void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw Fail{code, buf};
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---------------------------------------------------------------- PQVec encoder (feudal/PQVec.cc:18-120)
// PQVecEncoder::init keeps, per prefix length, the cheapest cost in bytes (mCosts) and a stack of blocks (mBlocks) that it repairs
// after every quality: the best last block for the new prefix replaces the tail of the stack (:62-83).  For each new quality it
// scans back over up to 254 predecessors widening [min,max] (:44-61).  As soon as the scanned range holds two different values the
// reference's bit width is >= 58 and the candidate cannot beat "one 3-byte block per quality" (see the file comment: such a block
// of n qualities costs >= 7n + 2 bytes beyond cost[j], n one-quality blocks cost 3n), and the range only widens further back: only
// candidates inside the current run of equal qualities can be chosen.  Strict '<' keeps the SHORTEST last block among equal costs.
struct PqBlock { uint8_t n, q; };
struct PqEncoder {
    std::vector<uint32_t> cost;
    std::vector<uint32_t> plateau;       // indices where the (non-decreasing) cost array takes a new value
    std::vector<PqBlock> blocks;
    static uint32_t block_size(uint32_t n_qs, uint32_t bits) { return (n_qs * bits + 17 + 7) >> 3; }      // PQVec.h:57-58
    // returns false on a quality above 63 (PQVec.cc:30-35)
    //
    // The reference scans back from quality i over candidates j = i, i-1, ... (last block = [j, i], at most 255 long) and keeps the
    // first strictly cheaper one (:44-61).  Inside a run of equal qualities every candidate costs cost[j] + 3, and cost[] is
    // non-decreasing (cost[i+1] = cost[lowest allowed j] + 3, and that j never moves back), so the cheapest candidate is the LOWEST
    // allowed j and the one the reference ends up with is the HIGHEST j of the same cost: the end of the cost plateau that holds
    // the lowest allowed j.  The lowest allowed j only moves forward, so a pointer into the list of plateaus finds it in O(1)
    // amortised — the reference spends up to 254 steps per quality here, which is most of its step 1.
    bool init(const uint8_t* q, uint32_t n) {
        cost.clear(); blocks.clear(); plateau.clear();
        cost.reserve(n + 1);
        cost.push_back(1);                                       // cost of an empty qv: the terminator (:24)
        plateau.push_back(0);
        uint32_t run_start = 0, pl = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint8_t v = q[i];
            if (v > 63) return false;
            if (i && q[i - 1] != v) run_start = i;
            const uint32_t j_lo = std::max(run_start, i >= 254 ? i - 254 : 0u);
            while (pl + 1 < plateau.size() && plateau[pl + 1] <= j_lo) ++pl;
            const uint32_t j_best = pl + 1 < plateau.size() ? std::min(plateau[pl + 1] - 1, i) : i;
            const uint32_t best = cost[j_lo] + block_size(1, 0), best_n = i - j_best + 1;
            if (best != cost.back()) plateau.push_back(i + 1);
            cost.push_back(best);
            uint32_t to_remove = best_n - 1;                     // (:62-83)
            if (!to_remove) {
                blocks.push_back(PqBlock{1, v});
            } else {
                while (to_remove > blocks.back().n) { to_remove -= blocks.back().n; blocks.pop_back(); }
                if (to_remove == blocks.back().n) blocks.back() = PqBlock{(uint8_t)best_n, v};
                else { blocks.back().n -= (uint8_t)to_remove; blocks.push_back(PqBlock{(uint8_t)best_n, v}); }
            }
        }
        return true;
    }
    // (:87-120) with nBits == 0: {nQs, low byte of (minQ << 3), its high byte}; then the 0 terminator
    size_t encode(uint8_t* out) const {
        uint8_t* o = out;
        for (const PqBlock& b : blocks) {
            const uint32_t bits = (uint32_t)b.q << 3;
            *o++ = b.n; *o++ = (uint8_t)bits; *o++ = (uint8_t)(bits >> 8);
        }
        *o++ = 0;
        return (size_t)(o - out);
    }
};

// ---------------------------------------------------------------- input files
struct Input {
    const char* data = nullptr;
    size_t size = 0;
    bool mapped = false;
    std::vector<char> owned;
    ~Input() { if (mapped && data) munmap(const_cast<char*>(data), size); }
    void open(const char* path) {
        const size_t L = strlen(path);
        if (L > 3 && !strcmp(path + L - 3, ".gz")) {             // as the reference does: a zcat pipe (ExtractReads.cc:377-380)
            std::string cmd = "zcat '" + std::string(path) + "'";
            FILE* f = popen(cmd.c_str(), "r");
            if (!f) fail(W2RAP_ERR_IO, "cannot run %s", cmd.c_str());
            size_t cap = 1u << 26;
            owned.resize(cap);
            size_t n = 0;
            for (;;) {
                if (n == owned.size()) owned.resize(owned.size() * 2);
                const size_t got = fread(owned.data() + n, 1, owned.size() - n, f);
                if (!got) break;
                n += got;
            }
            if (pclose(f) != 0) fail(W2RAP_ERR_IO, "%s failed", cmd.c_str());
            owned.resize(n);
            data = owned.data(); size = n;
            return;
        }
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) fail(W2RAP_ERR_IO, "cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fail(W2RAP_ERR_IO, "cannot stat %s", path); }
        size = (size_t)st.st_size;
        if (size) {
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); fail(W2RAP_ERR_IO, "cannot map %s: %s", path, strerror(errno)); }
            madvise(m, size, MADV_SEQUENTIAL);
            data = (const char*)m; mapped = true;
        }
        ::close(fd);
    }
};

// growable byte buffer without the zero-fill of std::vector::resize
struct Bytes {
    uint8_t* p = nullptr;
    size_t n = 0, cap = 0;
    Bytes() {}
    Bytes(const Bytes&) = delete;
    Bytes& operator=(const Bytes&) = delete;
    Bytes(Bytes&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    ~Bytes() { free(p); }
    uint8_t* room(size_t more) {                 // pointer to `more` writable bytes at the end (not yet counted in n)
        if (n + more > cap) {
            size_t c = std::max(cap * 2, n + more + 4096);
            uint8_t* q = (uint8_t*)realloc(p, c);
            if (!q) fail(W2RAP_ERR_OOM, "out of memory while parsing (%zu bytes)", c);
            p = q; cap = c;
        }
        return p + n;
    }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

// PQVec stream of one read straight from the FASTQ quality characters: the closed form of PqEncoder for what the reference's
// programme does with a run of R equal qualities — R / 255 blocks of 255 followed by one block of R % 255 (the first 255
// elements keep extending one block; element 256 finds the cheapest predecessor on the cost plateau that ends at itself and opens a
// block of 1, which then grows: PqEncoder::init, and tests/test_step1_ingest.py::test_run_length_form_equals_the_programme).
// Returns the bytes written (terminator included), 0 on a quality above 63.  out needs 3 * n + 1 bytes.
inline size_t pq_encode_runs(const char* q, size_t n, uint8_t* out) {
    uint8_t* o = out;
    size_t i = 0;
    while (i < n) {
        const char c = q[i];
        size_t j = i + 1;
        while (j < n && q[j] == c) ++j;
        const uint8_t v = (uint8_t)(c - 33);                       // (ExtractReads.cc:470-473)
        if (v > 63) return 0;
        const uint8_t b1 = (uint8_t)(v << 3), b2 = (uint8_t)(v >> 5);
        size_t R = j - i;
        while (R >= 255) { *o++ = 255; *o++ = b1; *o++ = b2; R -= 255; }
        if (R) { *o++ = (uint8_t)R; *o++ = b1; *o++ = b2; }
        i = j;
    }
    *o++ = 0;
    return (size_t)(o - out);
}

// base character -> 2-bit code; 4 = 'N' (becomes A, ExtractReads.cc:416-419); 0xff = anything else (the reference's
// CharToBaseMapper accepts ACGTacgt only, dna/Bases.h:200-205)
struct BaseLut {
    uint8_t v[256];
    BaseLut() { memset(v, 0xff, 256); v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; v['N'] = 4; }
};
const BaseLut kBaseLut;

// One file, parsed: flat bases / quality streams of its records, in file order.
struct Parsed {
    uint64_t n = 0;
    std::vector<uint32_t> len;           // per record
    std::vector<uint64_t> boff, qoff;    // n + 1
    std::vector<Bytes> tb, tq;           // per thread: packed bases, PQVec streams
    std::vector<uint64_t> t_first;       // first record of each thread (T + 1)
    std::vector<uint64_t> t_boff, t_qoff;       // offsets of each thread's buffers in the file-level numbering (T + 1)
    uint64_t n_bases = 0, n_converted = 0;
};

void run_threads(unsigned T, const std::function<void(unsigned)>& fn) {
    std::vector<std::thread> th;
    std::vector<Fail> errs(T, Fail{0, ""});
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t] { try { fn(t); } catch (const Fail& f) { errs[t] = f; } catch (const std::exception& e) { errs[t] = Fail{W2RAP_ERR_INTERNAL, e.what()}; } });
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < T; ++t) if (errs[t].code) throw errs[t];      // the first failing range, in file order
}

// Cuts `in` into T ranges of whole records (4 lines each, ExtractReads.cc:394-446) and parses them in parallel.
void parse_file(const Input& in, const char* name, unsigned T, Parsed& out) {
    const char* d = in.data;
    const size_t N = in.size;
    // 1. newlines per byte range
    std::vector<size_t> cut(T + 1);
    for (unsigned t = 0; t <= T; ++t) cut[t] = N * t / T;
    std::vector<uint64_t> nl(T + 1, 0);
    run_threads(T, [&](unsigned t) {
        uint64_t c = 0;
        const char* p = d + cut[t]; const char* e = d + cut[t + 1];
        while (p < e) { const char* q = (const char*)memchr(p, '\n', (size_t)(e - p)); if (!q) break; ++c; p = q + 1; }
        nl[t + 1] = c;
    });
    for (unsigned t = 0; t < T; ++t) nl[t + 1] += nl[t];
    uint64_t lines = nl[T] + ((N && d[N - 1] != '\n') ? 1 : 0);             // a last line without '\n' counts (getline returns it)
    if (lines % 4) fail(W2RAP_ERR_BAD_ARG, "incomplete record in %s (%llu lines)", name, (unsigned long long)lines);
    out.n = lines / 4;
    // 2. record ranges: thread t takes records [n*t/T, n*(t+1)/T); its first byte is found from the newline counts
    out.t_first.resize(T + 1);
    for (unsigned t = 0; t <= T; ++t) out.t_first[t] = out.n * t / T;
    std::vector<size_t> start(T + 1, N);
    run_threads(T, [&](unsigned t) {
        const uint64_t want_line = out.t_first[t] * 4;                        // this record starts after `want_line` newlines
        if (want_line == 0) { start[t] = 0; return; }
        unsigned c = (unsigned)(std::upper_bound(nl.begin(), nl.end(), want_line - 1) - nl.begin()) - 1;      // range holding newline #want_line
        if (c >= T) return;
        uint64_t seen = nl[c];
        const char* p = d + cut[c]; const char* e = d + N;
        while (p < e) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!q) break;
            if (++seen == want_line) { start[t] = (size_t)(q + 1 - d); return; }
            p = q + 1;
        }
    });
    start[T] = N;
    // 3. parse
    out.len.assign(out.n, 0);
    out.tb.clear(); out.tq.clear();
    out.tb.resize(T); out.tq.resize(T);
    std::vector<uint64_t> t_bases(T, 0), t_conv(T, 0);
    run_threads(T, [&](unsigned t) {
        const uint64_t r0 = out.t_first[t], r1 = out.t_first[t + 1];
        const char* p = d + start[t];
        const char* e = d + N;
        Bytes& B = out.tb[t];
        Bytes& Q = out.tq[t];
        B.room((size_t)((start[t + 1] - start[t]) / 8 + 64));
        Q.room((size_t)((start[t + 1] - start[t]) / 16 + 64));
        const uint8_t* lut = kBaseLut.v;
        auto line = [&](const char** b, size_t* n) {
            *b = p;
            const char* q = p < e ? (const char*)memchr(p, '\n', (size_t)(e - p)) : nullptr;
            if (q) { *n = (size_t)(q - p); p = q + 1; } else { *n = (size_t)(e - p); p = e; }
        };
        uint64_t nb = 0, conv = 0;
        for (uint64_t r = r0; r < r1; ++r) {
            const char *l0, *l1, *l2, *l3; size_t n0, n1, n2, n3;
            line(&l0, &n0); line(&l1, &n1); line(&l2, &n2); line(&l3, &n3);
            (void)l0; (void)n0; (void)l2; (void)n2;                         // the header and '+' lines are skipped unread (:394, :424)
            if (n1 != n3) fail(W2RAP_ERR_BAD_ARG, "inconsistent base/quality lengths in %s, record %llu: %zu bases, %zu quals", name, (unsigned long long)r, n1, n3);
            if (n1 > 65535) fail(W2RAP_ERR_BAD_ARG, "read of %zu bases in %s, record %llu: step 2 stores good lengths in 16 bits", n1, name, (unsigned long long)r);
            out.len[r] = (uint32_t)n1;
            // bases: 2 bits each, LSB-first, byte-aligned per read (feudal/FieldVec.h:765-769); 'N' -> 'A' (:416-419)
            uint8_t* bp = B.room((n1 + 3) / 4);
            const unsigned char* s1 = (const unsigned char*)l1;
            size_t i = 0;
            unsigned flags = 0;
            for (; i + 4 <= n1; i += 4) {
                const unsigned a = lut[s1[i]], b = lut[s1[i + 1]], c = lut[s1[i + 2]], dd = lut[s1[i + 3]];
                flags |= a | b | c | dd;
                *bp++ = (uint8_t)((a & 3) | (b & 3) << 2 | (c & 3) << 4 | (dd & 3) << 6);
            }
            if (i < n1) {
                unsigned acc = 0;
                for (unsigned k = 0; i < n1; ++i, ++k) { const unsigned a = lut[s1[i]]; flags |= a; acc |= (a & 3) << (2 * k); }
                *bp++ = (uint8_t)acc;
            }
            if (flags & ~3u) {                                              // an 'N' or a character the reference refuses: look again
                for (size_t k = 0; k < n1; ++k) {
                    if (lut[s1[k]] == 4) ++conv;
                    else if (lut[s1[k]] == 0xff)
                        fail(W2RAP_ERR_BAD_ARG, "character '%c' (0x%02x) in the bases of %s, record %llu (the reference accepts ACGTacgt and N)", s1[k] > 31 && s1[k] < 127 ? s1[k] : '?',
                             s1[k], name, (unsigned long long)r);
                }
            }
            B.n += (n1 + 3) / 4;
            nb += n1;
            // qualities: char - 33 (:470-473), compressed as PQVecEncoder does it
            const size_t qn = pq_encode_runs(l3, n1, Q.room(3 * n1 + 1));
            if (!qn) fail(W2RAP_ERR_BAD_ARG, "quality score above 63 in %s, record %llu (the reference: \"Your input reads are funny\")", name, (unsigned long long)r);
            Q.n += qn;
        }
        t_bases[t] = nb; t_conv[t] = conv;
    });
    for (unsigned t = 0; t < T; ++t) { out.n_bases += t_bases[t]; out.n_converted += t_conv[t]; }
    out.t_boff.assign(T + 1, 0); out.t_qoff.assign(T + 1, 0);
    for (unsigned t = 0; t < T; ++t) { out.t_boff[t + 1] = out.t_boff[t] + out.tb[t].size(); out.t_qoff[t + 1] = out.t_qoff[t] + out.tq[t].size(); }
}

// length of the PQVec stream at p (terminator included): constant blocks are 3 bytes (PQVec.h:124-127 for nBits == 0)
inline size_t pq_stream_bytes(const uint8_t* p) { size_t n = 0; while (p[n]) n += 3; return n + 1; }

void write_feudal(const std::string& path, const uint8_t* var, uint64_t var_bytes, const uint64_t* off, uint64_t n, const void* fixed, uint64_t fixed_bytes,
                  uint8_t sz_fixed, uint8_t sz_x) {
    // feudal/FeudalControlBlock.h:43-53: u32 n, u8 flags(=1), sizeof fixed, sizeof X, sizeof A(=1), u64 offset of the offsets table,
    // u64 offset of the fixed data; then var data, (n+1) ABSOLUTE u64 offsets, fixed data (:156-163)
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) fail(W2RAP_ERR_IO, "cannot write %s: %s", path.c_str(), strerror(errno));
    const uint64_t var_off = 24 + var_bytes, fixed_off = var_off + 8 * (n + 1);
    uint8_t hdr[24];
    const uint32_t n32 = (uint32_t)n;
    memcpy(hdr, &n32, 4); hdr[4] = 1; hdr[5] = sz_fixed; hdr[6] = sz_x; hdr[7] = 1; memcpy(hdr + 8, &var_off, 8); memcpy(hdr + 16, &fixed_off, 8);
    bool ok = fwrite(hdr, 1, 24, f) == 24 && (!var_bytes || fwrite(var, 1, var_bytes, f) == var_bytes);
    std::vector<uint64_t> abs_off(n + 1);
    for (uint64_t i = 0; i <= n; ++i) abs_off[i] = off[i] + 24;
    ok = ok && fwrite(abs_off.data(), 8, n + 1, f) == n + 1 && (!fixed_bytes || fwrite(fixed, 1, fixed_bytes, f) == fixed_bytes);
    if (fclose(f) != 0 || !ok) fail(W2RAP_ERR_IO, "short write to %s", path.c_str());
}

}  // namespace

extern "C" {

int w2rap_step1_abi_version(void) { return W2RAP_STEP1_ABI_VERSION; }

size_t w2rap_step1_pq_encode(const uint8_t* quals, uint32_t n, uint8_t* out) {
    PqEncoder e;
    if (!e.init(quals, n)) return 0;
    return e.encode(out);
}

size_t w2rap_step1_pq_encode_fastq(const char* quality_line, uint32_t n, uint8_t* out) { return pq_encode_runs(quality_line, n, out); }

void w2rap_step1_free(const w2rap_step1_params* p, w2rap_reads* out) {
    if (!out) return;
    void (*rel)(void*) = (p && p->release) ? p->release : free;
    if (out->bases) rel((void*)out->bases);
    if (out->base_off) rel((void*)out->base_off);
    if (out->len) rel((void*)out->len);
    if (out->quals) rel((void*)out->quals);
    if (out->qual_off) rel((void*)out->qual_off);
    memset(out, 0, sizeof *out);
}

int w2rap_step1_fastq_pair(const char* fastq1, const char* fastq2, const w2rap_step1_params* p, w2rap_reads* out, w2rap_step1_stats* st, char* err, size_t errlen) {
    if (out) memset(out, 0, sizeof *out);
    try {
        if (!fastq1 || !fastq2 || !p || !out) fail(W2RAP_ERR_BAD_ARG, "null argument");
        if (p->abi_version != W2RAP_STEP1_ABI_VERSION) fail(W2RAP_ERR_BAD_ARG, "ABI version %u, library has %d", p->abi_version, W2RAP_STEP1_ABI_VERSION);
        if ((p->alloc == nullptr) != (p->release == nullptr)) fail(W2RAP_ERR_BAD_ARG, "alloc and release must be given together");
        unsigned T = p->threads ? p->threads : std::max(1u, std::thread::hardware_concurrency());
        T = std::min(T, 256u);
        void* (*alloc)(size_t) = p->alloc ? p->alloc : malloc;
        const double t0 = now_s();
        Input in[2];
        in[0].open(fastq1); in[1].open(fastq2);
        const double t1 = now_s();
        Parsed ps[2];
        parse_file(in[0], fastq1, T, ps[0]);
        parse_file(in[1], fastq2, T, ps[1]);
        if (ps[0].n != ps[1].n)
            fail(W2RAP_ERR_BAD_ARG, "the files %s and %s appear to be paired, yet have different numbers of records (%llu, %llu)", fastq1, fastq2,
                 (unsigned long long)ps[0].n, (unsigned long long)ps[1].n);
        const double t2 = now_s();
        // ---- interleave: read 2i = record i of file 1, read 2i+1 = record i of file 2 (:475)
        const uint64_t np = ps[0].n, n = 2 * np;
        uint64_t* base_off = (uint64_t*)alloc((n + 1) * 8);
        uint64_t* qual_off = (uint64_t*)alloc((n + 1) * 8);
        uint32_t* len = (uint32_t*)alloc((n + 1) * 4);
        if (!base_off || !qual_off || !len) fail(W2RAP_ERR_OOM, "out of memory for the offset tables of %llu reads", (unsigned long long)n);
        out->base_off = base_off; out->qual_off = qual_off; out->len = len;
        // per-record sizes of each file, in file order (the quality streams are measured where they lie)
        for (int f = 0; f < 2; ++f) {
            Parsed& P = ps[f];
            P.boff.assign(P.n + 1, 0); P.qoff.assign(P.n + 1, 0);
            run_threads(T, [&](unsigned t) {
                uint64_t b = P.t_boff[t], q = P.t_qoff[t];
                const uint8_t* qs = P.tq[t].data();
                uint64_t ql = 0;
                for (uint64_t r = P.t_first[t]; r < P.t_first[t + 1]; ++r) {
                    P.boff[r] = b; P.qoff[r] = q;
                    b += (P.len[r] + 3) / 4;
                    const size_t s = pq_stream_bytes(qs + ql); ql += s; q += s;
                }
            });
            P.boff[P.n] = P.t_boff[T]; P.qoff[P.n] = P.t_qoff[T];
        }
        for (uint64_t i = 0; i < np; ++i) {                                    // (serial: two additions per pair)
            base_off[2 * i] = ps[0].boff[i] + ps[1].boff[i];     base_off[2 * i + 1] = ps[0].boff[i + 1] + ps[1].boff[i];
            qual_off[2 * i] = ps[0].qoff[i] + ps[1].qoff[i];     qual_off[2 * i + 1] = ps[0].qoff[i + 1] + ps[1].qoff[i];
            len[2 * i] = ps[0].len[i]; len[2 * i + 1] = ps[1].len[i];
        }
        base_off[n] = ps[0].boff[np] + ps[1].boff[np];
        qual_off[n] = ps[0].qoff[np] + ps[1].qoff[np];
        len[n] = 0;
        uint8_t* bases = (uint8_t*)alloc(base_off[n] + 32);
        uint8_t* quals = (uint8_t*)alloc(qual_off[n] + 32);
        if (!bases || !quals) fail(W2RAP_ERR_OOM, "out of memory for the read stores (%llu + %llu bytes)", (unsigned long long)base_off[n], (unsigned long long)qual_off[n]);
        out->bases = bases; out->quals = quals;
        memset(bases + base_off[n], 0, 32); memset(quals + qual_off[n], 0, 32);
        run_threads(T, [&](unsigned t) {
            for (int f = 0; f < 2; ++f) {
                const Parsed& P = ps[f];
                const uint8_t* sb = P.tb[t].data(); const uint8_t* sq = P.tq[t].data();
                for (uint64_t r = P.t_first[t]; r < P.t_first[t + 1]; ++r) {
                    const uint64_t k = 2 * r + f;
                    memcpy(bases + base_off[k], sb + (P.boff[r] - P.t_boff[t]), P.boff[r + 1] - P.boff[r]);
                    memcpy(quals + qual_off[k], sq + (P.qoff[r] - P.t_qoff[t]), P.qoff[r + 1] - P.qoff[r]);
                }
            }
        });
        out->n_reads = n;
        const double t3 = now_s();
        if (st) {
            st->n_pairs = np; st->n_bases = ps[0].n_bases + ps[1].n_bases; st->n_converted = ps[0].n_converted + ps[1].n_converted;
            st->qual_bytes = qual_off[n]; st->read_s = t1 - t0; st->parse_s = t2 - t1; st->merge_s = t3 - t2;
        }
        return W2RAP_OK;
    } catch (const Fail& f) {
        if (err && errlen) snprintf(err, errlen, "%s", f.msg.c_str());
        if (out) w2rap_step1_free(p, out);
        return f.code;
    } catch (const std::exception& e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        if (out) w2rap_step1_free(p, out);
        return W2RAP_ERR_OOM;
    }
}

int w2rap_step1_write_stores(const char* dir, const w2rap_reads* r, char* err, size_t errlen) {
    try {
        if (!dir || !r) fail(W2RAP_ERR_BAD_ARG, "null argument");
        const uint64_t n = r->n_reads;
        if (n >> 32) fail(W2RAP_ERR_BAD_ARG, "feudal files hold at most 2^32-1 elements");
        static const uint64_t zero = 0;
        const uint64_t* bo = n ? r->base_off : &zero; const uint64_t* qo = n ? r->qual_off : &zero;
        // fastb: X = 2-bit bases packed 4 per byte, fixed data = u32 length per read (sizeof fixed 4, sizeof X recorded as 16);
        // qualp: X = bytes of the PQVec stream, no fixed data (sizeof X 8) — the values the reference's writers record
        write_feudal(std::string(dir) + "/frag_reads_orig.fastb", r->bases, bo[n], bo, n, r->len, 4 * n, 4, 16);
        write_feudal(std::string(dir) + "/frag_reads_orig.qualp", r->quals, qo[n], qo, n, nullptr, 0, 0, 8);
        return W2RAP_OK;
    } catch (const Fail& f) {
        if (err && errlen) snprintf(err, errlen, "%s", f.msg.c_str());
        return f.code;
    }
}

}  // extern "C"
