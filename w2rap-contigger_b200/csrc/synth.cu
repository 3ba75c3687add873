// synth.cu — synthetic paired-end read sets generated ON the device, in the reference's read-store encodings
// (2-bit packed bases, PQVec qualities), for benchmarking at the BASELINE.json sizes where a host generator would
// dominate the run.  Model: i.i.d. ACGT genome with planted repeat families, optional SNP haplotype,
// 2 x read_len PE reads from N(500,50) fragments on both strands, Q37 body with a U(0,60) low-quality tail (Q2-15),
// 1% sporadic Q2-19, substitution errors drawn at 10^(-Q/10), no Ns.  Deterministic in (seed, read index).
#include <algorithm>
#include <cmath>
#include <random>

#include "../../include/w2rap_step2.h"
#include "device_reads.cuh"
#include "kernels.cuh"
#include "prims.cuh"

namespace w2r {

__host__ __device__ __forceinline__ uint64_t splitmix(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}

__global__ void k_syn_genome(uint8_t* g, uint64_t n, uint64_t seed) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n + 31) / 32; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = splitmix(seed ^ (i * 0x632be59bd9b4e019ull));
        for (int j = 0; j < 32 && i * 32 + j < n; ++j) g[i * 32 + j] = (r >> (2 * j)) & 3;
    }
}
struct RepeatJob { uint64_t src, dst; uint32_t len, rc; uint64_t seed; };
__global__ void k_syn_repeats(uint8_t* g, const uint8_t* pool, const RepeatJob* jobs, uint32_t njobs) {
    for (uint32_t j = blockIdx.x; j < njobs; j += gridDim.x) {
        RepeatJob jb = jobs[j];
        for (uint32_t i = threadIdx.x; i < jb.len; i += blockDim.x) {
            uint32_t s = jb.rc ? jb.len - 1 - i : i;
            uint32_t b = pool[jb.src + s];
            if (jb.rc) b = 3 - b;
            uint64_t r = splitmix(jb.seed ^ i);
            if ((r & 0xffff) < 655) b = (b + 1 + ((r >> 16) % 3)) & 3;   // ~1% divergence between copies
            g[jb.dst + i] = (uint8_t)b;
        }
    }
}
__global__ void k_syn_hap(const uint8_t* g, uint8_t* h, uint64_t n, uint64_t seed, uint32_t het_per_10k) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = splitmix(seed ^ (i + 0x1234567ull));
        uint32_t b = g[i];
        if ((r % 10000u) < het_per_10k) b = (b + 1 + ((r >> 32) % 3)) & 3;
        h[i] = (uint8_t)b;
    }
}

__constant__ uint32_t c_perr[64];   // P(error | Q) scaled to 2^32

// Greedy PQVec block encoder (the same partition rule as oracle/readsim_support.c mode 0): close the block when the delta
// width would grow and the block already holds >= 8 quals.  out == nullptr: size only.  Returns bytes incl. terminator.
__device__ uint32_t pq_encode_greedy(const uint8_t* q, uint32_t n, uint8_t* out) {
    uint32_t o = 0, i = 0;
    while (i < n) {
        uint32_t mn = q[i], mx = q[i], bits = 0, len = 1;
        while (i + len < n && len < 255) {
            uint32_t v = q[i + len], nmn = min(v, mn), nmx = max(v, mx);
            uint32_t rng = nmx - nmn + 1, nb = 0;
            while ((1u << nb) < rng) ++nb;
            if (nb > bits && len >= 8) break;
            mn = nmn; mx = nmx; bits = nb; ++len;
        }
        uint32_t nbytes = 1 + ((9 + len * bits + 7) >> 3);
        if (out) {
            out[o] = (uint8_t)len;
            uint64_t acc = bits | ((uint64_t)mn << 3);
            uint32_t have = 9, w = o + 1;
            for (uint32_t k = 0; k < len; ++k) {
                acc |= (uint64_t)(q[i + k] - mn) << have; have += bits;
                while (have >= 8) { out[w++] = (uint8_t)acc; acc >>= 8; have -= 8; }
            }
            if (have) out[w++] = (uint8_t)acc;
        }
        o += nbytes; i += len;
    }
    if (out) out[o] = 0;
    return o + 1;
}

struct SynArgs { const uint8_t* hap0; const uint8_t* hap1; uint64_t G; uint32_t L; uint64_t seed; uint64_t n_reads; uint64_t first_read; };

template <bool WRITE>
__global__ void __launch_bounds__(128) k_syn_reads(SynArgs a, uint32_t* __restrict__ qsize, const uint64_t* __restrict__ qual_off, uint8_t* __restrict__ bases,
                                                   uint8_t* __restrict__ quals) {
    uint8_t q[256], b[256];
    const uint32_t L = a.L, nbb = (L + 3) / 4;
    for (uint64_t lr = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; lr < a.n_reads; lr += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = lr + a.first_read;      // global read index: the read set is a function of (seed, index) only
        uint64_t pair = r >> 1;
        uint64_t h = splitmix(a.seed ^ (pair * 0x9e3779b97f4a7c15ull));
        const uint8_t* hap = (a.hap1 && (h & 1)) ? a.hap1 : a.hap0;
        bool flip = (h >> 1) & 1;
        // fragment length ~ N(500,50): sum of four uniforms on [-1,1) has sd sqrt(4/3)
        int64_t u = (int64_t)((h >> 8) & 0xffff) + ((h >> 24) & 0xffff) + ((h >> 40) & 0xffff) + ((splitmix(h) >> 3) & 0xffff) - 2 * 65536;
        int64_t flen = 500 + (int64_t)((double)u / 32768.0 * 50.0 / 1.1547);
        if (flen < (int64_t)L) flen = L;
        if ((uint64_t)flen > a.G) flen = (int64_t)a.G;
        uint64_t start = splitmix(h ^ 0xabcdefull) % (a.G - (uint64_t)flen + 1);
        bool second = (r & 1) != 0;
        bool from_end = second != flip;        // which fragment end this read starts at
        uint64_t rs = splitmix(a.seed ^ (r * 0xd1342543de82ef95ull) ^ 0x55aa);
        uint32_t tail = (uint32_t)(rs % 61u);
        for (uint32_t i = 0; i < L; ++i) {
            uint32_t base = from_end ? 3u - hap[start + (uint64_t)flen - 1 - i] : hap[start + i];
            uint64_t x = splitmix(rs + i + 1);
            uint32_t qv = 37;
            if (i >= L - tail) qv = 2 + (uint32_t)((x >> 8) % 14u);
            if ((x & 0xff) < 3 && ((x >> 40) & 3) != 3) qv = 2 + (uint32_t)((x >> 20) % 18u);   // ~1% sporadic low quality
            if ((uint32_t)(x >> 32) < c_perr[qv]) base = (base + 1 + (uint32_t)((x >> 12) % 3u)) & 3u;
            q[i] = (uint8_t)qv; b[i] = (uint8_t)base;
        }
        if (!WRITE) qsize[lr] = pq_encode_greedy(q, L, nullptr);
        else {
            pq_encode_greedy(q, L, quals + qual_off[lr]);
            uint8_t* pb = bases + lr * nbb;
            for (uint32_t j = 0; j < nbb; ++j) {
                uint32_t v = 0;
                for (uint32_t k = 0; k < 4 && 4 * j + k < L; ++k) v |= (uint32_t)b[4 * j + k] << (2 * k);
                pb[j] = (uint8_t)v;
            }
        }
    }
}
__global__ void k_syn_regular(uint64_t n, uint32_t L, uint64_t* base_off, uint32_t* len) {
    const uint64_t nbb = (L + 3) / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) { base_off[i] = i * nbb; if (i < n) len[i] = L; }
}

}  // namespace w2r

using namespace w2r;

extern "C" {

int w2rap_step2_synth(const w2rap_synth_params* sp, int device, w2rap_device_reads** handle, char* err, size_t errlen) {
    try {
        if (!sp || !handle) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null argument");
        if (sp->read_len < 61 || sp->read_len > 256) W2R_FAIL(W2RAP_ERR_BAD_ARG, "read_len must be in 61..256");
        if (sp->genome_len < 1000) W2R_FAIL(W2RAP_ERR_BAD_ARG, "genome_len must be >= 1000");
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); W2R_FAIL(W2RAP_ERR_NO_DEVICE, "no CUDA device is visible"); }
        if (device < 0) cudaGetDevice(&device);
        W2R_CUDA(cudaSetDevice(device));
        Ctx c; c.device = device;
        cudaDeviceProp pr; W2R_CUDA(cudaGetDeviceProperties(&pr, device)); c.sm_count = pr.multiProcessorCount;
        c.stream = 0;
        const uint64_t G = sp->genome_len;
        const uint32_t L = sp->read_len;
        uint64_t nr = sp->n_reads ? sp->n_reads : (G * sp->coverage / L);
        nr &= ~1ull;
        DBuf<uint8_t> hap0(G), hap1(sp->het_per_10k ? G : 0);
        W2R_LAUNCH(c, k_syn_genome, grid_for(c, (G + 31) / 32, 256), 256, 0, hap0.p, G, sp->seed);
        // repeat families: ~2% of the sequence in 300 bp-5 kb repeats at 5-50 copies
        {
            std::mt19937_64 rng(sp->seed * 7919 + 13);
            std::vector<RepeatJob> jobs;
            uint64_t budget = G / 50, pool_len = 0;
            while (budget > 1000 && G > 100000) {
                uint32_t len = 300 + (uint32_t)(rng() % 4700), copies = 5 + (uint32_t)(rng() % 46);
                if ((uint64_t)len * copies > budget) copies = (uint32_t)std::max<uint64_t>(2, budget / len);
                for (uint32_t k = 0; k < copies; ++k) jobs.push_back(RepeatJob{pool_len, rng() % (G - len), len, (uint32_t)(rng() & 1), rng()});
                pool_len += len;
                budget -= std::min<uint64_t>(budget, (uint64_t)len * copies);
            }
            // copies must not overlap: concurrent blocks write them, and the genome has to be deterministic in the seed
            std::sort(jobs.begin(), jobs.end(), [](const RepeatJob& a, const RepeatJob& b) { return a.dst < b.dst; });
            {
                std::vector<RepeatJob> keep;
                uint64_t end = 0;
                for (const RepeatJob& j : jobs) if (keep.empty() || j.dst >= end) { keep.push_back(j); end = j.dst + j.len; }
                jobs.swap(keep);
            }
            if (!jobs.empty()) {
                DBuf<uint8_t> pool(pool_len);
                W2R_LAUNCH(c, k_syn_genome, grid_for(c, (pool_len + 31) / 32, 256), 256, 0, pool.p, pool_len, sp->seed ^ 0x5eedf00dull);
                DBuf<RepeatJob> dj(jobs.size());
                W2R_CUDA(cudaMemcpy(dj.p, jobs.data(), jobs.size() * sizeof(RepeatJob), cudaMemcpyHostToDevice));
                W2R_LAUNCH(c, k_syn_repeats, (unsigned)std::min<size_t>(jobs.size(), 65535), 256, 0, hap0.p, pool.p, dj.p, (uint32_t)jobs.size());
                W2R_CUDA(cudaDeviceSynchronize());
            }
        }
        if (sp->het_per_10k) W2R_LAUNCH(c, k_syn_hap, grid_for(c, G, 256), 256, 0, hap0.p, hap1.p, G, sp->seed ^ 0x4e7ull, sp->het_per_10k);
        uint32_t perr[64];
        for (int q = 0; q < 64; ++q) { double p = std::pow(10.0, -q / 10.0); perr[q] = p >= 1.0 ? 0xffffffffu : (uint32_t)(p * 4294967296.0); }
        W2R_CUDA(cudaMemcpyToSymbol(c_perr, perr, sizeof(perr)));

        w2rap_device_reads* h = new w2rap_device_reads();
        DeviceReads& d = h->d;
        try {
            d.device = device; d.n = nr; d.n_bases = nr * L; d.max_len = L; d.n_inst_upper = nr * (uint64_t)(L - 59); d.n_kreads = nr;
            const uint64_t nbb = (L + 3) / 4;
            d.bases_bytes = nr * nbb;
            W2R_CUDA(cudaMalloc((void**)&d.bases, d.bases_bytes + 32));
            W2R_CUDA(cudaMalloc((void**)&d.base_off, (nr + 1) * 8));
            W2R_CUDA(cudaMalloc((void**)&d.qual_off, (nr + 1) * 8));
            W2R_CUDA(cudaMalloc((void**)&d.len, (nr + 1) * 4));
            W2R_LAUNCH(c, k_syn_regular, grid_for(c, nr + 1, 256), 256, 0, nr, L, d.base_off, d.len);
            SynArgs a{hap0.p, sp->het_per_10k ? hap1.p : nullptr, G, L, sp->seed, nr, sp->first_read};
            DBuf<uint32_t> qsize(nr);
            DBuf<unsigned long long> tot(1);
            W2R_LAUNCH(c, (k_syn_reads<false>), grid_for(c, nr, 128), 128, 0, a, qsize.p, (const uint64_t*)nullptr, (uint8_t*)nullptr, (uint8_t*)nullptr);
            exclusive_scan<uint32_t, unsigned long long>(c, qsize.p, nr, (unsigned long long*)d.qual_off, tot.p);
            unsigned long long qb = 0;
            W2R_CUDA(cudaMemcpy(&qb, tot.p, 8, cudaMemcpyDeviceToHost));
            W2R_CUDA(cudaMemcpy(d.qual_off + nr, &qb, 8, cudaMemcpyHostToDevice));
            d.quals_bytes = qb;
            W2R_CUDA(cudaMalloc((void**)&d.quals, qb + 32));
            W2R_CUDA(cudaMemset(d.quals + qb, 0, 32));
            W2R_CUDA(cudaMemset(d.bases + d.bases_bytes, 0, 32));
            W2R_LAUNCH(c, (k_syn_reads<true>), grid_for(c, nr, 128), 128, 0, a, (uint32_t*)nullptr, (const uint64_t*)d.qual_off, d.bases, d.quals);
            W2R_CUDA(cudaDeviceSynchronize());
        } catch (...) {
            cudaFree(d.bases); cudaFree(d.quals); cudaFree(d.base_off); cudaFree(d.qual_off); cudaFree(d.len);
            delete h;
            throw;
        }
        *handle = h;
    } catch (const Error& e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.msg.c_str());
        return e.code;
    }
    return W2RAP_OK;
}

int w2rap_step2_download_reads(w2rap_device_reads* handle, w2rap_reads* out, char* err, size_t errlen) {
    try {
        if (!handle || !out) W2R_FAIL(W2RAP_ERR_BAD_ARG, "null argument");
        const DeviceReads& d = handle->d;
        W2R_CUDA(cudaSetDevice(d.device));
        memset(out, 0, sizeof(*out));
        uint8_t *b = nullptr, *q = nullptr; uint64_t *bo = nullptr, *qo = nullptr; uint32_t* ln = nullptr;
        W2R_CUDA(cudaMallocHost((void**)&b, d.bases_bytes + 32));
        W2R_CUDA(cudaMallocHost((void**)&q, d.quals_bytes + 32));
        W2R_CUDA(cudaMallocHost((void**)&bo, (d.n + 1) * 8));
        W2R_CUDA(cudaMallocHost((void**)&qo, (d.n + 1) * 8));
        W2R_CUDA(cudaMallocHost((void**)&ln, (d.n + 1) * 4));
        W2R_CUDA(cudaMemcpy(b, d.bases, d.bases_bytes + 32, cudaMemcpyDeviceToHost));
        W2R_CUDA(cudaMemcpy(q, d.quals, d.quals_bytes + 32, cudaMemcpyDeviceToHost));
        W2R_CUDA(cudaMemcpy(bo, d.base_off, (d.n + 1) * 8, cudaMemcpyDeviceToHost));
        W2R_CUDA(cudaMemcpy(qo, d.qual_off, (d.n + 1) * 8, cudaMemcpyDeviceToHost));
        if (d.n) W2R_CUDA(cudaMemcpy(ln, d.len, d.n * 4, cudaMemcpyDeviceToHost));
        out->n_reads = d.n; out->bases = b; out->quals = q; out->base_off = bo; out->qual_off = qo; out->len = ln;
    } catch (const Error& e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.msg.c_str());
        return e.code;
    }
    return W2RAP_OK;
}

// Host read sets come either from w2rap_step2_download_reads (pinned) or from w2rap_read_fastb_qualp (malloc).
void w2rap_step2_free_host_reads(w2rap_reads* r) {
    if (!r) return;
    const void* ptrs[5] = {r->bases, r->quals, r->base_off, r->qual_off, r->len};
    for (const void* p : ptrs) {
        if (!p) continue;
        if (cudaFreeHost((void*)p) != cudaSuccess) { cudaGetLastError(); free((void*)p); }
    }
    memset(r, 0, sizeof(*r));
}

}  // extern "C"
