// prims.cuh — hand-written device primitives: exclusive scan and a stable LSD radix sort of an index permutation by
// multi-word keys.  Used for edge ordering, vertex numbering and offset tables (all small next to the counting stage).
#pragma once
#include "rt.cuh"

namespace w2r {

// ---------------------------------------------------------------- exclusive scan (TIn -> TOut), three kernels
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class TOut>
__device__ __forceinline__ TOut block_exclusive_scan(TOut v, TOut* total, TOut* smem /* >= 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    TOut incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { TOut t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        TOut w = lane < nwarps ? smem[lane] : TOut(0);
        TOut wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { TOut t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
        smem[lane] = wi - w;                       // exclusive warp offsets
        if (lane == 31) smem[32] = wi;             // block total
    }
    __syncthreads();
    TOut r = smem[warp] + incl - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

template <class TIn, class TOut>
__global__ void k_scan_tile_sums(const TIn* __restrict__ in, uint64_t n, TOut* __restrict__ tile_sums) {
    __shared__ TOut sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += (TOut)in[base + i];
    TOut total;
    block_exclusive_scan<TOut>(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of `m` tile sums in place; writes the grand total to *total_out
template <class TOut>
__global__ void k_scan_sums_inplace(TOut* __restrict__ sums, uint64_t m, TOut* __restrict__ total_out) {
    __shared__ TOut sm[33];
    const uint64_t per = (m + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = (uint64_t)threadIdx.x * per, hi = (lo + per < m) ? lo + per : m;
    TOut s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += sums[i];
    TOut total;
    TOut off = block_exclusive_scan<TOut>(s, &total, sm);
    for (uint64_t i = lo; i < hi; ++i) { TOut v = sums[i]; sums[i] = off; off += v; }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

template <class TIn, class TOut>
__global__ void k_scan_apply(const TIn* __restrict__ in, uint64_t n, const TOut* __restrict__ tile_offsets, TOut* __restrict__ out) {
    __shared__ TOut sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? (TOut)in[base + i] : TOut(0); s += v[i]; }
    TOut total;
    TOut off = block_exclusive_scan<TOut>(s, &total, sm) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = off; off += v[i]; }
}

// out[i] = sum of in[0..i), i in [0,n); *total (device pointer, may be null) = sum of all.  out may alias in only if TIn==TOut.
template <class TIn, class TOut>
inline void exclusive_scan(Ctx& c, const TIn* in, uint64_t n, TOut* out, TOut* total_dev) {
    if (n == 0) { if (total_dev) W2R_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(TOut), c.stream)); return; }
    uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    TmpBuf<TOut> sums(c, tiles);
    W2R_LAUNCH(c, (k_scan_tile_sums<TIn, TOut>), (unsigned)tiles, SCAN_THREADS, 0, in, n, sums.p);
    W2R_LAUNCH(c, (k_scan_sums_inplace<TOut>), 1, 1024, 0, sums.p, tiles, total_dev);
    W2R_LAUNCH(c, (k_scan_apply<TIn, TOut>), (unsigned)tiles, SCAN_THREADS, 0, in, n, sums.p, out);
}

// ---------------------------------------------------------------- LSD radix sort of a permutation by multi-word keys
// keys are SoA: word w of element e is words[w][e].  perm is sorted so that (words[nw-1], ..., words[0]) ascend with
// words[nw-1] the MOST significant; stable.  Only bits [lo_bit, hi_bit) of each word take part.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void k_rs_iota(uint32_t* perm, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) perm[i] = i;
}

__global__ void k_rs_hist(const uint32_t* __restrict__ perm, uint32_t n, const uint64_t* __restrict__ word, int shift, uint32_t nb,
                          uint32_t* __restrict__ counts /* [256][nb] */) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int it = 0; it < RS_ITEMS; ++it) {
        uint32_t i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(word[perm[i]] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(uint32_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void k_rs_scatter(const uint32_t* __restrict__ perm, uint32_t n, const uint64_t* __restrict__ word, int shift, uint32_t nb,
                             const uint32_t* __restrict__ offsets /* [256][nb] exclusive */, uint32_t* __restrict__ out) {
    __shared__ uint32_t run[256];                 // running output cursor per digit for this block
    __shared__ uint16_t wcnt[RS_THREADS / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    run[threadIdx.x] = offsets[(uint32_t)threadIdx.x * nb + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int it = 0; it < RS_ITEMS; ++it) {
        for (int w = 0; w < RS_THREADS / 32; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        uint32_t i = base + it * RS_THREADS + threadIdx.x;
        bool valid = i < n;
        uint32_t e = valid ? perm[i] : 0u;
        uint32_t d = valid ? (uint32_t)(word[e] >> shift) & 255u : (256u + lane);   // invalid lanes match nobody
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) wcnt[warp][d] = (uint16_t)__popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += wcnt[w][d];
            out[run[d] + before + rank_in_warp] = e;
        }
        __syncthreads();
        {
            uint32_t tot = 0;
            for (int w = 0; w < RS_THREADS / 32; ++w) tot += wcnt[w][threadIdx.x];
            run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

struct SortWord { const uint64_t* word; int lo_bit, hi_bit; };

// Sorts perm (n entries, must hold a permutation of element indices, e.g. iota) by the given words, least significant word
// first in `words`.  tmp must hold n entries.  On return the sorted permutation is in `perm`.
inline void radix_sort_perm(Ctx& c, uint32_t* perm, uint32_t* tmp, uint32_t n, const SortWord* words, int nwords) {
    if (n <= 1) return;
    uint32_t nb = (n + RS_TILE - 1) / RS_TILE;
    TmpBuf<uint32_t> counts(c, (size_t)256 * nb);
    uint32_t* src = perm;
    uint32_t* dst = tmp;
    for (int w = 0; w < nwords; ++w) {
        for (int shift = words[w].lo_bit; shift < words[w].hi_bit; shift += 8) {
            W2R_LAUNCH(c, k_rs_hist, nb, RS_THREADS, 0, src, n, words[w].word, shift, nb, counts.p);
            exclusive_scan<uint32_t, uint32_t>(c, counts.p, (uint64_t)256 * nb, counts.p, nullptr);
            W2R_LAUNCH(c, k_rs_scatter, nb, RS_THREADS, 0, src, n, words[w].word, shift, nb, counts.p, dst);
            uint32_t* t = src; src = dst; dst = t;
        }
    }
    if (src != perm) W2R_CUDA(cudaMemcpyAsync(perm, src, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
}

}  // namespace w2r
