// places.cuh — step-3 input on the device: RepathInMemory's `places` (paths/long/large/Repath.cc:46-72).
//
// Reference: for every read path x (hbv edge ids, after FixPaths): nkmers = sum of edges[x[j]].size() - (K-1) (:57-59); the path is
// dropped if nkmers + (K-1) < K2 (:60); y = x reversed with every edge replaced by its involution (:61-62); the smaller of x and y
// in std::vector<int> order is the place (:63); all places are sorted and made unique (:69-71).
//
// Here: k_place_measure decides keep/orientation per read, a scan lays the kept places out, k_place_fill materialises them in
// canonical orientation.  Duplicates (a 60x read set repeats every place dozens of times) are removed FIRST, by sorting on a 64-bit
// content hash and comparing neighbours by content (a neighbour pair with equal hash and different content is counted; the host
// then retries with another salt, so the result is exact, not probabilistic).  Only the unique places are sorted
// lexicographically: a stable LSD radix sort over element positions, last position first, an absent element (shorter vector)
// sorting before every edge id — exactly std::vector<int>::operator<.
//
// The per-element logic is W2R_HD so that tests/hostcheck runs the same functions serially on the CPU.
#pragma once
#include <stdint.h>

#include "kmer.cuh"

namespace w2r {

struct PathsView { const uint64_t* off; const int32_t* edges; uint64_t n; };
struct PlacesView { const uint64_t* off; const int32_t* edges; uint64_t n; };

// keep/orientation of one path.  Returns the place length (0 = dropped); *flip = the inverse path is the smaller one.
// hcanon[h] = (canonical edge index << 1) | is-rc, as k_hbv_edges writes it.
W2R_HD uint32_t place_measure(const int32_t* x, uint64_t n, const uint32_t* hcanon, const uint32_t* edge_len, const int32_t* inv, uint32_t K2, bool* flip) {
    long long nk = 0;
    for (uint64_t j = 0; j < n; ++j) nk += (long long)edge_len[hcanon[x[j]] >> 1] - (long long)(K - 1);
    *flip = false;
    if (nk + (long long)(K - 1) < (long long)K2) return 0;            // (an empty path implies K-1 bases: dropped for every legal K2)
    for (uint64_t j = 0; j < n; ++j) {
        const int32_t yj = inv[x[n - 1 - j]];
        if (yj != x[j]) { *flip = yj < x[j]; break; }
    }
    return (uint32_t)n;
}
W2R_HD int32_t place_element(const int32_t* x, uint64_t n, const int32_t* inv, bool flip, uint64_t j) { return flip ? inv[x[n - 1 - j]] : x[j]; }

W2R_HD uint64_t place_mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
W2R_HD uint64_t place_hash(const int32_t* p, uint64_t n, uint64_t salt) {
    uint64_t h = place_mix(salt ^ (n * 0x9e3779b97f4a7c15ull));
    for (uint64_t j = 0; j < n; ++j) h = place_mix(h ^ ((uint64_t)(uint32_t)p[j] + 0x632be59bd9b4e019ull * (j + 1)));
    return h;
}
W2R_HD bool place_equal(const PlacesView& s, uint64_t a, uint64_t b) {
    const uint64_t na = s.off[a + 1] - s.off[a], nb = s.off[b + 1] - s.off[b];
    if (na != nb) return false;
    const int32_t* pa = s.edges + s.off[a]; const int32_t* pb = s.edges + s.off[b];
    for (uint64_t j = 0; j < na; ++j) if (pa[j] != pb[j]) return false;
    return true;
}
// radix key of place `a` at element position `pos`: 0 = the vector has ended (sorts first), else id + 1
W2R_HD uint64_t place_key(const PlacesView& s, uint64_t a, uint32_t pos) {
    const uint64_t n = s.off[a + 1] - s.off[a];
    return pos < n ? (uint64_t)(uint32_t)s.edges[s.off[a] + pos] + 1ull : 0ull;
}

#if defined(__CUDACC__)
__global__ void k_place_measure(PathsView pv, const uint32_t* __restrict__ hcanon, const uint32_t* __restrict__ edge_len, const int32_t* __restrict__ inv, uint32_t K2,
                                uint32_t* __restrict__ plen, uint32_t* __restrict__ kept, uint8_t* __restrict__ flip) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < pv.n; r += (uint64_t)gridDim.x * blockDim.x) {
        bool f;
        const uint32_t n = place_measure(pv.edges + pv.off[r], pv.off[r + 1] - pv.off[r], hcanon, edge_len, inv, K2, &f);
        plen[r] = n; kept[r] = n ? 1u : 0u; flip[r] = f ? 1 : 0;
    }
}
// the kept places, canonical orientation, in read order: place pidx[r] occupies out_edges[poff[r] .. poff[r] + plen[r])
__global__ void k_place_fill(PathsView pv, const int32_t* __restrict__ inv, const uint32_t* __restrict__ plen, const uint8_t* __restrict__ flip,
                             const uint32_t* __restrict__ pidx, const uint64_t* __restrict__ poff, uint64_t* __restrict__ out_off, int32_t* __restrict__ out_edges) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < pv.n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t n = plen[r];
        if (!n) continue;
        const int32_t* x = pv.edges + pv.off[r];
        out_off[pidx[r]] = poff[r];
        for (uint32_t j = 0; j < n; ++j) out_edges[poff[r] + j] = place_element(x, n, inv, flip[r] != 0, j);
    }
}
__global__ void k_place_lens(const uint64_t* __restrict__ off, uint64_t n, uint32_t* __restrict__ len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) len[i] = (uint32_t)(off[i + 1] - off[i]);
}
__global__ void k_place_hash(PlacesView s, uint64_t salt, uint64_t* __restrict__ h) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (uint64_t)gridDim.x * blockDim.x)
        h[i] = place_hash(s.edges + s.off[i], s.off[i + 1] - s.off[i], salt);
}
// perm: the places sorted by hash.  first[i] = place perm[i] differs from its predecessor; a differing pair with equal hashes is a
// collision (equal places might then not be neighbours: the caller re-salts).
__global__ void k_place_first(PlacesView s, const uint32_t* __restrict__ perm, const uint64_t* __restrict__ h, uint32_t* __restrict__ first,
                              unsigned long long* __restrict__ collisions) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t f = 1;
        if (i) {
            const uint32_t a = perm[i], b = perm[i - 1];
            if (place_equal(s, a, b)) f = 0;
            else if (h[a] == h[b]) atomicAdd(collisions, 1ull);
        }
        first[i] = f;
    }
}
// representatives of the unique places (hash order) and the longest of them
__global__ void k_place_select(PlacesView s, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ first, const uint32_t* __restrict__ excl,
                               uint32_t* __restrict__ rep, unsigned int* __restrict__ maxlen) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!first[i]) continue;
        const uint32_t a = perm[i];
        rep[excl[i]] = a;
        atomicMax(maxlen, (unsigned int)(s.off[a + 1] - s.off[a]));
    }
}
__global__ void k_place_key(PlacesView s, const uint32_t* __restrict__ rep, uint64_t U, uint32_t pos, uint64_t* __restrict__ key) {
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < U; u += (uint64_t)gridDim.x * blockDim.x) key[u] = place_key(s, rep[u], pos);
}
// order[i] = index (into rep) of the i-th place of the result; its length
__global__ void k_place_out_len(PlacesView s, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ order, uint64_t U, uint32_t* __restrict__ olen) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = rep[order ? order[i] : i];
        olen[i] = (uint32_t)(s.off[a + 1] - s.off[a]);
    }
}
__global__ void k_place_gather(PlacesView s, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ order, uint64_t U, const uint64_t* __restrict__ ooff,
                               int32_t* __restrict__ oedges) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = rep[order ? order[i] : i];
        const uint64_t n = s.off[a + 1] - s.off[a];
        const int32_t* src = s.edges + s.off[a];
        int32_t* dst = oedges + ooff[i];
        for (uint64_t j = 0; j < n; ++j) dst[j] = src[j];
    }
}
#endif

}  // namespace w2r
