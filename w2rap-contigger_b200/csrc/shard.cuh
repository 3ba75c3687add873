// shard.cuh — which k-mer goes where: hash partition of a k-mer and the rank that owns a partition.
// Host/device; shared by the partition kernel, the host orchestration and the CPU tests of the sharded protocol.
#pragma once
#include "kmer.cuh"

namespace w2r {

// partition = top logP bits of the k-mer hash
W2R_HD uint32_t part_of_hash(uint64_t h, uint32_t logP) { return logP ? (uint32_t)(h >> (64 - logP)) : 0u; }
// slot inside the counting region: the bits below the partition bits
W2R_HD uint64_t region_slot_of_hash(uint64_t h, uint32_t logP, uint32_t logR) { return (logP ? (h << logP) : h) >> (64 - logR); }
// rank r owns the contiguous partition range [r * P/world, (r+1) * P/world); world is a power of two <= P
W2R_HD uint32_t owner_of_partition(uint32_t part, uint32_t logP, uint32_t world) { return (uint32_t)(((uint64_t)part * world) >> logP); }

}  // namespace w2r
