/*
 * w2rap_step1.h — C ABI of the step-1 producer that feeds the B200 step 2.
 *
 * Replaces, for paired FASTQ input, the reference's ExtractReads FASTQ branch
 * (src/paths/long/large/ExtractReads.cc:372-479: four lines per record, 'N' -> 'A' (:416-419), quality = char - 33 (:470-473),
 * mates interleaved (:475)) and its quality compressor PQVecEncoder (src/feudal/PQVec.cc:18-120), and produces the flattened
 * read stores `w2rap_reads` (include/w2rap_step2.h) that w2rap_step2_run consumes directly — no MasterVec scatter into Mempool chunks,
 * no files in between.  w2rap_step1_write_stores writes the same two step files the reference writes after step 1
 * (frag_reads_orig.fastb / .qualp, src/modules/w2rap-contigger.cc:300-310), byte for byte.
 *
 * Host code only (libw2rap_step1.so has no CUDA dependency).  Plain pointers and sizes; no exception crosses the boundary.
 */
#ifndef W2RAP_STEP1_H
#define W2RAP_STEP1_H

#include <stddef.h>
#include <stdint.h>

#include "w2rap_step2.h"

#ifdef __cplusplus
extern "C" {
#endif

#define W2RAP_STEP1_ABI_VERSION 1

typedef struct w2rap_step1_params {
    uint32_t abi_version;        /* W2RAP_STEP1_ABI_VERSION */
    uint32_t threads;            /* 0 = all hardware threads */
    /* where the five arrays of the result live: null = malloc/free.  Pass w2rap_step2_host_alloc / w2rap_step2_host_free to have
     * the stores built straight in pinned memory (w2rap_step2_run then copies them at PCIe speed). */
    void* (*alloc)(size_t bytes);
    void (*release)(void* p);
} w2rap_step1_params;

typedef struct w2rap_step1_stats {
    uint64_t n_pairs;
    uint64_t n_bases;
    uint64_t n_converted;        /* 'N' turned into 'A' */
    uint64_t qual_bytes;         /* size of the PQVec streams */
    double read_s, parse_s, merge_s;   /* wall-clock: files into memory, parse + pack + compress, interleave into the result */
} w2rap_step1_stats;

/*
 * Reads a FASTQ pair (plain or .gz — .gz goes through `zcat`, as in the reference).  On success `out` holds 2 * n_pairs reads, mate 1
 * of pair i at index 2i and mate 2 at 2i+1.  Errors (the reference Scram()s on each): different record counts, incomplete record,
 * base/quality length mismatch, a base character outside [ACGTacgtN], a quality above 63 (PQVec.cc:30-35), a read of 65536 bases or
 * more.  `frac` subsampling (ExtractReads.cc:452-457) is not offered.
 */
int w2rap_step1_fastq_pair(const char* fastq1, const char* fastq2, const w2rap_step1_params* p, w2rap_reads* out, w2rap_step1_stats* stats_or_null,
                           char* err, size_t errlen);

/* Frees what w2rap_step1_fastq_pair put into `out` (with the same params) and zeroes it. */
void w2rap_step1_free(const w2rap_step1_params* p, w2rap_reads* out);

/* <dir>/frag_reads_orig.fastb and .qualp (feudal/FeudalControlBlock.h:43-53,156-163), byte-identical to the reference's. */
int w2rap_step1_write_stores(const char* dir, const w2rap_reads* reads, char* err, size_t errlen);

/* One PQVec stream from n qualities (each <= 63), exactly the bytes PQVecEncoder produces (PQVec.cc:18-120).  `out` needs
 * 3 * n + 1 bytes.  Returns the stream length including the terminating 0, or 0 if a quality is above 63. */
size_t w2rap_step1_pq_encode(const uint8_t* quals, uint32_t n, uint8_t* out);

/* The same stream from the n characters of a FASTQ quality line (phred + 33): the run-length form the producer uses — for a run of R
 * equal qualities the reference's programme always ends with R / 255 blocks of 255 and one of R % 255.  Test hook: must equal
 * w2rap_step1_pq_encode on the decoded qualities. */
size_t w2rap_step1_pq_encode_fastq(const char* quality_line, uint32_t n, uint8_t* out);

int w2rap_step1_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
