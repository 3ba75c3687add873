// step12_main.cc — FASTQ pair -> step files of steps 1 and 2, without the reference in between:
//   step12 <r1.fastq[.gz]> <r2.fastq[.gz]> <out_dir> <prefix> [--min_freq F] [--min_qual Q] [--threads T] [--device D] [--stores-only]
// = `w2rap-contigger -r r1,r2 -o out_dir -p prefix --to_step 2` for paired FASTQ input: libw2rap_step1.so reads the pair straight into
// pinned flattened stores (include/w2rap_step1.h), writes frag_reads_orig.fastb/.qualp (the reference's step-1 files, byte for byte),
// and libw2rap_step2.so builds the graph and the paths from the same buffers and writes PFX.small_K.hbv / .paths / small_K.freqs.
// `w2rap-contigger --from_step 3` continues from there.  --stores-only stops after step 1 (no GPU needed).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "w2rap_step1.h"
#include "w2rap_step2.h"

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s <r1.fastq[.gz]> <r2.fastq[.gz]> <out_dir> <prefix> [--min_freq F] [--min_qual Q] [--threads T] [--device D] [--stores-only]\n", argv[0]);
        return 2;
    }
    const char *fq1 = argv[1], *fq2 = argv[2], *dir = argv[3], *prefix = argv[4];
    w2rap_params p;
    memset(&p, 0, sizeof p);
    p.abi_version = W2RAP_STEP2_ABI_VERSION; p.K = W2RAP_K; p.min_qual = 7; p.min_freq = 4; p.device = -1; p.verbose = 1;
    p.want_paths = 1; p.apply_fixpaths = 1;            // the step files hold post-FixPaths paths (w2rap-contigger.cc:340-346)
    p.workdir = dir;
    unsigned threads = 0;
    bool stores_only = false;
    for (int i = 5; i < argc; ++i) {
        if (!strcmp(argv[i], "--min_freq") && i + 1 < argc) p.min_freq = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--min_qual") && i + 1 < argc) p.min_qual = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) p.device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--stores-only")) stores_only = true;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    char err[1024] = {0};
    // pinned stores when a GPU will read them (falls back to malloc if the pinned pool refuses, e.g. no device at all)
    w2rap_step1_params sp = {W2RAP_STEP1_ABI_VERSION, threads, nullptr, nullptr};
    const bool pinned = !stores_only && w2rap_step2_device_count() > 0;
    if (pinned) { sp.alloc = w2rap_step2_host_alloc; sp.release = w2rap_step2_host_free; }
    w2rap_reads reads;
    w2rap_step1_stats st;
    int rc = w2rap_step1_fastq_pair(fq1, fq2, &sp, &reads, &st, err, sizeof err);
    if (rc) { fprintf(stderr, "step 1 failed (%d): %s\n", rc, err); return 1; }
    printf("step 1: %llu pairs, %llu bases (%llu N -> A), %.1f MB of quality streams; parse %.2f s, interleave %.2f s\n", (unsigned long long)st.n_pairs,
           (unsigned long long)st.n_bases, (unsigned long long)st.n_converted, st.qual_bytes / 1e6, st.parse_s, st.merge_s);
    rc = w2rap_step1_write_stores(dir, &reads, err, sizeof err);
    if (rc) { fprintf(stderr, "writing the step-1 files failed (%d): %s\n", rc, err); w2rap_step1_free(&sp, &reads); return 1; }
    if (stores_only) { w2rap_step1_free(&sp, &reads); return 0; }
    w2rap_graph g;
    rc = w2rap_step2_run(&reads, &p, &g, err, sizeof err);
    w2rap_step1_free(&sp, &reads);
    if (rc) { fprintf(stderr, "step 2 failed (%d): %s\n", rc, err); return 1; }      // (no CPU path: without a B200 this is where it stops)
    const std::string base = std::string(dir) + "/" + prefix;
    rc = w2rap_write_hbv((base + ".small_K.hbv").c_str(), &g, err, sizeof err);
    if (!rc) rc = w2rap_write_paths((base + ".small_K.paths").c_str(), &g, err, sizeof err);
    if (rc) { fprintf(stderr, "writing the step-2 files failed (%d): %s\n", rc, err); w2rap_step2_free(&g); return 1; }
    printf("TIME, buildReadQGraph+FixPaths (B200), %.3f\n", g.timings.total_ms * 1e-3);
    w2rap_step2_free(&g);
    return 0;
}
