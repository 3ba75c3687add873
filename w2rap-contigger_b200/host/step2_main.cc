// step2_main.cc — file-level drop-in for step 2: `step2 <out_dir> <prefix> [--min_freq F] [--min_qual Q] [--device D]`
// reads <out_dir>/frag_reads_orig.fastb/.qualp (what `w2rap-contigger --to_step 1` leaves, w2rap-contigger.cc:315-316),
// writes <out_dir>/<prefix>.small_K.hbv, .small_K.paths (post-FixPaths) and small_K.freqs, after which
// `w2rap-contigger -o <out_dir> -p <prefix> --from_step 3` continues (w2rap-contigger.cc:352-358).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "w2rap_step2.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <out_dir> <prefix> [--min_freq F] [--min_qual Q] [--device D] [--quiet]\n", argv[0]); return 2; }
    w2rap_params p;
    memset(&p, 0, sizeof p);
    p.abi_version = W2RAP_STEP2_ABI_VERSION; p.K = W2RAP_K; p.min_qual = 7; p.min_freq = 4; p.device = -1; p.verbose = 1;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--min_freq") && i + 1 < argc) p.min_freq = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--min_qual") && i + 1 < argc) p.min_qual = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) p.device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) p.verbose = 0;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    char err[1024] = {0};
    w2rap_graph g;
    int rc = w2rap_step2_run_files(argv[1], argv[2], &p, &g, err, sizeof err);
    if (rc) { fprintf(stderr, "step2 failed (%d): %s\n", rc, err); return 1; }   // the reference exits non-zero on any failure
    printf("TIME, buildReadQGraph+FixPaths (B200), %.3f\n", g.timings.total_ms * 1e-3);
    w2rap_step2_free(&g);
    return 0;
}
