"""Diagnostic: wall time of w2rap_step2_run on the bounded CPU-baseline sample, cold and warm, with stage times."""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import w2r_testlib as T
args = argparse.Namespace(ref_genome_mbp=2.0, coverage=60, read_len=250)
rs, desc = bench.sample_reads(args)
p = T.default_params(apply_fixpaths=1, device=0)
for i in range(5):
    t0 = time.time()
    got = T.run_product(rs, p)
    t = got["timings"]
    print(i, "python wall %.1f ms" % ((time.time() - t0) * 1e3), {k: round(t[k], 2) for k in ("wall_ms", "host_pre_ms", "h2d_ms", "count_ms", "dict_ms", "adjacency_ms", "unipath_ms", "hbv_ms", "path_ms", "d2h_ms", "total_ms", "alloc_host_ms")}, flush=True)
