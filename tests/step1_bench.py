#!/usr/bin/env python
"""Step-1 producer against the reference's own step 1 on the same FASTQ pair, host cores only.
usage: step1_bench.py [n_pairs] [threads]   -> one JSON line; checks that both wrote identical .fastb/.qualp."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))      # (this script lives in tests/: it runs the reference binary under oracle/_ref as the comparison arm)
import w2r_testlib as T  # noqa: E402
from test_step1_ingest import S1Params, S1Stats  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    rng = np.random.default_rng(1)
    d = tempfile.mkdtemp(prefix="w2rap_step1_")
    L = 250
    lut = np.frombuffer(b"ACGT", np.uint8)
    fns = []
    for k in range(2):                                              # 2x250 reads, binned qualities with a decaying tail (the bench's quality model)
        fn = os.path.join(d, "r%d.fastq" % (k + 1)); fns.append(fn)
        with open(fn, "wb") as f:
            for c0 in range(0, n_pairs, 20000):
                m = min(20000, n_pairs - c0)
                s = lut[rng.integers(0, 4, (m, L))]
                q = np.full((m, L), 37, np.uint8)
                tail = rng.integers(0, 80, m)
                for i in range(m):
                    if tail[i]:
                        q[i, L - tail[i]:] = rng.choice(np.array([2, 12, 23, 27, 32], np.uint8), tail[i])
                q += 33
                f.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (c0 + i, s[i].tobytes(), q[i].tobytes()) for i in range(m)))
    lib = C.CDLL(os.path.join(ROOT, "w2rap-contigger_b200", "libw2rap_step1.so"))
    lib.w2rap_step1_fastq_pair.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(S1Params), C.POINTER(T.Reads), C.POINTER(S1Stats), C.c_char_p, C.c_size_t]
    lib.w2rap_step1_write_stores.argtypes = [C.c_char_p, C.POINTER(T.Reads), C.c_char_p, C.c_size_t]
    lib.w2rap_step1_free.argtypes = [C.POINTER(S1Params), C.POINTER(T.Reads)]
    p, r, st, err = S1Params(1, threads, None, None), T.Reads(), S1Stats(), C.create_string_buffer(1024)
    t0 = time.time()
    rc = lib.w2rap_step1_fastq_pair(fns[0].encode(), fns[1].encode(), C.byref(p), C.byref(r), C.byref(st), err, 1024)
    t_mine = time.time() - t0
    assert rc == 0, err.value
    mine = os.path.join(d, "mine"); os.makedirs(mine)
    t0 = time.time()
    assert lib.w2rap_step1_write_stores(mine.encode(), C.byref(r), err, 1024) == 0
    t_write = time.time() - t0
    lib.w2rap_step1_free(C.byref(p), C.byref(r))
    out = {"workload": "%d pairs 2x%d (%.0f Mbases), plain FASTQ" % (n_pairs, L, 2 * n_pairs * L / 1e6), "threads": threads,
           "b200_host_ingest_s": round(t_mine, 3), "of_which": {"map_files": round(st.read_s, 3), "parse_pack_compress": round(st.parse_s, 3), "interleave": round(st.merge_s, 3)},
           "write_step_files_s": round(t_write, 3), "gbases_per_s": round(st.n_bases / t_mine / 1e9, 3)}
    if os.path.exists(T.REF_BIN):
        ref = os.path.join(d, "ref"); os.makedirs(ref)
        t0 = time.time()
        subprocess.run([T.REF_BIN, "-t", str(threads), "-o", ref, "-p", "x", "-r", fns[0] + "," + fns[1], "--to_step", "1"], check=True, stdout=subprocess.DEVNULL)
        out["reference_step1_s"] = round(time.time() - t0, 3)
        out["files_identical"] = all(open(os.path.join(mine, f), "rb").read() == open(os.path.join(ref, f), "rb").read() for f in ("frag_reads_orig.fastb", "frag_reads_orig.qualp"))
        out["speedup_incl_file_write"] = round(out["reference_step1_s"] / (t_mine + t_write), 1)
    subprocess.run(["rm", "-rf", d])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
