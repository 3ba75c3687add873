"""Host-side unit tests of the kernels' device functions (tests/hostcheck/hostcheck.cpp) against the oracle.

The CUDA kernels are thin loops around host/device functions (csrc/{kmer,pqvec,extract,unipath,path}.cuh); hostcheck drives
the same functions serially in kernel order.  This is where logic errors are caught without a GPU; the `-m gpu` tests then
check the real kernels through the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(ROOT, "oracle", "_build", "libhostcheck.so")


@pytest.fixture(scope="session")
def hc(T):
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    deps = [src] + [os.path.join(ROOT, "w2rap-contigger_b200", "csrc", f) for f in ("kmer.cuh", "pqvec.cuh", "extract.cuh", "unipath.cuh", "path.cuh", "shard.cuh", "shardgraph.cuh", "slab_freelist.h", "places.cuh")] + [os.path.join(ROOT, "include", "w2rap_step2.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", SO, src], check=True)
    lib = C.CDLL(SO)
    lib.hc_count.argtypes = [C.POINTER(T.Reads), C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.hc_graph.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(T.Reads), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(T.Graph)]
    lib.hc_graph_sharded.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(T.Reads), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(T.Graph)]
    lib.hc_free.argtypes = [C.c_void_p]
    lib.hc_graph_free.argtypes = [C.POINTER(T.Graph)]
    return lib


def run_hostcheck(T, hc, rs, min_qual=7, min_freq=4, apply_fixpaths=0, cap=24, left_cap=8):
    ptr, n, ninst = C.c_void_p(), C.c_uint64(), C.c_uint64()
    reads = rs.c()
    assert hc.hc_count(C.byref(reads), min_qual, C.byref(ptr), C.byref(n), C.byref(ninst)) == 0
    allk = T._arr(ptr.value, n.value, T.KMER_REC_DTYPE)
    g = T.Graph()
    rc = hc.hc_graph(ptr, n, min_freq, C.byref(reads), 1, apply_fixpaths, cap, left_cap, C.byref(g))
    hc.hc_free(ptr)
    assert rc == 0, "hostcheck stage failure %d" % rc
    d = T.graph_to_dict(g)
    n_ovf = g.timings.reserved
    hc.hc_graph_free(C.byref(g))
    return allk, ninst.value, d, n_ovf


def check_against_oracle(T, hc, rs, **kw):
    allk, ninst, d, n_ovf = run_hostcheck(T, hc, rs, **kw)
    pk = {k: v for k, v in kw.items() if k in ("min_qual", "min_freq", "apply_fixpaths")}
    o2 = T.run_oracle(rs, T.default_params(dump_kmers=2, want_paths=0, **pk))
    assert ninst == o2["n_kmer_instances"]
    for f in ("w0", "w1", "count", "ctx"):
        assert np.array_equal(allk[f], o2["dump"][f]), "distinct k-mer field %s differs" % f
    o1 = T.run_oracle(rs, T.default_params(dump_kmers=1, **pk))
    for k in ("n_kmer_instances", "n_bases", "n_reads", "hist"):      # not produced by hostcheck
        d[k] = o1[k]
    T.assert_graph_equal(o1, d, "hostcheck vs oracle")
    return d, n_ovf


def test_device_functions_smoke(T, hc):
    check_against_oracle(T, hc, T.smoke_set(seed=1, genome=20000, cov=40))


def test_device_functions_rich(T, hc):
    d, _ = check_against_oracle(T, hc, T.rich_set(seed=2, genome=60000, cov=50, pq_mode=1))
    assert d["n_edges"] > 500


def test_device_functions_rich_varlen_fixpaths(T, hc):
    check_against_oracle(T, hc, T.rich_set(seed=5, genome=40000, cov=50, families=5, palindromes=4, plasmid=1500, vary_len=True), apply_fixpaths=1)


def test_device_functions_long_reads(T, hc):
    """600-base reads of varying length (several gaps per read: the resumable path walker is resumed many times)."""
    check_against_oracle(T, hc, T.rich_set(seed=9, genome=50000, cov=40, read_len=600, families=3, palindromes=2, plasmid=2000, vary_len=True), apply_fixpaths=1)


def test_device_functions_golden_circ(T, hc):
    """Circles of both length parities, palindromes; tiny staging rows force the overflow path."""
    rs = T.read_fastb_qualp(os.path.join(HERE, "golden", "circ"))
    d, n_ovf = check_against_oracle(T, hc, rs, cap=3, left_cap=1)
    assert n_ovf > 0
    check_against_oracle(T, hc, rs, min_freq=2, min_qual=10)


def test_device_functions_low_coverage_fragmented(T, hc):
    """min_freq high relative to coverage: a shattered graph with many tips, gaps and short edges."""
    check_against_oracle(T, hc, T.rich_set(seed=7, genome=30000, cov=12, families=4, palindromes=2, plasmid=800), min_freq=3)


def test_minimizer_partition_key_is_strand_symmetric(T, hc):
    """Every instance of a canonical k-mer must reach the same partition, whichever strand the read shows."""
    rs = T.rich_set(seed=11, genome=6000, cov=6, families=2, palindromes=2, plasmid=500)
    reads = rs.c()
    hc.hc_minimizer_symmetry.argtypes = [C.POINTER(T.Reads), C.POINTER(C.c_uint64)]
    hc.hc_minimizer_symmetry.restype = C.c_uint64
    n = C.c_uint64(0)
    assert hc.hc_minimizer_symmetry(C.byref(reads), C.byref(n)) == 0
    assert n.value > 1000


def test_super_kmer_records_round_trip(T, hc):
    """The map's 32-byte super-k-mer records (extract.cuh: skm_build) expanded by the reduce (skm_kmer_at) give exactly the
    (canonical k-mer, context) stream of the reference's leaf loop, for every read: short, exactly-K, variable-length and long ones."""
    hc.hc_skm_roundtrip.restype = C.c_uint64
    hc.hc_skm_roundtrip.argtypes = [C.POINTER(T.Reads), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    for rs, logP in ((T.rich_set(seed=4, genome=30000, cov=20, vary_len=True), 6), (T.rich_set(seed=9, genome=20000, cov=10, read_len=600, vary_len=True), 10),
                     (T.smoke_set(seed=1, genome=5000, cov=10, read_len=61), 3)):
        nrec, ninst = C.c_uint64(), C.c_uint64()
        reads = rs.c()
        bad = hc.hc_skm_roundtrip(C.byref(reads), 7, logP, C.byref(nrec), C.byref(ninst))
        assert bad == 0 and ninst.value > 0
        assert nrec.value <= ninst.value


@pytest.mark.parametrize("world,logP", [(2, 5), (4, 7), (8, 9)])
def test_sharded_graph_stage_simulated_ranks(T, hc, world, logP):
    """csrc/shardgraph.cuh with `world` simulated ranks: dictionary sharded by minimiser owner, neighbour queries -> ghost entries,
    local chains ranked per rank, one record per chain piece gathered and ranked, strands/edges/offsets derived per owner — must give
    the oracle's graph bit for bit (pruned contexts, edge + offset of every k-mer, edge bytes, vertices, paths), including circles
    that span ranks (the plasmid) and palindromes."""
    rs = T.rich_set(seed=11, genome=40000, cov=40, families=4, palindromes=3, plasmid=1500)
    ptr, n, ninst = C.c_void_p(), C.c_uint64(), C.c_uint64()
    reads = rs.c()
    assert hc.hc_count(C.byref(reads), 7, C.byref(ptr), C.byref(n), C.byref(ninst)) == 0
    g = T.Graph()
    rc = hc.hc_graph_sharded(ptr, n, 4, world, logP, C.byref(reads), 1, 1, 24, 8, C.byref(g))
    hc.hc_free(ptr)
    assert rc == 0, "sharded hostcheck failure %d" % rc
    d = T.graph_to_dict(g)
    n_pieces = g.timings.count_passes
    hc.hc_graph_free(C.byref(g))
    want = T.run_oracle(rs, T.default_params(dump_kmers=1, apply_fixpaths=1))
    for k in ("n_kmer_instances", "n_bases", "n_reads", "hist"):
        d[k] = want[k]
    T.assert_graph_equal(want, d, "sharded hostcheck (world %d) vs oracle" % world)
    assert n_pieces > 2 * want["n_edges"]           # the chains really were cut into pieces at rank boundaries


def test_quality_floor_on_random_quality_vectors(T, hc):
    """pq_good_length (the body of k_good_len: block headers, constant blocks, packed deltas compared with the floor several at a
    time) against the oracle's end-anchored scan (BuildReadQGraph.cc:962-987) on random quality vectors that use every delta width,
    both PQVec block partitions and several floors.  Compared through the k-mer instance total, which every good length enters."""
    rng = np.random.default_rng(5)
    n, L = 1500, 300
    codes = rng.integers(0, 4, (n, L), dtype=np.uint8)
    quals = np.zeros((n, L), np.uint8)
    for r in range(n):
        base, spread = rng.integers(2, 41), [1, 3, 7, 20, 40][r % 5]
        q = np.clip(base + rng.integers(0, spread + 1, L) - spread // 2, 0, 63)
        for _ in range(rng.integers(0, 4)):
            a = rng.integers(0, L - 70)
            q[a:a + rng.integers(40, 120)] = rng.integers(7, 41)
        quals[r] = q
    lens = rng.integers(1, L + 1, n).astype(np.uint32)
    for mode in (0, 1):
        rs = T.flatten_reads(codes, quals, lens, pq_mode=mode)
        reads = rs.c()
        for mq in (0, 7, 10, 20, 41):
            ptr, nn, ni = C.c_void_p(), C.c_uint64(), C.c_uint64()
            assert hc.hc_count(C.byref(reads), mq, C.byref(ptr), C.byref(nn), C.byref(ni)) == 0
            hc.hc_free(ptr)
            want = T.run_oracle(rs, T.default_params(min_qual=mq, want_paths=0, min_freq=1))
            assert ni.value == want["n_kmer_instances"], (mode, mq)


def test_device_slab_free_list(hc):
    """The bookkeeping of the device slab (csrc/slab_freelist.h: lowest-address first fit, coalescing free ranges) fuzzed against a
    byte-map model: thousands of random takes, give-backs and growth steps, invariants checked after every one."""
    hc.hc_slab_freelist_fuzz.restype = C.c_int
    hc.hc_slab_freelist_fuzz.argtypes = [C.c_uint64, C.c_uint32]
    for seed in range(1, 9):
        assert hc.hc_slab_freelist_fuzz(seed, 4000) == 0


@pytest.mark.parametrize("K2", [100, 200, 320])
def test_places_device_functions(T, hc, K2):
    """Step-3 places through the functions the kernels call, in the pipeline's order (hash sort -> neighbour
    comparison -> stable LSD sort over element positions), against the oracle's qsort + unique (pinned on the reference's log in
    test_oracle_golden.py).  Built on the ORACLE's graph, so only the places logic is under test."""
    rs = T.rich_set(seed=31, genome=60000, cov=50, families=4, palindromes=2, plasmid=1200)
    lib = T.oracle_lib()
    g = T.Graph()
    p = T.default_params(apply_fixpaths=1, places_K2=K2)
    assert lib.oracle_step2_run(C.byref(rs.c()), C.byref(p), C.byref(g)) == 0
    try:
        want = T.graph_to_dict(g)
        hc.hc_places.argtypes = [C.POINTER(T.Graph), C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        nk, npl, po, pe, coll = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        assert hc.hc_places(C.byref(g), K2, 7, C.byref(nk), C.byref(npl), C.byref(po), C.byref(pe), C.byref(coll)) == 0
        off = T._arr(po.value, npl.value + 1, "<u8")
        edges = T._arr(pe.value, int(off[-1]), "<i4")
        hc.hc_free(po); hc.hc_free(pe)
    finally:
        lib.oracle_step2_free(C.byref(g))
    assert coll.value == 0
    assert want["n_places"] > 100 and want["n_places_kept"] > want["n_places"]
    assert (nk.value, npl.value) == (want["n_places_kept"], want["n_places"])
    assert np.array_equal(off, want["place_off"]) and np.array_equal(edges, want["place_edges"])
