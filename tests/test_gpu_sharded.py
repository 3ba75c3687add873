"""Sharded step 2 on several B200s (one rank per GPU, NCCL): every rank must return the whole graph, bit-identical to the
single-GPU / oracle result, and the paths of its own read shard.  Ranks are driven from threads of this process (NCCL allows
one rank per thread); `bench.py` drives the same entry points with one process per GPU under torchrun."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_sharded(T, rs, world, **kw):
    lib = T.product_lib()
    if lib.w2rap_step2_device_count() < world:
        pytest.skip("needs %d B200s" % world)
    err0 = C.create_string_buffer(512)
    uid = (C.c_uint8 * 128)()
    assert lib.w2rap_step2_comm_unique_id(uid, err0, 512) == 0, err0.value
    n = rs.n - (rs.n % (2 * world))
    bounds = [(n // world) * r for r in range(world)] + [rs.n]      # pairs stay together; the last rank takes the remainder
    shards = [rs.subset(np.arange(bounds[r], bounds[r + 1])) for r in range(world)]
    results, errors = [None] * world, [None] * world

    def worker(r):
        err = C.create_string_buffer(512)
        comm = C.c_void_p()
        if lib.w2rap_step2_comm_init(uid, world, r, r, C.byref(comm), err, 512):
            errors[r] = err.value
            return
        p = T.default_params(device=r, **kw)
        g = T.Graph()
        reads = shards[r].c()
        rc = lib.w2rap_step2_run_sharded(C.byref(reads), C.byref(p), comm, C.byref(g), err, 512)
        if rc:
            errors[r] = err.value
        else:
            results[r] = T.graph_to_dict(g)
            lib.w2rap_step2_free(C.byref(g))
        lib.w2rap_step2_comm_destroy(comm)

    th = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]      # daemon: a hung rank must not block exit
    [t.start() for t in th]
    deadline = time.time() + 240
    [t.join(max(0.0, deadline - time.time())) for t in th]
    assert not any(t.is_alive() for t in th), "sharded run did not finish within 240 s (ranks alive: %s; errors so far: %s)" % ([t.is_alive() for t in th], errors)
    assert all(e is None for e in errors), errors
    return results, bounds


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_oracle(T, world):
    rs = T.rich_set(seed=12, genome=80000, cov=50, families=5, palindromes=3, plasmid=1500)
    want = T.run_oracle(rs, T.default_params(dump_kmers=1, apply_fixpaths=1))
    res, bounds = run_sharded(T, rs, world, dump_kmers=1, apply_fixpaths=1)
    for r, got in enumerate(res):
        # the graph, histogram, dictionary and whole-job counters are identical on every rank
        T.assert_graph_equal(want, dict(got, n_reads=want["n_reads"], n_bases=want["n_bases"]), "rank %d graph" % r, check_paths=False)
        # paths: this rank's shard
        lo, hi = bounds[r], bounds[r + 1]
        assert got["n_paths"] == hi - lo
        assert np.array_equal(got["path_offset"], want["path_offset"][lo:hi])
        a, b = int(want["path_off"][lo]), int(want["path_off"][hi])
        assert np.array_equal(got["path_edges"], want["path_edges"][a:b])
        assert np.array_equal(got["path_off"], want["path_off"][lo:hi + 1] - want["path_off"][lo])
    assert res[0]["timings"]["exchange_ms"] > 0


def test_sharded_forced_passes(T):
    """Hash-range counting passes under sharding: the pass filter, the per-pass exchange and the accumulation of solid records."""
    rs = T.rich_set(seed=15, genome=50000, cov=40, families=3, palindromes=2, plasmid=900)
    want = T.run_oracle(rs, T.default_params(dump_kmers=1))
    res, _ = run_sharded(T, rs, 2, dump_kmers=1, force_passes=3)
    for got in res:
        T.assert_graph_equal(want, dict(got, n_reads=want["n_reads"], n_bases=want["n_bases"]), "forced passes", check_paths=False)


def test_sharded_small_region_and_skew(T):
    """Tiny counting region (many groups, overflow fallbacks) and a high-multiplicity k-mer family under sharding."""
    rng = np.random.default_rng(5)
    rs = T.rich_set(seed=13, genome=20000, cov=30, families=2, palindromes=1, plasmid=700)
    want = T.run_oracle(rs, T.default_params(dump_kmers=1))
    res, _ = run_sharded(T, rs, 2, dump_kmers=1, table_slots=256)
    for got in res:
        T.assert_graph_equal(want, dict(got, n_reads=want["n_reads"], n_bases=want["n_bases"]), "small region", check_paths=False)


def test_graph_on_root_only(T):
    """graph_on_root_only: rank 0 receives the whole graph, the other ranks the counters, edge lengths, digests and their own paths."""
    rs = T.rich_set(seed=16, genome=40000, cov=40, families=3, palindromes=2, plasmid=900)
    want = T.run_oracle(rs, T.default_params(apply_fixpaths=1))
    res, bounds = run_sharded(T, rs, 2, apply_fixpaths=1, graph_on_root_only=1)
    T.assert_graph_equal(want, dict(res[0], n_reads=want["n_reads"], n_bases=want["n_bases"]), "rank 0 graph", check_paths=False, check_dump=False)
    other = res[1]
    assert other["n_edges"] == want["n_edges"] and other["n_vertices"] == want["n_vertices"] and other["n_edge_bases"] == want["n_edge_bases"]
    assert np.array_equal(other["edge_len"], want["edge_len"])
    assert other["edge_bases"].size == 0 and other["edge_vertices"].size == 0 and other["fwd_xlat"].size == 0 and other["involution"].size == 0
    assert other["digest_graph"] == res[0]["digest_graph"] and other["digest_paths"] == res[0]["digest_paths"]
    lo, hi = bounds[1], bounds[2]
    assert np.array_equal(other["path_offset"], want["path_offset"][lo:hi])
    a, b = int(want["path_off"][lo]), int(want["path_off"][hi])
    assert np.array_equal(other["path_edges"], want["path_edges"][a:b])


def test_sharded_places(T):
    """Step-3 places in a sharded run: every rank removes the duplicates of its shard, the unique places are all-gathered and merged;
    every rank returns the list of the whole job."""
    rs = T.rich_set(seed=17, genome=60000, cov=50, families=4, palindromes=2, plasmid=1200)
    want = T.run_oracle(rs, T.default_params(apply_fixpaths=1, places_K2=200))
    res, _ = run_sharded(T, rs, 2, apply_fixpaths=1, places_K2=200)
    assert want["n_places"] > 200
    for got in res:
        assert (got["n_places_kept"], got["n_places"]) == (want["n_places_kept"], want["n_places"])
        assert np.array_equal(got["place_off"], want["place_off"]) and np.array_equal(got["place_edges"], want["place_edges"])
