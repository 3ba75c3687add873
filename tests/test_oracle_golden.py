"""The CPU oracle (oracle/step2_oracle.c) against outputs of the UNMODIFIED reference binary.

tests/golden/{circ,rich,long} were produced by tests/golden/make_golden.py running oracle/_ref/w2rap-contigger
(--from_step 2 --to_step 2, -t 1).  Equality is modulo the reference's racy edge numbering: hbv edges are matched by
sequence; vertex ids, incidence, the k-mer histogram and every read path must then be identical; a path difference is
tolerated only when it is an extension tie between parallel equal-length edges.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["circ", "rich", "long"])
def test_oracle_matches_reference_files(T, case):
    d = os.path.join(GOLD, case)
    rs = T.read_fastb_qualp(d)
    o = T.run_oracle(rs, T.default_params(apply_fixpaths=1))
    rep = T.compare_with_reference(o, T.graph_from_reference_files(d))
    assert rep["edge_set_equal"], rep
    assert rep["hist_equal"], rep
    assert rep["vertices_equal"], rep
    assert rep["path_mismatches"] == [], rep
    assert rep["path_ties"] <= 3, rep
    # the reference's own stdout counters
    lines = open(os.path.join(d, "reference_stdout.txt")).read().splitlines()
    counted = int([l for l in lines if "kmers counted" in l][0].split()[0])
    solid = int([l for l in lines if "kmers with Freq" in l][0].split()[0])
    assert (o["n_distinct"], o["n_solid"]) == (counted, solid)
    # "pathed / spanning junctions" are printed before FixPaths
    o2 = T.run_oracle(rs, T.default_params(apply_fixpaths=0))
    pathed = [l for l in lines if "reads pathed" in l][0].replace(",", "").split()
    assert (o2["n_pathed"], o2["n_multipathed"]) == (int(pathed[0]), int(pathed[5]))


def test_circ_case_has_circles_and_palindromes(T):
    d = os.path.join(GOLD, "circ")
    o = T.run_oracle(T.read_fastb_qualp(d), T.default_params(want_paths=0))
    assert o["n_hbv_edges"] < 2 * o["n_edges"]          # palindromic edges have a single hbv id
    edges = T.unpack_edges(o)
    circles = [e for e in edges if len(e) > 60 and np.array_equal(e[:59], e[-59:])]
    assert len(circles) >= 8
    assert {len(e) & 1 for e in circles} == {0, 1}      # both length parities (middle-base rule vs outside-in compare)


def test_pqvec_decode_against_reference_encoder(T):
    """.qualp written by the reference's step 1 from a FASTQ whose qualities we know."""
    d = os.path.join(GOLD, "step1")
    rs = T.read_fastb_qualp(d)
    exp = np.load(os.path.join(d, "expected.npz"))
    assert np.array_equal(rs.len, exp["lens"])
    lib = T.oracle_lib()
    off = 0
    buf = np.zeros(70000, np.uint8)
    for r in range(rs.n):
        L = int(rs.len[r])
        n = lib.oracle_pq_decode(rs.quals.ctypes.data + int(rs.qual_off[r]), buf.ctypes.data)
        assert n == L
        assert np.array_equal(buf[:L], exp["quals"][off:off + L])
        assert np.array_equal(rs.read_codes(r), exp["bases"][off:off + L])
        off += L


def test_pq_encoders_roundtrip(T):
    rng = np.random.default_rng(3)
    lib = T.oracle_lib()
    out = np.zeros(2000, np.uint8)
    dec = np.zeros(1000, np.uint8)
    for mode in (0, 1):
        for trial in range(200):
            n = int(rng.integers(0, 400))
            kind = trial % 4
            q = (np.full(n, 37) if kind == 0 else rng.integers(0, 64, n) if kind == 1 else
                 np.clip(30 + rng.integers(-2, 3, n), 0, 63) if kind == 2 else rng.integers(2, 4, n)).astype(np.uint8)
            nb = lib.sim_pq_encode(q.ctypes.data if n else out.ctypes.data, n, out.ctypes.data, mode)
            assert out[nb - 1] == 0
            m = lib.oracle_pq_decode(out.ctypes.data, dec.ctypes.data)
            assert m == n and np.array_equal(dec[:n], q)


def test_oracle_edge_cases(T):
    """Empty store, reads shorter than K, reads whose quality never reaches K good bases, exactly-K good reads."""
    rng = np.random.default_rng(9)
    o = T.run_oracle(T.ReadSet(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8),
                               np.zeros(1, np.uint64)))
    assert o["n_edges"] == 0 and o["n_paths"] == 0 and o["n_vertices"] == 0
    g = rng.integers(0, 4, 400, dtype=np.uint8)
    codes = np.zeros((40, 250), np.uint8)
    quals = np.full((40, 250), 30, np.uint8)
    lens = np.full(40, 250, np.uint32)
    for i in range(40):
        codes[i] = g[i:i + 250]
    lens[0] = 10            # shorter than K: a single gap part
    lens[1] = 59
    quals[2, :] = 2         # never good
    quals[3, 60:] = 2       # exactly K good quals -> good_len == K -> contributes nothing (strict >)
    quals[4, 61:] = 2       # K+1 good
    rs = T.flatten_reads(codes, quals, lens)
    o = T.run_oracle(rs, T.default_params(min_freq=1))
    inst = (250 - 59) * 35 + 2
    assert o["n_kmer_instances"] == inst
    assert o["n_paths"] == 40 and o["path_off"][1] == o["path_off"][0]   # read 0 has no path
