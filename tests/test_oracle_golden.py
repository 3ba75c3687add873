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


def places_from(paths, hlen, inv, K2):
    """RepathInMemory's `places` (paths/long/large/Repath.cc:46-72) restated in Python: paths = per read a list of hbv edge ids,
    hlen = bases per hbv edge, inv = the involution.  Returns (paths that passed the K2 test, sorted unique places)."""
    kept = []
    for x in paths:
        x = [int(e) for e in x]
        if sum(int(hlen[e]) - 59 for e in x) + 59 < K2:          # :57-60 (an empty path has 0 k-mers and is dropped too)
            continue
        y = [int(inv[e]) for e in reversed(x)]                   # :61-62
        kept.append(tuple(x if x < y else y))                    # :63
    return len(kept), sorted(set(kept))                          # :69-71


def places_of_result(o, K2):
    hlen = np.zeros(o["n_hbv_edges"], np.int64)
    hlen[o["fwd_xlat"]] = o["edge_len"]
    hlen[o["rev_xlat"]] = o["edge_len"]
    paths = [o["path_edges"][int(o["path_off"][r]):int(o["path_off"][r + 1])] for r in range(o["n_paths"])]
    return places_from(paths, hlen, o["involution"], K2)


@pytest.mark.parametrize("case", ["circ", "rich", "long"])
def test_places_against_the_reference_log(T, case):
    """Step-3 input.  The reference never writes its places; it prints how many paths passed the K2 test and how many
    unique places remain (tests/golden/<case>/reference_step3_places.txt, made by make_golden_places.py from the reference binary run
    on its own step-2 files, for six values of K2).  (1) The Python restatement applied to THOSE files must reproduce both counts
    exactly; (2) the C oracle's list must be the restatement's on the oracle's own graph; (3) the oracle's counts can differ from the
    log only by the extension ties its paths differ by (<= 3 reads, see test_oracle_matches_reference_files)."""
    d = os.path.join(GOLD, case)
    rows = [tuple(int(x) for x in l.split()) for l in open(os.path.join(d, "reference_step3_places.txt")).read().splitlines() if l.strip()]
    assert len(rows) >= 5
    ref = T.graph_from_reference_files(d)
    seqs = [e.tobytes() for e in ref["hbv"]["edges"]]
    ids = {s: i for i, s in enumerate(seqs)}
    inv = np.array([ids[T.revcomp(e).tobytes()] for e in ref["hbv"]["edges"]], np.int64)      # hbv.Involution (HyperBasevector.cc:648-660)
    hlen = np.array([len(e) for e in ref["hbv"]["edges"]], np.int64)
    rs = T.read_fastb_qualp(d)
    for K2, n_paths, kept, unique in rows:
        assert len(ref["paths"]) == n_paths
        nk, places = places_from(ref["paths"], hlen, inv, K2)
        assert (nk, len(places)) == (kept, unique), (K2, nk, len(places), kept, unique)
        o = T.run_oracle(rs, T.default_params(apply_fixpaths=1, places_K2=K2))
        nk2, want = places_of_result(o, K2)
        got = [tuple(int(e) for e in o["place_edges"][int(o["place_off"][i]):int(o["place_off"][i + 1])]) for i in range(o["n_places"])]
        assert o["n_places_kept"] == nk2 and got == want
        assert abs(o["n_places_kept"] - kept) <= 3 and abs(o["n_places"] - unique) <= 3
