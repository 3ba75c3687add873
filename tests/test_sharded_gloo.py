"""World-size-2 (and 4) run of the sharded step-2 protocol on CPU over gloo (tests/sharded_gloo_worker.py) against the oracle: record routing
and counting by owners, then the sharded graph stage (neighbour queries, chain-end records, circles across ranks, strand and edge
reductions, dictionary slices) with every exchange a real collective."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_protocol_over_gloo(T, tmp_path, world):
    import test_hostcheck  # noqa: F401  (builds oracle/_build/libhostcheck.so through its fixture when stale)
    so = os.path.join(T.ROOT, "oracle", "_build", "libhostcheck.so")
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, src], check=True)
    seed = 21
    port = 29500 + world + (os.getpid() % 500)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HERE, "sharded_gloo_worker.py"), str(tmp_path), str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rs = T.rich_set(seed=seed, genome=30000, cov=40, families=3, palindromes=2, plasmid=900)
    want = T.run_oracle(rs, T.default_params(dump_kmers=1, apply_fixpaths=1))
    want2 = T.run_oracle(rs, T.default_params(dump_kmers=2, want_paths=0))
    total_inst = 0
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        total_inst += int(z["n_total_inst"])
        for f in ("w0", "w1", "count", "ctx"):
            assert np.array_equal(z["allk"][f], want2["dump"][f]), f           # every rank sees the same, complete count table
        for k in ("edge_len", "edge_off", "edge_bases", "edge_vertices", "fwd_xlat", "rev_xlat"):
            assert np.array_equal(z[k], want[k]), (rank, k)                     # identical graph on every rank
        for f in ("w0", "w1", "ctx", "edge", "offset"):
            assert np.array_equal(z["dump"][f], want["dump"][f]), f
        lo, hi = int(z["lo"]), int(z["hi"])
        assert np.array_equal(z["path_offset"], want["path_offset"][lo:hi])    # paths of the rank's own shard
        a, b = int(want["path_off"][lo]), int(want["path_off"][hi])
        assert np.array_equal(z["path_edges"], want["path_edges"][a:b])
    assert total_inst == want["n_kmer_instances"]
    z = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    assert int(z["n_pieces"]) > 2 * want["n_edges"] and int(z["n_cut_rounds"]) == 1      # chains were cut at rank boundaries; the plasmid circle spans ranks
