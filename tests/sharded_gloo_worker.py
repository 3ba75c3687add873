"""Worker of tests/test_sharded_gloo.py: one rank of the sharded step-2 PROTOCOL on CPU over torch.distributed (gloo).

The device stages are stood in for by the host-compiled device functions (tests/hostcheck); what is under test is the
multi-rank logic: read sharding by index, routing of super-k-mer records to the owner of their minimiser partition (the product's own
mini_part / owner_of_partition, csrc/extract.cuh, csrc/shard.cuh), counting by owners, all-gather of the counted k-mers, identical graph on
every rank, paths by shard.  The GPU implementation of the same protocol (NCCL) is tested in tests/test_gpu_sharded.py.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import w2r_testlib as T  # noqa: E402


def main():
    out_dir, seed = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    hc = C.CDLL(os.path.join(T.ROOT, "oracle", "_build", "libhostcheck.so"))
    hc.hc_extract_records.argtypes = [C.POINTER(T.Reads), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    hc.hc_count_records.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    hc.hc_graph.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(T.Reads), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(T.Graph)]
    hc.hc_free.argtypes = [C.c_void_p]
    hc.hc_graph_free.argtypes = [C.POINTER(T.Graph)]

    rs = T.rich_set(seed=seed, genome=30000, cov=40, families=3, palindromes=2, plasmid=900)     # every rank derives the same set
    n = rs.n - (rs.n % (2 * world))
    bounds = [(n // world) * r for r in range(world)] + [rs.n]
    shard = rs.subset(np.arange(bounds[rank], bounds[rank + 1]))
    logP = 7
    # map: records of my shard, with the owner rank of each
    recs_p, own_p, nrec = C.c_void_p(), C.c_void_p(), C.c_uint64()
    reads = shard.c()
    assert hc.hc_extract_records(C.byref(reads), 7, logP, world, C.byref(recs_p), C.byref(own_p), C.byref(nrec)) == 0
    recs = T._arr(recs_p.value, 4 * nrec.value, "<u8").reshape(-1, 4)      # super-k-mer records (csrc/extract.cuh: SkmRec)
    owner = T._arr(own_p.value, nrec.value, "<u4")
    hc.hc_free(recs_p); hc.hc_free(own_p)
    # swizzle: all-to-all by owner
    outgoing = [recs[owner == d] for d in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, outgoing)
    mine = np.ascontiguousarray(np.concatenate([gathered[s][rank] for s in range(world)]))
    # reduce: count what I own
    dptr, nd = C.c_void_p(), C.c_uint64()
    assert hc.hc_count_records(mine.ctypes.data, len(mine), C.byref(dptr), C.byref(nd)) == 0
    counted = T._arr(dptr.value, nd.value, T.KMER_REC_DTYPE)
    hc.hc_free(dptr)
    # all-gather the counted k-mers; every rank builds the whole graph, then paths its shard
    allc = [None] * world
    dist.all_gather_object(allc, counted)
    allk = np.concatenate(allc)
    allk = allk[np.lexsort((allk["w1"], allk["w0"]))]
    assert len(np.unique(allk[["w0", "w1"]])) == len(allk), "a k-mer was counted by two owners"
    g = T.Graph()
    assert hc.hc_graph(allk.ctypes.data, len(allk), 4, C.byref(reads), 1, 1, 24, 8, C.byref(g)) == 0
    d = T.graph_to_dict(g)
    hc.hc_graph_free(C.byref(g))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=bounds[rank], hi=bounds[rank + 1], n_total_inst=int(sum(int((((x[:, 3] >> np.uint64(56)) & np.uint64(31)) + np.uint64(1)).sum()) for x in outgoing)),
             **{k: d[k] for k in ("hist", "edge_len", "edge_off", "edge_bases", "edge_vertices", "fwd_xlat", "rev_xlat", "path_offset", "path_off",
                                  "path_edges", "dump")}, allk=allk)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
