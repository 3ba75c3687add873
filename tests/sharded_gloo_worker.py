"""Worker of tests/test_sharded_gloo.py: one rank of the sharded step-2 PROTOCOL on CPU over torch.distributed (gloo).

The device stages are stood in for by the host-compiled device functions (tests/hostcheck); what is under test is the
multi-rank logic: read sharding by index, routing of super-k-mer records to the owner of their minimiser partition (the product's own
mini_part / owner_of_partition, csrc/extract.cuh, csrc/shard.cuh), counting by owners, then the SHARDED graph stage (csrc/shardgraph.cuh):
neighbour queries and ghost entries, chain-end records gathered and ranked, circles that span ranks, strand flags and edge bases
reduced, dictionary slices built per rank — identical graph on every rank, paths by shard.  The GPU implementation of the same protocol (NCCL) is tested in tests/test_gpu_sharded.py.
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
    # ---- the SHARDED graph stage (csrc/shardgraph.cuh; pipeline.cu: graph_stage_sharded): every rank keeps the solid k-mers it
    # counted; neighbour queries, ghost contexts, chain-end records, strand flags, edge bases and dictionary slices are exchanged
    # over gloo, the per-rank phases are the host-compiled device functions (tests/hostcheck: SgRank)
    V = C.c_void_p
    for name, res, args in (("sg_new", V, [V, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]), ("sg_delete", None, [V]),
                            ("sg_queries", C.c_uint64, [V, C.c_uint32, C.POINTER(V)]), ("sg_answer", None, [V, V, C.c_uint64, V]),
                            ("sg_insert_ghosts", None, [V, C.c_uint32, V]), ("sg_adjacency", None, [V]), ("sg_ctx_answer", None, [V, V, C.c_uint64, V]),
                            ("sg_apply_ghost_ctx", None, [V, C.c_uint32, V]), ("sg_links", C.c_int, [V]),
                            ("sg_rank_and_pieces", C.c_int, [V, C.POINTER(V), C.POINTER(C.c_uint64)]), ("sg_set_pieces", C.c_int64, [V, V, C.c_uint64, V]),
                            ("sg_cycle_nodes", C.c_uint64, [V, C.POINTER(V)]), ("sg_apply_cuts", None, [V, V, C.c_uint64]), ("sg_strands", C.c_int, [V, V]),
                            ("sg_edges", C.c_uint64, [V, V, C.POINTER(V)]), ("sg_entries", C.c_uint64, [V, C.c_uint32, C.POINTER(V)]),
                            ("sg_sizeof", C.c_uint32, [C.c_int]), ("sg_owner", C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]),
                            ("sg_finish", C.c_int, [V, V, V, C.c_uint64, C.POINTER(T.Reads), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(T.Graph)])):
        getattr(hc, name).restype = res
        getattr(hc, name).argtypes = args

    def buf(ptr, nbytes):
        return np.frombuffer((C.c_char * nbytes).from_address(ptr), np.uint8).copy() if nbytes else np.zeros(0, np.uint8)

    def alltoall(chunks):                       # chunks[d] -> rank d; returns what every rank sent me, by source
        everyone = [None] * world
        dist.all_gather_object(everyone, chunks)
        return [everyone[s][rank] for s in range(world)]

    def allgather(x):
        everyone = [None] * world
        dist.all_gather_object(everyone, x)
        return everyone

    solid = counted[counted["count"] >= 4]
    # (the counting owner of a k-mer is the owner of its minimiser partition: the graph stage keeps that sharding)
    assert all(hc.sg_owner(int(k["w0"]), int(k["w1"]), logP, world) == rank for k in solid[:200])
    n_solid = int(sum(allgather(len(solid))))
    h = hc.sg_new(solid.ctypes.data, len(solid), world, rank, logP)
    # round 1: neighbour queries -> owners -> slots back -> ghost entries; adjacency; round 1b: the ghosts' pruned contexts
    qs = []
    for d in range(world):
        ptr = V()
        n = hc.sg_queries(h, d, C.byref(ptr))
        qs.append(buf(ptr.value, 16 * n))
    asked = alltoall(qs)
    slots_for = []
    for s_ in range(world):
        n = len(asked[s_]) // 16
        rep = np.zeros(n, np.uint32)
        hc.sg_answer(h, asked[s_].ctypes.data, n, rep.ctypes.data)
        slots_for.append(rep)
    replies = alltoall(slots_for)
    for d in range(world):
        hc.sg_insert_ghosts(h, d, np.ascontiguousarray(replies[d]).ctypes.data)
    hc.sg_adjacency(h)
    dist.barrier()
    ctx_for = []
    for s_ in range(world):
        ctx = np.zeros(len(slots_for[s_]), np.uint32)
        hc.sg_ctx_answer(h, slots_for[s_].ctypes.data, len(ctx), ctx.ctypes.data)
        ctx_for.append(ctx)
    ctxs = alltoall(ctx_for)
    for d in range(world):
        hc.sg_apply_ghost_ctx(h, d, np.ascontiguousarray(ctxs[d]).ctypes.data)
    assert hc.sg_links(h) == 0
    # round 2: chain-end records gathered and ranked on every rank; circles that span ranks are cut and the ranking repeats
    psz, csz, esz = hc.sg_sizeof(0), hc.sg_sizeof(1), hc.sg_sizeof(2)
    n_cut_rounds = 0
    for iteration in range(3):
        assert iteration < 2
        ptr, n = V(), C.c_uint64()
        assert hc.sg_rank_and_pieces(h, C.byref(ptr), C.byref(n)) == 0
        parts = allgather(buf(ptr.value, psz * n.value))
        off = np.concatenate([[0], np.cumsum([len(x) // psz for x in parts])]).astype(np.uint64)
        allp = np.ascontiguousarray(np.concatenate(parts))
        un = hc.sg_set_pieces(h, allp.ctypes.data, len(allp) // psz, off.ctypes.data)
        assert un >= 0, un
        if un == 0:
            break
        ptr = V()
        n = hc.sg_cycle_nodes(h, C.byref(ptr))
        nodes = np.ascontiguousarray(np.concatenate(allgather(buf(ptr.value, csz * n))))
        hc.sg_apply_cuts(h, nodes.ctypes.data, len(nodes) // csz)
        n_cut_rounds += 1
    n_pieces = len(allp) // psz
    # strands: keep flags max-reduced; edges: bases OR-reduced
    keep = np.zeros(n_pieces, np.uint8)
    assert hc.sg_strands(h, keep.ctypes.data) == 0
    keep = np.ascontiguousarray(np.maximum.reduce(allgather(keep)))
    ptr = V()
    nb = hc.sg_edges(h, keep.ctypes.data, C.byref(ptr))
    eb = np.bitwise_or.reduce(allgather(buf(ptr.value, nb)))
    C.memmove(ptr.value, np.ascontiguousarray(eb).ctypes.data, nb)
    # the pathing dictionary: finished entries to the rank that builds their hash slice, slices (here: their entry lists) all-gathered
    ents = []
    for d in range(world):
        ptr = V()
        n = hc.sg_entries(h, d, C.byref(ptr))
        ents.append(buf(ptr.value, esz * n))
        hc.hc_free(ptr)
    mine_slice = np.ascontiguousarray(np.concatenate(alltoall(ents)))
    slices = [np.ascontiguousarray(x) for x in allgather(mine_slice)]
    sl_ptr = (V * world)(*[x.ctypes.data for x in slices])
    sl_n = (C.c_uint64 * world)(*[len(x) // esz for x in slices])
    g = T.Graph()
    assert hc.sg_finish(h, sl_ptr, sl_n, n_solid, C.byref(reads), 1, 1, 24, 8, C.byref(g)) == 0
    d = T.graph_to_dict(g)
    hc.hc_graph_free(C.byref(g))
    hc.sg_delete(h)
    allc = allgather(counted)
    allk = np.concatenate(allc)
    allk = allk[np.lexsort((allk["w1"], allk["w0"]))]
    assert len(np.unique(allk[["w0", "w1"]])) == len(allk), "a k-mer was counted by two owners"
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=bounds[rank], hi=bounds[rank + 1], n_total_inst=int(sum(int((((x[:, 3] >> np.uint64(56)) & np.uint64(31)) + np.uint64(1)).sum()) for x in outgoing)),
             n_pieces=n_pieces, n_cut_rounds=n_cut_rounds,
             **{k: d[k] for k in ("hist", "edge_len", "edge_off", "edge_bases", "edge_vertices", "fwd_xlat", "rev_xlat", "path_offset", "path_off",
                                  "path_edges", "dump")}, allk=allk)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
