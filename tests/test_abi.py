"""The C-ABI library loads, exports every symbol include/w2rap_step2.h declares, and refuses to compute without a B200."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "w2rap_step2.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(w2rap_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_testlib_agree(T):
    assert declared_symbols() == sorted(T.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(T):
    lib = T.product_lib()
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export " + s
    assert lib.w2rap_step2_abi_version() == 2
    assert b"sm_100a" in lib.w2rap_step2_build_info()


def test_struct_sizes_match_header(T, tmp_path):
    """ctypes mirrors vs the real header, measured by compiling a C program against include/w2rap_step2.h."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "w2rap_step2.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(w2rap_reads),sizeof(w2rap_params),sizeof(w2rap_kmer_rec),sizeof(w2rap_timings),sizeof(w2rap_graph),'
                   'sizeof(w2rap_synth_params),offsetof(w2rap_graph,timings),offsetof(w2rap_graph,path_edges));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(T.Reads), C.sizeof(T.Params), C.sizeof(T.KmerRec), C.sizeof(T.Timings), C.sizeof(T.Graph), C.sizeof(T.SynthParams),
            T.Graph.timings.offset, T.Graph.path_edges.offset]
    assert got == want


def test_no_cpu_fallback(T):
    """Without a usable B200 the compute entry points must fail loudly (W2RAP_ERR_NO_DEVICE), never compute on the CPU."""
    lib = T.product_lib()
    if lib.w2rap_step2_device_count() > 0:
        pytest.skip("a B200 is visible")
    rs = T.smoke_set(seed=1, genome=3000, cov=10)
    with pytest.raises(RuntimeError, match=r"failed \(2\)"):
        T.run_product(rs)
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    reads = rs.c()
    assert lib.w2rap_step2_upload(C.byref(reads), -1, C.byref(h), err, 256) == 2
    sp = T.SynthParams(100000, 250, 10, 1, 0, 0, 0)
    assert lib.w2rap_step2_synth(C.byref(sp), -1, C.byref(h), err, 256) == 2


def test_bad_arguments_are_rejected_before_any_device_work(T):
    lib = T.product_lib()
    rs = T.smoke_set(seed=1, genome=3000, cov=10)
    g = T.Graph()
    err = C.create_string_buffer(256)
    p = T.default_params()
    p.K = 31
    reads = rs.c()
    assert lib.w2rap_step2_run(C.byref(reads), C.byref(p), C.byref(g), err, 256) == 1
    assert b"K=31" in err.value
    p = T.default_params(min_freq=0)
    assert lib.w2rap_step2_run(C.byref(reads), C.byref(p), C.byref(g), err, 256) == 1


def test_host_format_writers_roundtrip(T, tmp_path):
    """w2rap_write_fastb/qualp/hbv/paths/freqs produce the reference's formats (parsed back by the independent Python readers)."""
    lib = T.product_lib()
    rs = T.rich_set(seed=4, genome=8000, cov=30, families=2, palindromes=1, plasmid=500)
    err = C.create_string_buffer(256)
    reads = rs.c()
    fb, qp = str(tmp_path / "frag_reads_orig.fastb").encode(), str(tmp_path / "frag_reads_orig.qualp").encode()
    assert lib.w2rap_write_fastb(fb, C.byref(reads), err, 256) == 0
    assert lib.w2rap_write_qualp(qp, C.byref(reads), err, 256) == 0
    back = T.read_fastb_qualp(str(tmp_path))
    for f in ("bases", "base_off", "len", "quals", "qual_off"):
        assert np.array_equal(getattr(back, f), getattr(rs, f)), f
    r2 = T.Reads()
    assert lib.w2rap_read_fastb_qualp(fb, qp, C.byref(r2), err, 256) == 0
    assert r2.n_reads == rs.n
    assert np.array_equal(T._arr(r2.len, rs.n, "<u4"), rs.len)
    assert np.array_equal(T._arr(r2.quals, int(rs.qual_off[-1]), "u1"), rs.quals[:int(rs.qual_off[-1])])
    lib.w2rap_step2_free_host_reads(C.byref(r2))
    # graph writers, fed with the oracle's graph
    o = T.run_oracle(rs, T.default_params(apply_fixpaths=1))
    g = T.Graph()
    keep = {}
    for name, ctype in (("edge_off", "<u8"), ("edge_len", "<u4"), ("edge_bases", "u1"), ("fwd_xlat", "<i4"), ("rev_xlat", "<i4"),
                        ("path_offset", "<i4"), ("path_off", "<u8"), ("path_edges", "<i4")):
        keep[name] = np.ascontiguousarray(o[name], dtype=ctype)
        setattr(g, name, keep[name].ctypes.data)
    keep["ev"] = np.ascontiguousarray(o["edge_vertices"].reshape(-1), dtype="<i4")
    g.edge_vertices = keep["ev"].ctypes.data
    g.n_edges, g.n_vertices, g.n_hbv_edges, g.n_paths = o["n_edges"], o["n_vertices"], o["n_hbv_edges"], o["n_paths"]
    for i in range(101):
        g.hist[i] = int(o["hist"][i])
    assert lib.w2rap_write_hbv(str(tmp_path / "x.small_K.hbv").encode(), C.byref(g), err, 256) == 0, err.value
    assert lib.w2rap_write_paths(str(tmp_path / "x.small_K.paths").encode(), C.byref(g), err, 256) == 0
    assert lib.w2rap_write_freqs(str(tmp_path / "small_K.freqs").encode(), C.byref(g), err, 256) == 0
    ref = T.graph_from_reference_files(str(tmp_path))
    assert ref["hbv"]["trailing"] == 0 and ref["hbv"]["K"] == 60
    rep = T.compare_with_reference(o, ref)
    assert rep["edge_set_equal"] and rep["vertices_equal"] and rep["hist_equal"] and rep["path_mismatches"] == [] and rep["path_ties"] == 0
    # byte-level: to_edge_obj lists follow the AddEdge order as well
    seqs, left, right = T.hbv_view(o)
    for v, lst in enumerate(ref["hbv"]["to_edge_obj"]):
        assert all(right[e] == v for e in lst)
        assert list(lst) == sorted(lst, key=lambda e: (left[e], e))
