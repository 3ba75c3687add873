"""Shared test / bench support: synthetic reads, reference file formats, oracle and C-ABI bindings.

TEST INFRASTRUCTURE.  The oracle (oracle/step2_oracle.c) and the reference binary (oracle/_ref) are checkers only;
the product is the CUDA library behind include/w2rap_step2.h, bound here with ctypes exactly as any host would.
"""
import ctypes as C
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libstep2_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "w2rap-contigger")
PKG_DIR = os.path.join(ROOT, "w2rap-contigger_b200")
PRODUCT_SO = os.path.join(PKG_DIR, "libw2rap_step2.so")
K = 60

# ---------------------------------------------------------------- ctypes mirrors of include/w2rap_step2.h


class Reads(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("bases", C.c_void_p), ("base_off", C.c_void_p), ("len", C.c_void_p),
                ("quals", C.c_void_p), ("qual_off", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("K", C.c_uint32), ("min_qual", C.c_uint32), ("min_freq", C.c_uint32),
                ("want_paths", C.c_uint32), ("apply_fixpaths", C.c_uint32), ("dump_kmers", C.c_uint32),
                ("device", C.c_int32), ("workdir", C.c_char_p), ("table_slots", C.c_uint64), ("verbose", C.c_uint32),
                ("force_passes", C.c_uint32), ("graph_on_root_only", C.c_uint32), ("places_K2", C.c_uint32)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "count_ms", "solid_ms", "adjacency_ms", "unipath_ms", "hbv_ms",
                                         "path_ms", "d2h_ms", "total_ms", "count_kernel_ms", "region_ms", "exchange_ms", "host_pre_ms", "host_post_ms", "wall_ms")] + \
               [(n, C.c_uint32) for n in ("count_launches", "kernel_launches", "count_passes", "reserved")] + \
               [("dict_ms", C.c_float), ("graph_exchange_ms", C.c_float), ("exchange_bytes", C.c_uint64), ("n_records", C.c_uint64), ("kernel_ms", C.c_float * 16), ("alloc_host_ms", C.c_float), ("reserved2", C.c_uint32), ("count_exchange_bytes", C.c_uint64),
                ("places_ms", C.c_float), ("reserved3", C.c_uint32)]


KERNEL_NAMES = ["k_good_len", "k_minimizer_map", "k_scatter_records", "k_count_smem", "k_insert_solid", "k_adjacency", "k_links",
                "k_splitter_walk", "k_splitter_finish", "k_emit_edges", "unused:10", "k_path_reads",
                "sharded:queries+ghosts", "sharded:pieces", "sharded:strands+edges", "sharded:gather dictionary"]


class KmerRec(C.Structure):
    _fields_ = [("w0", C.c_uint64), ("w1", C.c_uint64), ("count", C.c_uint32), ("ctx", C.c_uint32),
                ("edge", C.c_uint32), ("offset", C.c_uint32)]


KMER_REC_DTYPE = np.dtype([("w0", "<u8"), ("w1", "<u8"), ("count", "<u4"), ("ctx", "<u4"), ("edge", "<u4"),
                           ("offset", "<u4")])


class Graph(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_kmer_instances", C.c_uint64),
                ("n_distinct", C.c_uint64), ("n_solid", C.c_uint64), ("hist", C.c_uint64 * 101),
                ("n_edges", C.c_uint64), ("n_edge_bases", C.c_uint64), ("edge_off", C.c_void_p),
                ("edge_len", C.c_void_p), ("edge_bases", C.c_void_p),
                ("n_vertices", C.c_uint64), ("n_hbv_edges", C.c_uint64), ("edge_vertices", C.c_void_p),
                ("fwd_xlat", C.c_void_p), ("rev_xlat", C.c_void_p), ("involution", C.c_void_p),
                ("n_paths", C.c_uint64), ("n_path_edges", C.c_uint64), ("path_offset", C.c_void_p),
                ("path_off", C.c_void_p), ("path_edges", C.c_void_p), ("n_pathed", C.c_uint64),
                ("n_multipathed", C.c_uint64), ("digest_graph", C.c_uint64), ("digest_paths", C.c_uint64),
                ("n_dump", C.c_uint64), ("dump", C.c_void_p),
                ("n_places_kept", C.c_uint64), ("n_places", C.c_uint64), ("n_place_edges", C.c_uint64), ("place_off", C.c_void_p), ("place_edges", C.c_void_p),
                ("timings", Timings), ("_owner", C.c_void_p)]


class SynthParams(C.Structure):
    _fields_ = [("genome_len", C.c_uint64), ("read_len", C.c_uint32), ("coverage", C.c_uint32), ("seed", C.c_uint64),
                ("het_per_10k", C.c_uint32), ("reserved", C.c_uint32), ("n_reads", C.c_uint64), ("first_read", C.c_uint64)]


ABI_VERSION = 2


def default_params(min_qual=7, min_freq=4, want_paths=1, apply_fixpaths=0, dump_kmers=0, workdir=None,
                   table_slots=0, device=-1, verbose=0, force_passes=0, graph_on_root_only=0, places_K2=0):
    return Params(ABI_VERSION, K, min_qual, min_freq, want_paths, apply_fixpaths, dump_kmers, device,
                  workdir.encode() if workdir else None, table_slots, verbose, force_passes, graph_on_root_only, places_K2)


def _arr(ptr, n, dtype):
    """Copy n elements of dtype from a C pointer into a fresh numpy array."""
    n = int(n)
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def graph_to_dict(g):
    """Convert a filled w2rap_graph into plain numpy data (copies; safe to free the graph afterwards)."""
    ne, npth = int(g.n_edges), int(g.n_paths)
    edge_off = _arr(g.edge_off, ne + 1, "<u8") if g.edge_off else np.zeros(1, dtype="<u8")       # (null: graph_on_root_only on another rank)
    d = dict(
        n_reads=int(g.n_reads), n_bases=int(g.n_bases), n_kmer_instances=int(g.n_kmer_instances),
        n_distinct=int(g.n_distinct), n_solid=int(g.n_solid), hist=np.array(list(g.hist), dtype=np.uint64),
        n_edges=ne, n_edge_bases=int(g.n_edge_bases), edge_off=edge_off, edge_len=_arr(g.edge_len, ne, "<u4"),
        edge_bases=_arr(g.edge_bases, int(edge_off[-1]) if ne and g.edge_off else 0, "u1"),
        n_vertices=int(g.n_vertices), n_hbv_edges=int(g.n_hbv_edges),
        edge_vertices=_arr(g.edge_vertices, 4 * ne if g.edge_vertices else 0, "<i4").reshape(-1, 4),
        fwd_xlat=_arr(g.fwd_xlat, ne, "<i4"), rev_xlat=_arr(g.rev_xlat, ne, "<i4"),
        involution=_arr(g.involution, g.n_hbv_edges, "<i4"), digest_graph=int(g.digest_graph), digest_paths=int(g.digest_paths),
        n_paths=npth, n_path_edges=int(g.n_path_edges), path_offset=_arr(g.path_offset, npth, "<i4"),
        path_off=_arr(g.path_off, npth + 1 if npth else 0, "<u8"), path_edges=_arr(g.path_edges, g.n_path_edges, "<i4"),
        n_pathed=int(g.n_pathed), n_multipathed=int(g.n_multipathed),
        dump=_arr(g.dump, g.n_dump, KMER_REC_DTYPE),
        n_places_kept=int(g.n_places_kept), n_places=int(g.n_places), place_off=_arr(g.place_off, g.n_places + 1 if g.place_off else 0, "<u8"),
        place_edges=_arr(g.place_edges, g.n_place_edges, "<i4"),
        timings={n: (list(getattr(g.timings, n)) if n == "kernel_ms" else getattr(g.timings, n)) for n, _ in Timings._fields_},
    )
    return d


# ---------------------------------------------------------------- building the checkers

def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build_oracle(force=False):
    """gcc-compile the CPU restatement + read-simulation support into oracle/_build/libstep2_oracle.so."""
    srcs = [os.path.join(ORACLE_DIR, "step2_oracle.c"), os.path.join(ORACLE_DIR, "readsim_support.c")]
    if not force and os.path.exists(ORACLE_SO) and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in srcs):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    _run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-Wall", "-o", ORACLE_SO] + srcs)
    return ORACLE_SO


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_step2_run.argtypes = [C.POINTER(Reads), C.POINTER(Params), C.POINTER(Graph)]
        lib.oracle_step2_run.restype = C.c_int
        lib.oracle_step2_free.argtypes = [C.POINTER(Graph)]
        lib.oracle_step2_free.restype = None
        lib.sim_flatten_reads.restype = C.c_size_t
        lib.sim_flatten_reads.argtypes = [C.c_uint64, C.c_uint32] + [C.c_void_p] * 7 + [C.c_int]
        lib.sim_pq_encode.restype = C.c_size_t
        lib.sim_pq_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        lib.oracle_pq_decode.restype = C.c_size_t
        lib.oracle_pq_decode.argtypes = [C.c_void_p, C.c_void_p]
        _oracle = lib
    return _oracle


_product = None


def product_lib():
    """The product: the CUDA library behind include/w2rap_step2.h.  Never falls back to anything."""
    global _product
    if _product is None:
        if not os.path.exists(PRODUCT_SO):
            raise RuntimeError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % PRODUCT_SO)
        lib = C.CDLL(PRODUCT_SO)
        E = [C.c_char_p, C.c_size_t]
        lib.w2rap_step2_abi_version.restype = C.c_int
        lib.w2rap_step2_build_info.restype = C.c_char_p
        lib.w2rap_step2_device_count.restype = C.c_int
        lib.w2rap_step2_run.argtypes = [C.POINTER(Reads), C.POINTER(Params), C.POINTER(Graph)] + E
        lib.w2rap_step2_upload.argtypes = [C.POINTER(Reads), C.c_int, C.POINTER(C.c_void_p)] + E
        lib.w2rap_step2_run_resident.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Graph)] + E
        lib.w2rap_step2_release.argtypes = [C.c_void_p]
        lib.w2rap_step2_release.restype = None
        lib.w2rap_step2_free.argtypes = [C.POINTER(Graph)]
        lib.w2rap_step2_free.restype = None
        lib.w2rap_step2_run_files.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), C.POINTER(Graph)] + E
        for n in ("w2rap_write_hbv", "w2rap_write_paths", "w2rap_write_freqs"):
            getattr(lib, n).argtypes = [C.c_char_p, C.POINTER(Graph)] + E
        lib.w2rap_step2_synth.argtypes = [C.POINTER(SynthParams), C.c_int, C.POINTER(C.c_void_p)] + E
        lib.w2rap_step2_download_reads.argtypes = [C.c_void_p, C.POINTER(Reads)] + E
        lib.w2rap_step2_free_host_reads.argtypes = [C.POINTER(Reads)]
        lib.w2rap_step2_free_host_reads.restype = None
        for n in ("w2rap_write_fastb", "w2rap_write_qualp"):
            getattr(lib, n).argtypes = [C.c_char_p, C.POINTER(Reads)] + E
        lib.w2rap_read_fastb_qualp.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Reads)] + E
        lib.w2rap_step2_comm_unique_id.argtypes = [C.c_void_p] + E
        lib.w2rap_step2_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)] + E
        lib.w2rap_step2_comm_destroy.argtypes = [C.c_void_p]
        lib.w2rap_step2_comm_destroy.restype = None
        lib.w2rap_step2_run_sharded.argtypes = [C.POINTER(Reads), C.POINTER(Params), C.c_void_p, C.POINTER(Graph)] + E
        lib.w2rap_step2_run_sharded_resident.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.POINTER(Graph)] + E
        _product = lib
    return _product


ABI_SYMBOLS = ["w2rap_step2_host_alloc", "w2rap_step2_host_free", "w2rap_step2_abi_version", "w2rap_step2_build_info", "w2rap_step2_device_count", "w2rap_step2_run",
               "w2rap_step2_upload", "w2rap_step2_run_resident", "w2rap_step2_release", "w2rap_step2_free",
               "w2rap_step2_run_files", "w2rap_write_hbv", "w2rap_write_paths", "w2rap_write_freqs",
               "w2rap_step2_synth", "w2rap_step2_download_reads", "w2rap_step2_free_host_reads",
               "w2rap_write_fastb", "w2rap_write_qualp", "w2rap_read_fastb_qualp",
               "w2rap_step2_comm_unique_id", "w2rap_step2_comm_init", "w2rap_step2_comm_destroy", "w2rap_step2_run_sharded",
               "w2rap_step2_run_sharded_resident"]


# ---------------------------------------------------------------- flattened read sets

class ReadSet:
    """Host-side flattened read store (numpy owners + the ctypes view passed through the C ABI)."""

    def __init__(self, bases, base_off, lens, quals, qual_off):
        # 32 bytes of slack behind both byte stores: the device functions (also run on the host by tests/hostcheck) read packed
        # bases with aligned 8-byte loads that may touch a few bytes past the last base.  The C ABI itself needs no padding.
        self._bases_store = np.concatenate([np.ascontiguousarray(bases, dtype=np.uint8), np.zeros(32, np.uint8)])
        self._quals_store = np.concatenate([np.ascontiguousarray(quals, dtype=np.uint8), np.zeros(32, np.uint8)])
        self.bases = self._bases_store[:len(self._bases_store) - 32]
        self.base_off = np.ascontiguousarray(base_off, dtype=np.uint64)
        self.len = np.ascontiguousarray(lens, dtype=np.uint32)
        self.quals = self._quals_store[:len(self._quals_store) - 32]
        self.qual_off = np.ascontiguousarray(qual_off, dtype=np.uint64)
        self.n = len(self.len)
        if self.len.size == 0:
            self.len = np.zeros(1, np.uint32)[:0]

    def c(self):
        return Reads(self.n, self.bases.ctypes.data, self.base_off.ctypes.data,
                     self.len.ctypes.data if self.n else np.zeros(1, np.uint32).ctypes.data,
                     self.quals.ctypes.data, self.qual_off.ctypes.data)

    @property
    def n_bases(self):
        return int(self.len.sum())

    def read_codes(self, r):
        nb = int(self.len[r])
        b = self.bases[int(self.base_off[r]):int(self.base_off[r]) + (nb + 3) // 4]
        return (((b[:, None] >> (np.arange(4) * 2)[None, :]) & 3).reshape(-1)[:nb]).astype(np.uint8)

    def subset(self, idx):
        """A new ReadSet holding the reads `idx` (array of indices), re-flattened."""
        idx = np.asarray(idx, dtype=np.int64)
        bl = (self.base_off[idx + 1] - self.base_off[idx]).astype(np.int64)
        ql = (self.qual_off[idx + 1] - self.qual_off[idx]).astype(np.int64)
        bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.uint64)
        qo = np.concatenate([[0], np.cumsum(ql)]).astype(np.uint64)
        bases = np.concatenate([self.bases[int(self.base_off[i]):int(self.base_off[i + 1])] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
        quals = np.concatenate([self.quals[int(self.qual_off[i]):int(self.qual_off[i + 1])] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
        return ReadSet(bases, bo, self.len[idx], quals, qo)


def flatten_reads(codes, quals, lens, pq_mode=0):
    """codes/quals: [n, stride] uint8 matrices; lens: per-read lengths.  Returns a ReadSet."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    quals = np.ascontiguousarray(quals, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(lens)
    stride = codes.shape[1] if n else 1
    assert quals.shape == codes.shape
    bases_out = np.zeros(int(((lens.astype(np.int64) + 3) // 4).sum()) + 1, np.uint8)
    quals_out = np.zeros(int((2 * lens.astype(np.int64) + 4).sum()) + 1, np.uint8)
    base_off = np.zeros(n + 1, np.uint64)
    qual_off = np.zeros(n + 1, np.uint64)
    nq = oracle_lib().sim_flatten_reads(n, stride, codes.ctypes.data, quals.ctypes.data, lens.ctypes.data,
                                        bases_out.ctypes.data, base_off.ctypes.data, quals_out.ctypes.data,
                                        qual_off.ctypes.data, pq_mode)
    return ReadSet(bases_out[:int(base_off[-1])], base_off, lens, quals_out[:nq], qual_off)


# ---------------------------------------------------------------- synthetic genomes and reads

def revcomp(codes):
    return (3 - codes[::-1]).astype(np.uint8)


def make_genome(rng, size, n_repeat_families=0, n_palindromes=0, repeat_div=0.01):
    """i.i.d. ACGT plus planted repeat families (3-30 copies, 300-3000 bp, ~1% diverged) and 120-bp palindromes."""
    g = rng.integers(0, 4, size, dtype=np.uint8)
    for _ in range(n_repeat_families):
        L = int(rng.integers(300, 3001))
        if L * 3 >= size:
            continue
        fam = rng.integers(0, 4, L, dtype=np.uint8)
        for _c in range(int(rng.integers(3, 31))):
            p = int(rng.integers(0, size - L))
            cp = fam.copy()
            mut = rng.random(L) < repeat_div
            cp[mut] = (cp[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
            g[p:p + L] = cp if rng.random() < 0.5 else revcomp(cp)
    for _ in range(n_palindromes):
        half = rng.integers(0, 4, 60, dtype=np.uint8)
        p = int(rng.integers(0, size - 120))
        g[p:p + 60] = half
        g[p + 60:p + 120] = revcomp(half)
    return g


def add_snps(rng, g, rate):
    h = g.copy()
    mut = rng.random(len(g)) < rate
    h[mut] = (h[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
    return h


def simulate_reads(rng, replicons, n_pairs, read_len=250, frag_mean=500, frag_sd=50, tail_max=60, sporadic=0.01,
                   errors=True, vary_len=False):
    """PE reads from a list of (sequence, is_circular, weight) replicons.  Returns codes, quals [2*n_pairs, read_len], lens.

    Qualities: Q37 body, the last U(0,tail_max) bases Q2-15, `sporadic` of the bases Q2-19; substitution errors are drawn at
    10^(-Q/10); no Ns.  Pairs are interleaved (read 2i, 2i+1), read 2 is the reverse strand of the fragment end.
    """
    n = 2 * n_pairs
    codes = np.zeros((n, read_len), np.uint8)
    w = np.array([r[2] for r in replicons], dtype=float)
    which = rng.choice(len(replicons), size=n_pairs, p=w / w.sum())
    flen = np.clip(np.rint(rng.normal(frag_mean, frag_sd, n_pairs)).astype(np.int64), read_len, None)
    strand = rng.random(n_pairs) < 0.5
    for ri, (seq, circ, _w) in enumerate(replicons):
        sel = np.nonzero(which == ri)[0]
        if len(sel) == 0:
            continue
        L = len(seq)
        ext = np.concatenate([seq, seq[:2000]]) if circ else seq
        fl = np.minimum(flen[sel], len(ext) - 1)
        hi = (L if circ else np.maximum(L - fl, 1))
        start = (rng.random(len(sel)) * hi).astype(np.int64)
        idx1 = start[:, None] + np.arange(read_len)[None, :]
        idx2 = (start + fl - 1)[:, None] - np.arange(read_len)[None, :]
        idx1 = np.clip(idx1, 0, len(ext) - 1)
        idx2 = np.clip(idx2, 0, len(ext) - 1)
        r1 = ext[idx1]
        r2 = 3 - ext[idx2]
        sw = strand[sel]
        a = np.where(sw[:, None], r2, r1)
        b = np.where(sw[:, None], r1, r2)
        codes[2 * sel] = a
        codes[2 * sel + 1] = b
    quals = np.full((n, read_len), 37, np.uint8)
    tail = rng.integers(0, tail_max + 1, n)
    pos = np.arange(read_len)[None, :]
    in_tail = pos >= (read_len - tail)[:, None]
    quals[in_tail] = rng.integers(2, 16, int(in_tail.sum()), dtype=np.uint8)
    sp = rng.random((n, read_len)) < sporadic
    quals[sp] = rng.integers(2, 20, int(sp.sum()), dtype=np.uint8)
    if errors:
        perr = 10.0 ** (-quals.astype(float) / 10.0)
        err = rng.random((n, read_len)) < perr
        codes[err] = (codes[err] + rng.integers(1, 4, int(err.sum()), dtype=np.uint8)) % 4
    lens = np.full(n, read_len, np.uint32)
    if vary_len:
        lens = rng.integers(30, read_len + 1, n).astype(np.uint32)
    return codes, quals, lens


def smoke_set(seed=1, genome=20000, cov=40, read_len=250):
    rng = np.random.default_rng(seed)
    g = make_genome(rng, genome)
    n_pairs = genome * cov // (2 * read_len)
    return flatten_reads(*simulate_reads(rng, [(g, False, 1.0)], n_pairs, read_len))


def rich_set(seed=2, genome=100000, cov=60, read_len=250, families=6, palindromes=4, het=0.01, plasmid=3000, pq_mode=0,
             vary_len=False):
    """Repeat families + planted palindromes + a SNP haplotype + a circular plasmid: exercises branches, bubbles,
    palindromic k-mers, smooth circles and gapped paths."""
    rng = np.random.default_rng(seed)
    g = make_genome(rng, genome, families, palindromes)
    reps = [(g, False, float(genome))]
    if het > 0:
        reps = [(g, False, genome / 2.0), (add_snps(rng, g, het), False, genome / 2.0)]
    if plasmid:
        reps.append((rng.integers(0, 4, plasmid, dtype=np.uint8), True, float(plasmid) * 3))
    total = sum(r[2] for r in reps)
    n_pairs = int(total * cov // (2 * read_len))
    return flatten_reads(*simulate_reads(rng, reps, n_pairs, read_len, vary_len=vary_len), pq_mode=pq_mode)


# ---------------------------------------------------------------- reference file formats

def write_feudal(path, var_data, offsets, fixed, sz_fixed, sz_x, sz_a=1):
    """feudal/FeudalControlBlock.h:43-53,156-163: 24-B header, var data, (N+1) absolute u64 offsets, fixed data."""
    n = len(offsets) - 1
    var_off = 24 + len(var_data)
    fixed_off = var_off + 8 * (n + 1)
    with open(path, "wb") as f:
        f.write(struct.pack("<IBBBBQQ", n & 0xffffffff, 1, sz_fixed, sz_x, sz_a, var_off, fixed_off))
        f.write(np.ascontiguousarray(var_data, np.uint8).tobytes())
        f.write((np.asarray(offsets, np.uint64) + np.uint64(24)).astype("<u8").tobytes())
        f.write(fixed)


def write_fastb_qualp(dirname, rs):
    os.makedirs(dirname, exist_ok=True)
    write_feudal(os.path.join(dirname, "frag_reads_orig.fastb"), rs.bases[:int(rs.base_off[-1])], rs.base_off,
                 rs.len.astype("<u4").tobytes(), 4, 16)
    write_feudal(os.path.join(dirname, "frag_reads_orig.qualp"), rs.quals[:int(rs.qual_off[-1])], rs.qual_off, b"", 0, 8)


def read_feudal(path):
    d = np.fromfile(path, dtype=np.uint8)
    n, flags, szf, szx, sza, var_off, fixed_off = struct.unpack("<IBBBBQQ", d[:24].tobytes())
    offs = d[var_off:fixed_off].view("<u8").astype(np.uint64)
    return d[24:var_off].copy(), offs - np.uint64(24), d[fixed_off:].copy()


def read_fastb_qualp(dirname):
    bases, boff, fixed = read_feudal(os.path.join(dirname, "frag_reads_orig.fastb"))
    quals, qoff, _ = read_feudal(os.path.join(dirname, "frag_reads_orig.qualp"))
    return ReadSet(bases, boff, fixed.view("<u4"), quals, qoff)


def parse_hbv(path):
    """BINWRITE stream of a HyperBasevector: K, from_, from_edge_obj_, to_edge_obj_, edges_."""
    d = open(path, "rb").read()
    assert d[:8] == b"BINWRITE", d[:8]
    pos = 8
    (k,) = struct.unpack_from("<i", d, pos); pos += 4

    def vecvec(pos):
        (n,) = struct.unpack_from("<Q", d, pos); pos += 8
        out = []
        for _ in range(n):
            (m,) = struct.unpack_from("<Q", d, pos); pos += 8
            out.append(np.frombuffer(d, "<i4", m, pos).copy()); pos += 4 * m
        return out, pos
    frm, pos = vecvec(pos)
    feo, pos = vecvec(pos)
    teo, pos = vecvec(pos)
    (ne,) = struct.unpack_from("<Q", d, pos); pos += 8
    edges = []
    for _ in range(ne):
        (sz,) = struct.unpack_from("<I", d, pos); pos += 4
        nb = (sz + 3) // 4
        b = np.frombuffer(d, "u1", nb, pos); pos += nb
        edges.append((((b[:, None] >> (np.arange(4) * 2)[None, :]) & 3).reshape(-1)[:sz]).astype(np.uint8))
    return dict(K=k, from_=frm, from_edge_obj=feo, to_edge_obj=teo, edges=edges, trailing=len(d) - pos)


def parse_paths(path):
    """paths/long/ReadPath.cc:6-20: u64 n; per read i32 offset, u16 len, len x i32."""
    d = open(path, "rb").read()
    (n,) = struct.unpack_from("<Q", d, 0)
    pos = 8
    offs = np.zeros(n, np.int32)
    paths = []
    for i in range(n):
        o, m = struct.unpack_from("<iH", d, pos); pos += 6
        offs[i] = o
        paths.append(np.frombuffer(d, "<i4", m, pos).copy()); pos += 4 * m
    assert pos == len(d)
    return offs, paths


def parse_freqs(path):
    h = np.zeros(101, np.uint64)
    for line in open(path):
        i, c = line.split(",")
        h[int(i)] = int(c)
    return h


def run_reference_step2(dirname, threads=1, min_freq=4, min_qual=7, timeout=3600, binary=None):
    """Runs the reference's own step 2 (+FixPaths, +dump) on <dirname>/frag_reads_orig.{fastb,qualp}."""
    binary = binary or REF_BIN
    if not os.path.exists(binary):
        raise RuntimeError("reference binary missing: make -C oracle")
    env = dict(os.environ, OMP_PROC_BIND="spread", MALLOC_PER_THREAD="1")
    out = _run([binary, "-t", str(threads), "-o", dirname, "-p", "x", "-r", "dummy", "--from_step", "2", "--to_step", "2",
                "--dump_perf", "1", "--min_freq", str(min_freq), "--min_qual", str(min_qual)], env=env, timeout=timeout)
    perf = {}
    pf = os.path.join(dirname, "x.perf")
    if os.path.exists(pf):
        for line in open(pf):
            t = [x.strip() for x in line.split(",")]
            if len(t) >= 3 and t[0] == "TIME":
                perf[t[1]] = float(t[2])
    return out, perf


# ---------------------------------------------------------------- running the oracle / the product

def run_oracle(rs, params=None):
    params = params or default_params()
    g = Graph()
    rc = oracle_lib().oracle_step2_run(C.byref(rs.c()), C.byref(params), C.byref(g))
    if rc != 0:
        raise RuntimeError("oracle_step2_run failed: %d" % rc)
    d = graph_to_dict(g)
    oracle_lib().oracle_step2_free(C.byref(g))
    return d


def run_product(rs, params=None):
    """Through the C ABI with host buffers: exactly what the reference-side wrapper calls."""
    params = params or default_params()
    lib = product_lib()
    g = Graph()
    err = C.create_string_buffer(1024)
    rc = lib.w2rap_step2_run(C.byref(rs.c()), C.byref(params), C.byref(g), err, len(err))
    if rc != 0:
        raise RuntimeError("w2rap_step2_run failed (%d): %s" % (rc, err.value.decode(errors="replace")))
    d = graph_to_dict(g)
    lib.w2rap_step2_free(C.byref(g))
    return d


# ---------------------------------------------------------------- comparison

def unpack_edges(d):
    out = []
    for i in range(d["n_edges"]):
        nb = int(d["edge_len"][i])
        b = d["edge_bases"][int(d["edge_off"][i]):int(d["edge_off"][i]) + (nb + 3) // 4]
        out.append((((b[:, None] >> (np.arange(4) * 2)[None, :]) & 3).reshape(-1)[:nb]).astype(np.uint8))
    return out


def graph_from_reference_files(dirname, prefix="x"):
    """Rebuild (canonical edges sorted by sequence, vertices, paths relabelled) from the reference's .hbv/.paths so that it
    can be compared with a w2rap_graph dict.  Returns dict with hbv edge sequences keyed the same way."""
    hbv = parse_hbv(os.path.join(dirname, prefix + ".small_K.hbv"))
    offs, paths = parse_paths(os.path.join(dirname, prefix + ".small_K.paths"))
    ne = len(hbv["edges"])
    left = np.full(ne, -1, np.int64)
    right = np.full(ne, -1, np.int64)
    for v, (tos, eos) in enumerate(zip(hbv["from_"], hbv["from_edge_obj"])):
        for w, e in zip(tos, eos):
            left[e] = v
            right[e] = w
    return dict(hbv=hbv, left=left, right=right, path_offset=offs, paths=paths,
                hist=parse_freqs(os.path.join(dirname, "small_K.freqs")))


def hbv_view(d):
    """From a w2rap_graph dict: per hbv edge id -> (sequence bytes, left vertex, right vertex)."""
    edges = unpack_edges(d)
    n = d["n_hbv_edges"]
    seqs = [None] * n
    left = np.zeros(n, np.int64)
    right = np.zeros(n, np.int64)
    for i, e in enumerate(edges):
        f, r = int(d["fwd_xlat"][i]), int(d["rev_xlat"][i])
        seqs[f] = e.tobytes()
        left[f], right[f] = d["edge_vertices"][i, 0], d["edge_vertices"][i, 1]
        if r != f:
            seqs[r] = revcomp(e).tobytes()
            left[r], right[r] = d["edge_vertices"][i, 2], d["edge_vertices"][i, 3]
    return seqs, left, right


def compare_with_reference(d, ref, post_fixpaths=True):
    """d: w2rap_graph dict (oracle or product, apply_fixpaths=1); ref: graph_from_reference_files().
    Equality modulo edge relabelling; returns a report dict (mismatch counts; tie-explained path differences)."""
    rep = {}
    seqs, left, right = hbv_view(d)
    rseqs = [e.tobytes() for e in ref["hbv"]["edges"]]
    rep["n_hbv_edges"] = (len(seqs), len(rseqs))
    ours = {s: i for i, s in enumerate(seqs)}
    rep["edge_set_equal"] = (len(ours) == len(seqs)) and set(ours) == set(rseqs) and len(rseqs) == len(seqs)
    rep["hist_equal"] = bool(np.array_equal(d["hist"][1:], ref["hist"][1:]))
    if not rep["edge_set_equal"]:
        return rep
    r2o = np.array([ours[s] for s in rseqs], dtype=np.int64)      # reference hbv id -> our hbv id
    rep["n_vertices"] = (d["n_vertices"], len(ref["hbv"]["from_"]))
    rep["vertices_equal"] = bool(np.array_equal(left[r2o], ref["left"]) and np.array_equal(right[r2o], ref["right"]))
    # paths
    n = len(ref["paths"])
    bad = []
    ties = 0
    po, pe, poff = d["path_off"], d["path_edges"], d["path_offset"]
    elen = np.array([len(s) for s in seqs])
    for r in range(n):
        mine = pe[int(po[r]):int(po[r + 1])]
        theirs = r2o[ref["paths"][r]] if len(ref["paths"][r]) else np.zeros(0, np.int64)
        if len(mine) == len(theirs) and np.array_equal(mine, theirs) and poff[r] == ref["path_offset"][r]:
            continue
        # tolerated: extension tie between parallel edges of equal length
        ok = len(mine) == len(theirs) and poff[r] == ref["path_offset"][r]
        if ok:
            for a, b in zip(mine, theirs):
                if a != b and not (left[a] == left[b] and right[a] == right[b] and elen[a] == elen[b]):
                    ok = False
                    break
        if ok:
            ties += 1
        else:
            bad.append(r)
    rep["path_ties"] = ties
    rep["path_mismatches"] = bad
    return rep


def compare_with_reference_fast(d, ref):
    """compare_with_reference for millions of reads: the path comparison is done on flattened arrays; only reads that differ are
    looked at one by one (tolerated: extension ties between parallel equal-length edges)."""
    rep = {}
    seqs, left, right = hbv_view(d)
    rseqs = [e.tobytes() for e in ref["hbv"]["edges"]]
    ours = {s: i for i, s in enumerate(seqs)}
    rep["n_hbv_edges"] = (len(seqs), len(rseqs))
    rep["edge_set_equal"] = len(ours) == len(seqs) and len(rseqs) == len(seqs) and set(ours) == set(rseqs)
    rep["hist_equal"] = bool(np.array_equal(d["hist"][1:], ref["hist"][1:]))
    if not rep["edge_set_equal"]:
        return rep
    r2o = np.array([ours[s] for s in rseqs], dtype=np.int64)
    rep["vertices_equal"] = bool(np.array_equal(left[r2o], ref["left"]) and np.array_equal(right[r2o], ref["right"]))
    n = len(ref["paths"])
    rlens = np.fromiter((len(p) for p in ref["paths"]), dtype=np.int64, count=n)
    rflat = r2o[np.concatenate(ref["paths"]).astype(np.int64)] if rlens.sum() else np.zeros(0, np.int64)
    po = d["path_off"].astype(np.int64)
    olens = po[1:] - po[:-1]
    same_len = (olens == rlens) & (d["path_offset"] == ref["path_offset"])
    roff = np.concatenate([[0], np.cumsum(rlens)])
    differs = ~same_len
    if same_len.any():
        # element-wise comparison of the reads whose lengths agree
        idx = np.nonzero(same_len)[0]
        reps = np.repeat(idx, rlens[idx])
        within = np.arange(len(reps)) - np.repeat(np.cumsum(rlens[idx]) - rlens[idx], rlens[idx])
        neq = d["path_edges"][po[reps] + within] != rflat[roff[reps] + within]
        differs[np.unique(reps[neq])] = True
    elen = np.array([len(s) for s in seqs])
    ties, bad = 0, []
    for r in np.nonzero(differs)[0]:
        mine = d["path_edges"][po[r]:po[r + 1]]
        theirs = rflat[roff[r]:roff[r + 1]]
        ok = len(mine) == len(theirs) and d["path_offset"][r] == ref["path_offset"][r]
        if ok:
            for a, b in zip(mine, theirs):
                if a != b and not (left[a] == left[b] and right[a] == right[b] and elen[a] == elen[b]):
                    ok = False
                    break
        if ok:
            ties += 1
        else:
            bad.append(int(r))
    rep["path_ties"] = ties
    rep["path_mismatches"] = bad
    rep["n_reads"] = n
    return rep


def assert_graph_equal(a, b, what="graph", check_paths=True, check_dump=True):
    """Exact equality of two w2rap_graph dicts (oracle vs product): both use the sorted-by-sequence edge order."""
    for k in ("n_reads", "n_bases", "n_kmer_instances", "n_distinct", "n_solid", "n_edges", "n_edge_bases", "n_vertices",
              "n_hbv_edges"):
        assert a[k] == b[k], "%s: %s differs: %s vs %s" % (what, k, a[k], b[k])
    for k in ("hist", "edge_len", "edge_off", "edge_bases", "edge_vertices", "fwd_xlat", "rev_xlat", "involution"):
        assert np.array_equal(a[k], b[k]), "%s: array %s differs" % (what, k)
    if check_dump and (len(a["dump"]) or len(b["dump"])):
        assert len(a["dump"]) == len(b["dump"]), "%s: dump sizes differ %d vs %d" % (what, len(a["dump"]), len(b["dump"]))
        for f in ("w0", "w1", "count", "ctx", "edge", "offset"):
            assert np.array_equal(a["dump"][f], b["dump"][f]), "%s: dump field %s differs" % (what, f)
    if check_paths:
        for k in ("n_paths", "n_path_edges", "n_pathed", "n_multipathed"):
            assert a[k] == b[k], "%s: %s differs: %s vs %s" % (what, k, a[k], b[k])
        for k in ("path_offset", "path_off", "path_edges"):
            if not np.array_equal(a[k], b[k]):
                bad = np.nonzero(a["path_offset"] != b["path_offset"])[0]
                raise AssertionError("%s: array %s differs (first offset mismatch at reads %s)" % (what, k, bad[:5]))
        # step-3 places (params.places_K2): counts, then the lists
        for k in ("n_places_kept", "n_places"):
            assert a[k] == b[k], "%s: %s differs: %s vs %s" % (what, k, a[k], b[k])
        for k in ("place_off", "place_edges"):
            assert np.array_equal(a[k], b[k]), "%s: array %s differs" % (what, k)


if __name__ == "__main__":
    build_oracle(force=True)
    print("built", ORACLE_SO)
    sys.exit(0)
