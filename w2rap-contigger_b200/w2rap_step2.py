"""Host-side mirror of the reference's step-2 entry point over the C ABI (include/w2rap_step2.h).

The reference is C++ and has no Python; this module exists for the tests and the benchmark, and mirrors
`buildReadQGraph(reads, quals, doFillGaps, doJoinOverlaps, minQual, minFreq, minFreq2Fract, maxGapSize, pHBV, pPaths, K,
workdir, tmpdir, disk_batches)` (paths/long/BuildReadQGraph.h:24-29): same argument meaning, same abort-on-error behaviour
(an exception here where the reference calls FatalErr / CRD::exit).  The C++ drop-in lives in host/BuildReadQGraph_b200.cc.

There is no CPU fallback: without libw2rap_step2.so or without a B200 every call raises.
"""
import ctypes as C
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tests"))
import w2r_testlib as _T  # noqa: E402  (ctypes struct mirrors + numpy conversion, shared with the tests)


class Step2Error(RuntimeError):
    pass


def build_read_q_graph(reads, min_qual=7, min_freq=4, want_paths=True, K=60, workdir=None, fix_paths=False, device=-1,
                       do_fill_gaps=False, do_join_overlaps=False, verbose=False):
    """reads: tests.w2r_testlib.ReadSet (flattened vecbvec + VecPQVec).  Returns the graph/paths as a dict of numpy arrays."""
    if do_fill_gaps or do_join_overlaps:
        raise Step2Error("fillGaps / joinOverlaps are off on the reference's only call site (w2rap-contigger.cc:336-338) and are not built")
    p = _T.default_params(min_qual=min_qual, min_freq=min_freq, want_paths=int(bool(want_paths)), apply_fixpaths=int(bool(fix_paths)),
                          workdir=workdir, device=device, verbose=int(bool(verbose)))
    p.K = K
    try:
        return _T.run_product(reads, p)
    except RuntimeError as e:
        raise Step2Error(str(e))


def device_count():
    return _T.product_lib().w2rap_step2_device_count()
