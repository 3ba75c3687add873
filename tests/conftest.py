import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a B200: without one they are skipped, not failed (the product has no CPU path to fall back to)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import w2r_testlib
        n = w2r_testlib.product_lib().w2rap_step2_device_count()
        why = "no sm_100 device visible"
    except Exception as e:      # library not built: nothing to run the GPU tests with
        n, why = 0, "CUDA library unavailable: %s" % e
    if n == 0:
        skip = pytest.mark.skip(reason=why)
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def T():
    import w2r_testlib
    w2r_testlib.build_oracle()
    return w2r_testlib
