import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    import astropaint_b200 as ap
    ap.paint_bucket.verbose = False
    yield


@pytest.fixture(scope="session")
def nside():
    return 64


@pytest.fixture(scope="module")
def test_catalog():
    from astropaint_b200 import Catalog
    return Catalog("test")


@pytest.fixture(scope="function")
def test_canvas(test_catalog, nside):
    from astropaint_b200 import Canvas
    return Canvas(test_catalog, nside)


def apply_func_to_cutout(patch, mult_by, add_to):
    return mult_by * patch + add_to


def map_close(got, want, rtol=1e-6):
    """The parity bar of BASELINE.json: max-abs difference relative to the map's max magnitude, and
    per-pixel relative difference wherever the reference magnitude exceeds 1e-12 x max."""
    import numpy as np
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "NaN pattern differs"
    g, w = np.where(nan_g, 0.0, got), np.where(nan_w, 0.0, want)
    scale = np.abs(w).max()
    if scale == 0:
        assert np.abs(g).max() == 0
        return 0.0
    err = np.abs(g - w).max() / scale
    assert err <= rtol, f"max-abs error {err} > {rtol}"
    big = np.abs(w) > 1e-12 * scale
    rel = np.abs(g[big] - w[big]) / np.abs(w[big])
    assert rel.max() <= rtol, f"per-pixel relative error {rel.max()} > {rtol}"
    return err


def pytest_sessionfinish(session, exitstatus):
    """With a -DAP_DEBUG_BOUNDS library (tools/debug_bounds.sh) the run fails if any index was out of range."""
    if os.environ.get("AP_REPORT_BOUNDS"):
        from astropaint_b200 import _capi
        n = _capi.lib().ap_debug_bounds_errors()
        print(f"\n[ap] out-of-range indices counted by the instrumented build: {n}")
        if n != 0:
            session.exitstatus = 1
