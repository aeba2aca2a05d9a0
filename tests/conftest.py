import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session (fails loudly without a GPU
    or without the built libenf_b200.so -- there is no fallback to test)."""
    import enf_b200
    c = enf_b200.default_context()
    yield c


# ---- the error metric every parity test uses (stated in DESIGN.md §parity) -------------
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def rel_err(got, ref, floor=0.0):
    """max_i |got_i - ref_i| / (|ref_i| + scale),  scale = RMS(ref) (1 if ref == 0).

    A mixed relative/absolute measure: plain element-wise relative error is
    ill-posed where the reference value passes through zero (e.g. ladj ~ 0, or
    center_stretch near x = 0 where the reference itself loses digits,
    SURVEY §7).  Equivalent to numpy.allclose(rtol=tol, atol=tol*RMS(ref))."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = float(np.sqrt(np.mean(ref * ref)))
    scale = max(scale, floor)      # floor: for outputs that are mathematically zero (e.g. the gradient of an
    if not np.isfinite(scale) or scale == 0.0:   # outermost Householder stack: |Hy| = |y|), where ref is rounding noise
        scale = 1.0
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + scale)))


def strict_rel_err(got, ref):
    """Plain element-wise relative error max |got - ref| / |ref| over the elements with |ref| > RMS(ref):
    the un-softened companion of rel_err (it ignores only the elements near zero, where a relative error is
    ill-posed).  Reported in every assertion message; the Float32 budget is checked on rel_err."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    m = np.abs(ref) > np.sqrt(np.mean(ref * ref))
    if not m.any():
        return 0.0
    return float(np.max(np.abs(got - ref)[m] / np.abs(ref)[m]))


def assert_close(got, ref, dtype, what="", factor=1.0, floor=0.0):
    tol = TOL[np.dtype(dtype)] * factor
    e = rel_err(got, ref, floor)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e} (plain relative error on |ref| > RMS: {strict_rel_err(got, ref):.3e})"
    return e
