"""CPU tests of the boundary and the host-side logic: the C-ABI library loads and
exports every symbol include/enf_b200.h declares (no compute calls without a
GPU), and the Python mirror of the reference interface (flattening, inversion,
parameter packing, gradient unpacking, batching, optimizer) agrees with the
oracle's restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from chains import build
from oracle import enf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "enf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(enf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import enf_b200 as E
    from enf_b200 import _lib as L
    lib = E.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/enf_b200.h but not exported by libenf_b200.so"
    bound = {n for n, _, _ in L.SYMBOLS}
    assert bound == set(names), (bound ^ set(names))
    assert lib.enf_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device enf_init must fail loudly (error code + message)."""
    import enf_b200 as E
    lib = E.lib()
    n = C.c_int(-1)
    rc = lib.enf_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.enf_init(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.enf_last_error(None)
    with pytest.raises(E.EnfError):
        E.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "euclidiannormalizingflows.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the literal oracle", "").replace("CPU oracle", ""), os.path.join(dirpath, f)


def test_flatten_inverse_pack_match_oracle():
    import enf_b200 as E
    spec = ["cc", "ji", "hh3", "ss", "cs", "jo", "hhv"]
    D = 5
    fo = build(O, spec, D, np.random.default_rng(1))
    fe = build(E, spec, D, np.random.default_rng(1))
    lo, le = O.flatten(fo), E.flatten(fe)
    assert [type(a).__name__ for a in lo] == [type(a).__name__ for a in le]
    io, ie = O.flatten(O.inverse(fo)), E.flatten(E.inverse(fe))
    assert [type(a).__name__ for a in io] == [type(a).__name__ for a in ie]
    for a, b in zip(io, ie):
        for n in a.fields:
            np.testing.assert_allclose(np.asarray(getattr(a, n)), np.asarray(getattr(b, n)), rtol=1e-15)
    for a, b in zip(E.flatten(E.inverse(E.inverse(fe))), le):        # inverse(inverse(f)) == f up to 1/(1/a) rounding
        assert type(a) is type(b)
        for n in a.fields:
            np.testing.assert_allclose(np.asarray(getattr(a, n)), np.asarray(getattr(b, n)), rtol=1e-15)
    assert E.inverse(E.JohnsonTrafo(1.0, 2.0, 3.0, 4.0)) == E.JohnsonTrafoInv(1.0, 2.0, 3.0, 4.0)
    packed = E.pack_params(le, D, np.float64)
    assert packed.size == 3 * D + 4 * D + 3 * D + 2 * D + 3 * D + 4 * D + D
    np.testing.assert_array_equal(packed[:D], lo[0].a)
    np.testing.assert_array_equal(packed[7 * D:10 * D], np.asarray(lo[2].V).ravel(order="F"))
    g = E.unpack_grads(fe, np.arange(packed.size, dtype=np.float64), D)
    leaf = g
    while "inner" in leaf:
        leaf = leaf["inner"]
    np.testing.assert_array_equal(leaf["a"], np.arange(D))                 # innermost op comes first in the packing
    last = g
    while "outer" in last:
        last = last["outer"]
    assert last["V"].shape == (D, 1)                                       # vector V -> D x 1 gradient (householder_trafo.jl:39)
    # scalar fields are expanded for the ABI and un-broadcast (summed) in the gradient
    s = E.CenterStretch(4.0, 2.0, 3.0)
    np.testing.assert_array_equal(E.pack_params([s], 3, np.float64), [4, 4, 4, 2, 2, 2, 3, 3, 3])
    assert E.unpack_grads(s, np.arange(9.0), 3) == {"a": 3.0, "b": 12.0, "c": 21.0}
    # result type = float(promote_type(...)) (src/center_stretch.jl:5)
    assert E.result_dtype(E.CenterStretch(np.float32(1), np.float32(1), np.float32(0)), np.float32) == np.float32
    assert E.result_dtype(E.CenterStretch(1.0, 1.0, 0.0), np.float32) == np.float64
    assert E.result_dtype(E.CenterStretch(7, 2, 4), np.float32) == np.float64   # Python ints carry no float type: f64 like the oracle default


def test_batching_and_optimizer_match_oracle():
    import enf_b200 as E
    for n, nb in ((10, 3), (10, 4), (100000, 1000), (7, 2), (1000, 7)):
        assert E.batch_ranges(n, nb) == O.batch_ranges(n, nb)
    rng = np.random.default_rng(2)
    spec = ["cc", "hh2", "ss"]
    fo, fe = build(O, spec, 3, np.random.default_rng(3)), build(E, spec, 3, np.random.default_rng(3))
    so, se = O.optim_setup(O.ADAGrad(), fo), E.setup(E.ADAGrad(), fe)
    for _ in range(3):
        g = {"outer": {"outer": {"a": rng.standard_normal(3), "b": rng.standard_normal(3)},
                       "inner": {"V": rng.standard_normal((3, 2))}},
             "inner": {"a": rng.standard_normal(3), "b": rng.standard_normal(3), "c": rng.standard_normal(3)}}
        so, fo = O.optim_update(O.ADAGrad(), so, fo, g)
        se, fe = E.update(E.ADAGrad(), se, fe, g)
    for a, b in zip(O.flatten(fo), E.flatten(fe)):
        for n in a.fields:
            np.testing.assert_allclose(np.asarray(getattr(a, n)), np.asarray(getattr(b, n)), rtol=1e-14)
    V = E.flatten(fe)[1].V
    np.testing.assert_allclose((V * V).sum(0), 1.0, rtol=1e-13)


def test_shard_helpers():
    from enf_b200 import dist
    for n, w in ((10, 3), (1000, 8), (7, 8), (0, 2)):
        parts = [dist.shard_columns(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    ranges = O.batch_ranges(103, 10)
    for w in (1, 2, 4):
        cover = [dist.shard_batches(ranges, r, w) for r in range(w)]
        for bi, (s, e) in enumerate(ranges):
            assert sum(c[bi][1] - c[bi][0] for c in cover) == e - s
