"""CPU tests of the oracle: pinned to the reference's known-answer values and to
the properties the reference's own tests check (test/*.jl), with torch-float64
autograd / finite differences standing in for ForwardDiff and Zygote."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from chains import build, flat_grads
from oracle import enf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- known-answer values: test/test_center_stretch.jl:18-19, test/test_johnson_trafo.jl:21-22
def test_golden_known_answers():
    y = O.center_stretch(np.float32(1), 7, 2, 4)
    assert y.dtype == np.float32 and abs(float(y) - 11.927293) < 2e-6          # `≈` on Float32
    y = O.center_contract(np.float32(12), 7, 2, 4)
    assert y.dtype == np.float32 and abs(float(y) - 1.063464) < 2e-7
    assert O.johnsontrafo(0.3, 1, 3, -4, 0.5) == pytest.approx(9.544817734776984, rel=1e-15)
    assert O.johnsontrafo_inv(0.3, 1, 3, -4, 0.5) == pytest.approx(-4.1177281942392545, rel=1e-15)
    # type promotion: test/test_johnson_trafo.jl:18-19
    assert np.asarray(O.johnsontrafo(0.5, 1, 2, 3, 4)).dtype == np.float64


def test_center_stretch_roundtrip_and_ladj():
    """test/test_center_stretch.jl:21-26."""
    X = np.random.default_rng(0).standard_normal(1000)
    Y = O.center_stretch(X, 7, 2, 4)
    np.testing.assert_allclose(O.center_contract(Y, 7, 2, 4), X, rtol=1.5e-8, atol=1e-12)
    x = torch.tensor(4.2, dtype=torch.float64, requires_grad=True)
    O.center_contract(x, 4, 2, 3).backward()
    assert float(O.center_contract_ladj(4.2, 4, 2, 3)) == pytest.approx(np.log(abs(float(x.grad))), rel=1e-12)
    y0 = float(O.center_contract(4.2, 4, 2, 3))
    y = torch.tensor(y0, dtype=torch.float64, requires_grad=True)
    O.center_stretch(y, 4, 2, 3).backward()
    assert -float(O.center_contract_ladj(4.2, 4, 2, 3)) == pytest.approx(np.log(abs(float(y.grad))), rel=1e-9)


def test_johnson_roundtrip_and_ladj():
    """test/test_johnson_trafo.jl:24-29."""
    K = np.random.default_rng(1).standard_normal(10000)
    Z = O.johnsontrafo_inv(K, -2, 1, 0, 2.5)
    np.testing.assert_allclose(O.johnsontrafo(Z, -2, 1, 0, 2.5), K, rtol=1.5e-8, atol=1e-12)
    for fn, ladj in ((O.johnsontrafo, O.johnsontrafo_ladj), (O.johnsontrafo_inv, O.johnsontrafo_inv_ladj)):
        x = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
        fn(x, 4.2, 4, 2, 3).backward()
        assert float(ladj(0.5, 4.2, 4, 2, 3)) == pytest.approx(np.log(abs(float(x.grad))), rel=1e-12)


def _dense_householder(v):
    return np.eye(len(v)) - 2 * np.outer(v, v) / (v @ v)


def test_householder_properties():
    """test/test_householder_trafo.jl:18-25,36-43."""
    rng = np.random.default_rng(2)
    v, x, X, V = rng.random(5), rng.random(5), rng.random((5, 3)), rng.random((5, 3))
    np.testing.assert_allclose(O.householder_trafo(v, x), _dense_householder(v) @ x, rtol=1e-12)
    np.testing.assert_allclose(O.householder_trafo(v, O.householder_trafo(v, x)), x, rtol=1e-12)
    np.testing.assert_allclose(O.householder_trafo(v, X), _dense_householder(v) @ X, rtol=1e-12)
    np.testing.assert_allclose(O.householder_trafo(v, X), np.stack([O.householder_trafo(v, c) for c in X.T], 1), rtol=1e-14)
    # chained = H_K ... H_1 : defines the reflection order
    H = np.eye(5)
    for i in range(3):
        H = _dense_householder(V[:, i]) @ H
    np.testing.assert_allclose(O.chained_householder_trafo(V, x), H @ x, rtol=1e-12)
    np.testing.assert_allclose(O.chained_householder_trafo(V[:, ::-1], O.chained_householder_trafo(V, x)), x, rtol=1e-12)
    y, l = O.with_logabsdet_jacobian(O.HouseholderTrafo(V), X)
    assert l.shape == (3,) and not l.any()


def test_householder_rrules_match_autograd():
    """test/test_householder_trafo.jl:27-33,49-55 (ForwardDiff there, torch autograd here)."""
    rng = np.random.default_rng(3)
    for shape in ((5,), (5, 4)):
        x = rng.random(shape)
        dO = rng.standard_normal(shape)
        v = rng.random(5)
        vt = torch.tensor(v, requires_grad=True)
        xt = torch.tensor(x, requires_grad=True)
        (O.householder_trafo(vt, xt) * torch.tensor(dO)).sum().backward()
        np.testing.assert_allclose(O.householder_trafo_pullback_v(v, x, dO), vt.grad.numpy(), rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(O.householder_trafo_pullback_x(v, x, dO), xt.grad.numpy(), rtol=1e-10, atol=1e-13)
        V = rng.random((5, 3))
        Vt = torch.tensor(V, requires_grad=True)
        xt = torch.tensor(x, requires_grad=True)
        yt = O.chained_householder_trafo(Vt, xt)
        (yt * torch.tensor(dO)).sum().backward()
        X2 = x if x.ndim == 2 else x[:, None]
        d2 = dO if dO.ndim == 2 else dO[:, None]
        y2 = yt.detach().numpy() if x.ndim == 2 else yt.detach().numpy()[:, None]
        np.testing.assert_allclose(O.chained_householder_trafo_pullback_V(V, X2, y2, d2), Vt.grad.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(O.chained_householder_trafo_pullback_x(V, x, None, dO), xt.grad.numpy(), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("code", ["cs", "cc", "jo", "ji", "ss", "hh3"])
def test_inverse_and_wlaj_against_jacobian(code):
    """InverseFunctions.test_inverse and ChangesOfVariables.test_with_logabsdet_jacobian
    (test/test_center_stretch.jl:49-62, test/test_johnson_trafo.jl:56-69)."""
    D = 3
    f = build(O, [code], D, np.random.default_rng(4))
    rng = np.random.default_rng(5)
    for _ in range(3):
        x = rng.standard_normal(D) * 2
        y, ladj = O.with_logabsdet_jacobian(f, x)
        np.testing.assert_allclose(O.apply(O.inverse(f), y), x, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(O.apply(O.inverse(O.inverse(f)), x), y, rtol=1e-12)
        xt = torch.tensor(x, dtype=torch.float64)
        J = torch.autograd.functional.jacobian(lambda z: O.apply(_torchify(f), z), xt).numpy()
        assert float(ladj) == pytest.approx(np.linalg.slogdet(J)[1], abs=1e-10)
    # matrix path == column-wise path (test/test_center_stretch.jl:64-70)
    X = rng.standard_normal((D, 4))
    Y, L = O.with_logabsdet_jacobian(f, X)
    cols = [O.with_logabsdet_jacobian(f, X[:, j]) for j in range(4)]
    np.testing.assert_allclose(Y, np.stack([c[0] for c in cols], 1), rtol=1e-14)
    np.testing.assert_allclose(L, np.array([float(c[1]) for c in cols]), rtol=1e-13, atol=1e-15)
    X2, L2 = O.with_logabsdet_jacobian(O.inverse(f), Y)
    np.testing.assert_allclose(X2, X, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(L2, -L, rtol=1e-9, atol=1e-11)


def _torchify(f):
    return O._to_torch_params(f)


def test_composed_chain_semantics():
    """ChangesOfVariables / InverseFunctions ComposedFunction rules (un-vendored):
    inner first, ladjs add; inverse reverses and inverts."""
    rng = np.random.default_rng(6)
    D = 4
    a = build(O, ["jo"], D, rng)
    b = build(O, ["ss"], D, rng)
    f = O.compose(b, a)            # b ∘ a
    X = rng.standard_normal((D, 5))
    ya, la = O.with_logabsdet_jacobian(a, X)
    yb, lb = O.with_logabsdet_jacobian(b, ya)
    y, l = O.with_logabsdet_jacobian(f, X)
    np.testing.assert_array_equal(y, yb)
    np.testing.assert_allclose(l, la + lb, rtol=1e-15)
    inv = O.inverse(f)
    assert isinstance(inv.outer, O.JohnsonTrafoInv) and isinstance(inv.inner, O.ScaleShiftTrafo)


def test_zygote_primal_quirk_and_gradient():
    """rrule(similar_fill) returns zeros as primal (src/abstract_trafo.jl:30-33):
    under Zygote the ScaleShift ladj value is dropped, its gradient is kept."""
    rng = np.random.default_rng(7)
    f = build(O, ["jo", "ss"], 3, rng)
    X = rng.standard_normal((3, 50))
    true = float(O.mvnormal_negll_trafo(f, X))
    vz, g = O.mvnormal_negll_trafograd(f, X)
    vt, g2 = O.mvnormal_negll_trafograd(f, X, zygote_primal=False)
    ss = O.flatten(f)[1]
    assert vt == pytest.approx(true, rel=1e-14)
    assert vz == pytest.approx(true + np.log(np.abs(ss.a)).sum(), rel=1e-13)
    for (_, a), (_, b) in zip(flat_grads(g, f), flat_grads(g2, f)):
        np.testing.assert_array_equal(a, b)
    # finite-difference check of one gradient entry
    eps = 1e-6
    ss2 = O.ScaleShiftTrafo(ss.a.copy(), ss.b.copy())
    ss2.a[1] += eps
    f2 = O.Composed(ss2, f.inner)
    fd = (float(O.mvnormal_negll_trafo(f2, X)) - true) / eps
    assert g["outer"]["a"][1] == pytest.approx(fd, rel=1e-4)


def test_batching_and_optimizer_restatement():
    assert O.batch_ranges(10, 3) == [(0, 3), (3, 6), (6, 9), (9, 10)]     # round(10/3)=3, extra short batch
    assert O.batch_ranges(10, 4) == [(0, 2), (2, 4), (4, 6), (6, 8), (8, 10)]   # round(2.5)=2 (ties to even)
    assert len(O.batch_ranges(100000, 1000)) == 1000
    rng = np.random.default_rng(8)
    Xw = rng.standard_normal((2, 2000))
    f_true = O.compose(O.ScaleShiftTrafo(np.array([1.3, 0.4]), np.array([2.5, -1.2])),
                       O.HouseholderTrafo(np.array([1.0, 0.3])))
    X = O.apply(f_true, Xw)
    init = O.compose(O.inverse(O.HouseholderTrafo(rng.standard_normal(2))), O.ScaleShiftTrafo(np.ones(2), np.zeros(2)))
    r = O.optimize_whitening(X, init, O.ADAGrad(), nbatches=10, nepochs=4)
    h = r["negll_history"]
    assert len(h) == 40 and h[-1] < h[0]
    V = O.flatten(r["result"])[1].V
    assert np.linalg.norm(V) == pytest.approx(1.0, rel=1e-12)             # functor re-normalises (householder_trafo.jl:134-146)
    r2 = O.optimize_whitening(X, r["result"], O.ADAGrad(), nbatches=10, nepochs=1,
                              optstate=r["optimizer_state"], negll_history=h)
    assert len(r2["negll_history"]) == 50


def test_c_restatement_matches_numpy_oracle():
    """oracle/libenf_ref_cpu.so (the timed CPU baseline) == the numpy oracle."""
    so = os.path.join(ROOT, "oracle", "libenf_ref_cpu.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(so)
    D, N = 6, 1000
    for dt, suf, tol in ((np.float64, "f64", 1e-12), (np.float32, "f32", 2e-4)):
        f = build(O, ["cc", "hh3", "jo", "ss", "cs", "ji"], D, np.random.default_rng(9), dt)
        leaves = O.flatten(f)
        X = (np.random.default_rng(10).standard_normal((N, D)) * 1.2).astype(dt)     # (N, D) C-order == D x N column-major
        kinds = (C.c_int * 6)(1, 5, 2, 4, 0, 3)
        Ks = (C.c_int * 6)(0, 3, 0, 0, 0, 0)
        ps = []
        for lf in leaves:
            if isinstance(lf, O.HouseholderTrafo):
                ps.append(np.ascontiguousarray(np.asarray(lf.V, dtype=dt).T).ravel())
            else:
                ps.append(np.concatenate([np.asarray(getattr(lf, n), dtype=dt) for n in lf.fields]))
        PT = C.POINTER(C.c_float if dt == np.float32 else C.c_double)
        parr = (PT * 6)(*[p.ctypes.data_as(PT) for p in ps])
        Y = np.empty_like(X)
        L = np.empty(N, dtype=dt)
        rc = getattr(lib, "ref_forward_ladj_" + suf)(D, C.c_int64(N), 6, kinds, Ks, parr, X.ctypes.data_as(PT),
                                                     Y.ctypes.data_as(PT), L.ctypes.data_as(PT))
        assert rc == 0
        y_ref, l_ref = O.with_logabsdet_jacobian(f, X.T.astype(np.float64))
        assert np.max(np.abs(Y.T - y_ref) / (np.abs(y_ref) + 1)) < tol
        assert np.max(np.abs(L - l_ref) / (np.abs(l_ref) + 1)) < tol
        out = C.c_double()
        rc = getattr(lib, "ref_negll_" + suf)(D, C.c_int64(N), 6, kinds, Ks, parr, X.ctypes.data_as(PT), C.byref(out))
        assert rc == 0
        assert out.value == pytest.approx(float(O.mvnormal_negll_trafo(f, X.T.astype(np.float64))), rel=tol * 10)


def test_golden_fixtures_reproduce():
    """tests/golden/*.npz pin the oracle: regenerate with tests/golden/make_golden.py."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name, (chain, D, N, seed) in mg.CASES.items():
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        f = build(O, chain, D, np.random.default_rng(seed))
        Y, ladj = O.with_logabsdet_jacobian(f, z["X"])
        np.testing.assert_allclose(Y, z["Y"], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(ladj, z["ladj"], rtol=1e-12, atol=1e-13)
        assert float(O.mvnormal_negll_trafo(f, z["X"])) == pytest.approx(float(z["negll"]), rel=1e-13)
