"""The algebra the CUDA kernels use (tests/device_model.py: overflow-free forms,
hoisted constants, raw-sum gradient accumulators + host finish) against the
literal oracle, on the CPU in float64."""
import numpy as np
import pytest

import device_model as M
from oracle import enf_oracle as O


def _params(rng, D):
    P = dict(a=rng.uniform(.5, 3, D), b=rng.uniform(.5, 1.5, D), c=rng.uniform(-1, 1, D))
    J = dict(gamma=rng.uniform(-1, 1, D), delta=rng.uniform(1, 3, D), xi=rng.uniform(-1, 1, D), lam=rng.uniform(.5, 2, D))
    return P, J


def _ops(rng, D):
    P1, J1 = _params(rng, D)
    P2, J2 = _params(rng, D)
    return [("cc", O.CenterContract(**P1)), ("jo", O.JohnsonTrafo(**J1)), ("hh", O.HouseholderTrafo(rng.standard_normal((D, 3)))),
            ("ss", O.ScaleShiftTrafo(rng.uniform(.5, 2, D) * rng.choice([-1, 1], D), rng.standard_normal(D))),
            ("cs", O.CenterStretch(**P2)), ("ji", O.JohnsonTrafoInv(**J2)), ("hh", O.HouseholderTrafo(rng.standard_normal(D))),
            ("ss", O.ScaleShiftTrafo(rng.uniform(.5, 2, D), rng.standard_normal(D)))]


def _fwd(kind, o, x):
    if kind == "cc": return M.cc_fwd(x, o.a, o.b, o.c)
    if kind == "cs": return M.cs_fwd(x, o.a, o.b, o.c)
    if kind == "jo": return M.jo_fwd(x, o.gamma, o.delta, o.xi, o.lam)
    if kind == "ji": return M.ji_fwd(x, o.gamma, o.delta, o.xi, o.lam)
    if kind == "ss": return M.ss_fwd(x, o.a, o.b)
    return M.hh_fwd(x, o.V)


def _flat(g, f):
    if isinstance(f, O.Composed):
        return _flat(g["inner"], f.inner) + _flat(g["outer"], f.outer)
    return [g]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_forward_and_gradient_algebra(seed):
    rng = np.random.default_rng(seed)
    D, N = 4, 60
    ops = _ops(rng, D)
    chain = O.compose(*[o for _, o in reversed(ops)])
    X = rng.standard_normal((D, N)) * 1.5
    y_ref, l_ref = O.with_logabsdet_jacobian(chain, X)
    xs, l = [X], np.zeros(N)
    for k, o in ops:
        y, lk = _fwd(k, o, xs[-1])
        xs.append(y)
        l = l + lk
    np.testing.assert_allclose(xs[-1], y_ref, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(l, l_ref, rtol=1e-11, atol=1e-12)

    _, gref = O.mvnormal_negll_trafograd(chain, X, zygote_primal=False)
    gl = _flat(gref, chain)
    G = xs[-1].copy()
    for i in reversed(range(len(ops))):
        k, o = ops[i]
        xin, xout = xs[i], xs[i + 1]
        if k == "cc": G, raw = M.cc_bwd(xin, xout, G, o.a, o.b, o.c); g = M.cc_finish(raw, N, o.a, o.b, o.c)
        elif k == "cs": G, raw = M.cs_bwd(xin, xout, G, o.a, o.b, o.c); g = M.cs_finish(raw, N, o.a, o.b, o.c)
        elif k == "jo": G, raw = M.jo_bwd(xin, xout, G, o.gamma, o.delta, o.xi, o.lam); g = M.jo_finish(raw, N, o.gamma, o.delta, o.xi, o.lam)
        elif k == "ji": G, raw = M.ji_bwd(xin, xout, G, o.gamma, o.delta, o.xi, o.lam); g = M.ji_finish(raw, N, o.gamma, o.delta, o.xi, o.lam)
        elif k == "ss": G, raw = M.ss_bwd(xin, G, o.a, o.b); g = M.ss_finish(raw, N, o.a, o.b)
        else:
            G, raw, z = M.hh_bwd(xout, G, o.V)
            g = M.hh_finish(raw, N, o.V)
            np.testing.assert_allclose(z, xin, rtol=1e-10, atol=1e-12)      # `@assert z ≈ x` (householder_trafo.jl:101)
        for name, val in g.items():
            ref = np.asarray(gl[i][name]).reshape(np.shape(val))
            scale = np.abs(ref).max() + 1e-9
            assert np.abs(val / N - ref).max() / scale < 1e-10, (i, k, name)


def test_center_stretch_form_has_no_overflow():
    """Where the literal Float32 formula overflows (x=20, a=7, b=2),
    the device form stays finite and equals the float64 truth."""
    x = np.array([[20.0]])
    a, b, c = np.array([7.0]), np.array([2.0]), np.array([4.0])
    with np.errstate(over="ignore", invalid="ignore"):
        lit32 = O.center_stretch(np.float32(20), np.float32(7), np.float32(2), np.float32(4))
    assert not np.isfinite(lit32)
    y, _ = M.cs_fwd(x.astype(np.float32), a.astype(np.float32), b.astype(np.float32), c.astype(np.float32))
    assert np.isfinite(y).all()
    np.testing.assert_allclose(y[0, 0], float(O.center_stretch(20.0, 7.0, 2.0, 4.0)), rtol=1e-6)


@pytest.mark.parametrize("seed", [0, 1])
def test_affine_chain_gradient_from_second_moments(seed):
    """Householder/ScaleShift chains: (negll, grads) from [[S, m], [m^T, N]] == oracle autograd on the batch."""
    rng = np.random.default_rng(seed)
    D, N = 6, 57
    V1, V2 = rng.normal(size=(D, 3)), rng.normal(size=(D, 2))
    a1, b1 = rng.uniform(0.5, 2, D) * rng.choice([-1.0, 1.0], D), rng.uniform(-1, 1, D)
    a2, b2 = rng.uniform(0.5, 2, D), rng.uniform(-1, 1, D)
    x = rng.normal(size=(D, N)) * 1.7 + rng.uniform(-1, 1, (D, 1))
    f = O.compose(O.ScaleShiftTrafo(a2, b2), O.HouseholderTrafo(V2), O.ScaleShiftTrafo(a1, b1), O.HouseholderTrafo(V1))
    ops = [("hh", V1), ("ss", a1, b1), ("hh", V2), ("ss", a2, b2)]
    xh = np.vstack([x, np.ones((1, N))])
    Shat = xh @ xh.T
    for zp in (False, True):
        ref_l, ref_g = O.mvnormal_negll_trafograd(f, x, zygote_primal=zp)
        got_l, got_g = M.affine_moments_finish(ops, Shat, N, zygote_primal=zp)
        assert abs(got_l - ref_l) < 1e-12 * max(1.0, abs(ref_l))
        ref_flat = _flat(ref_g, f)
        assert len(ref_flat) == len(got_g)
        for r, g in zip(ref_flat, got_g):
            assert set(r) == set(g)
            for k in r:
                np.testing.assert_allclose(np.asarray(g[k]).reshape(np.asarray(r[k]).shape), r[k], rtol=1e-10, atol=1e-12)
