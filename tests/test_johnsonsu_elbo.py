"""JohnsonSU distribution operations (src/johnson_trafo.jl:1-26,109-129) and the variational objective
of examples/nf_variational_1d.jl:25-69.  CPU tests pin the oracle restatement (the reference's own statistical test,
closed-form moments against quadrature, cdf/quantile round trips, ELBO gradient against finite differences); GPU tests
compare the CUDA path with the oracle."""
import numpy as np
import pytest

from conftest import assert_close, rel_err
from oracle import enf_oracle as O


def example_chain(ns, dtype=np.float64):
    """initial_trafo of examples/nf_variational_1d.jl:72-79"""
    one = np.ones(1, dtype=dtype)
    fwd = ns.compose(ns.JohnsonTrafo(0 * one, 5 * one, 0 * one, 5 * one), ns.inverse(ns.CenterStretch(0 * one, one.copy(), 0 * one)),
                     ns.JohnsonTrafo(0 * one, 5 * one, 0 * one, 5 * one), ns.inverse(ns.CenterStretch(0 * one, one.copy(), 0 * one)))
    return ns.inverse(fwd)


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_johnsonsu_matches_the_reference_statistical_test():
    """test/test_johnson_trafo.jl:12-16: first absolute moment of rand(JohnsonSU) vs johnsontrafo_inv.(randn), rtol 1e-2, 1e6 draws."""
    rng = np.random.default_rng(0)
    d = O.JohnsonSU()
    a = np.abs(d.rand(rng, 10 ** 6)).mean()
    b = np.abs(O.johnsontrafo_inv(rng.standard_normal(10 ** 6), d.gamma, d.delta, d.xi, d.lam)).mean()
    assert abs(a - b) <= 1e-2 * max(a, b)


@pytest.mark.parametrize("p", [(10.0, 3.5, 10.0, 1.0), (-0.7, 1.3, 0.4, 2.0), (0.0, 2.0, -1.0, 0.5)])
def test_oracle_johnsonsu_closed_forms(p):
    d = O.JohnsonSU(*p)
    lo, hi = d.quantile(1e-13), d.quantile(1 - 1e-13)
    x = np.linspace(lo, hi, 2_000_001)
    pdf = d.pdf(x)
    assert abs(np.trapezoid(pdf, x) - 1) < 1e-8
    assert abs(np.trapezoid(x * pdf, x) - d.mean()) < 1e-6 * (abs(d.mean()) + 1)                  # src/johnson_trafo.jl:24
    assert abs(np.trapezoid((x - d.mean()) ** 2 * pdf, x) - d.var()) < 1e-5 * d.var()             # :26
    assert abs(d.cdf(d.median()) - 0.5) < 1e-14                                                   # :25
    q = np.array([1e-6, 0.01, 0.3, 0.5, 0.9, 1 - 1e-6])
    np.testing.assert_allclose(d.cdf(d.quantile(q)), q, rtol=1e-10)                               # :121,129
    xs = d.quantile(np.linspace(0.001, 0.999, 101))
    np.testing.assert_allclose(np.log(d.pdf(xs)), d.logpdf(xs), rtol=1e-12, atol=1e-12)           # :123
    np.testing.assert_allclose(np.log(d.cdf(xs)), d.logcdf(xs), rtol=1e-10)                       # :124
    np.testing.assert_allclose(d.ccdf(xs) + d.cdf(xs), 1.0, rtol=0, atol=1e-15)                   # :125
    body = pdf > 1e-3 * pdf.max()                                                                 # (differences of a cdf near 1 lose digits)
    np.testing.assert_allclose(np.gradient(d.cdf(x), x)[body][::5000], pdf[body][::5000], rtol=1e-5)   # pdf = cdf'


def test_oracle_elbo_gradient_against_finite_differences():
    """nELBO_trafograd (torch autograd in Zygote's role) == central differences of the literal nELBO."""
    rng = np.random.default_rng(1)
    f = example_chain(O)
    b = rng.standard_normal((50, 1))
    xi = np.vstack([b, -b])
    v, g = O.nELBO_trafograd(f, xi)
    assert abs(v - float(O.nELBO(f, xi))) < 1e-14
    leaves = O.flatten(f)
    from chains import flat_grads
    for (name, ga), lf in zip(flat_grads(g, f), [l for l in leaves for _ in l.fields]):
        field = name.split(".")[1]
        h = 1e-6
        old = getattr(lf, field).copy()
        setattr(lf, field, old + h); vp = float(O.nELBO(f, xi))
        setattr(lf, field, old - h); vm = float(O.nELBO(f, xi))
        setattr(lf, field, old)
        assert abs((vp - vm) / (2 * h) - ga[0]) < 1e-6 * (abs(ga[0]) + 1), name


# ------------------------------------------------------------------ CUDA path (GPU)
OPS = ["pdf", "logpdf", "cdf", "logcdf", "ccdf", "logccdf", "quantile"]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("p", [(10.0, 3.5, 10.0, 1.0), (-0.7, 1.3, 0.4, 2.0)])
def test_johnsonsu_operations_match_oracle(ctx, dtype, p):
    import enf_b200 as E
    d, do = E.JohnsonSU(*p), O.JohnsonSU(*p)
    rng = np.random.default_rng(3)
    N = 100_003                                                            # ragged: vector body + scalar tail
    x = do.quantile(rng.uniform(1e-4, 1 - 1e-4, N)).astype(dtype)         # points where the distribution lives
    pr = rng.uniform(1e-4, 1 - 1e-4, N).astype(dtype)
    Xd, Pd = E.B200Matrix.from_host(x[None, :], ctx), E.B200Matrix.from_host(pr[None, :], ctx)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    for op in OPS:
        arg, argd = (pr, Pd) if op == "quantile" else (x, Xd)
        got = getattr(d, op)(argd).to_host()[0]
        ref = getattr(do, op)(arg.astype(np.float64))
        assert got.dtype == dtype
        if op in ("logccdf", "ccdf", "quantile"):
            # 1 - cdf (the reference's literal form, src/johnson_trafo.jl:125-126) and the normal quantile are conditioned
            # like 1/(1 - cdf) and 1/pdf: measured on the part of the range where that factor is below 100
            ok = (do.cdf(x.astype(np.float64)) < 0.99) if op != "quantile" else (np.abs(pr - 0.5) < 0.49)
            assert rel_err(got[ok], ref[ok]) <= tol * 100, (op, rel_err(got[ok], ref[ok]))
        else:
            assert rel_err(got, ref) <= tol * (10 if dtype == np.float64 else 1), (op, rel_err(got, ref))
    # host arrays and scalars go through the same kernel; unaligned device views take the scalar path
    np.testing.assert_allclose(d.cdf(x[:7]), do.cdf(x[:7].astype(np.float64)), rtol=1e-5 if dtype == np.float32 else 1e-12)
    assert abs(float(d.pdf(dtype(do.median()))) - do.pdf(do.median())) <= 1e-5 * do.pdf(do.median())
    v = d.logpdf(Xd.cols(1, 1001)).to_host()[0]
    assert rel_err(v, do.logpdf(x[1:1001].astype(np.float64))) <= tol * 10
    # the reference's statistical test through the CUDA path (test/test_johnson_trafo.jl:12-16)
    r = np.random.default_rng(0)
    a = np.abs(np.asarray(d.rand(r, 10 ** 6, dtype=dtype, ctx=ctx), dtype=np.float64)).mean()
    b = np.abs(O.johnsontrafo_inv(r.standard_normal(10 ** 6), *p)).mean()
    assert abs(a - b) <= 1e-2 * max(a, b)
    assert abs(d.mean() - do.mean()) < 1e-14 and abs(d.var() - do.var()) < 1e-12 * do.var() and d.median() == do.median()
    with pytest.raises(E.EnfError):
        E.JohnsonSU(0, 0, 0, 1).pdf(Xd)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_elbo_value_and_gradient_match_oracle(ctx, dtype):
    """examples/nf_variational_1d.jl:29-47 on its own chain and batch shape: product API on a 1 x (2 batchsize) matrix ==
    the literal example on the (2 batchsize) x 1 matrix; plus a D = 3 chain with a ScaleShift (Zygote-primal quirk) and a
    two-component target with widths."""
    import enf_b200 as E
    from chains import both, flat_grads
    tol = 1e-5 if dtype == np.float32 else 1e-12
    rng = np.random.default_rng(5)
    fe, fo = example_chain(E, dtype), example_chain(O, dtype)
    b = rng.standard_normal(100).astype(dtype)
    xi = np.concatenate([b, -b])
    v_ref, g_ref = O.nELBO_trafograd(fo, xi.astype(np.float64)[:, None])
    Xd = E.B200Matrix.from_host(xi[None, :], ctx)
    v = E.nELBO(fe, Xd)
    v2, g = E.nELBO_trafograd(fe, Xd)
    assert abs(v - v_ref) <= tol * (abs(v_ref) + 1) and abs(v2 - v_ref) <= tol * (abs(v_ref) + 1)
    g_rms = float(np.sqrt(np.mean(np.concatenate([r.ravel() for _, r in flat_grads(g_ref, fo)]) ** 2)))
    for (k, a), (_, r) in zip(flat_grads(g, fe), flat_grads(g_ref, fo)):
        # antithetic pairs +-xi through a chain that starts symmetric (a = c = 0, gamma = xi = 0) make some gradients
        # mathematically zero: those are measured against the size of the whole gradient, the others against themselves
        zero = float(np.abs(r).max()) < 1e-10 * g_rms
        assert_close(a, r.reshape(a.shape), dtype, f"elbo grad {k}", floor=g_rms if zero else 0.0)
    # D = 3 (samples are columns), ragged N, target with widths, Zygote-primal loss value
    fo3, fe3 = both(["cc", "jo", "hh2", "ss"], 3, 9, dtype)
    X3 = rng.standard_normal((3, 1001)).astype(dtype)
    tgt_o = O.GaussMixture((0.6, 0.4), (-1.0, 2.0), (0.7, 1.5))
    tgt_e = E.GaussMixture((0.6, 0.4), (-1.0, 2.0), (0.7, 1.5))
    import torch
    ft = O._to_torch_params(fo3)
    Xt = torch.tensor(X3.astype(np.float64))
    z, ladj = O.with_logabsdet_jacobian(ft, Xt)
    val = -((tgt_o.logpdf(z).sum() + ladj.sum()) / X3.shape[1] - 0.5 * (O.LOG2PI + 1) * 3)     # samples are columns: N = 1001, D = 3
    val.backward()
    g3_ref = O._grads_of(ft)
    v3, g3 = E.nELBO_trafograd(fe3, E.B200Matrix.from_host(X3, ctx), tgt_e, zygote_primal=False)
    assert abs(v3 - float(val)) <= tol * (abs(float(val)) + 1)
    for (k, a), (_, r) in zip(flat_grads(g3, fe3), flat_grads(g3_ref, fo3)):
        assert_close(a, r.reshape(a.shape), dtype, f"elbo grad D=3 {k}")
    ss_a = np.asarray(O.flatten(fo3)[-1].a, dtype=np.float64)
    v3z, _ = E.nELBO_trafograd(fe3, E.B200Matrix.from_host(X3, ctx), tgt_e)                  # default: Zygote's loss value
    assert abs((v3z - v3) - np.log(np.abs(ss_a)).sum()) <= tol * 10
    with pytest.raises(E.EnfError):
        E.nELBO(fe3, E.B200Matrix.from_host(X3, ctx), E.GaussMixture((1.0,), (0.0,), (-1.0,)))


@pytest.mark.gpu
def test_optimise_elbo_matches_oracle_loop(ctx):
    """optimise_ELBO (examples/nf_variational_1d.jl:49-69): same nELBO history as the oracle's loop on the same draws."""
    import enf_b200 as E
    rng = np.random.default_rng(7)
    batches = [rng.standard_normal(100) for _ in range(25)]
    r_ref = O.optimise_ELBO(example_chain(O), O.ADAGrad(), batches)
    r = E.optimise_ELBO(example_chain(E), E.ADAGrad(), batches=batches, ctx=ctx)
    h, h_ref = np.array(r["nelbo_history"]), np.array(r_ref["nelbo_history"])
    assert h.shape == h_ref.shape == (25,)
    assert np.max(np.abs(h - h_ref) / (np.abs(h_ref) + 1)) < 1e-9
    assert h[-5:].mean() < h[:5].mean()                                   # the bound improves
    for a, b in zip(E.flatten(r["result"]), O.flatten(r_ref["result"])):
        for n in a.fields:
            assert np.max(np.abs(np.asarray(getattr(a, n)) - np.asarray(getattr(b, n)))) < 1e-8
