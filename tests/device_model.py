"""numpy model of the DEVICE algorithm (test infrastructure).

The CUDA kernels do not evaluate the reference's formulas literally: they use
algebraically equivalent overflow-free forms, hoist per-row constants to the
host, and accumulate *raw sums* on the device that the host maps to parameter
gradients.  This file states that algorithm in numpy so the algebra can be
checked against the literal oracle (oracle/enf_oracle.py) on the CPU, before
any GPU time is spent.  csrc/enf_math.cuh and csrc/enf_abi.cu transcribe it.

Conventions: G = N * dL/d(output) (un-normalised cotangent), lb = N * dL/dladj
= -1.  All functions take one op's per-row parameters as length-D vectors and
x as D x N.
"""
import numpy as np

LB = -1.0
LOG2E = 1.4426950408889634


# ---------------------------------------------------------------- forward
def cs_fwd(x, a, b, c):
    """CenterStretch forward + ladj (stable form of src/center_stretch.jl:4-8,39-43)."""
    a, b, c = a[:, None], b[:, None], c[:, None]
    A = np.exp(b * a)
    ax = np.abs(x)
    w = np.exp2(-b * LOG2E * ax)                  # e^{-b|x|}
    m = 0.5 * A - 0.5 * A * w                     # (1 - w) A / 2
    g = np.sqrt(m * m + w) + m                    # e^{b(|u| - |x|)}: root of the reference's quadratic / e^{b|x|}
    au = ax + np.log(g) / b
    y = np.copysign(au, x) + c
    # S(u) = nd / nn with numerator and denominator scaled by g^2 A (no division by g)
    wg = w * g
    p2 = g * g + w * w
    nd = p2 + (2 / A) * wg
    nn = p2 + ((1 + A * A) / A) * wg
    # one log per lane-vector: log(prod nn / prod nd); here per element
    ladj = np.log(nn / nd)                        # = -log S(u)
    return y, ladj.sum(0)


def cc_fwd(x, a, b, c):
    """CenterContract forward + ladj (src/center_stretch.jl:11-22,63-67)."""
    a, b, c = a[:, None], b[:, None], c[:, None]
    A = np.exp(b * a)
    u = x - c
    au = np.abs(u)
    w = np.exp2(-b * LOG2E * au)
    n1 = 1 + A * w
    n2 = A + w
    r12 = 1 / (n1 * n2)
    y = np.copysign(au + np.log(n1 * n1 * r12) / b, u)
    ladj = np.log((n2 + w * n1) * r12)
    return y, ladj.sum(0)


def jo_fwd(x, gamma, delta, xi, lam):
    """JohnsonTrafo forward + ladj (src/johnson_trafo.jl:29-32,39-42,76-80)."""
    gamma, delta, xi, lam = gamma[:, None], delta[:, None], xi[:, None], lam[:, None]
    il = 1 / lam
    z = x * il - xi * il
    s = 1 + z * z
    r = 1 / np.sqrt(s)
    ash = np.copysign(np.log(np.abs(z) + s * r), z)
    y = gamma + delta * ash
    ladj = np.log(np.abs(delta * il)) + np.log(r)
    return y, ladj.sum(0)


def ji_fwd(x, gamma, delta, xi, lam):
    """JohnsonTrafoInv forward + ladj (src/johnson_trafo.jl:34-37,101-105)."""
    gamma, delta, xi, lam = gamma[:, None], delta[:, None], xi[:, None], lam[:, None]
    idl = 1 / delta
    s = x * idl - gamma * idl
    e = np.exp2(LOG2E * s)
    ei = 1 / e
    sh = 0.5 * (e - ei)
    ch = 0.5 * (e + ei)
    y = lam * sh + xi
    ladj = np.log(np.abs(lam * idl)) + np.log(ch)
    return y, ladj.sum(0)


def ss_fwd(x, a, b):
    return x * a[:, None] + b[:, None], np.full(x.shape[1], np.log(np.abs(a)).sum())


def hh_fwd(x, V):
    """Householder with pre-scaled v' = v*sqrt(2/v'v): y = x - (v'.x) v'."""
    if V.ndim == 1:
        V = V[:, None]
    for k in range(V.shape[1]):
        v = V[:, k]
        vp = v * np.sqrt(2.0 / (v @ v))
        x = x - vp[:, None] * (vp @ x)[None, :]
    return x, np.zeros(x.shape[1])


# ---------------------------------------------------------------- backward
# each returns (Gx, raw) ; finish_*() maps raw sums -> parameter gradients
def _cc_parts(w, A, b):
    """sigma_1 = 1/n1, sigma_2 = w/n2 from ONE reciprocal R = 1/(n1 n2 n3):  S = sigma_1 + sigma_2, 1/S,
    ds = sigma_2 - sigma_1, and E_i = b sigma_i (1 - sigma_i) / S = (b A w R) n_(3-i)^2 (no 1 - sigma cancellation):
    nEd = -(E1 - E2), Es = E1 + E2."""
    aw = A * w
    n1 = aw + 1
    n2 = A + w
    n3 = n2 + w * n1
    p12 = n1 * n2
    R = 1 / (p12 * n3)
    t = n3 * R
    hb = (aw * b) * R
    q1, q2 = n1 * n1, n2 * n2
    return n3 * t, p12 * (p12 * R), (w * n1 - n2) * t, hb * (q1 - q2), hb * (q1 + q2)


def cc_bwd(x, y, G, a, b, c):
    """x: the op's input, y: its output (|y| replaces the recomputed forward value).  The b integrand is
    accumulated times b (cc_finish divides)."""
    a, b, c = a[:, None], b[:, None], c[:, None]
    A = np.exp(b * a)
    u = x - c
    au = np.abs(u)
    sg = np.where(np.signbit(u), -1.0, 1.0).astype(u.dtype)
    S, iS, ds, nEd, Es = _cc_parts(np.exp2(-b * LOG2E * au), A, b)
    sgG = sg * G
    Gx = G * S + sg * nEd                      # LB = -1
    ya = au * S + a * ds - np.abs(y)
    ra = sgG * ds + Es
    rb = sgG * ya + (au * nEd + a * Es)
    return Gx, (Gx.sum(1), ra.sum(1), rb.sum(1))


def cc_finish(raw, N, a, b, c):
    R1, R2, R3 = raw
    return {"a": R2, "b": R3 / b, "c": -R1}


def cs_bwd(x, y, G, a, b, c):
    """Implicit differentiation of the inverse of CenterContract at u = y - c.  Raw sums: dc, -da, -b db."""
    a, b, c = a[:, None], b[:, None], c[:, None]
    A = np.exp(b * a)
    au = np.abs(y - c)
    sg = np.where(np.signbit(x), -1.0, 1.0).astype(x.dtype)
    S, iS, ds, nEd, Es = _cc_parts(np.exp2(-b * LOG2E * au), A, b)
    Gy = G - sg * nEd                          # total cotangent on y
    Gx = Gy * iS
    sgGx = sg * Gx
    cb = au * S + a * ds - np.abs(x)
    ra = sgGx * ds + Es
    rb = sgGx * cb + (au * nEd + a * Es)
    return Gx, (G.sum(1), ra.sum(1), rb.sum(1))


def cs_finish(raw, N, a, b, c):
    R1, R2, R3 = raw
    return {"a": -R2, "b": -R3 / b, "c": R1}


def jo_bwd(x, y, G, gamma, delta, xi, lam):
    gamma, delta, xi, lam = gamma[:, None], delta[:, None], xi[:, None], lam[:, None]
    il = 1 / lam
    z = x * il - xi * il
    r = 1 / np.sqrt(1 + z * z)
    gz = G * delta * r - LB * z * r * r
    return gz * il, (G.sum(1), (G * y).sum(1), gz.sum(1), (z * gz).sum(1))


def jo_finish(raw, N, gamma, delta, xi, lam):
    S1, S2, S3, S4 = raw
    return {"gamma": S1, "delta": (S2 - gamma * S1) / delta + LB * N / delta, "xi": -S3 / lam, "lam": -(S4 + LB * N) / lam}


def ji_bwd(x, y, G, gamma, delta, xi, lam):
    gamma, delta, xi, lam = gamma[:, None], delta[:, None], xi[:, None], lam[:, None]
    idl = 1 / delta
    s = x * idl - gamma * idl
    sh = (y - xi) / lam
    q = 1 + sh * sh
    rq = 1 / np.sqrt(q)
    gs = G * lam * q * rq + LB * sh * rq
    return gs * idl, (gs.sum(1), (s * gs).sum(1), G.sum(1), (G * y).sum(1))


def ji_finish(raw, N, gamma, delta, xi, lam):
    S1, S2, S3, S4 = raw
    return {"gamma": -S1 / delta, "delta": -(S2 + LB * N) / delta, "xi": S3, "lam": (S4 - xi * S3) / lam + LB * N / lam}


def ss_bwd(x, G, a, b):
    return G * a[:, None], ((G * x).sum(1), G.sum(1))


def ss_finish(raw, N, a, b):
    R0, R1 = raw
    return {"a": R0 + LB * N / a, "b": R1}


def hh_bwd(y, G, V):
    """y: the op's OUTPUT (the reverse sweep recovers every reflection's input
    by re-applying it, src/householder_trafo.jl:88-103)."""
    vec = V.ndim == 1
    if vec:
        V = V[:, None]
    K = V.shape[1]
    acc1 = np.zeros_like(V, dtype=np.float64)
    acc2 = np.zeros(K)
    z, Dl = y, G
    for k in reversed(range(K)):
        v = V[:, k]
        vp = v * np.sqrt(2.0 / (v @ v))
        po = vp @ z                      # v'.z_out = -(v'.z_in)
        q = vp @ Dl
        z = z - vp[:, None] * po[None, :]
        p = -po
        acc1[:, k] = (p[None, :] * Dl + q[None, :] * z).sum(1)
        acc2[k] = (p * q).sum()
        Dl = Dl - vp[:, None] * q[None, :]
    return Dl, (acc1, acc2), z


def hh_finish(raw, N, V):
    acc1, acc2 = raw
    vec = V.ndim == 1
    Vm = V[:, None] if vec else V
    n = (Vm * Vm).sum(0)
    # acc1 was accumulated with z = reflection *input*; p' uses the input too.
    dV = -np.sqrt(2.0 / n)[None, :] * acc1 + (2.0 / n * acc2)[None, :] * Vm
    return {"V": dV[:, 0] if vec else dV}


# ---------------------------------------------------------------- affine chains from second moments
def affine_moments_finish(ops, Shat, N, zygote_primal=False):
    """Loss and gradients of a Householder/ScaleShift-only chain from the batch's second moments
    Shat = [[S, m], [m^T, N]] (csrc/enf_abi.cu: finish_moments transcribes this).

    ops: list of ("ss", a, b) / ("hh", V) in application order (V: D x K, columns applied in order).
    Returns (negll, [grad dicts in application order])."""
    D = Shat.shape[0] - 1
    B = np.hstack([np.eye(D), np.zeros((D, 1))])          # x_i = B_i [x; 1]
    Z = Shat[:D, :] / N                                    # becomes B_n Shat / N
    w = Shat[D, :] / N                                     # homogeneous weight of every column of Shat / N
    lconst = 0.0

    def reflect(M, v, s):
        t = v @ M
        return M - s * np.outer(v, t), t

    for op in ops:
        if op[0] == "ss":
            a, b = op[1], op[2]
            B = a[:, None] * B
            B[:, D] += b
            Z = a[:, None] * Z + np.outer(b, w)
            lconst += np.log(np.abs(a)).sum()
        else:
            V = op[1]
            for r in range(V.shape[1]):
                v = V[:, r]
                s = 2.0 / (v @ v)
                B, _ = reflect(B, v, s)
                Z, _ = reflect(Z, v, s)
    sum_y = 0.5 * N * (Z * B).sum()
    negll = (sum_y + 0.5 * np.log(2 * np.pi) * N * D - N * (0.0 if zygote_primal else lconst)) / N
    grads = [None] * len(ops)
    for i in reversed(range(len(ops))):
        op = ops[i]
        if op[0] == "ss":
            a, b = op[1], op[2]
            gb = Z[:, D].copy()
            B = B.copy()
            B[:, D] -= b
            B = B / a[:, None]
            ga = (Z * B).sum(1) - 1.0 / a
            Z = a[:, None] * Z
            grads[i] = {"a": ga, "b": gb}
        else:
            V = op[1]
            gV = np.zeros_like(V)
            for r in reversed(range(V.shape[1])):
                v = V[:, r]
                n = v @ v
                s = 2.0 / n
                B, tB = reflect(B, v, s)          # B is now the reflection's input; v^T B_in = -tB
                tZ = v @ Z
                u = -tB
                gV[:, r] = -s * (Z @ u + B @ tZ) + (4.0 / n ** 2) * (tZ @ u) * v
                Z = Z - s * np.outer(v, tZ)
            grads[i] = {"V": gV}
    return negll, grads
