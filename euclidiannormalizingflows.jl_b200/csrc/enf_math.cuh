// Per-element device math of the trafo ops: forward value + ladj term and the
// backward pass (input cotangent + raw-sum integrands).
//
// The reference formulas (cited per function, paths relative to the reference
// repository) are evaluated in algebraically equivalent, overflow-free forms
// whose per-row constants (e^{ba}, 1/b, 1/lambda, ...) are hoisted to the host
// (enf_abi.cu: derive_constants).  tests/device_model.py states the same algebra
// in numpy and tests/test_device_model.py checks it against the literal oracle.
//
// Conventions: G = N * dL/d(output); LB = N * dL/dladj = -1 (the seeds of
// src/optimize_whitening.jl:12,19-20 with the 1/N pulled out).
#pragma once
#include <cuda_runtime.h>

namespace enf {

// ---------------------------------------------------------------- primitives
template <typename T> struct Prim;

template <> struct Prim<float> {
    static constexpr float LN2 = 0.69314718055994531f;
    static constexpr float LOG2E = 1.4426950408889634f;
    static __device__ __forceinline__ float ex2(float x) {
        float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float lg2(float x) {
        float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float ln(float x) { return lg2(x) * LN2; }
    static __device__ __forceinline__ float rcp(float x) {
        float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float rsq(float x) {
        float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float sqrt_(float x) {
        float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float csign(float mag, float sgn) { return copysignf(mag, sgn); }
    // asinh(z) given s = 1 + z^2 and r = rsqrt(s)
    static __device__ __forceinline__ float asinh_(float z, float s, float r) {
        return copysignf(ln(fabsf(z) + s * r), z);
    }
    // ln(1/sqrt(1+z^2)) given s = 1 + z^2, r = rsqrt(s)
    static __device__ __forceinline__ float ln_rsq(float z, float s, float r) { return ln(r); }
    static __device__ __forceinline__ void sinhcosh(float s, float& sh, float& ch) {
        float e = ex2(s * LOG2E);
        float ei = rcp(e);
        sh = 0.5f * (e - ei);
        ch = 0.5f * (e + ei);
        // (e - 1/e)/2 cancels for small |s|: odd Taylor polynomial there
        float s2 = s * s;
        float p = fma_(s2, fma_(s2, fma_(s2, fma_(s2, 2.7557319e-6f, 1.9841270e-4f), 8.3333333e-3f), 0.16666667f), 1.0f);
        sh = (fabsf(s) < 0.4f) ? s * p : sh;
    }
};

template <> struct Prim<double> {
    static constexpr double LN2 = 0.69314718055994530942;
    static constexpr double LOG2E = 1.44269504088896340736;
    static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
    static __device__ __forceinline__ double ln(double x) { return log(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rsq(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double csign(double mag, double sgn) { return copysign(mag, sgn); }
    static __device__ __forceinline__ double asinh_(double z, double s, double r) { return asinh(z); }
    static __device__ __forceinline__ double ln_rsq(double z, double s, double r) { return -0.5 * log1p(z * z); }
    static __device__ __forceinline__ void sinhcosh(double s, double& sh, double& ch) {
        sh = sinh(s);
        ch = cosh(s);
    }
};

// ---------------------------------------------------------------- forward
// Every *_fwd returns y and ADDS the element's ladj term to `l` (row constants
// such as log|delta/lambda| and sum(log|a|) are added once per sample by the
// caller: ChainDesc::ladj_const).

// CenterStretch: src/center_stretch.jl:4-8 (value), :39-43 (ladj = -center_contract_ladj(y)).
// constants: nb2 = -b*log2(e), A = e^{ba}, ib = 1/b, c
template <typename T>
__device__ __forceinline__ T cs_fwd(T x, T nb2, T A, T ib, T c, T& l) {
    using P = Prim<T>;
    T ax = P::abs_(x);
    T w0 = P::ex2(nb2 * ax);                    // e^{-b|x|}
    T m = P::fma_(-A, w0, A);                   // (1 - w0) e^{ba}
    T g = T(0.5) * (P::sqrt_(P::fma_(m, m, T(4) * w0)) + m);   // e^{b(|u|-|x|)}
    T au = P::fma_(P::ln(g), ib, ax);           // |u| = |y - c|
    T wu = w0 * P::rcp(g);                      // e^{-b|u|}
    T n1 = P::fma_(A, wu, T(1));
    T n2 = A + wu;
    T num = P::fma_(wu, n1, n2);                // S = num/(n1 n2)
    l += P::ln(n1 * n2 * P::rcp(num));          // -log S
    return P::csign(au, x) + c;
}

// CenterContract: src/center_stretch.jl:11-15 (value), :17-22,63-67 (ladj).
template <typename T>
__device__ __forceinline__ T cc_fwd(T x, T nb2, T A, T ib, T c, T& l) {
    using P = Prim<T>;
    T u = x - c;
    T au = P::abs_(u);
    T w = P::ex2(nb2 * au);                     // e^{-b|u|}
    T n1 = P::fma_(A, w, T(1));
    T n2 = A + w;
    T L1 = P::ln(n1), L2 = P::ln(n2);
    T L3 = P::ln(P::fma_(w, n1, n2));
    l += L3 - L1 - L2;                          // log S
    return P::csign(P::fma_(L1 - L2, ib, au), u);
}

// JohnsonTrafo: src/johnson_trafo.jl:29-32 (value), :39-42,49-52,76-80 (ladj).
// constants: il = 1/lambda, c0 = -xi/lambda, gamma, delta
template <typename T>
__device__ __forceinline__ T jo_fwd(T x, T il, T c0, T gamma, T delta, T& l) {
    using P = Prim<T>;
    T z = P::fma_(x, il, c0);
    T s = P::fma_(z, z, T(1));
    T r = P::rsq(s);
    l += P::ln_rsq(z, s, r);                      // -log(1+z^2)/2
    return P::fma_(delta, P::asinh_(z, s, r), gamma);
}

// JohnsonTrafoInv: src/johnson_trafo.jl:34-37 (value), :101-105 (ladj = -johnsontrafo_ladj(y)).
// constants: idl = 1/delta, c0 = -gamma/delta, lambda, xi
template <typename T>
__device__ __forceinline__ T ji_fwd(T x, T idl, T c0, T lam, T xi, T& l) {
    using P = Prim<T>;
    T s = P::fma_(x, idl, c0);
    T sh, ch;
    P::sinhcosh(s, sh, ch);
    l += P::ln(ch);                             // log sqrt(1 + sinh^2)
    return P::fma_(lam, sh, xi);
}

// ---------------------------------------------------------------- backward
// Every *_bwd takes the op's INPUT x and the output cotangent G, returns the
// input cotangent and writes the raw-sum integrands r[...] (summed over samples
// on the device, mapped to parameter gradients by enf_abi.cu: finish_grads).

// CenterContract.  raw: r0 -> -dc, r1 -> da, r2 -> db.  extra constants a, b.
template <typename T>
__device__ __forceinline__ T cc_bwd(T x, T G, T nb2, T A, T ib, T c, T a, T b, T* r) {
    using P = Prim<T>;
    T u = x - c;
    T au = P::abs_(u);
    T sg = u < T(0) ? T(-1) : T(1);
    T w = P::ex2(nb2 * au);
    T n1 = P::fma_(A, w, T(1));
    T n2 = A + w;
    T s1 = P::rcp(n1);
    T s2 = w * P::rcp(n2);
    T S = s1 + s2;
    T d1 = s1 * (T(1) - s1);
    T d2 = s2 * (T(1) - s2);
    T ya = P::fma_(P::ln(n1 * P::rcp(n2)), ib, au);
    T Su = b * (d1 - d2);
    T Sa = -b * (d1 + d2);
    T Sb = (au - a) * d1 - (au + a) * d2;
    T ya_a = s2 - s1;
    T ya_b = (s1 * (au - a) + s2 * (au + a) - ya) * ib;
    T iS = P::rcp(S);
    T Gx = G * S - sg * Su * iS;                // LB = -1
    r[0] = Gx;
    r[1] = sg * G * ya_a - Sa * iS;
    r[2] = sg * G * ya_b - Sb * iS;
    return Gx;
}

// CenterStretch (implicit inverse of CenterContract).  raw: r0 -> dc, r1 -> da, r2 -> db.
template <typename T>
__device__ __forceinline__ T cs_bwd(T x, T G, T nb2, T A, T ib, T c, T a, T b, T* r) {
    using P = Prim<T>;
    T ax = P::abs_(x);
    T sg = x < T(0) ? T(-1) : T(1);
    T w0 = P::ex2(nb2 * ax);
    T m = P::fma_(-A, w0, A);
    T g = T(0.5) * (P::sqrt_(P::fma_(m, m, T(4) * w0)) + m);
    T au = P::fma_(P::ln(g), ib, ax);
    T wu = w0 * P::rcp(g);
    T n1 = P::fma_(A, wu, T(1));
    T n2 = A + wu;
    T s1 = P::rcp(n1);
    T s2 = wu * P::rcp(n2);
    T S = s1 + s2;
    T d1 = s1 * (T(1) - s1);
    T d2 = s2 * (T(1) - s2);
    T Su = b * (d1 - d2);
    T Sa = -b * (d1 + d2);
    T Sb = (au - a) * d1 - (au + a) * d2;
    T Ca = s2 - s1;
    T Cb = (s1 * (au - a) + s2 * (au + a) - ax) * ib;
    T iS = P::rcp(S);
    T Gy = G + sg * Su * iS;                    // G - LB*sg*Su/S
    T Gx = Gy * iS;
    r[0] = G;
    r[1] = -Gx * sg * Ca + Sa * iS;
    r[2] = -Gx * sg * Cb + Sb * iS;
    return Gx;
}

// JohnsonTrafo.  raw: r0 = G, r1 = G asinh z, r2 = gz, r3 = z gz.
template <typename T>
__device__ __forceinline__ T jo_bwd(T x, T G, T il, T c0, T gamma, T delta, T* r) {
    using P = Prim<T>;
    T z = P::fma_(x, il, c0);
    T s = P::fma_(z, z, T(1));
    T rr = P::rsq(s);
    T ash = P::asinh_(z, s, rr);
    T gz = P::fma_(G * delta, rr, z * rr * rr); // G delta r - LB z r^2
    r[0] = G;
    r[1] = G * ash;
    r[2] = gz;
    r[3] = z * gz;
    return gz * il;
}

// JohnsonTrafoInv.  raw: r0 = gs, r1 = s gs, r2 = G, r3 = G sinh s.
template <typename T>
__device__ __forceinline__ T ji_bwd(T x, T G, T idl, T c0, T lam, T xi, T* r) {
    using P = Prim<T>;
    T s = P::fma_(x, idl, c0);
    T sh, ch;
    P::sinhcosh(s, sh, ch);
    T gs = P::fma_(G * lam, ch, -sh * P::rcp(ch));  // G lam cosh + LB tanh
    r[0] = gs;
    r[1] = s * gs;
    r[2] = G;
    r[3] = G * sh;
    return gs * idl;
}

}  // namespace enf
