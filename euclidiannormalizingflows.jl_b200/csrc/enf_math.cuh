// Device math of the trafo ops on one 16-byte vector of a sample: forward value +
// ladj contribution, and the backward pass (input cotangent + raw-sum integrands).
//
// The reference formulas (cited per function, paths relative to the reference
// repository) are evaluated in algebraically equivalent, overflow-free forms
// whose per-row constants (e^{ba}, 1/b, 1/lambda, ...) are hoisted to the host
// (enf_abi.cu: derive_constants).  tests/device_model.py states the same algebra
// in numpy and tests/test_device_model.py checks it against the literal oracle.
//
// The kernels are bound by the special-function (MUFU) pipe and by instruction
// issue, not by HBM (profiles/): every op is written to minimise both.
//   * logs of Jacobian factors are not taken per element: the factors of the
//     elements a lane owns are multiplied and ONE log is taken per lane
//     (log_of_product; a range flag makes the kernel redo the rare tile whose product over/underflowed),
//   * ladj is accumulated in the unit of the hardware log (log2 for Float32) and
//     scaled once per sample,
//   * 1/b, ln2/b, e^{ba}/4, (1+e^{2ba})/2 ... come in as constants.
//
// Conventions: G = N * dL/d(output); LB = N * dL/dladj = -1 (the seeds of
// src/optimize_whitening.jl:12,19-20 with the 1/N pulled out).
#pragma once
#include <cuda_runtime.h>

namespace enf {

// ---------------------------------------------------------------- primitives
template <typename T> struct Prim;

// Float32: MUFU approximations (ex2/lg2/rcp/rsqrt/sqrt, ~2^-22 relative), logs in base 2.
template <> struct Prim<float> {
    static constexpr float LGU = 0.69314718055994531f;      // natural log of the log base: ln x = LGU * lg x
    static constexpr float INV_LGU = 1.4426950408889634f;
    static constexpr float LG_OF_2 = 1.0f;                   // lg(2)
    static constexpr float LG_SAFE = 100.0f;                 // |lg(product)| below this: no over/underflow happened
    static __device__ __forceinline__ float ex2(float x) {
        float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
    static __device__ __forceinline__ float lg(float x) {
        float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
    }
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
    static __device__ __forceinline__ float sign_(float x) { return x < 0.f ? -1.f : 1.f; }
    static __device__ __forceinline__ float csign(float mag, float sgn) { return copysignf(mag, sgn); }
    // v * sign(s) as a sign-bit XOR (one LOP3 on the ALU pipe instead of a compare, a select and a multiply)
    static __device__ __forceinline__ float xsign(float v, float s) {
        return __uint_as_float(__float_as_uint(v) ^ (__float_as_uint(s) & 0x80000000u));
    }
    static __device__ __forceinline__ float xnsign(float v, float s) {   // -v * sign(s)
        return __uint_as_float(__float_as_uint(v) ^ (~__float_as_uint(s) & 0x80000000u));
    }
    // lg(|z| + sqrt(1+z^2)) * sign(z), given s = 1 + z^2 and r = rsqrt(s)   [asinh(z) / LGU]
    static __device__ __forceinline__ float asinh_lg(float z, float s, float r) {
        return copysignf(lg(fma_(s, r, fabsf(z))), z);
    }
    // sinh / cosh of sa * LGU (sa = argument in units of the exponential base)
    static __device__ __forceinline__ void sinhcosh(float sa, float& sh, float& ch) {
        const float e = ex2(sa);
        const float ei = rcp(e);
        sh = 0.5f * (e - ei);
        ch = 0.5f * (e + ei);
        // (e - 1/e)/2 cancels for small arguments: odd Taylor polynomial in sa (coefficients ln2^k / k!)
        const float s2 = sa * sa;
        const float p = fma_(s2, fma_(s2, fma_(s2, fma_(s2, 1.0178086e-7f, 1.5252734e-5f), 1.3333558e-3f), 5.5504109e-2f),
                             0.69314718f);
        sh = (fabsf(sa) < 0.55f) ? sa * p : sh;
    }
};

// Float64: hand-written branch-free exp / log / rsqrt / sqrt / reciprocal (natural logs) for the value ranges the chain
// kernels produce -- positive normal arguments for log / sqrt / rsqrt / rcp (Jacobian-factor products are range-checked
// by the callers), |x| < 700 for exp -- instead of libm's general-purpose versions: asinh(double) alone is 384 SASS
// instructions, log 104, exp2 / sqrt / division 80 each (special cases, denormals, correctly rounded results), and every
// chain-kernel instantiation inlines them several times.  Accuracy ~2e-16 relative (the Float64 parity budget is 1e-12).
//   rsqrt / rcp : MUFU.RSQ64H / MUFU.RCP64H seed (rsqrt.approx.ftz.f64, rcp.approx.ftz.f64: ~20 bits) + two Newton steps
//   exp2(t), t <= 0 : t = n + f (magic-number rounding), degree-12 Taylor polynomial of 2^f on [-1/2, 1/2], exponent
//                     added with integer arithmetic; t is clamped at -1020
//   exp(x)      : Cody-Waite x = n ln2 + f with a two-part ln2, degree-13 polynomial of e^f
//   log(x)      : x = 2^e m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); log m = 2 atanh(s) as a degree-10 polynomial in s^2
#ifndef ENF_F64_LIBM
#define ENF_F64_LIBM 0   // 1: libm everywhere (cross-check of the hand-written versions)
#endif
__device__ __forceinline__ double d_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double d_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = fma(fma(-hx * r, r, 0.5), r, r);
    r = fma(fma(-hx * r, r, 0.5), r, r);
    return r;
}
__device__ __forceinline__ double d_sqrt(double x) {
    const double r = d_rsqrt(x);
    const double s = x * r;
    return fma(fma(-s, s, x), 0.5 * r, s);      // one more Newton step on the square root itself
}
__device__ __forceinline__ double d_poly_exp(double f, const double c1) {
    // 1 + c1 f + (c1 f)^2/2! + ... written in g = c1 f (|g| <= 0.35): degree 13
    const double g = c1 * f;
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, g, 2.0876756987868099e-09);        // 1/12!
    p = fma(p, g, 2.5052108385441720e-08);        // 1/11!
    p = fma(p, g, 2.7557319223985893e-07);        // 1/10!
    p = fma(p, g, 2.7557319223985888e-06);        // 1/9!
    p = fma(p, g, 2.4801587301587302e-05);        // 1/8!
    p = fma(p, g, 1.9841269841269841e-04);        // 1/7!
    p = fma(p, g, 1.3888888888888889e-03);        // 1/6!
    p = fma(p, g, 8.3333333333333332e-03);        // 1/5!
    p = fma(p, g, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, g, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, g, 0.5);
    p = fma(p, g, 1.0);
    return fma(p, g, 1.0);
}
__device__ __forceinline__ double d_scale2(double p, double r_magic) {
    // p * 2^n with n in the low mantissa bits of r_magic = n + 1.5 * 2^52
    const int n = __double2loint(r_magic);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
__device__ __forceinline__ double d_exp2_neg(double t) {             // 2^t, t <= 0
    t = fmax(t, -1020.0);
    const double r = t + 6755399441055744.0;
    const double f = t - (r - 6755399441055744.0);
    return d_scale2(d_poly_exp(f, 0.69314718055994530942), r);
}
__device__ __forceinline__ double d_exp(double x) {                  // e^x, |x| < 700
    const double r = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const double n = r - 6755399441055744.0;
    double f = fma(n, -6.93147180369123816490e-01, x);               // ln2 hi (32 bits), lo: exact products for |n| < 2^20
    f = fma(n, -1.90821492927058770002e-10, f);
    return d_scale2(d_poly_exp(f, 1.0), r);
}
__device__ __forceinline__ double d_log(double x) {                  // natural log, x positive and normal
    int hi = __double2hiint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                             // m in [1, 2)
    double m = __hiloint2double(hi, __double2loint(x));
    const bool big = m > 1.4142135623730951;
    m = big ? 0.5 * m : m;
    e = big ? e + 1 : e;
    const double f = m - 1.0;
    const double s = f * d_rcp(2.0 + f);
    const double z = s * s;
    double p = 9.5238095238095233e-02;          // 2/21
    p = fma(p, z, 1.0526315789473684e-01);      // 2/19
    p = fma(p, z, 1.1764705882352941e-01);      // 2/17
    p = fma(p, z, 1.3333333333333333e-01);      // 2/15
    p = fma(p, z, 1.5384615384615385e-01);      // 2/13
    p = fma(p, z, 1.8181818181818182e-01);      // 2/11
    p = fma(p, z, 2.2222222222222221e-01);      // 2/9
    p = fma(p, z, 2.8571428571428570e-01);      // 2/7
    p = fma(p, z, 4.0000000000000002e-01);      // 2/5
    p = fma(p, z, 6.6666666666666663e-01);      // 2/3
    const double de = double(e);
    // log m = 2 s + s z p ;  log x = e ln2_hi + (e ln2_lo + log m)
    const double lm = fma(s * z, p, 2.0 * s);
    return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, lm));
}

template <> struct Prim<double> {
    static constexpr double LGU = 1.0;
    static constexpr double INV_LGU = 1.0;
    static constexpr double LG_OF_2 = 0.69314718055994530942;
    static constexpr double LG_SAFE = 600.0;
#if ENF_F64_LIBM
    static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
    static __device__ __forceinline__ double lg(double x) { return log(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rsq(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double asinh_lg(double z, double s, double r) { return asinh(z); }
    static __device__ __forceinline__ void sinhcosh(double sa, double& sh, double& ch) {
        sh = sinh(sa);
        ch = cosh(sa);
    }
#else
    static __device__ __forceinline__ double ex2(double x) { return d_exp2_neg(x); }     // every call site has x <= 0
    static __device__ __forceinline__ double lg(double x) { return d_log(x); }
    static __device__ __forceinline__ double rcp(double x) { return d_rcp(x); }
    static __device__ __forceinline__ double rsq(double x) { return d_rsqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return d_sqrt(x); }
    // asinh z = sign(z) log(|z| + sqrt(1 + z^2)), given s = 1 + z^2 and r = rsqrt(s)
    static __device__ __forceinline__ double asinh_lg(double z, double s, double r) {
        return copysign(d_log(fma(s, r, fabs(z))), z);
    }
    static __device__ __forceinline__ void sinhcosh(double sa, double& sh, double& ch) {
        const double e = d_exp(fabs(sa));
        const double ei = d_rcp(e);
        ch = 0.5 * (e + ei);
        // (e - 1/e)/2 cancels for small arguments: odd Taylor polynomial below |sa| = 0.3 (next term 0.3^16/17! < 2e-23)
        const double s2 = sa * sa;
        double p = 7.6471637318198164e-13;          // 1/15!
        p = fma(p, s2, 1.6059043836821613e-10);     // 1/13!
        p = fma(p, s2, 2.5052108385441720e-08);     // 1/11!
        p = fma(p, s2, 2.7557319223985888e-06);     // 1/9!
        p = fma(p, s2, 1.9841269841269841e-04);     // 1/7!
        p = fma(p, s2, 8.3333333333333332e-03);     // 1/5!
        p = fma(p, s2, 1.6666666666666666e-01);     // 1/3!
        const double small = fma(sa * s2, p, sa);
        sh = fabs(sa) < 0.3 ? small : copysign(0.5 * (e - ei), sa);
    }
#endif
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double sign_(double x) { return x < 0.0 ? -1.0 : 1.0; }
    static __device__ __forceinline__ double csign(double mag, double sgn) { return copysign(mag, sgn); }
    static __device__ __forceinline__ double xsign(double v, double s) {
        return __longlong_as_double(__double_as_longlong(v) ^ (__double_as_longlong(s) & (long long)0x8000000000000000ull));
    }
    static __device__ __forceinline__ double xnsign(double v, double s) {
        return __longlong_as_double(__double_as_longlong(v) ^ (~__double_as_longlong(s) & (long long)0x8000000000000000ull));
    }
};

// Packed FP32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 issue two FP32 operations per instruction).  The chain kernels
// are bound by instruction issue and by the XU pipe, not by the FMA pipe, so pairing the FP32 work of adjacent rows
// halves its share of the issue slots.
#ifndef ENF_F32X2
#define ENF_F32X2 1
#endif
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 ld2(const float* p) { return make_float2(p[0], p[1]); }

// 2^t for t <= 0 on the FMA pipe (packed pairs), for the kernels whose special-function (XU) pipe is the limiter:
// Cody-Waite split t = n + f with the round-to-nearest "magic number" trick (n lands in the low mantissa bits of
// r = t + 1.5 * 2^23), a degree-6 minimax polynomial for 2^f on [-1/2, 1/2] (relative error 1e-7 in Float32 Horner form,
// as good as ex2.approx) and the exponent added with integer arithmetic.  t is clamped at -126 (the result then is 2^-126,
// the flush-to-zero boundary of the MUFU version).  MEASURED (profiles/README.md, round 2): CenterStretch alone 0.88 ->
// 0.93 of the HBM peak, but the C3 chain 0.70 -> 0.64: the extra 6.5 issue slots per element cost more than the freed XU
// cycles give back (the chain kernel is issue-limited as much as XU-limited).  Off by default.
#ifndef ENF_EX2_POLY
#define ENF_EX2_POLY 0
#endif
__device__ __forceinline__ float2 ex2_neg_poly2(float2 t) {
    const float2 magic = make_float2(12582912.f, 12582912.f);
    t.x = fmaxf(t.x, -126.f);
    t.y = fmaxf(t.y, -126.f);
    const float2 r = add2(t, magic);
    const float2 nf = add2(r, make_float2(-12582912.f, -12582912.f));
    const float2 f = fma2(nf, make_float2(-1.f, -1.f), t);
    float2 p = make_float2(0.00015345810970757157f, 0.00015345810970757157f);
    p = fma2(p, f, make_float2(0.0013399930903688073f, 0.0013399930903688073f));
    p = fma2(p, f, make_float2(0.009618489071726799f, 0.009618489071726799f));
    p = fma2(p, f, make_float2(0.05550328642129898f, 0.05550328642129898f));
    p = fma2(p, f, make_float2(0.24022646248340607f, 0.24022646248340607f));
    p = fma2(p, f, make_float2(0.6931471824645996f, 0.6931471824645996f));
    p = fma2(p, f, make_float2(1.f, 1.f));
    p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23));
    p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23));
    return p;
}

// lg(prod_i p[i]).  Fast form (SAFE = false): ONE log of the product; if the product
// left the safe range (over/underflow), `bad` is raised and the caller recomputes the
// whole tile with SAFE = true (sum of per-element logs).  No branch on the fast path:
// a branch here would serialise the warp on the MUFU latency of every vector.
template <typename T, int N, bool SAFE>
__device__ __forceinline__ T log_of_product(const T (&p)[N], bool& bad) {
    using P = Prim<T>;
    if (SAFE || N == 1) {
        T L = P::lg(p[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) L += P::lg(p[i]);
        return L;
    }
    T prod = p[0];
#pragma unroll
    for (int i = 1; i < N; ++i) prod *= p[i];
    const T L = P::lg(prod);
    bad = bad || !(P::abs_(L) < P::LG_SAFE);
    return L;
}

// lg(prod_i n[i] / prod_i d[i])
template <typename T, int N, bool SAFE>
__device__ __forceinline__ T log_of_ratio(const T (&n)[N], const T (&d)[N], bool& bad) {
    using P = Prim<T>;
    if (SAFE) {
        T L = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) L += P::lg(n[i] * P::rcp(d[i]));   // log of the ratio: no cancellation of two large logs
        return L;
    }
    T pn = n[0], pd = d[0];
#pragma unroll
    for (int i = 1; i < N; ++i) { pn *= n[i]; pd *= d[i]; }
    const T L = P::lg(pn * P::rcp(pd));
    bad = bad || !(P::abs_(L) < P::LG_SAFE);
    return L;
}

// product of the N values (deferred logs: the caller multiplies the Jacobian factors of ALL vectors a lane
// owns of one sample and takes a single log per op, see elem_fwd_from)
template <typename T, int N>
__device__ __forceinline__ T product_of(const T (&p)[N]) {
    T prod = p[0];
#pragma unroll
    for (int i = 1; i < N; ++i) prod *= p[i];
    return prod;
}

// ---------------------------------------------------------------- forward
// Each *_fwd_v transforms GR consecutive elements that belong to ONE sample and
// adds their ladj contribution, in lg units, to `l` (row constants such as
// log|delta/lambda| and sum(log|a|) are added once per sample by the caller).
//
// constants (one value per row; K[k][e] = constant k of element e):
//   CS / CC:  0 nb2 = -b log2(e)   1 A = e^{ba}   2 ib2 = LGU/b   3 c   4 a   5 b   6 A/2   7 2/A   8 (1+A^2)/A
//   JO:       0 1/lambda   1 -xi/lambda   2 gamma   3 delta*LGU   4 delta
//   JI:       0 U/delta    1 -gamma U/delta   2 lambda   3 xi   4 1/delta   5 1/lambda   (U = log2(e) for f32, 1 for f64)

// CenterStretch: src/center_stretch.jl:4-8 (value), :39-43 (ladj = -center_contract_ladj(y)).
// With w = e^{-b|x|}, A = e^{ba}: e^{b(|u|-|x|)} = g = sqrt(m^2 + w) + m, m = (1-w)A/2 (the positive root of
// the reference's quadratic, divided through by e^{b|x|} so nothing overflows), u = y - c.
// -ladj = log S(u), S = sigma(b(u-a)) + sigma(-b(u+a)) = [P + (2/A) w g] / [P + ((1+A^2)/A) w g], P = g^2 + w^2
// (numerator and denominator scaled by g^2 A, which cancels in the ratio: no division by g is needed).
// DEFER: do not take the log here; multiply the factors into pn (and pd) instead.
template <typename T, int GR, bool LADJ, bool SAFE, bool DEFER = false>
__device__ __forceinline__ void cs_fwd_v(T* v, const T* nb2, const T* Ah, const T* ib2, const T* c, const T* k1,
                                         const T* k2, T& l, bool& bad, T* pn = nullptr, T* pd = nullptr) {
    using P = Prim<T>;
#if ENF_F32X2
    if constexpr (sizeof(T) == 4 && GR == 4 && !SAFE) {
        // two rows per FP32 instruction (same formulas as the scalar loop below)
        float2 pn2 = make_float2(1.f, 1.f), pd2 = make_float2(1.f, 1.f);
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const float2 x = make_float2(v[e], v[e + 1]);
            const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
            const float2 t = mul2(ld2(nb2 + e), ax);
#if ENF_EX2_POLY
            const float2 w = ex2_neg_poly2(t);
#else
            const float2 w = make_float2(P::ex2(t.x), P::ex2(t.y));
#endif
            const float2 ah = ld2(Ah + e);
            const float2 m = fma2(make_float2(-ah.x, -ah.y), w, ah);
            const float2 q = fma2(m, m, w);
            const float2 g = add2(make_float2(P::sqrt_(q.x), P::sqrt_(q.y)), m);
            const float2 au = fma2(make_float2(P::lg(g.x), P::lg(g.y)), ld2(ib2 + e), ax);
            const float2 y = add2(make_float2(copysignf(au.x, x.x), copysignf(au.y, x.y)), ld2(c + e));
            v[e] = y.x;
            v[e + 1] = y.y;
            if (LADJ) {
                const float2 wg = mul2(w, g);
                const float2 p2 = fma2(g, g, mul2(w, w));
                // (the first pair starts the products: the packed-multiply intrinsic is not folded against a constant 1)
                pd2 = e == 0 ? fma2(ld2(k1 + e), wg, p2) : mul2(pd2, fma2(ld2(k1 + e), wg, p2));
                pn2 = e == 0 ? fma2(ld2(k2 + e), wg, p2) : mul2(pn2, fma2(ld2(k2 + e), wg, p2));
            }
        }
        if (LADJ) {
            const float nnp = pn2.x * pn2.y, ndp = pd2.x * pd2.y;
            if (DEFER) {
                *pn *= nnp;
                *pd *= ndp;
            } else {
                const float L = P::lg(nnp * P::rcp(ndp));
                bad = bad || !(P::abs_(L) < P::LG_SAFE);
                l += L;
            }
        }
        return;
    }
#endif
    T nn[GR], nd[GR];
#pragma unroll
    for (int e = 0; e < GR; ++e) {
        const T x = v[e];
        const T ax = P::abs_(x);
        const T w = P::ex2(nb2[e] * ax);                          // e^{-b|x|}
        const T m = P::fma_(-Ah[e], w, Ah[e]);                    // (1 - w) e^{ba} / 2
        const T g = P::sqrt_(P::fma_(m, m, w)) + m;               // e^{b(|u|-|x|)}
        v[e] = P::csign(P::fma_(P::lg(g), ib2[e], ax), x) + c[e];
        if (LADJ) {
            const T wg = w * g;
            const T p2 = P::fma_(g, g, w * w);
            nd[e] = P::fma_(k1[e], wg, p2);                       // S = nd / nn
            nn[e] = P::fma_(k2[e], wg, p2);
        }
    }
    if (LADJ) {
        if (DEFER) {
            *pn *= product_of<T, GR>(nn);
            *pd *= product_of<T, GR>(nd);
        } else {
            l += log_of_ratio<T, GR, SAFE>(nn, nd, bad);                      // -log S
        }
    }
}

// CenterContract: src/center_stretch.jl:11-15 (value), :17-22,63-67 (ladj).
template <typename T, int GR, bool LADJ, bool SAFE, bool DEFER = false>
__device__ __forceinline__ void cc_fwd_v(T* v, const T* nb2, const T* A, const T* ib2, const T* c, T& l, bool& bad,
                                         T* pn = nullptr) {
    using P = Prim<T>;
#if ENF_F32X2
    if constexpr (sizeof(T) == 4 && GR == 4 && !SAFE) {
        const float2 m1 = make_float2(-1.f, -1.f), one = make_float2(1.f, 1.f);
        float2 pn2 = one, ls2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const float2 u = fma2(ld2(c + e), m1, make_float2(v[e], v[e + 1]));
            const float2 au = make_float2(fabsf(u.x), fabsf(u.y));
            const float2 t = mul2(ld2(nb2 + e), au);
            const float2 w = make_float2(P::ex2(t.x), P::ex2(t.y));
            const float2 a = ld2(A + e);
            const float2 n1 = fma2(a, w, one), n2 = add2(a, w);
            const float2 L1 = make_float2(P::lg(n1.x), P::lg(n1.y)), L2 = make_float2(P::lg(n2.x), P::lg(n2.y));
            const float2 y = fma2(fma2(L2, m1, L1), ld2(ib2 + e), au);
            v[e] = copysignf(y.x, u.x);
            v[e + 1] = copysignf(y.y, u.y);
            if (LADJ) {
                pn2 = e == 0 ? fma2(w, n1, n2) : mul2(pn2, fma2(w, n1, n2));   // S = n3 / (n1 n2)
                ls2 = e == 0 ? add2(L1, L2) : add2(ls2, add2(L1, L2));
            }
        }
        if (LADJ) {
            const float prod = pn2.x * pn2.y;
            l -= ls2.x + ls2.y;
            if (DEFER) {
                *pn *= prod;
            } else {
                const float L = P::lg(prod);
                bad = bad || !(P::abs_(L) < P::LG_SAFE);
                l += L;
            }
        }
        return;
    }
#endif
    T n3[GR];
    T ls = T(0);
#pragma unroll
    for (int e = 0; e < GR; ++e) {
        const T u = v[e] - c[e];
        const T au = P::abs_(u);
        const T w = P::ex2(nb2[e] * au);                          // e^{-b|u|}
        const T n1 = P::fma_(A[e], w, T(1));
        const T n2 = A[e] + w;
        const T L1 = P::lg(n1), L2 = P::lg(n2);
        v[e] = P::csign(P::fma_(L1 - L2, ib2[e], au), u);
        if (LADJ) {
            n3[e] = P::fma_(w, n1, n2);                           // S = n3 / (n1 n2)
            ls -= L1 + L2;
        }
    }
    if (LADJ) {
        if (DEFER) {
            l += ls;
            *pn *= product_of<T, GR>(n3);
        } else {
            l += ls + log_of_product<T, GR, SAFE>(n3, bad);                   // log S
        }
    }
}

// JohnsonTrafo: src/johnson_trafo.jl:29-32 (value), :39-42,49-52,76-80 (ladj).
template <typename T, int GR, bool LADJ, bool SAFE, bool DEFER = false>
__device__ __forceinline__ void jo_fwd_v(T* v, const T* il, const T* c0, const T* gamma, const T* delta2, T& l,
                                         bool& bad, T* pn = nullptr) {
    using P = Prim<T>;
#if ENF_F32X2
    if constexpr (sizeof(T) == 4 && GR == 4 && !SAFE) {
        float2 pr2 = make_float2(1.f, 1.f);
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const float2 z = fma2(make_float2(v[e], v[e + 1]), ld2(il + e), ld2(c0 + e));
            const float2 s = fma2(z, z, make_float2(1.f, 1.f));
            const float2 r = make_float2(P::rsq(s.x), P::rsq(s.y));
            const float2 t = fma2(s, r, make_float2(fabsf(z.x), fabsf(z.y)));          // |z| + sqrt(1 + z^2)
            const float2 a = make_float2(copysignf(P::lg(t.x), z.x), copysignf(P::lg(t.y), z.y));
            const float2 y = fma2(ld2(delta2 + e), a, ld2(gamma + e));
            v[e] = y.x;
            v[e + 1] = y.y;
            if (LADJ) pr2 = e == 0 ? r : mul2(pr2, r);
        }
        if (LADJ) {
            const float prod = pr2.x * pr2.y;
            if (DEFER) {
                *pn *= prod;
            } else {
                const float L = P::lg(prod);
                bad = bad || !(P::abs_(L) < P::LG_SAFE);
                l += L;
            }
        }
        return;
    }
#endif
    T f[GR];
#pragma unroll
    for (int e = 0; e < GR; ++e) {
        const T z = P::fma_(v[e], il[e], c0[e]);
        const T s = P::fma_(z, z, T(1));
        const T r = P::rsq(s);
        v[e] = P::fma_(delta2[e], P::asinh_lg(z, s, r), gamma[e]);
        if (LADJ) f[e] = sizeof(T) == 4 ? r : s;
    }
    if (LADJ) {
        // -log(1+z^2)/2 :  lg(prod r) for f32 (r = rsqrt(s) is already there), -lg(prod s)/2 for f64
        if (DEFER) {
            *pn *= product_of<T, GR>(f);
        } else {
            const T L = log_of_product<T, GR, SAFE>(f, bad);
            l += sizeof(T) == 4 ? L : T(-0.5) * L;
        }
    }
}

// JohnsonTrafoInv: src/johnson_trafo.jl:34-37 (value), :101-105 (ladj = -johnsontrafo_ladj(y)).
template <typename T, int GR, bool LADJ, bool SAFE, bool DEFER = false>
__device__ __forceinline__ void ji_fwd_v(T* v, const T* k0, const T* k1, const T* lam, const T* xi, T& l, bool& bad,
                                         T* pn = nullptr) {
    using P = Prim<T>;
#if ENF_F32X2
    if constexpr (sizeof(T) == 4 && GR == 4 && !SAFE) {
        const float2 half = make_float2(0.5f, 0.5f), mhalf = make_float2(-0.5f, -0.5f);
        float2 pc2 = make_float2(1.f, 1.f);
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const float2 sa = fma2(make_float2(v[e], v[e + 1]), ld2(k0 + e), ld2(k1 + e));
            const float2 ex = make_float2(P::ex2(sa.x), P::ex2(sa.y));
            const float2 ei = make_float2(P::rcp(ex.x), P::rcp(ex.y));
            const float2 he = mul2(ex, half);
            float2 sh = fma2(ei, mhalf, he);
            const float2 ch = fma2(ei, half, he);
            // (e - 1/e)/2 cancels for small arguments: odd Taylor polynomial in sa (same as Prim<float>::sinhcosh)
            const float2 s2 = mul2(sa, sa);
            float2 pl = fma2(s2, make_float2(1.0178086e-7f, 1.0178086e-7f), make_float2(1.5252734e-5f, 1.5252734e-5f));
            pl = fma2(s2, pl, make_float2(1.3333558e-3f, 1.3333558e-3f));
            pl = fma2(s2, pl, make_float2(5.5504109e-2f, 5.5504109e-2f));
            pl = fma2(s2, pl, make_float2(0.69314718f, 0.69314718f));
            const float2 sp = mul2(sa, pl);
            sh.x = fabsf(sa.x) < 0.55f ? sp.x : sh.x;
            sh.y = fabsf(sa.y) < 0.55f ? sp.y : sh.y;
            const float2 y = fma2(ld2(lam + e), sh, ld2(xi + e));
            v[e] = y.x;
            v[e + 1] = y.y;
            pc2 = e == 0 ? ch : mul2(pc2, ch);
        }
        if (LADJ) {
            const float prod = pc2.x * pc2.y;
            if (DEFER) {
                *pn *= prod;
            } else {
                const float L = P::lg(prod);
                bad = bad || !(P::abs_(L) < P::LG_SAFE);
                l += L;
            }
        }
        return;
    }
#endif
    T chs[GR];
#pragma unroll
    for (int e = 0; e < GR; ++e) {
        const T sa = P::fma_(v[e], k0[e], k1[e]);
        T sh, ch;
        P::sinhcosh(sa, sh, ch);
        v[e] = P::fma_(lam[e], sh, xi[e]);
        chs[e] = ch;
    }
    if (LADJ) {
        if (DEFER) *pn *= product_of<T, GR>(chs);
        else l += log_of_product<T, GR, SAFE>(chs, bad);                     // log sqrt(1 + sinh^2)
    }
}

// A pair of float rows as one value: the backward formulas below are written once and instantiated for T = float,
// double and F2 (two rows per FFMA2 / FMUL2 / FADD2; MUFU per half).
struct F2 {
    float2 v;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ F2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit F2(float2 a) : v(a) {}
};
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return F2(add2(a.v, b.v)); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return F2(mul2(a.v, b.v)); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return F2(fma2(b.v, make_float2(-1.f, -1.f), a.v)); }
__device__ __forceinline__ F2 operator-(F2 a) { return F2(make_float2(-a.v.x, -a.v.y)); }
template <>
struct Prim<F2> {
    using S = Prim<float>;
    static constexpr float LGU = S::LGU, INV_LGU = S::INV_LGU;
    static __device__ __forceinline__ F2 ex2(F2 x) { return F2(S::ex2(x.v.x), S::ex2(x.v.y)); }
    static __device__ __forceinline__ F2 rcp(F2 x) { return F2(S::rcp(x.v.x), S::rcp(x.v.y)); }
    static __device__ __forceinline__ F2 rsq(F2 x) { return F2(S::rsq(x.v.x), S::rsq(x.v.y)); }
    static __device__ __forceinline__ F2 fma_(F2 a, F2 b, F2 c) { return F2(fma2(a.v, b.v, c.v)); }
    static __device__ __forceinline__ F2 abs_(F2 x) { return F2(fabsf(x.v.x), fabsf(x.v.y)); }
    static __device__ __forceinline__ F2 sign_(F2 x) { return F2(x.v.x < 0.f ? -1.f : 1.f, x.v.y < 0.f ? -1.f : 1.f); }
    static __device__ __forceinline__ F2 xsign(F2 v, F2 s) { return F2(S::xsign(v.v.x, s.v.x), S::xsign(v.v.y, s.v.y)); }
    static __device__ __forceinline__ F2 xnsign(F2 v, F2 s) { return F2(S::xnsign(v.v.x, s.v.x), S::xnsign(v.v.y, s.v.y)); }
};

// ---------------------------------------------------------------- backward
// Every *_bwd takes the op's INPUT x, its OUTPUT y (both are at hand in the reverse sweep: y is the
// input of the next op) and the output cotangent G; it returns the input cotangent and writes the
// raw-sum integrands r[...] (summed over samples on the device, mapped to parameter gradients by
// enf_abi.cu: finish).  Using y avoids recomputing the forward transcendental (log / asinh / sinh).
//
// CenterContract / CenterStretch share the sigmoid algebra of src/center_stretch.jl:17-22.  With w = e^{-b|u|},
// A = e^{ba}:  n1 = 1 + A w, n2 = A + w, n3 = n2 + w n1, ONE reciprocal R = 1/(n1 n2 n3), t = n3 R = 1/(n1 n2):
//   sigma_1 = 1/n1, sigma_2 = w/n2, S = sigma_1 + sigma_2 = n3 t, ds = sigma_2 - sigma_1 = (w n1 - n2) t,
//   sigma_i' = sigma_i (1 - sigma_i):  sigma_1' = A w / n1^2, sigma_2' = A w / n2^2   (no 1 - sigma cancellation),
//   E_i = b sigma_i' / S = (b A w R) n_(3-i)^2;  Ed = E1 - E2 = b S_u / S (times sign u), Es = E1 + E2 = -b S_a / S ... (1/b) of
//   the reference's dladj/du, dladj/da; dladj/db = (|u| Ed - a Es) / b.
// The b-gradient integrand is accumulated times b (one multiplication per row on the host instead of one per element).
template <typename T>
struct CcParts { T S, iS, ds, nEd, Es; };   // nEd = -Ed

template <typename T, bool INV>
__device__ __forceinline__ CcParts<T> cc_parts(T w, T A, T b) {
    using P = Prim<T>;
    const T aw = A * w;
    const T n1 = aw + T(1);
    const T n2 = A + w;
    const T n3 = P::fma_(w, n1, n2);
    const T p12 = n1 * n2;
    const T R = P::rcp(p12 * n3);
    const T t = n3 * R;
    const T hb = (aw * b) * R;
    const T q1 = n1 * n1, q2 = n2 * n2;
    CcParts<T> o;
    o.S = n3 * t;
    o.ds = P::fma_(w, n1, -n2) * t;
    o.nEd = hb * (q1 - q2);
    o.Es = hb * (q1 + q2);
    if (INV) o.iS = p12 * (p12 * R);
    return o;
}

// CenterContract.  raw: r0 -> -dc, r1 -> da, r2 -> b db.
template <typename T>
__device__ __forceinline__ T cc_bwd(T x, T y, T G, T nb2, T A, T ib2, T c, T a, T b, T* r) {
    using P = Prim<T>;
    const T u = x - c;
    const T au = P::abs_(u);
    const CcParts<T> k = cc_parts<T, false>(P::ex2(nb2 * au), A, b);
    const T sgG = P::xsign(G, u);                                   // sign(u) G
    const T Gx = P::fma_(G, k.S, P::xsign(k.nEd, u));               // G S + LB sign(u) b S_u / S,  LB = -1
    const T ya = P::fma_(au, k.S, a * k.ds) - P::abs_(y);           // b d|y|/db
    r[0] = Gx;
    r[1] = P::fma_(sgG, k.ds, k.Es);
    r[2] = P::fma_(sgG, ya, P::fma_(au, k.nEd, a * k.Es));
    return Gx;
}

// CenterStretch (implicit inverse of CenterContract: its output y plays the contract's input).
// raw: r0 -> dc, r1 -> -da, r2 -> -b db.
template <typename T>
__device__ __forceinline__ T cs_bwd(T x, T y, T G, T nb2, T A, T ib2, T c, T a, T b, T* r) {
    using P = Prim<T>;
    const T u = y - c;
    const T au = P::abs_(u);
    const CcParts<T> k = cc_parts<T, true>(P::ex2(nb2 * au), A, b);
    const T Gy = G + P::xnsign(k.nEd, x);                           // G - LB sign(x) b S_u / S
    const T Gx = Gy * k.iS;
    const T sgGx = P::xsign(Gx, x);
    const T cb = P::fma_(au, k.S, a * k.ds) - P::abs_(x);
    r[0] = G;
    r[1] = P::fma_(sgGx, k.ds, k.Es);
    r[2] = P::fma_(sgGx, cb, P::fma_(au, k.nEd, a * k.Es));
    return Gx;
}

// JohnsonTrafo.  raw: r0 = G, r1 = G y, r2 = gz, r3 = z gz   (G asinh z = G (y - gamma)/delta: finished on the host)
template <typename T>
__device__ __forceinline__ T jo_bwd(T x, T y, T G, T il, T c0, T delta, T* r) {
    using P = Prim<T>;
    const T z = P::fma_(x, il, c0);
    const T rr = P::rsq(P::fma_(z, z, T(1)));
    const T gz = P::fma_(G * delta, rr, z * rr * rr); // G delta r - LB z r^2
    r[0] = G;
    r[1] = G * y;
    r[2] = gz;
    r[3] = z * gz;
    return gz * il;
}

// JohnsonTrafoInv.  raw: r0 = gs, r1 = s gs, r2 = G, r3 = G y   (sinh s = (y - xi)/lambda)
template <typename T>
__device__ __forceinline__ T ji_bwd(T x, T y, T G, T k0, T k1, T lam, T xi, T idl, T ilam, T* r) {
    using P = Prim<T>;
    const T s = P::fma_(x, k0, k1) * P::LGU;          // argument in natural units
    const T sh = (y - xi) * ilam;
    const T q = P::fma_(sh, sh, T(1));
    const T rq = P::rsq(q);                           // 1/cosh s
    const T gs = P::fma_(G * lam, q * rq, -sh * rq);  // G lam cosh + LB tanh
    r[0] = gs;
    r[1] = s * gs;
    r[2] = G;
    r[3] = G * y;
    return gs * idl;
}

}  // namespace enf
