// batched element-wise operations of the reference's JohnsonSU distribution object
// (src/johnson_trafo.jl:1-26,109-129): pdf, logpdf, cdf, logcdf, ccdf, logccdf and quantile for scalar parameters
// (gamma, delta, xi, lambda) over N values.  The distribution is the pull-back of the standard normal through
// johnsontrafo (src/johnson_trafo.jl:29-32): every operation is the Johnson scalar kernel composed with a standard
// normal pdf / cdf / quantile, exactly as the reference writes them:
//   pdf      = deriv_johnsontrafo(x) * pdf(Normal(), johnsontrafo(x))                       (:120)
//   cdf      = cdf(Normal(), johnsontrafo(x))                                               (:121)
//   logpdf   = log(deriv_johnsontrafo(x) * pdf(Normal(), johnsontrafo(x)))                   (:123)
//   logcdf   = logcdf(Normal(), johnsontrafo(x))                                            (:124)
//   ccdf     = 1 - cdf,  logccdf = log(1 - cdf)                                             (:125-126)
//   quantile = johnsontrafo_inv(quantile(Normal(), p))                                      (:129)
// logpdf is evaluated as the sum of the logs (finite where the reference's literal log-of-a-product underflows to -Inf);
// logcdf follows StatsFuns.normlogcdf (log(erfcx(-z/sqrt2)/2) - z^2/2 in the lower tail, log1p(-erfc(z/sqrt2)/2) else).
// Memory-bound element-wise pass: 16-byte vector accesses where the pointers allow, grid sized to the SM count.
#include <cuda_runtime.h>

#include "enf_launch.h"

namespace enf {
namespace {

template <typename T> struct M;
template <> struct M<float> {
    static __device__ __forceinline__ float asinh_(float x) { return asinhf(x); }
    static __device__ __forceinline__ float sinh_(float x) { return sinhf(x); }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float normcdf_(float x) { return normcdff(x); }
    static __device__ __forceinline__ float normcdfinv_(float x) { return normcdfinvf(x); }
    static __device__ __forceinline__ float erfc_(float x) { return erfcf(x); }
    static __device__ __forceinline__ float erfcx_(float x) { return erfcxf(x); }
};
template <> struct M<double> {
    static __device__ __forceinline__ double asinh_(double x) { return asinh(x); }
    static __device__ __forceinline__ double sinh_(double x) { return sinh(x); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double normcdf_(double x) { return normcdf(x); }
    static __device__ __forceinline__ double normcdfinv_(double x) { return normcdfinv(x); }
    static __device__ __forceinline__ double erfc_(double x) { return erfc(x); }
    static __device__ __forceinline__ double erfcx_(double x) { return erfcx(x); }
};

template <typename T, int OP>
__device__ __forceinline__ T jsu_eval(T x, T gamma, T delta, T xi, T lambda) {
    using F = M<T>;
    constexpr T HALF_LOG_2PI = T(0.91893853320467274178);
    constexpr T INV_SQRT2 = T(0.70710678118654752440);
    if (OP == JSU_QUANTILE) {
        // src/johnson_trafo.jl:129 with johnsontrafo_inv of :34-37
        const T z = F::normcdfinv_(x);
        return lambda * F::sinh_((z - gamma) / delta) + xi;
    }
    const T u = (x - xi) / lambda;
    const T z = gamma + delta * F::asinh_(u);                       // johnsontrafo, :29-32
    const T s = T(1) + u * u;
    switch (OP) {
        case JSU_PDF:                                               // deriv_johnsontrafo (:39-42) * std-normal pdf
            return (delta / lambda) * (T(1) / F::sqrt_(s)) * F::exp_(-T(0.5) * z * z - HALF_LOG_2PI);
        case JSU_LOGPDF:
            return F::log_(F::abs_(delta / lambda)) - T(0.5) * F::log_(s) - T(0.5) * z * z - HALF_LOG_2PI;
        case JSU_CDF: return F::normcdf_(z);
        case JSU_LOGCDF:
            return z < T(-1) ? F::log_(F::erfcx_(-z * INV_SQRT2) * T(0.5)) - T(0.5) * z * z
                             : F::log1p_(-F::erfc_(z * INV_SQRT2) * T(0.5));
        case JSU_CCDF: return T(1) - F::normcdf_(z);                // literally 1 - cdf (:125)
        default: return F::log_(T(1) - F::normcdf_(z));             // JSU_LOGCCDF: literally log(1 - cdf) (:126)
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) johnsonsu_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t N, T gamma, T delta,
                                                        T xi, T lambda, int vec_ok) {
    constexpr int VE = 16 / int(sizeof(T));
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    const int64_t t0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nvec = vec_ok ? N / VE : 0;
    for (int64_t i = t0; i < nvec; i += stride) {
        T v[VE];
        if constexpr (sizeof(T) == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(x) + i);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            const double2 t = __ldcs(reinterpret_cast<const double2*>(x) + i);
            v[0] = t.x; v[1] = t.y;
        }
#pragma unroll
        for (int e = 0; e < VE; ++e) v[e] = jsu_eval<T, OP>(v[e], gamma, delta, xi, lambda);
        if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float4*>(out) + i, make_float4(v[0], v[1], v[2], v[3]));
        else __stcs(reinterpret_cast<double2*>(out) + i, make_double2(v[0], v[1]));
    }
    for (int64_t i = nvec * VE + t0; i < N; i += stride) out[i] = jsu_eval<T, OP>(x[i], gamma, delta, xi, lambda);
}

template <typename T>
cudaError_t launch_t(int op, const void* x, void* out, int64_t N, const double* p, int sm_count, cudaStream_t st) {
    const int vec_ok = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const int64_t want = (N + 256 * 4 - 1) / (256 * 4);
    const unsigned grid = unsigned(want < int64_t(sm_count) * 8 ? (want < 1 ? 1 : want) : int64_t(sm_count) * 8);
    const T g = T(p[0]), d = T(p[1]), xi = T(p[2]), l = T(p[3]);
    const T* xt = static_cast<const T*>(x);
    T* ot = static_cast<T*>(out);
    switch (op) {
        case JSU_PDF: johnsonsu_kernel<T, JSU_PDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_LOGPDF: johnsonsu_kernel<T, JSU_LOGPDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_CDF: johnsonsu_kernel<T, JSU_CDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_LOGCDF: johnsonsu_kernel<T, JSU_LOGCDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_CCDF: johnsonsu_kernel<T, JSU_CCDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_LOGCCDF: johnsonsu_kernel<T, JSU_LOGCCDF><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        case JSU_QUANTILE: johnsonsu_kernel<T, JSU_QUANTILE><<<grid, 256, 0, st>>>(xt, ot, N, g, d, xi, l, vec_ok); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_johnsonsu(int dtype, int op, const void* x, void* out, int64_t N, const double* params4, int sm_count,
                             cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    return dtype == 0 ? launch_t<float>(op, x, out, N, params4, sm_count, st) : launch_t<double>(op, x, out, N, params4, sm_count, st);
}

}  // namespace enf
