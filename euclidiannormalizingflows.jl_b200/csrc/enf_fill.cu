// Synthetic benchmark input: i.i.d. N(0,1) samples from a counter-based generator
// (Philox4x32-10 keyed by the seed, counter = global element index), so the value
// of element (i, j) does not depend on how the columns are sharded over GPUs
// (SURVEY §8d).  The reference draws from Julia's unseeded global RNG
// (examples/nf_example_2d.jl:9); this only fixes the *distribution*.
#include "enf_launch.h"

namespace enf {
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename T>
__global__ void fill_normal_kernel(T* x, int64_t n_elems, int64_t elem0, uint64_t seed) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const uint64_t ctr = uint64_t(elem0 + i);
        uint32_t r[4];
        philox4x32_10(uint32_t(ctr), uint32_t(ctr >> 32), 0u, 0u, uint32_t(seed), uint32_t(seed >> 32), r);
        // Box-Muller on two uniforms in (0,1)
        const double u1 = (double(r[0]) * 4294967296.0 + double(r[2]) + 0.5) * (1.0 / 18446744073709551616.0);
        const double u2 = (double(r[1]) * 4294967296.0 + double(r[3]) + 0.5) * (1.0 / 18446744073709551616.0);
        if (sizeof(T) == 4) {
            const float rad = sqrtf(-2.0f * logf(float(u1) > 0.f ? float(u1) : 1e-30f));
            x[i] = T(rad * cospif(2.0f * float(u2)));
        } else {
            x[i] = T(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
        }
    }
}

template <typename TD, typename TS>
__global__ void convert_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t n) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) dst[i] = TD(src[i]);
}

}  // namespace

cudaError_t launch_convert(int dst_dtype, void* dst, int src_dtype, const void* src, int64_t n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dst_dtype == 1 && src_dtype == 0)
        convert_kernel<double, float><<<unsigned(blocks), threads, 0, st>>>(static_cast<double*>(dst), static_cast<const float*>(src), n);
    else if (dst_dtype == 0 && src_dtype == 1)
        convert_kernel<float, double><<<unsigned(blocks), threads, 0, st>>>(static_cast<float*>(dst), static_cast<const double*>(src), n);
    else
        return cudaMemcpyAsync(dst, src, size_t(n) * (dst_dtype == 0 ? 4 : 8), cudaMemcpyDeviceToDevice, st);
    return cudaGetLastError();
}

cudaError_t launch_fill_normal(int dtype, void* x, int D, int64_t N, int64_t col0, uint64_t seed, cudaStream_t st) {
    const int64_t n = int64_t(D) * N;
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype == 0)
        fill_normal_kernel<float><<<unsigned(blocks), threads, 0, st>>>(static_cast<float*>(x), n, col0 * D, seed);
    else
        fill_normal_kernel<double><<<unsigned(blocks), threads, 0, st>>>(static_cast<double*>(x), n, col0 * D, seed);
    return cudaGetLastError();
}

}  // namespace enf
