// Accuracy of the hand-written Float64 primitives of csrc/enf_math.cuh against CUDA's libm (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I euclidiannormalizingflows.jl_b200/csrc -o build/tmp/f64check tools/f64_math_check.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "enf_math.cuh"
using namespace enf;
__global__ void k(const double* x, double* o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double u = x[i];                      // uniform in (0,1)
    const double pos = exp((u - 0.5) * 600.0);  // positive, 1e-130 .. 1e130
    const double neg = -u * 1000.0;             // exp2 argument
    const double sym = (u - 0.5) * 80.0;        // exp / sinh argument
    double* r = o + size_t(i) * 8;
    auto rel = [](double a, double b) { return fabs(a - b) / fabs(b); };
    r[0] = rel(d_rcp(pos), 1.0 / pos);
    r[1] = rel(d_rsqrt(pos), 1.0 / sqrt(pos));
    r[2] = rel(d_sqrt(pos), sqrt(pos));
    r[3] = rel(d_exp2_neg(neg), exp2(neg));
    r[4] = rel(d_exp(sym), exp(sym));
    r[5] = fabs(d_log(pos) - log(pos)) / (fabs(log(pos)) + 1.0);
    double sh, ch; Prim<double>::sinhcosh(sym * 0.5, sh, ch);
    r[6] = fmax(rel(ch, cosh(sym * 0.5)), fabs(sh - sinh(sym * 0.5)) / (fabs(sinh(sym * 0.5)) + 1e-300));
    const double z = (u - 0.5) * 2e3, s = fma(z, z, 1.0);
    r[7] = fabs(Prim<double>::asinh_lg(z, s, d_rsqrt(s)) - asinh(z)) / (fabs(asinh(z)) + 1.0);
}
int main() {
    const int n = 1 << 22;
    std::vector<double> h(n);
    unsigned long long st = 88172645463325252ull;
    for (auto& v : h) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = ((st >> 11) + 0.5) / 9007199254740992.0; }
    double *dx, *dout;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dout, size_t(n) * 64);
    cudaMemcpy(dx, h.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, dout, n);
    std::vector<double> o(size_t(n) * 8);
    if (cudaMemcpy(o.data(), dout, size_t(n) * 64, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    const char* names[8] = {"rcp", "rsqrt", "sqrt", "exp2(t<=0)", "exp", "log (abs/(|log|+1))", "sinh/cosh", "asinh (abs/(|.|+1))"};
    for (int j = 0; j < 8; ++j) { double m = 0; for (int i = 0; i < n; ++i) m = fmax(m, o[size_t(i) * 8 + j]); printf("%-24s max err %.3e\n", names[j], m); }
    return 0;
}
