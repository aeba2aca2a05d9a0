// the optimize_whitening loop (src/optimize_whitening.jl:36-43) kept on the device.
//
// Per gradient step the host only launches two kernels (no synchronisation, no copies):
//   chain_grad_kernel   the fused loss + gradient pass over the batch          (enf_chain.cuh)
//   fit_update_kernel   fixed-order sum of the per-CTA partials -> gradients (the device twin of enf_abi.cu:
//                       finish) -> optimizer update -> re-normalisation of Householder columns -> loss history
//                       entry -> per-row constants + ladj row constants of the NEXT step (fit_derive)
// (sharded batches: reduce_partials + ncclAllReduce of the sums sit between the two).
// One epoch of launches is captured into a CUDA graph and replayed for the remaining epochs.
//
// Optimizer: ADAGrad with the Optimisers.jl 0.2 rule (un-vendored dependency of the reference, PARITY UNPINNED:
// state starts at eps; acc += g^2; x -= eta g / (sqrt(acc) + eps)), followed by what the reference's functor
// does on every struct rebuild: every HouseholderTrafo column is normalised (src/householder_trafo.jl:134-146).
#include <cuda_runtime.h>

#include "enf_chain.cuh"
#include "enf_launch.h"
#include "enf_p2p.cuh"

namespace enf {

namespace {

constexpr double LOG2E_D = 1.4426950408889634073599246810019;
constexpr double LN2_D = 0.69314718055994530942;
constexpr double LOG2PI_D = 1.8378770664093454835606594728112;

template <typename T>
__device__ __forceinline__ void put(T* c, size_t i, double v) { c[i] = T(v); }

// one CTA; thread per padded row for the elementwise ops, block reductions for the Householder norms
template <typename T>
__device__ void fit_derive(const FitDesc& fd, const double* __restrict__ params, T* __restrict__ consts,
                           double* __restrict__ lconst /* [2]: ss, other */, double* s_red) {
    const int D = fd.D, Dp = fd.Dp;
    const double lgu = sizeof(T) == 4 ? LN2_D : 1.0;
    const double exu = sizeof(T) == 4 ? LOG2E_D : 1.0;
    double my_ss = 0.0, my_other = 0.0;
    for (int o = 0; o < fd.n_ops; ++o) {
        const FitOp op = fd.ops[o];
        const double* p = params + op.poff;
        T* cb = consts + op.coff;
        if (op.kind == OP_HH) {
            for (int k = 0; k < op.K; ++k) {
                const double* v = p + size_t(k) * D;
                double part = 0.0;
                for (int j = threadIdx.x; j < D; j += blockDim.x) part += v[j] * v[j];
                s_red[threadIdx.x] = part;
                __syncthreads();
                for (int st = 128; st > 0; st >>= 1) {
                    if (threadIdx.x < st) s_red[threadIdx.x] += s_red[threadIdx.x + st];
                    __syncthreads();
                }
                const double sc = sqrt(2.0 / s_red[0]);
                __syncthreads();
                for (int r = threadIdx.x; r < Dp; r += blockDim.x) {
                    const bool real = fd.packed || r < D;
                    const int i = fd.packed ? r % D : r;
                    put(cb, size_t(k) * Dp + r, real ? v[i] * sc : 0.0);
                }
            }
            continue;
        }
        for (int r = threadIdx.x; r < Dp; r += blockDim.x) {
            const bool real = fd.packed || r < D;
            const int i = fd.packed ? r % D : r;
            const bool count = r < D;
            switch (op.kind) {
                case OP_CS:
                case OP_CC: {
                    const double a = real ? p[i] : 0.0, b = real ? p[D + i] : 1.0, c = real ? p[2 * D + i] : 0.0;
                    const double A = exp(b * a);
                    put(cb, 0 * size_t(Dp) + r, -b * LOG2E_D);
                    put(cb, 1 * size_t(Dp) + r, A);
                    put(cb, 2 * size_t(Dp) + r, lgu / b);
                    put(cb, 3 * size_t(Dp) + r, c);
                    put(cb, 4 * size_t(Dp) + r, a);
                    put(cb, 5 * size_t(Dp) + r, b);
                    put(cb, 6 * size_t(Dp) + r, 0.5 * A);
                    put(cb, 7 * size_t(Dp) + r, 2.0 / A);
                    put(cb, 8 * size_t(Dp) + r, (1.0 + A * A) / A);
                    break;
                }
                case OP_JO: {
                    const double gm = real ? p[i] : 0.0, dl = real ? p[D + i] : 1.0, xi = real ? p[2 * D + i] : 0.0,
                                 lm = real ? p[3 * D + i] : 1.0;
                    put(cb, 0 * size_t(Dp) + r, 1.0 / lm);
                    put(cb, 1 * size_t(Dp) + r, -xi / lm);
                    put(cb, 2 * size_t(Dp) + r, gm);
                    put(cb, 3 * size_t(Dp) + r, dl * lgu);
                    put(cb, 4 * size_t(Dp) + r, dl);
                    if (count) my_other += log(fabs(dl / lm));
                    break;
                }
                case OP_JI: {
                    const double gm = real ? p[i] : 0.0, dl = real ? p[D + i] : 1.0, xi = real ? p[2 * D + i] : 0.0,
                                 lm = real ? p[3 * D + i] : 1.0;
                    put(cb, 0 * size_t(Dp) + r, exu / dl);
                    put(cb, 1 * size_t(Dp) + r, -gm * exu / dl);
                    put(cb, 2 * size_t(Dp) + r, lm);
                    put(cb, 3 * size_t(Dp) + r, xi);
                    put(cb, 4 * size_t(Dp) + r, 1.0 / dl);
                    put(cb, 5 * size_t(Dp) + r, 1.0 / lm);
                    if (count) my_other += log(fabs(lm / dl));
                    break;
                }
                default: {  // OP_SS
                    const double a = real ? p[i] : 1.0, b = real ? p[D + i] : 0.0;
                    put(cb, 0 * size_t(Dp) + r, a);
                    put(cb, 1 * size_t(Dp) + r, b);
                    if (count) my_ss += log(fabs(a));
                    break;
                }
            }
        }
    }
    // row constants of ladj
    for (int which = 0; which < 2; ++which) {
        __syncthreads();
        s_red[threadIdx.x] = which == 0 ? my_ss : my_other;
        __syncthreads();
        for (int st = 128; st > 0; st >>= 1) {
            if (threadIdx.x < st) s_red[threadIdx.x] += s_red[threadIdx.x + st];
            __syncthreads();
        }
        if (threadIdx.x == 0) lconst[which] = s_red[0];
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) fit_derive_kernel(const __grid_constant__ FitDesc fd, const double* __restrict__ params,
                                                         T* __restrict__ consts, double* __restrict__ lconst) {
    __shared__ double s_red[256];
    fit_derive<T>(fd, params, consts, lconst, s_red);
}

// raw per-row sum `slot` of an op for row i (packed layouts keep VE/D copies of every row)
__device__ __forceinline__ double row_sum(const FitDesc& fd, const double* sums, int roff, int slot, int i) {
    const double* base = sums + size_t(roff + slot) * fd.Dp;
    if (!fd.packed) return base[i];
    double s = 0.0;
    for (int r = i; r < fd.Dp; r += fd.D) s += base[r];
    return s;
}

// one CTA.  sums: n_raw raw sums + 1 (sample count of the batch, summed over the group).
// partials != nullptr: first add up the gradient kernel's per-CTA partial sums (fixed order) -- the single-GPU
// step is then two launches: chain_grad_kernel and this one.  At the end the constants of the NEXT step are
// derived from the updated parameters.
template <typename T>
__global__ void __launch_bounds__(256) fit_update_kernel(const __grid_constant__ FitDesc fd, double* __restrict__ sums,
                                                         const double* __restrict__ partials, int n_blocks, double count,
                                                         double* __restrict__ lconst, double* __restrict__ params,
                                                         double* __restrict__ state, double eta, double eps, int flags,
                                                         double* __restrict__ history, long long* __restrict__ step_ctr,
                                                         T* __restrict__ consts, const __grid_constant__ P2PDesc p2p) {
    const int D = fd.D;
    __shared__ double s_red[256];
    // Small problems (the D = 1 / D = 2 fits of BASELINE configs[0-1]) work out of shared memory: a step of this kernel
    // used to be a chain of ~10 dependent L2 round trips (sums -> gradients -> state -> parameters -> constants).
    constexpr int FIT_SMEM_PARAMS = 1024, FIT_SMEM_SUMS = 2048;
    __shared__ double s_params[FIT_SMEM_PARAMS], s_state[FIT_SMEM_PARAMS], s_sums[FIT_SMEM_SUMS];
    __shared__ double s_lc[2];
    __shared__ long long s_step;
    int n_params = 0;
    for (int o = 0; o < fd.n_ops; ++o) {
        const int e = fd.ops[o].poff + (fd.ops[o].kind == OP_HH ? fd.ops[o].K * D : (fd.ops[o].kind == OP_SS ? 2 : fd.ops[o].kind == OP_JO || fd.ops[o].kind == OP_JI ? 4 : 3) * D);
        n_params = e > n_params ? e : n_params;
    }
    const bool small = n_params <= FIT_SMEM_PARAMS && fd.n_raw + 1 <= FIT_SMEM_SUMS;
    double* const g_params = params;
    double* const g_state = state;
    double* const g_sums = sums;
    // Programmatic dependent launch (see chain_grad_kernel): this CTA may be resident while the gradient kernel still
    // runs.  Parameters, optimizer state, ladj constants and the step counter were written by the PREVIOUS update kernel,
    // which is complete (the gradient kernel waited for it before it let this kernel launch): fetch them now, under the
    // gradient kernel's run time; the partial sums may only be read after the wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (small) {
        for (int i = threadIdx.x; i < n_params; i += blockDim.x) {
            s_params[i] = g_params[i];
            s_state[i] = g_state[i];
        }
        params = s_params;
        state = s_state;
        sums = s_sums;
    }
    if (threadIdx.x == 0) {
        s_lc[0] = lconst[0];
        s_lc[1] = lconst[1];
        s_step = *step_ctr;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (partials == nullptr && small) {          // sums already reduced (and all-reduced) in global memory by the caller
        for (int i = threadIdx.x; i <= fd.n_raw; i += blockDim.x) s_sums[i] = g_sums[i];
        __syncthreads();
    }
    if (partials != nullptr) {
        // fixed order, eight lanes per output with four loads in flight each (as reduce_partials_kernel): one thread
        // walking all CTAs is a chain of n_blocks dependent L2 round trips on the critical path of every step
        const int j = threadIdx.x & 7;
        for (int i0 = 0; i0 < fd.n_raw; i0 += blockDim.x >> 3) {
            const int i = i0 + (threadIdx.x >> 3);
            double s = 0.0;
            if (i < fd.n_raw) {
                int b = j;
                for (; b + 24 < n_blocks; b += 32) {
                    const double p0 = partials[size_t(b) * fd.n_raw + i], p1 = partials[size_t(b + 8) * fd.n_raw + i];
                    const double p2 = partials[size_t(b + 16) * fd.n_raw + i], p3 = partials[size_t(b + 24) * fd.n_raw + i];
                    s += p0;
                    s += p1;
                    s += p2;
                    s += p3;
                }
                for (; b < n_blocks; b += 8) s += partials[size_t(b) * fd.n_raw + i];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (i < fd.n_raw && j == 0) sums[i] = s;
        }
        if (threadIdx.x == 0) sums[fd.n_raw] = count;
        __syncthreads();
        // sharded batch: sum the raw sums (and the sample count) over the ranks through NVLink peer memory, in this kernel
        if (p2p.nranks > 1) {
            p2p_allreduce_block(p2p, sums, fd.n_raw + 1);
            __syncthreads();
        }
    }
    const double Nd = sums[fd.n_raw];
    const double LB = -1.0;
    if (threadIdx.x == 0) {
        const double lc = s_lc[1] + ((flags & 1) ? 0.0 : s_lc[0]);      // ENF_NEGLL_ZYGOTE_PRIMAL drops the ScaleShift ladj value
        // the step counter lives on the device so that one captured epoch (a CUDA graph) can be replayed
        const long long step = s_step;
        history[step] = (sums[fd.n_raw - 2] + 0.5 * LOG2PI_D * Nd * D - (sums[fd.n_raw - 1] + Nd * lc)) / Nd;
        *step_ctr = step + 1;
    }
    for (int o = 0; o < fd.n_ops; ++o) {
        const FitOp op = fd.ops[o];
        double* p = params + op.poff;
        double* st = state + op.poff;
        if (op.kind == OP_HH) {
            for (int k = 0; k < op.K; ++k) {
                double* v = p + size_t(k) * D;
                double part = 0.0;
                for (int j = threadIdx.x; j < D; j += blockDim.x) part += v[j] * v[j];
                s_red[threadIdx.x] = part;
                __syncthreads();
                for (int s2 = 128; s2 > 0; s2 >>= 1) {
                    if (threadIdx.x < s2) s_red[threadIdx.x] += s_red[threadIdx.x + s2];
                    __syncthreads();
                }
                const double n = s_red[0];
                __syncthreads();
                const double acc2 = sums[size_t(fd.n_rowslots) * fd.Dp + op.soff + k];
                // gradient + ADAGrad step, then the norm of the updated column
                part = 0.0;
                for (int i = threadIdx.x; i < D; i += blockDim.x) {
                    const double g = (-sqrt(2.0 / n) * row_sum(fd, sums, op.roff, k, i) + (2.0 / n) * acc2 * v[i]) / Nd;
                    const double a = st[size_t(k) * D + i] + g * g;
                    st[size_t(k) * D + i] = a;
                    const double nv = v[i] - eta * g / (sqrt(a) + eps);
                    v[i] = nv;
                    part += nv * nv;
                }
                s_red[threadIdx.x] = part;
                __syncthreads();
                for (int s2 = 128; s2 > 0; s2 >>= 1) {
                    if (threadIdx.x < s2) s_red[threadIdx.x] += s_red[threadIdx.x + s2];
                    __syncthreads();
                }
                const double inv = 1.0 / sqrt(s_red[0]);     // functor rebuild: normalize!(column)
                __syncthreads();
                for (int i = threadIdx.x; i < D; i += blockDim.x) v[i] *= inv;
                __syncthreads();
            }
            continue;
        }
        for (int i = threadIdx.x; i < D; i += blockDim.x) {
            double g[4] = {0.0, 0.0, 0.0, 0.0};
            const double r0 = row_sum(fd, sums, op.roff, 0, i), r1 = row_sum(fd, sums, op.roff, 1, i);
            int nf = 2;
            switch (op.kind) {
                case OP_CS: { const double r2 = row_sum(fd, sums, op.roff, 2, i); g[0] = -r1; g[1] = -r2 / p[D + i]; g[2] = r0; nf = 3; break; }
                case OP_CC: { const double r2 = row_sum(fd, sums, op.roff, 2, i); g[0] = r1; g[1] = r2 / p[D + i]; g[2] = -r0; nf = 3; break; }
                case OP_JO: {
                    const double r2 = row_sum(fd, sums, op.roff, 2, i), r3 = row_sum(fd, sums, op.roff, 3, i);
                    const double gm = p[i], dl = p[D + i], lm = p[3 * D + i];
                    g[0] = r0; g[1] = (r1 - gm * r0) / dl + LB * Nd / dl; g[2] = -r2 / lm; g[3] = -(r3 + LB * Nd) / lm; nf = 4;
                    break;
                }
                case OP_JI: {
                    const double r2 = row_sum(fd, sums, op.roff, 2, i), r3 = row_sum(fd, sums, op.roff, 3, i);
                    const double dl = p[D + i], xi = p[2 * D + i], lm = p[3 * D + i];
                    g[0] = -r0 / dl; g[1] = -(r1 + LB * Nd) / dl; g[2] = r2; g[3] = (r3 - xi * r2) / lm + LB * Nd / lm; nf = 4;
                    break;
                }
                default: { g[0] = r0 + LB * Nd / p[i]; g[1] = r1; nf = 2; break; }   // OP_SS
            }
            for (int f = 0; f < nf; ++f) {
                const double gg = g[f] / Nd;
                const double a = st[size_t(f) * D + i] + gg * gg;
                st[size_t(f) * D + i] = a;
                p[size_t(f) * D + i] -= eta * gg / (sqrt(a) + eps);
            }
        }
        __syncthreads();
    }
    fit_derive<T>(fd, params, consts, lconst, s_red);     // constants for the next step
    if (small) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_params; i += blockDim.x) {
            g_params[i] = s_params[i];
            g_state[i] = s_state[i];
        }
        for (int i = threadIdx.x; i <= fd.n_raw; i += blockDim.x) g_sums[i] = s_sums[i];
    }
}

}  // namespace

cudaError_t launch_fit_derive(int dtype, const FitDesc& fd, const double* params, void* consts, double* lconst,
                              cudaStream_t st) {
    if (dtype == 0) fit_derive_kernel<float><<<1, 256, 0, st>>>(fd, params, static_cast<float*>(consts), lconst);
    else fit_derive_kernel<double><<<1, 256, 0, st>>>(fd, params, static_cast<double*>(consts), lconst);
    return cudaGetLastError();
}

// p2p != nullptr (and partials != nullptr): the batch is sharded over the ranks of p2p; the kernel all-reduces the sums itself
cudaError_t launch_fit_update(int dtype, const FitDesc& fd, double* sums, const double* partials, int n_blocks, double count,
                              double* lconst, double* params, double* state, double eta, double eps, int flags,
                              double* history, long long* step_ctr, void* consts, cudaStream_t st, const P2PDesc* p2p, bool pdl) {
    P2PDesc none = {};
    const P2PDesc& pd = p2p ? *p2p : none;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (dtype == 0)
        return cudaLaunchKernelEx(&cfg, fit_update_kernel<float>, fd, sums, partials, n_blocks, count, lconst, params, state, eta, eps,
                                  flags, history, step_ctr, static_cast<float*>(consts), pd);
    return cudaLaunchKernelEx(&cfg, fit_update_kernel<double>, fd, sums, partials, n_blocks, count, lconst, params, state, eta, eps,
                              flags, history, step_ctr, static_cast<double*>(consts), pd);
}

}  // namespace enf
