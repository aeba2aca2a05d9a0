// loss and parameter gradients of Householder/ScaleShift-only chains at large D
// (mvnormal_negll_trafograd, src/optimize_whitening.jl:7-22, through src/householder_trafo.jl:88-124).
//
// Such a chain is an affine map y = W x + c, and the whitening loss is quadratic in y, so the loss and every
// parameter gradient depend on the batch only through its second moments
//        S = sum_j x_j x_j^T (D x D),   m = sum_j x_j,   N.
// The reverse sweep of the reference (one pass over the D x N batch per reflection) collapses into ONE
// reduction over the samples, which is a GEMM with the SAMPLES as the contraction dimension - this is the
// tensor-core reduction over N.  A D x N column-major sample matrix is exactly the MN-major operand of that
// GEMM (row index contiguous), for A and for B, so TMA feeds tcgen05.mma without a transpose.  The chain
// rule from (S, m, N) to (negll, dV, da, db) costs O(K D^2), independent of N (enf_abi.cu: finish_moments).
//
// Float32 accuracy on TF32 tensor cores: with x = xh + xl (xh = tf32(x)) the kernel accumulates
//        P = Xh Xh^T + Xh (2 Xl)^T            (2 MMAs per k-step instead of 3)
// and the reduction kernel symmetrises, (P + P^T)/2 = Xh Xh^T + Xh Xl^T + Xl Xh^T = X X^T - Xl Xl^T (2^-22).
// The tensor core truncates when it adds into the f32 accumulator, so TMEM is drained into a per-CTA f32
// partial in global memory (L2-resident) every MO_FLUSH_STAGES stages; partials are summed in f64.
//
// Warp roles (448 threads, one CTA per SM): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer,
// warps 2-5 hi/lo split + row sums (m), warps 6-13 TMEM drain.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cooperative_groups.h>

#include <cstdlib>

#include "enf_chain.cuh"
#include "enf_launch.h"
#include "enf_tc.cuh"

namespace enf {
namespace {

constexpr int MO_KC = 32;             // samples per pipeline stage = 4 UMMA k-steps (K = 8 for tf32)
constexpr int MO_THREADS = 448;     // warp 0 TMA, warp 1 MMA, warps 2-5 split, warps 6-13 drain
constexpr int MO_SPLIT_THREADS = 128;
constexpr int MO_EPI_WARPS = 8;     // two warps per TMEM lane quarter, each drains half of the columns
#ifndef ENF_MO_FLUSH_STAGES
#define ENF_MO_FLUSH_STAGES 16        // 512 samples (128 truncating accumulations) per TMEM drain
#endif
constexpr int MO_FLUSH_STAGES = ENF_MO_FLUSH_STAGES;

template <int ND>
struct MomSmem {
    static constexpr int NG = ND / 32;                  // groups of 32 rows (one 128-byte swizzle row per sample)
    static constexpr int GROUP_BYTES = MO_KC * 128;     // [sample][32 rows] : swizzle atoms of 4 samples
    static constexpr int X_BYTES = NG * GROUP_BYTES;    // 32 KB at ND = 256
    static constexpr int STAGE_BYTES = 2 * X_BYTES;     // xh | 2 xl
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 6 ? 6 : (192 * 1024) / STAGE_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int NH = ND / 128;                 // 128-row halves of the accumulator (UMMA M = 128)
    static constexpr uint32_t TMEM_COLS = uint32_t(NH) * ND;   // 512 at ND = 256, 128 at ND = 128
    static constexpr int REPS = MO_SPLIT_THREADS / (ND / 4);   // splitter threads that share the same 4 rows
};

// MN-major shared-memory matrix descriptor for tf32 (cute::UMMA::SmemDescriptor).  32-bit MN-major operands only
// exist in the SWIZZLE_128B_BASE32B layout (layout type 1; canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in
// 16-byte units, address swizzle Swizzle<2,5,2>): 32 consecutive rows (128 bytes) per sample, 4 samples per
// 512-byte atom whose 32-byte chunks are XOR-ed with (sample & 3) - what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; the next 32 rows are `lbo_bytes` further, the next 4 samples 512 bytes.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t(lbo_bytes >> 4) << 16;
    d |= uint64_t(512 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(1) << 61;
    return d;
}

#ifndef ENF_MO_TRUNC
#define ENF_MO_TRUNC 0   // 1: feed x itself as Xh (hardware truncation) and write only the remainder: faster at D=128, less accurate
#endif
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void red_add(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// part_s: [gridDim.x][ND*ND] float (P^T of this CTA), part_m: [gridDim.x][REPS][ND] double
template <int ND>
__global__ void __launch_bounds__(MO_THREADS, 1)
moments_kernel(const __grid_constant__ CUtensorMap map_x, float* __restrict__ part_s, double* __restrict__ part_m,
               int64_t N, int stages_per_cta) {
    using S = MomSmem<ND>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // TMA landed              (1 + tx)
    uint64_t* split = full + S::STAGES;                                // xh / 2xl written        (4 warps)
    uint64_t* empty = split + S::STAGES;                               // MMAs of the stage done  (tcgen05.commit)
    uint64_t* acc_full = empty + S::STAGES;                            // flush period accumulated
    uint64_t* acc_empty = acc_full + 1;                                // TMEM drained            (4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_stages = (N + MO_KC - 1) / MO_KC;
    const int64_t st0 = int64_t(blockIdx.x) * stages_per_cta;
    int64_t st1 = st0 + stages_per_cta;
    if (st1 > total_stages) st1 = total_stages;
    const int my_stages = int(st1 - st0);                              // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], MO_SPLIT_THREADS / 32);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, MO_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: one box of 32 rows x 32 samples per row group =====
        if (lane == 0) {
            for (int it = 0; it < my_stages; ++it) {
                const int s = it % S::STAGES;
                if (it >= S::STAGES) mbar_wait(&empty[s], uint32_t(it / S::STAGES - 1) & 1u);
                unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                mbar_expect_tx(&full[s], S::X_BYTES);
                const int col = int((st0 + it) * MO_KC);               // samples beyond N are zero-filled by TMA
#pragma unroll
                                // the batch is read exactly once: evict-first, so that it does not push the CTAs' partial sums (which the
                // drains update every MO_FLUSH_STAGES stages) out of L2
                for (int g = 0; g < S::NG; ++g)
                    tma_load_2d_hint(st + g * S::GROUP_BYTES, &map_x, g * 32, col, &full[s], L2_EVICT_FIRST);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: P[h] += Xh[h] Xh^T + Xh[h] (2 Xl)^T, both operands MN-major =====
        constexpr uint32_t idesc = make_idesc_tf32(128, ND) | (1u << 15) | (1u << 16);
        int nflush = 0;
        for (int it = 0; it < my_stages; ++it) {
            const int s = it % S::STAGES;
            const uint32_t ph = uint32_t(it / S::STAGES) & 1u;
            const int fs = it % MO_FLUSH_STAGES;
            if (fs == 0 && it > 0) {
                mbar_wait(acc_empty, uint32_t(nflush - 1) & 1u);     // the previous period has been drained
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(&full[s], ph);
            mbar_wait(&split[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool last = (fs == MO_FLUSH_STAGES - 1) || (it == my_stages - 1);
            if (lane == 0) {
                const uint32_t xh = smem_u32(smem + size_t(s) * S::STAGE_BYTES), xl = xh + S::X_BYTES;
#pragma unroll
                for (int j = 0; j < MO_KC / 8; ++j) {
                    const uint64_t dbh = make_desc_mn_sw128(xh + j * 1024, S::GROUP_BYTES);
                    const uint64_t dbl = make_desc_mn_sw128(xl + j * 1024, S::GROUP_BYTES);
#pragma unroll
                    for (int h = 0; h < S::NH; ++h) {
                        const uint64_t da = make_desc_mn_sw128(xh + h * 4 * S::GROUP_BYTES + j * 1024, S::GROUP_BYTES);
                        umma_tf32(tmem_base + uint32_t(h * ND), da, dbh, idesc, (fs | j) != 0);
                        umma_tf32(tmem_base + uint32_t(h * ND), da, dbl, idesc, 1);
                    }
                }
                umma_commit(&empty[s]);
                if (last) umma_commit(acc_full);
            }
            if (last) ++nflush;
            __syncwarp();
        }
    } else if (warp < 6) {
        // ===== splitters: x -> xh (in place), 2 (x - xh) (second buffer); row sums for m =====
        constexpr int PAIRS = ND / 4;                                  // (row group, 16-byte chunk) pairs
        const int t = threadIdx.x - 64;
        const int pair = t % PAIRS, rep = t / PAIRS;
        const int g = pair >> 3, c = pair & 7;
        double macc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int it = 0; it < my_stages; ++it) {
            const int s = it % S::STAGES;
            mbar_wait(&full[s], uint32_t(it / S::STAGES) & 1u);
            unsigned char* gb = smem + size_t(s) * S::STAGE_BYTES + g * S::GROUP_BYTES;
            float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll 4
            for (int k = rep; k < MO_KC; k += S::REPS) {
                // 128B swizzle with 32-byte atoms: 32-byte chunk c/2 of sample row k lives at chunk ((c/2) ^ (k & 3))
                float4* px = reinterpret_cast<float4*>(gb + k * 128 + ((((c >> 1) ^ (k & 3)) << 5) | ((c & 1) << 4)));
                const float4 v = *px;
#if ENF_MO_TRUNC
                // the tensor core ignores the low 13 mantissa bits of a tf32 operand: x itself serves as Xh = trunc(x),
                // only the remainder is written
                const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
#else
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                *px = h;
#endif
                *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(px) + S::X_BYTES) =
                    make_float4(tf32_hi(2.f * (v.x - h.x)), tf32_hi(2.f * (v.y - h.y)), tf32_hi(2.f * (v.z - h.z)),
                                tf32_hi(2.f * (v.w - h.w)));
                r0 += v.x; r1 += v.y; r2 += v.z; r3 += v.w;
            }
            macc[0] += double(r0); macc[1] += double(r1); macc[2] += double(r2); macc[3] += double(r3);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
        double* pm = part_m + (size_t(blockIdx.x) * S::REPS + rep) * ND + g * 32 + c * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) pm[e] = macc[e];
    } else {
        // ===== drain: TMEM -> registers -> this CTA's partial P in global memory (store, then red.add) =====
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int chalf = (warp - 6) >> 2;                             // which half of the 32-column chunks
        constexpr int NCC = ND / 32 / 2;                               // chunks per warp and accumulator half
        const int nfl = (my_stages + MO_FLUSH_STAGES - 1) / MO_FLUSH_STAGES;
        float* mine = part_s + size_t(blockIdx.x) * ND * ND;
        for (int f = 0; f < nfl; ++f) {
            mbar_wait(acc_full, uint32_t(f) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int h = 0; h < S::NH; ++h) {
                // the partial is stored TRANSPOSED (P^T[col][row]; the reduction symmetrises anyway): for one
                // accumulator column the 32 lanes (rows) are contiguous, so every store / red.add is one 128-byte line
                float* colp = mine + size_t(h * 128 + quarter * 32 + lane);
#pragma unroll 1
                for (int cc = chalf * NCC; cc < (chalf + 1) * NCC; ++cc) {
                    float v[32];
                    tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(h * ND + cc * 32), v);
                    float* dst = colp + size_t(cc * 32) * ND;
                    if (f == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) dst[size_t(i) * ND] = v[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) red_add(dst + size_t(i) * ND, v[i]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// sums[(D+1)*(D+1)] = [[S, m], [m^T, N]] (row-major, float64, fixed summation order), sums[(D+1)^2] = N
__global__ void moments_reduce_kernel(const float* __restrict__ part_s, const double* __restrict__ part_m, int n_cta,
                                      int reps, int D, int64_t N, double* __restrict__ sums) {
    const int D1 = D + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        sums[size_t(D) * D1 + D] = double(N);
        sums[size_t(D1) * D1] = double(N);
    }
    if (idx < D * D) {
        const int a = idx / D, b = idx % D;
        double s = 0.0;
        for (int c = 0; c < n_cta; ++c) {
            const float* p = part_s + size_t(c) * D * D;
            s += 0.5 * (double(p[size_t(a) * D + b]) + double(p[size_t(b) * D + a]));
        }
        sums[size_t(a) * D1 + b] = s;
    } else if (idx < D * D + D) {
        const int a = idx - D * D;
        double s = 0.0;
        for (int c = 0; c < n_cta * reps; ++c) s += part_m[size_t(c) * D + a];
        sums[size_t(a) * D1 + D] = s;
        sums[size_t(D) * D1 + a] = s;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Chain rule on the device: second moments -> (negll, gradients), float64.
//
// One thread-block CLUSTER of MC_CL CTAs; CTA r keeps rows [r RB, (r+1) RB) of B (x_i = B_i [x; 1]) and of Z
// (moments of the cotangent, Z_n = B_n S^/N) in shared memory.  A reflection needs the column sums v^T B and
// v^T Z over ALL rows: every CTA publishes its partial sums in its own shared memory, one cluster barrier, and
// every CTA adds the MC_CL partials it reads through distributed shared memory (fixed order: bitwise
// reproducible).  Everything else is row-local.  Same algebra as enf_abi.cu: finish_moments and
// tests/device_model.py: affine_moments_finish.
constexpr int MC_CL = 8;           // CTAs per cluster
constexpr double MO_LOG2PI = 1.8378770664093454835606594728112;
constexpr int MC_THREADS = 256;

struct MomOp { int kind, K, poff, noff; };   // noff: offset of this op's reflections in the v.v array
struct MomChain { int n_ops, D; MomOp ops[MAX_OPS]; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[0] = negll, out[1 .. 1+P) = gradients (packed like the parameters)
__global__ void __cluster_dims__(MC_CL, 1, 1) __launch_bounds__(MC_THREADS, 1)
moments_chainrule_kernel(const __grid_constant__ MomChain mc, const double* __restrict__ params,
                         const double* __restrict__ norms, const double* __restrict__ sums, double lconst,
                         double* __restrict__ out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = int(cluster.block_rank());
    const int D = mc.D, D1 = D + 1, RB = D / MC_CL, r0 = rank * RB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = MC_THREADS / 32;
    extern __shared__ double sm[];
    double* Bs = sm;                              // [RB][D1]
    double* Zs = Bs + size_t(RB) * D1;            // [RB][D1]
    double* part = Zs + size_t(RB) * D1;          // [2 buffers][2 matrices][D1] partial column sums (read by peers)
    double* tB = part + 4 * D1;                   // [D1] v^T B
    double* tZ = tB + D1;                         // [D1] v^T Z
    double* red = tZ + D1;                        // [NW + 1] block reduction scratch / this CTA's loss partial
    const double Nd = sums[size_t(D) * D1 + D];

    for (int i = tid; i < RB * D1; i += MC_THREADS) {
        const int k = i / D1, c = i % D1;
        Bs[i] = (c == r0 + k) ? 1.0 : 0.0;
        Zs[i] = sums[size_t(r0 + k) * D1 + c] / Nd;
    }
    __syncthreads();

    int nsync = 0;                                // reflections processed so far -> partial buffer parity
    // column sums over all rows of the cluster: tB = v^T B, tZ = v^T Z
    auto column_sums = [&](const double* v) {
        double* mine = part + (nsync & 1) * 2 * D1;
        for (int c = tid; c < D1; c += MC_THREADS) {
            double sb = 0.0, sz = 0.0;
            for (int k = 0; k < RB; ++k) {
                const double vk = v[r0 + k];
                sb += vk * Bs[k * D1 + c];
                sz += vk * Zs[k * D1 + c];
            }
            mine[c] = sb;
            mine[D1 + c] = sz;
        }
        cluster.sync();
        for (int c = tid; c < D1; c += MC_THREADS) {
            double sb = 0.0, sz = 0.0;
            for (int r = 0; r < MC_CL; ++r) {
                const double* peer = cluster.map_shared_rank(mine, r);
                sb += peer[c];
                sz += peer[D1 + c];
            }
            tB[c] = sb;
            tZ[c] = sz;
        }
        ++nsync;
        __syncthreads();
    };

    // ---- forward: B_n, Z_n
    for (int o = 0; o < mc.n_ops; ++o) {
        const MomOp op = mc.ops[o];
        const double* p = params + op.poff;
        if (op.kind == OP_SS) {
            for (int i = tid; i < RB * D1; i += MC_THREADS) {
                const int k = i / D1, c = i % D1;
                const double a = p[r0 + k], b = p[D + r0 + k];
                const double w = sums[size_t(D) * D1 + c] / Nd;
                Bs[i] = a * Bs[i] + (c == D ? b : 0.0);
                Zs[i] = a * Zs[i] + b * w;
            }
            __syncthreads();
        } else {
            for (int r = 0; r < op.K; ++r) {
                const double* v = p + size_t(r) * D;
                const double s = 2.0 / norms[op.noff + r];      // v.v, float64, from the host
                column_sums(v);
                for (int i = tid; i < RB * D1; i += MC_THREADS) {
                    const int k = i / D1, c = i % D1;
                    const double f = s * v[r0 + k];
                    Bs[i] -= f * tB[c];
                    Zs[i] -= f * tZ[c];
                }
                __syncthreads();
            }
        }
    }
    // ---- loss: sum_j |y_j|^2 / 2 = N/2 <Z_n, B_n>
    {
        double acc = 0.0;
        for (int i = tid; i < RB * D1; i += MC_THREADS) acc += Zs[i] * Bs[i];
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += red[w];
            red[NW] = t;
        }
        cluster.sync();
        if (rank == 0 && tid == 0) {
            double t = 0.0;
            for (int r = 0; r < MC_CL; ++r) t += cluster.map_shared_rank(red, r)[NW];
            out[0] = 0.5 * t + 0.5 * MO_LOG2PI * D - lconst;
        }
    }
    // ---- reverse sweep
    double* g_all = out + 1;
    for (int o = mc.n_ops - 1; o >= 0; --o) {
        const MomOp op = mc.ops[o];
        const double* p = params + op.poff;
        double* g = g_all + op.poff;
        if (op.kind == OP_SS) {
            for (int k = warp; k < RB; k += NW) {                    // one warp per row
                const double a = p[r0 + k], b = p[D + r0 + k], ia = 1.0 / a;
                double* rb = Bs + k * D1;
                double* rz = Zs + k * D1;
                const double gb = rz[D];
                __syncwarp();
                double acc = 0.0;
                for (int c = lane; c < D1; c += 32) {
                    const double bin = (rb[c] - (c == D ? b : 0.0)) * ia;
                    rb[c] = bin;
                    acc += rz[c] * bin;
                    rz[c] *= a;
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    g[r0 + k] = acc - ia;
                    g[D + r0 + k] = gb;
                }
            }
            __syncthreads();
        } else {
            for (int r = op.K - 1; r >= 0; --r) {
                const double* v = p + size_t(r) * D;
                const double n = norms[op.noff + r], s = 2.0 / n;
                column_sums(v);                                       // of B_out and Z_out
                double vcv = 0.0;                                     // v^T C v = -(v^T Z_out).(v^T B_out)
                for (int c = lane; c < D1; c += 32) vcv -= tZ[c] * tB[c];
                vcv = warp_sum(vcv);
                for (int k = warp; k < RB; k += NW) {
                    const double vk = v[r0 + k], f = s * vk;
                    double* rb = Bs + k * D1;
                    double* rz = Zs + k * D1;
                    double cv = 0.0, ctv = 0.0;
                    for (int c = lane; c < D1; c += 32) {
                        const double bin = rb[c] - f * tB[c];         // B: output -> input of this reflection
                        rb[c] = bin;
                        const double z = rz[c];
                        cv -= z * tB[c];                              // v^T B_in = -tB
                        ctv += bin * tZ[c];
                        rz[c] = z - f * tZ[c];
                    }
                    cv = warp_sum(cv);
                    ctv = warp_sum(ctv);
                    if (lane == 0) g[size_t(r) * D + r0 + k] = -s * (cv + ctv) + (4.0 / (n * n)) * vcv * vk;
                }
                __syncthreads();
            }
        }
    }
    cluster.sync();   // nobody exits while a peer may still read its partial sums
}

}  // namespace

bool moments_supported(int dtype, int D) { return dtype == 0 && (D == 128 || D == 256); }

size_t moments_partial_bytes(int D, int sm_count) {
    return size_t(sm_count) * D * D * sizeof(float) + size_t(sm_count) * 4 * D * sizeof(double);
}

// d_part: moments_partial_bytes(D, sm_count) bytes of scratch; d_sums: (D+1)^2 + 1 doubles
cudaError_t launch_moments(int D, const void* x, int64_t N, void* d_part, double* d_sums, int sm_count, cudaStream_t st) {
    float* part_s = static_cast<float*>(d_part);
    double* part_m = reinterpret_cast<double*>(part_s + size_t(sm_count) * D * D);
    const int64_t total = (N + MO_KC - 1) / MO_KC;
    int n_cta = 0, reps = MO_SPLIT_THREADS / (D / 4);
    if (N > 0) {
        CUtensorMap mx;
        if (!make_map(&mx, x, uint64_t(N), uint64_t(D), MO_KC, 32, true)) return cudaErrorInvalidValue;
        const int per = int((total + sm_count - 1) / sm_count);
        n_cta = int((total + per - 1) / per);
        cudaError_t e = cudaSuccess;
#define ENF_MOMENTS_LAUNCH(ND)                                                                                     \
    {                                                                                                              \
        const int smem = MomSmem<ND>::TOTAL;                                                                       \
        static bool set[64] = {};                                                                                  \
        int dev = 0;                                                                                               \
        cudaGetDevice(&dev);                                                                                       \
        if (!set[dev & 63]) {                                                                                      \
            e = cudaFuncSetAttribute(moments_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);       \
            if (e != cudaSuccess) return e;                                                                        \
            set[dev & 63] = true;                                                                                  \
        }                                                                                                          \
        moments_kernel<ND><<<n_cta, MO_THREADS, smem, st>>>(mx, part_s, part_m, N, per);                           \
    }
        if (D == 256) ENF_MOMENTS_LAUNCH(256)
        else if (D == 128) ENF_MOMENTS_LAUNCH(128)
        else return cudaErrorInvalidValue;
#undef ENF_MOMENTS_LAUNCH
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int n = D * D + D;
    moments_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part_s, part_m, n_cta, reps, D, N, d_sums);
    return cudaGetLastError();
}


// sums -> out[0] = negll, out[1 .. 1+P) = gradients; kinds/Ks/poffs: the chain's ops in application order,
// d_params: packed float64 parameters on the device, d_norms: v.v of every reflection in application order
cudaError_t launch_moments_chainrule(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs, const double* d_params,
                                     const double* d_norms, const double* d_sums, double lconst, double* d_out, cudaStream_t st) {
    MomChain mc;
    mc.n_ops = n_ops;
    mc.D = D;
    int noff = 0;
    for (int o = 0; o < n_ops; ++o) {
        mc.ops[o] = MomOp{kinds[o], Ks[o], poffs[o], noff};
        if (kinds[o] == OP_HH) noff += Ks[o];
    }
    const int D1 = D + 1, RB = D / MC_CL;
    const size_t smem = (size_t(2) * RB * D1 + 6 * D1 + MC_THREADS / 32 + 1) * sizeof(double);
    static size_t set[64] = {};   // largest dynamic shared-memory size enabled so far, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (set[dev & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(moments_chainrule_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        set[dev & 63] = smem;
    }
    moments_chainrule_kernel<<<MC_CL, MC_THREADS, smem, st>>>(mc, d_params, d_norms, d_sums, lconst, d_out);
    return cudaGetLastError();
}

}  // namespace enf
