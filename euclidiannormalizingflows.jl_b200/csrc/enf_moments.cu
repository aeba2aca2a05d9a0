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
// Warp roles (576 threads, one CTA per SM): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer,
// warps 2-9 hi/lo split + row sums (m), warps 10-17 TMEM drain.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "enf_chain.cuh"
#include "enf_launch.h"
#include "enf_tc.cuh"

namespace enf {
namespace {

constexpr int MO_KC = 32;             // samples per pipeline stage = 4 UMMA k-steps (K = 8 for tf32)
constexpr int MO_THREADS = 576;     // warp 0 TMA, warp 1 MMA, warps 2-9 split, warps 10-17 drain
constexpr int MO_SPLIT_THREADS = 256;
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
    static constexpr int NBUF = ND == 256 ? 1 : 2;      // accumulators: at ND = 128 the drain of one overlaps the MMAs into the other
    static constexpr uint32_t TMEM_COLS = uint32_t(NBUF) * NH * ND;   // 512 at ND = 256, 256 at ND = 128
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
    uint64_t* split = full + S::STAGES;                                // xh / 2xl written        (8 warps)
    uint64_t* empty = split + S::STAGES;                               // MMAs of the stage done  (tcgen05.commit)
    uint64_t* acc_full = empty + S::STAGES;                            // [NBUF] flush period accumulated
    uint64_t* acc_empty = acc_full + S::NBUF;                          // [NBUF] TMEM drained     (8 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + S::NBUF);

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
        for (int b = 0; b < S::NBUF; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], MO_EPI_WARPS);
        }
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
        // ===== MMA issuer: P[h] += Xh[h] Xh^T + Xh[h] (2 Xl)^T, both operands MN-major.  ONE thread runs the whole loop
        // with one wait per stage (split[s]; the splitters have seen the TMA land): every cycle the issuer spends in waits,
        // re-convergence or commits is a cycle the tensor pipe idles (tools/mma_rate.cu) =====
        constexpr uint32_t idesc = make_idesc_tf32(128, ND) | (1u << 15) | (1u << 16);
        if (lane == 0) {
            for (int it = 0; it < my_stages; ++it) {
                const int s = it % S::STAGES;
                const int fs = it % MO_FLUSH_STAGES, period = it / MO_FLUSH_STAGES, buf = period % S::NBUF;
                if (fs == 0 && period >= S::NBUF)
                    mbar_wait(&acc_empty[buf], uint32_t(period / S::NBUF - 1) & 1u);   // this accumulator has been drained
                mbar_wait(&split[s], uint32_t(it / S::STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool last = (fs == MO_FLUSH_STAGES - 1) || (it == my_stages - 1);
                const uint32_t xh = smem_u32(smem + size_t(s) * S::STAGE_BYTES), xl = xh + S::X_BYTES;
                const uint32_t acc = tmem_base + uint32_t(buf * S::NH * ND);
#pragma unroll
                for (int j = 0; j < MO_KC / 8; ++j) {
                    const uint64_t dbh = make_desc_mn_sw128(xh + j * 1024, S::GROUP_BYTES);
                    const uint64_t dbl = make_desc_mn_sw128(xl + j * 1024, S::GROUP_BYTES);
#pragma unroll
                    for (int h = 0; h < S::NH; ++h) {
                        const uint64_t da = make_desc_mn_sw128(xh + h * 4 * S::GROUP_BYTES + j * 1024, S::GROUP_BYTES);
                        umma_tf32(acc + uint32_t(h * ND), da, dbh, idesc, (fs | j) != 0);
                        umma_tf32(acc + uint32_t(h * ND), da, dbl, idesc, 1);
                    }
                }
                umma_commit(&empty[s]);
                if (last) umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 2 + MO_SPLIT_THREADS / 32) {
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
        const int chalf = (warp - 2 - MO_SPLIT_THREADS / 32) >> 2;     // which half of the 32-column chunks
        constexpr int NCC = ND / 32 / 2;                               // chunks per warp and accumulator half
        const int nfl = (my_stages + MO_FLUSH_STAGES - 1) / MO_FLUSH_STAGES;
        float* mine = part_s + size_t(blockIdx.x) * ND * ND;
        for (int f = 0; f < nfl; ++f) {
            const int buf = f % S::NBUF;
            mbar_wait(&acc_full[buf], uint32_t(f / S::NBUF) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int h = 0; h < S::NH; ++h) {
                // the partial is stored TRANSPOSED (P^T[col][row]; the reduction symmetrises anyway): for one
                // accumulator column the 32 lanes (rows) are contiguous, so every store / red.add is one 128-byte line
                float* colp = mine + size_t(h * 128 + quarter * 32 + lane);
#pragma unroll 1
                for (int cc = chalf * NCC; cc < (chalf + 1) * NCC; ++cc) {
                    float v[32];
                    tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * S::NH * ND + h * ND + cc * 32), v);
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
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// sums[(D+1)*(D+1)] = [[S, m], [m^T, N]] (row-major, float64, fixed summation order), sums[(D+1)^2] = N
// ND: row count the moments kernel ran with (D = 64 runs the 128-row kernel, rows 64.. are TMA zero fill)
__global__ void moments_reduce_kernel(const float* __restrict__ part_s, const double* __restrict__ part_m, int n_cta,
                                      int reps, int D, int ND, int64_t N, double* __restrict__ sums) {
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
            const float* p = part_s + size_t(c) * ND * ND;
            s += 0.5 * (double(p[size_t(a) * ND + b]) + double(p[size_t(b) * ND + a]));
        }
        sums[size_t(a) * D1 + b] = s;
    } else if (idx < D * D + D) {
        const int a = idx - D * D;
        double s = 0.0;
        for (int c = 0; c < n_cta * reps; ++c) s += part_m[size_t(c) * ND + a];
        sums[size_t(a) * D1 + D] = s;
        sums[size_t(D) * D1 + a] = s;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Chain rule on the device: second moments -> (negll, gradients), float64.
//
// With x^ = [x; 1] every intermediate of the chain is x_i = B_i x^ and the cotangent moments are Z_i (Z_n = B_n S^/N,
// Z_(i-1) = A_i^T Z_i).  Both are D x (D+1) matrices on which every op acts COLUMN by column: a reflection needs
// v^T B[:, c] and v^T Z[:, c] of the same column only.  So the columns are dealt out to independent CTAs (MC_NC
// columns each, one thread per row, the 2 MC_NC entries of a row in registers) and the sweep needs no communication
// between CTAs at all: what couples the columns - the row dot products of the parameter gradients and the loss - is
// additive, every CTA writes its partial (negll, gradient) vector and a second kernel adds them in a fixed order.
// Per reflection a CTA does one block-wide reduction of 2 MC_NC values.  Same algebra as enf_abi.cu: finish_moments
// and tests/device_model.py: affine_moments_finish.
constexpr int MC_NC = 2;           // columns of B and Z per CTA (129 CTAs at D = 256: one block reduction of 4 values per reflection)
constexpr double MO_LOG2PI = 1.8378770664093454835606594728112;

struct MomOp { int kind, K, poff, noff; };   // noff: offset of this op's reflections in the v.v array
struct MomChain { int n_ops, D; MomOp ops[MAX_OPS]; };

// part: [gridDim.x][1 + P]  (partial sum_j |y_j|^2 / (2N), partial gradients)
__global__ void __launch_bounds__(256) moments_chainrule_kernel(const __grid_constant__ MomChain mc, const double* __restrict__ params,
                                                                const double* __restrict__ norms, const double* __restrict__ sums,
                                                                int n_params, double* __restrict__ part) {
    const int D = mc.D, D1 = D + 1, k = threadIdx.x, c0 = blockIdx.x * MC_NC;
    const int warp = k >> 5, lane = k & 31, NW = blockDim.x >> 5;
    __shared__ double s_w[2][8][2 * MC_NC];      // per-warp partial sums, double-buffered (one barrier per reduction)
    const double Nd = sums[size_t(D) * D1 + D];
    double B[MC_NC], Z[MC_NC], w[MC_NC];
    int jD = -1;                                  // which of my columns is the homogeneous one (c == D), if any
#pragma unroll
    for (int j = 0; j < MC_NC; ++j) {
        const int c = c0 + j;
        const bool valid = c < D1;
        B[j] = (c == k) ? 1.0 : 0.0;
        Z[j] = valid ? sums[size_t(k) * D1 + c] / Nd : 0.0;
        w[j] = valid ? sums[size_t(D) * D1 + c] / Nd : 0.0;
        if (c == D) jD = j;
    }
    int nred = 0;
    // t[0..NC) = v^T B, t[NC..2NC) = v^T Z over the rows (all threads get all sums, fixed summation order)
    auto column_sums = [&](double vk, double (&t)[2 * MC_NC]) {
        double p[2 * MC_NC];
#pragma unroll
        for (int j = 0; j < MC_NC; ++j) {
            p[j] = vk * B[j];
            p[MC_NC + j] = vk * Z[j];
        }
#pragma unroll
        for (int i = 0; i < 2 * MC_NC; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p[i] += __shfl_xor_sync(0xffffffffu, p[i], o);
        }
        double (*buf)[2 * MC_NC] = s_w[nred & 1];
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 2 * MC_NC; ++i) buf[warp][i] = p[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2 * MC_NC; ++i) {
            double a = 0.0;
            for (int ww = 0; ww < NW; ++ww) a += buf[ww][i];
            t[i] = a;
        }
        ++nred;
    };
    // ---- forward: B_n, Z_n
    for (int o = 0; o < mc.n_ops; ++o) {
        const MomOp op = mc.ops[o];
        const double* p = params + op.poff;
        if (op.kind == OP_SS) {
            const double a = p[k], b = p[D + k];
#pragma unroll
            for (int j = 0; j < MC_NC; ++j) {
                B[j] = a * B[j] + (j == jD ? b : 0.0);
                Z[j] = a * Z[j] + b * w[j];
            }
        } else {
            for (int r = 0; r < op.K; ++r) {
                const double vk = p[size_t(r) * D + k];
                const double f = (2.0 / norms[op.noff + r]) * vk;
                double t[2 * MC_NC];
                column_sums(vk, t);
#pragma unroll
                for (int j = 0; j < MC_NC; ++j) {
                    B[j] -= f * t[j];
                    Z[j] -= f * t[MC_NC + j];
                }
            }
        }
    }
    double* mine = part + size_t(blockIdx.x) * (1 + n_params);
    {   // ---- loss: sum_j |y_j|^2 / (2N) = <Z_n, B_n> / 2, this CTA's columns
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < MC_NC; ++j) acc += Z[j] * B[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        double (*buf)[2 * MC_NC] = s_w[nred & 1];
        if (lane == 0) buf[warp][0] = acc;
        __syncthreads();
        if (k == 0) {
            double a = 0.0;
            for (int ww = 0; ww < NW; ++ww) a += buf[ww][0];
            mine[0] = 0.5 * a;
        }
        ++nred;
    }
    // ---- reverse sweep
    double* g_all = mine + 1;
    for (int o = mc.n_ops - 1; o >= 0; --o) {
        const MomOp op = mc.ops[o];
        const double* p = params + op.poff;
        double* g = g_all + op.poff;
        if (op.kind == OP_SS) {
            const double a = p[k], b = p[D + k], ia = 1.0 / a;
            double gb = 0.0, acc = 0.0;
#pragma unroll
            for (int j = 0; j < MC_NC; ++j) {
                if (j == jD) gb = Z[j];                               // db = Z[:, D] (only the CTA that owns that column)
                const double bin = (B[j] - (j == jD ? b : 0.0)) * ia;
                B[j] = bin;
                acc += Z[j] * bin;
                Z[j] *= a;
            }
            g[k] = acc - (blockIdx.x == 0 ? ia : 0.0);   // the ladj term -1/a enters once
            g[D + k] = gb;
        } else {
            for (int r = op.K - 1; r >= 0; --r) {
                const double vk = p[size_t(r) * D + k];
                const double n = norms[op.noff + r], s = 2.0 / n, f = s * vk;
                double t[2 * MC_NC];
                column_sums(vk, t);                                   // of B_out and Z_out
                double cv = 0.0, ctv = 0.0, vcv = 0.0;
#pragma unroll
                for (int j = 0; j < MC_NC; ++j) {
                    const double bin = B[j] - f * t[j];               // B: output -> input of this reflection
                    B[j] = bin;
                    cv -= Z[j] * t[j];                                // v^T B_in = -(v^T B_out)
                    ctv += bin * t[MC_NC + j];
                    vcv -= t[MC_NC + j] * t[j];
                    Z[j] -= f * t[MC_NC + j];
                }
                g[size_t(r) * D + k] = -s * (cv + ctv) + (4.0 / (n * n)) * vcv * vk;
            }
        }
    }
}

// out[0] = negll, out[1 .. 1+P) = gradients: fixed-order sum of the per-CTA partials
__global__ void moments_chainrule_reduce_kernel(const double* __restrict__ part, int n_cta, int n_params, int D, double lconst,
                                                const double* __restrict__ lconst_dev, int n_lconst, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_params) return;
    double a = 0.0;
    for (int c = 0; c < n_cta; ++c) a += part[size_t(c) * (1 + n_params) + i];
    if (i == 0) {
        double lc = lconst;                       // ladj row constants: host value, or one device slot per op
        if (lconst_dev != nullptr) {
            lc = 0.0;
            for (int o = 0; o < n_lconst; ++o) lc += lconst_dev[o];
        }
        a += 0.5 * MO_LOG2PI * D - lc;
    }
    out[i] = a;
}

// Device-side optimizer step for second-moment chains (enf_optimize_whitening): ADAGrad per Optimisers 0.2
// (acc += g^2; x -= eta g / (sqrt(acc) + eps)) on the gradients the chain-rule kernels left in out[1..], then the
// HouseholderTrafo functor rebuild (src/householder_trafo.jl:134-146: every column normalised), the v.v array and
// the ScaleShift ladj constants of the NEXT step (one slot per op), and the loss history.  One CTA per Householder
// column / per ScaleShift op.
__global__ void __launch_bounds__(256) moments_update_kernel(const __grid_constant__ MomChain mc, const double* __restrict__ out,
                                                             double* __restrict__ params, double* __restrict__ norms,
                                                             double* __restrict__ state, double eta, double eps, int flags,
                                                             double* __restrict__ lconst_dev, double* __restrict__ history,
                                                             long long* __restrict__ step_ctr) {
    const int D = mc.D, tid = threadIdx.x;
    __shared__ double s_red[256];
    auto block_sum = [&](double v) {
        s_red[tid] = v;
        __syncthreads();
        for (int s2 = 128; s2 > 0; s2 >>= 1) {
            if (tid < s2) s_red[tid] += s_red[tid + s2];
            __syncthreads();
        }
        const double r = s_red[0];
        __syncthreads();
        return r;
    };
    if (blockIdx.x == 0 && tid == 0) {
        const long long step = *step_ctr;   // on the device so that a captured epoch can be replayed
        history[step] = out[0];
        *step_ctr = step + 1;
    }
    int o = 0, k = int(blockIdx.x);          // this CTA's unit: column k of Householder op o, or ScaleShift op o
    for (; o < mc.n_ops; ++o) {
        const int cnt = mc.ops[o].kind == OP_HH ? mc.ops[o].K : 1;
        if (k < cnt) break;
        k -= cnt;
    }
    if (o >= mc.n_ops) return;
    const MomOp op = mc.ops[o];
    const double* g_all = out + 1;
    if (op.kind == OP_SS) {
        double* p = params + op.poff;
        double* st = state + op.poff;
        const double* g = g_all + op.poff;
        double part = 0.0;
        for (int i = tid; i < 2 * D; i += blockDim.x) {
            const double gg = g[i];
            const double a = st[i] + gg * gg;
            st[i] = a;
            const double nv = p[i] - eta * gg / (sqrt(a) + eps);
            p[i] = nv;
            if (i < D) part += log(fabs(nv));
        }
        const double lc = block_sum(part);
        if (tid == 0) lconst_dev[o] = (flags & 1) ? 0.0 : lc;   // ENF_NEGLL_ZYGOTE_PRIMAL drops the ScaleShift ladj value
    } else {
        double* v = params + op.poff + size_t(k) * D;
        double* st = state + op.poff + size_t(k) * D;
        const double* g = g_all + op.poff + size_t(k) * D;
        double part = 0.0;
        for (int i = tid; i < D; i += blockDim.x) {
            const double gg = g[i];
            const double a = st[i] + gg * gg;
            st[i] = a;
            const double nv = v[i] - eta * gg / (sqrt(a) + eps);
            v[i] = nv;
            part += nv * nv;
        }
        const double inv = 1.0 / sqrt(block_sum(part));
        part = 0.0;
        for (int i = tid; i < D; i += blockDim.x) {
            const double nv = v[i] * inv;
            v[i] = nv;
            part += nv * nv;
        }
        const double n2 = block_sum(part);
        if (tid == 0) norms[op.noff + k] = n2;
    }
}

}  // namespace

bool moments_supported(int dtype, int D) { return dtype == 0 && (D == 64 || D == 128 || D == 256); }

size_t moments_partial_bytes(int D, int sm_count) {
    const int ND = D < 128 ? 128 : D;
    return size_t(sm_count) * ND * ND * sizeof(float) + size_t(sm_count) * 8 * ND * sizeof(double);
}

// d_part: moments_partial_bytes(D, sm_count) bytes of scratch; d_sums: (D+1)^2 + 1 doubles
cudaError_t launch_moments(int D, const void* x, int64_t N, void* d_part, double* d_sums, int sm_count, cudaStream_t st) {
    const int ND = D < 128 ? 128 : D;   // D = 64: the 128-row kernel; row groups 2, 3 lie outside the tensor map -> zero fill
    float* part_s = static_cast<float*>(d_part);
    double* part_m = reinterpret_cast<double*>(part_s + size_t(sm_count) * ND * ND);
    const int64_t total = (N + MO_KC - 1) / MO_KC;
    int n_cta = 0, reps = MO_SPLIT_THREADS / (ND / 4);
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
        if (ND == 256) ENF_MOMENTS_LAUNCH(256)
        else if (ND == 128) ENF_MOMENTS_LAUNCH(128)
        else return cudaErrorInvalidValue;
#undef ENF_MOMENTS_LAUNCH
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int n = D * D + D;
    moments_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part_s, part_m, n_cta, reps, D, ND, N, d_sums);
    return cudaGetLastError();
}


// sums -> out[0] = negll, out[1 .. 1+P) = gradients; kinds/Ks/poffs: the chain's ops in application order,
// d_params: packed float64 parameters on the device, d_norms: v.v of every reflection in application order
namespace {
MomChain make_mom_chain(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs) {
    MomChain mc;
    mc.n_ops = n_ops;
    mc.D = D;
    int noff = 0;
    for (int o = 0; o < n_ops; ++o) {
        mc.ops[o] = MomOp{kinds[o], Ks[o], poffs[o], noff};
        if (kinds[o] == OP_HH) noff += Ks[o];
    }
    return mc;
}
}  // namespace

cudaError_t launch_moments_update(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs, const double* d_out,
                                  double* d_params, double* d_norms, double* d_state, double eta, double eps, int flags,
                                  double* d_lconst, double* d_history, long long* d_step, cudaStream_t st) {
    const MomChain mc = make_mom_chain(D, n_ops, kinds, Ks, poffs);
    int units = 0;
    for (int o = 0; o < n_ops; ++o) units += kinds[o] == OP_HH ? Ks[o] : 1;
    moments_update_kernel<<<units, 256, 0, st>>>(mc, d_out, d_params, d_norms, d_state, eta, eps, flags, d_lconst, d_history, d_step);
    return cudaGetLastError();
}

size_t moments_chainrule_part_bytes(int D, int n_params) {
    return size_t((D + 1 + MC_NC - 1) / MC_NC) * size_t(1 + n_params) * sizeof(double);
}

// d_part: moments_chainrule_part_bytes() of scratch; lconst / d_lconst: the ladj row constant (host value, or the sum of
// the n_ops per-op device slots when d_lconst != nullptr)
cudaError_t launch_moments_chainrule(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs, int n_params,
                                     const double* d_params, const double* d_norms, const double* d_sums, double lconst,
                                     const double* d_lconst, double* d_part, double* d_out, cudaStream_t st) {
    const MomChain mc = make_mom_chain(D, n_ops, kinds, Ks, poffs);
    const int n_cta = (D + 1 + MC_NC - 1) / MC_NC;
    moments_chainrule_kernel<<<n_cta, D, 0, st>>>(mc, d_params, d_norms, d_sums, n_params, d_part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    moments_chainrule_reduce_kernel<<<(n_params + 1 + 255) / 256, 256, 0, st>>>(d_part, n_cta, n_params, D, lconst, d_lconst, n_ops, d_out);
    return cudaGetLastError();
}

}  // namespace enf
