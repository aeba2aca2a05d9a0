// F4 of SURVEY §2.3: deep Householder stacks at large D on the 5th-generation tensor cores.
//
// A chain that consists only of HouseholderTrafo and ScaleShiftTrafo ops is an affine
// map y = W x + c.  The host folds the whole chain (every reflection of
// src/householder_trafo.jl:71-78 in column order, every muladd of
// src/scale_shift_trafo.jl:16) into W (D x D) and c in float64; the device evaluates
//        Y^T [n x D] = X^T [n x D] . W^T [D x D] + c
// as one GEMM per 128-sample tile with tcgen05.mma (kind::tf32, accumulators in TMEM).
// The samples are the M dimension: a D x N column-major sample matrix IS the K-major A
// operand (sample-major, row index contiguous), so TMA feeds it without any transpose.
//
// Float32 accuracy from TF32 tensor cores (3xTF32): with x = xh + xl, W = Wh + Wl
// (h = top 19 bits, l = remainder) the kernel accumulates xh.Wh + xl.Wh + xh.Wl in the
// f32 accumulator (the dropped xl.Wl term is 2^-22 relative).  Wh / Wl are split on
// the host; xl is produced per K-chunk by the four worker warps between the TMA
// arrival and the MMA issue (layout preserving: same swizzled offset, other buffer).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2-5 = xh/xl split of every K chunk, warps 6-9 = epilogue
// (tcgen05.ld -> + c -> swizzled staging box -> TMA store; ladj = const).  Two TMEM
// accumulators, so the epilogue of tile t overlaps the MMAs of tile t+1.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "enf_chain.cuh"
#include "enf_launch.h"
#include "enf_tc.cuh"

namespace enf {
namespace {

constexpr int AF_TILE_M = 128;     // samples per tile (UMMA M)
#ifndef ENF_AF_SPLITTERS
#define ENF_AF_SPLITTERS 128
#endif
constexpr int AF_SPLITTERS = ENF_AF_SPLITTERS;               // threads that split x into tf32 hi / lo (warps 2 ..)
constexpr int AF_THREADS = 64 + AF_SPLITTERS + 128;          // warp 0 TMA, warp 1 MMA, split warps, 4 epilogue warps
constexpr int AF_EPI_WARPS = 4;

template <int ND, int KC>
struct AffineSmem {
    static constexpr int X_BYTES = AF_TILE_M * KC * 4;      // 16 KB at KC = 32
    static constexpr int W_BYTES = ND * KC * 4;             // 32 KB at ND = 256, KC = 32
    static constexpr int STAGES = (192 * 1024) / (2 * X_BYTES + 2 * W_BYTES) > 6 ? 6 : (192 * 1024) / (2 * X_BYTES + 2 * W_BYTES);
    static constexpr int STAGE_BYTES = 2 * X_BYTES + 2 * W_BYTES;
    static constexpr int OUT_BYTES = 32 * AF_KC * 4;        // one epilogue staging box: 32 rows x 32 cols
    static constexpr int OUT_OFF = STAGES * STAGE_BYTES;    // [epilogue warp][2] staging boxes
    static constexpr int BAR_OFF = OUT_OFF + AF_EPI_WARPS * 2 * OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;      // barriers + slack for 1024-byte alignment
    static constexpr uint32_t ACC_COLS = ND;                // TMEM columns per accumulator
    static constexpr uint32_t TMEM_COLS = 2 * ND < 32 ? 32 : 2 * ND;   // two accumulators (512 at ND = 256)
};

// One CTA per SM, persistent over 128-sample tiles.
template <int ND, int KC>
__global__ void __launch_bounds__(AF_THREADS, 1)
affine_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                   const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_y,
                   const float* __restrict__ bias, float* __restrict__ ladj, float ladj_const, int64_t N) {
    using S = AffineSmem<ND, KC>;
    constexpr int NKC = ND / KC;                        // K chunks per tile (K = D = ND)
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // TMA landed             (count 1 + tx)
    uint64_t* split = full + S::STAGES;                                // xh / xl written        (count 2 warps)
    uint64_t* empty = split + S::STAGES;                               // MMAs of the stage done (tcgen05.commit)
    uint64_t* acc_full = empty + S::STAGES;                            // [2] tile accumulated
    uint64_t* acc_empty = acc_full + 2;                                // [2] epilogue drained   (count 4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (N + AF_TILE_M - 1) / AF_TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], AF_SPLITTERS / 32);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], AF_EPI_WARPS);
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
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % S::STAGES;
                    if (it >= uint32_t(S::STAGES)) mbar_wait(&empty[s], ((it / S::STAGES) - 1) & 1);
                    unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                    mbar_expect_tx(&full[s], S::X_BYTES + 2 * S::W_BYTES);
                    tma_load_2d(st, &map_x, kc * KC, int(tile * AF_TILE_M), &full[s]);             // x chunk  [128 x 32]
                    tma_load_2d(st + 2 * S::X_BYTES, &map_wh, kc * KC, 0, &full[s]);               // Wh chunk [ND x 32]
                    tma_load_2d(st + 2 * S::X_BYTES + S::W_BYTES, &map_wl, kc * KC, 0, &full[s]);  // Wl chunk
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer.  ONE thread runs the whole loop with a single wait per chunk (split[s]: the splitters have seen the
        // TMA transaction of the stage, W chunks included, complete): the tensor core's instruction queue is shallow, so
        // every cycle the issuing thread spends in a second wait, a warp re-convergence or a commit is a cycle the tensor
        // pipe idles (tools/mma_rate.cu; the warp-wide form of this loop cost ~450 clk per chunk next to 1536 clk of MMAs) =====
        constexpr uint32_t idesc = make_idesc_tf32(AF_TILE_M, ND);
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t buf = tcount & 1;
                if (tcount >= 2) mbar_wait(&acc_empty[buf], ((tcount >> 1) - 1) & 1);   // epilogue of tile t-2 drained this buffer
                const uint32_t acc = tmem_base + buf * S::ACC_COLS;
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % S::STAGES;
                    mbar_wait(&split[s], (it / S::STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                    const uint64_t dxh = make_desc_kmajor<KC>(st), dxl = make_desc_kmajor<KC>(st + S::X_BYTES);
                    const uint64_t dwh = make_desc_kmajor<KC>(st + 2 * S::X_BYTES);
                    const uint64_t dwl = make_desc_kmajor<KC>(st + 2 * S::X_BYTES + S::W_BYTES);
#pragma unroll
                    for (int j = 0; j < KC / 8; ++j) {             // UMMA K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        umma_tf32(acc, dxh + adv, dwh + adv, idesc, (kc | j) != 0);
                        umma_tf32(acc, dxl + adv, dwh + adv, idesc, 1);
                        umma_tf32(acc, dxh + adv, dwl + adv, idesc, 1);
                    }
                    umma_commit(&empty[s]);                           // frees the stage when these MMAs retire
                    if (kc == NKC - 1) umma_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp < 2 + AF_SPLITTERS / 32) {
        // ===== splitters: x -> xh (round-to-nearest tf32, in place) and xl = x - xh (second buffer) =====
        const int wt = threadIdx.x - 64;                             // 0..63
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % S::STAGES;
                mbar_wait(&full[s], (it / S::STAGES) & 1);
                float4* xs = reinterpret_cast<float4*>(smem + size_t(s) * S::STAGE_BYTES);
                float4* xl = reinterpret_cast<float4*>(smem + size_t(s) * S::STAGE_BYTES + S::X_BYTES);
#pragma unroll 4
                for (int i = 0; i < S::X_BYTES / 16 / AF_SPLITTERS; ++i) {
                    const float4 v = xs[wt + i * AF_SPLITTERS];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    xs[wt + i * AF_SPLITTERS] = h;
                    // the remainder is rounded to tf32 here so that the tensor core's operand truncation is exact
                    xl[wt + i * AF_SPLITTERS] = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+ c) -> swizzled staging box -> TMA store =====
        const int quarter = warp & 3;                                // TMEM lane quarter this warp may access
        unsigned char* stage_out = smem + S::OUT_OFF + size_t(warp - 2 - AF_SPLITTERS / 32) * 2 * S::OUT_BYTES;
        uint32_t tcount = 0, nbox = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1;
            mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row0 = tile * AF_TILE_M + quarter * 32;   // this warp's 32 sample rows; lane = row in the box
            const uint32_t taddr = tmem_base + buf * S::ACC_COLS + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < ND / 32; ++c, ++nbox) {
                float v[32];
                tmem_ld32(taddr + uint32_t(c * 32), v);
                // the staging box used two stores ago must have been read by its TMA store
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float4* box = reinterpret_cast<float4*>(stage_out + (nbox & 1) * S::OUT_BYTES);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + j * 4);
                    // SWIZZLE_128B: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
                    box[lane * 8 + (j ^ (lane & 7))] =
                        make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_y, box, c * 32, int(row0));      // rows beyond N are clipped by the tensor map
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (ladj != nullptr && row0 + lane < N) __stcs(ladj + row0 + lane, ladj_const);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}


}  // namespace


bool affine_supported(int dtype, int D, const ChainDesc& d) {
    if (dtype != 0 || !(D == 64 || D == 128 || D == 256)) return false;
    int n_refl = 0;
    for (int o = 0; o < d.n_ops; ++o) {
        if (d.ops[o].kind != OP_HH && d.ops[o].kind != OP_SS) return false;
        if (d.ops[o].kind == OP_HH) n_refl += d.ops[o].K;
    }
    return n_refl >= 8;   // shallow stacks are HBM-bound on the SIMT kernel already
}

// Fold the chain into y = W x + c (float64), split W into tf32 hi / lo.  kinds/Ks/params: the chain's ops in
// application order; params packed like the C ABI (float64 copy).
void affine_fold(int D, int n_ops, const int* kinds, const int* Ks, const double* const* params, std::vector<float>& wh,
                 std::vector<float>& wl, std::vector<float>& bias) {
    std::vector<double> W(size_t(D) * D, 0.0), c(D, 0.0), t(D);
    for (int i = 0; i < D; ++i) W[size_t(i) * D + i] = 1.0;
    for (int o = 0; o < n_ops; ++o) {
        const double* p = params[o];
        if (kinds[o] == OP_SS) {
            for (int i = 0; i < D; ++i) {
                for (int j = 0; j < D; ++j) W[size_t(i) * D + j] *= p[i];
                c[i] = c[i] * p[i] + p[D + i];
            }
        } else {
            for (int k = 0; k < Ks[o]; ++k) {
                const double* v = p + size_t(k) * D;
                double n = 0.0;
                for (int i = 0; i < D; ++i) n += v[i] * v[i];
                const double s = 2.0 / n;
                for (int j = 0; j < D; ++j) {          // t = v^T W
                    double a = 0.0;
                    for (int i = 0; i < D; ++i) a += v[i] * W[size_t(i) * D + j];
                    t[j] = a * s;
                }
                for (int i = 0; i < D; ++i)
                    for (int j = 0; j < D; ++j) W[size_t(i) * D + j] -= v[i] * t[j];
                double a = 0.0;
                for (int i = 0; i < D; ++i) a += v[i] * c[i];
                a *= s;
                for (int i = 0; i < D; ++i) c[i] -= v[i] * a;
            }
        }
    }
    wh.resize(size_t(D) * D);
    wl.resize(size_t(D) * D);
    bias.resize(D);
    for (size_t i = 0; i < W.size(); ++i) {
        const float f = float(W[i]);
        uint32_t bits;
        std::memcpy(&bits, &f, 4);
        bits = (bits + 0x1000u) & 0xFFFFE000u;          // round to tf32 (10-bit mantissa)
        float h;
        std::memcpy(&h, &bits, 4);
        wh[i] = h;
        const float l = float(W[i] - double(h));
        std::memcpy(&bits, &l, 4);
        bits = (bits + 0x1000u) & 0xFFFFE000u;          // remainder rounded to tf32 as well (operand truncation becomes exact)
        std::memcpy(&wl[i], &bits, 4);
    }
    for (int i = 0; i < D; ++i) bias[i] = float(c[i]);
}

// d_w: device buffer holding Wh | Wl | bias (2 D^2 + D floats)
cudaError_t launch_affine(int D, const float* d_w, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                          int sm_count, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    CUtensorMap mx, mh, ml, my;
    if (!make_map(&mx, x, uint64_t(N), uint64_t(D), AF_TILE_M) || !make_map(&mh, d_w, uint64_t(D), uint64_t(D), uint32_t(D)) ||
        !make_map(&ml, d_w + size_t(D) * D, uint64_t(D), uint64_t(D), uint32_t(D)) ||
        !make_map(&my, y, uint64_t(N), uint64_t(D), 32))
        return cudaErrorInvalidValue;
    const float* bias = d_w + 2 * size_t(D) * D;
    float* lf = static_cast<float*>(ladj);
    const float lc = float(ladj_const);
    const int64_t tiles = (N + AF_TILE_M - 1) / AF_TILE_M;
    const unsigned grid = unsigned(tiles < sm_count ? tiles : sm_count);
    cudaError_t e = cudaSuccess;
#define ENF_AFFINE_LAUNCH(ND)                                                                                          \
    {                                                                                                                  \
        const int smem = AffineSmem<ND, AF_KC>::TOTAL;                                                                 \
        static bool set[64] = {};                                                                                      \
        int dev = 0;                                                                                                   \
        cudaGetDevice(&dev);                                                                                           \
        if (!set[dev & 63]) {                                                                                          \
            e = cudaFuncSetAttribute(affine_gemm_kernel<ND, AF_KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            if (e != cudaSuccess) return e;                                                                            \
            set[dev & 63] = true;                                                                                      \
        }                                                                                                              \
        affine_gemm_kernel<ND, AF_KC><<<grid, AF_THREADS, smem, st>>>(mx, mh, ml, my, bias, lf, lc, N);                \
    }
    if (D == 256) ENF_AFFINE_LAUNCH(256)
    else if (D == 128) ENF_AFFINE_LAUNCH(128)
    else if (D == 64) ENF_AFFINE_LAUNCH(64)
    else return cudaErrorInvalidValue;
#undef ENF_AFFINE_LAUNCH
    return cudaGetLastError();
}

}  // namespace enf
