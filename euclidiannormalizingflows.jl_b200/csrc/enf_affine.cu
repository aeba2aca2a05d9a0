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
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_tf32(AF_TILE_M, ND);
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1;
            if (tcount >= 2) mbar_wait(&acc_empty[buf], ((tcount >> 1) - 1) & 1);   // epilogue of tile t-2 drained this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + buf * S::ACC_COLS;
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % S::STAGES;
                const uint32_t ph = (it / S::STAGES) & 1;
                mbar_wait(&full[s], ph);
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
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
                __syncwarp();
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

// Same pipeline with the two correction products of the 3xTF32 scheme issued as kind::f16 (bf16) MMAs at twice the
// TF32 rate:  y = tf32(xh).tf32(Wh) + bf16(x - xh).bf16(W) + bf16(x).bf16(W - Wh).  The corrections are 2^-12 of the
// main term, so 8 mantissa bits on each of their factors keep the total error at ~2^-20: 4 TF32 + 4 BF16 MMAs per
// K chunk instead of 12 TF32 MMAs = 2/3 of the tensor time.  bf16 tiles are [rows][32] with 64-byte rows, SWIZZLE_64B.
template <int ND>
struct AffineBfSmem {
    static constexpr int XH_BYTES = AF_TILE_M * 32 * 4;     // 16 KB  tf32(xh), 128-byte rows (SWIZZLE_128B)
    static constexpr int XB_BYTES = AF_TILE_M * 32 * 2;     //  8 KB  bf16 tile, 64-byte rows (SWIZZLE_64B)
    static constexpr int WH_BYTES = ND * 32 * 4;
    static constexpr int WB_BYTES = ND * 32 * 2;
    static constexpr int XLB_OFF = XH_BYTES, XB_OFF = XH_BYTES + XB_BYTES, WH_OFF = XH_BYTES + 2 * XB_BYTES;
    static constexpr int WHB_OFF = WH_OFF + WH_BYTES, WLB_OFF = WHB_OFF + WB_BYTES;
    static constexpr int STAGE_BYTES = WLB_OFF + WB_BYTES;  // 96 KB at ND = 256
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 6 ? 6 : (192 * 1024) / STAGE_BYTES;
    static constexpr int OUT_BYTES = 32 * AF_KC * 4;
    static constexpr int OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = OUT_OFF + AF_EPI_WARPS * 2 * OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr uint32_t ACC_COLS = ND;
    static constexpr uint32_t TMEM_COLS = 2 * ND < 32 ? 32 : 2 * ND;
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {   // round-to-nearest-even, lo in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int ND>
__global__ void __launch_bounds__(AF_THREADS, 1)
affine_bf_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                 const __grid_constant__ CUtensorMap map_whb, const __grid_constant__ CUtensorMap map_wlb,
                 const __grid_constant__ CUtensorMap map_y,
                   const float* __restrict__ bias, float* __restrict__ ladj, float ladj_const, int64_t N) {
    using S = AffineBfSmem<ND>;
    constexpr int KC = 32;
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
                    mbar_expect_tx(&full[s], S::XH_BYTES + S::WH_BYTES + 2 * S::WB_BYTES);
                    tma_load_2d(st, &map_x, kc * KC, int(tile * AF_TILE_M), &full[s]);             // x chunk        [128 x 32] f32
                    tma_load_2d(st + S::WH_OFF, &map_wh, kc * KC, 0, &full[s]);                    // tf32(W) chunk  [ND x 32] f32
                    tma_load_2d(st + S::WHB_OFF, &map_whb, kc * KC, 0, &full[s]);                  // bf16(W) chunk  [ND x 32] bf16
                    tma_load_2d(st + S::WLB_OFF, &map_wlb, kc * KC, 0, &full[s]);                  // bf16(W - Wh) chunk
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_tf32(AF_TILE_M, ND);
        constexpr uint32_t idesc_b = make_idesc_bf16(AF_TILE_M, ND);
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1;
            if (tcount >= 2) mbar_wait(&acc_empty[buf], ((tcount >> 1) - 1) & 1);   // epilogue of tile t-2 drained this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + buf * S::ACC_COLS;
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % S::STAGES;
                const uint32_t ph = (it / S::STAGES) & 1;
                mbar_wait(&full[s], ph);
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                    const uint64_t dxh = make_desc_kmajor<32>(st), dwh = make_desc_kmajor<32>(st + S::WH_OFF);
                    // bf16 tiles: 64-byte rows, SWIZZLE_64B, 8-row groups 512 bytes apart (the byte layout of make_desc_kmajor<16>)
                    const uint64_t dxlb = make_desc_kmajor<16>(st + S::XLB_OFF), dxb = make_desc_kmajor<16>(st + S::XB_OFF);
                    const uint64_t dwhb = make_desc_kmajor<16>(st + S::WHB_OFF), dwlb = make_desc_kmajor<16>(st + S::WLB_OFF);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                  // UMMA K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        umma_tf32(acc, dxh + adv, dwh + adv, idesc, (kc | j) != 0);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                  // UMMA K = 16 bf16 = 32 bytes
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        umma_bf16(acc, dxlb + adv, dwhb + adv, idesc_b, 1);
                        umma_bf16(acc, dxb + adv, dwlb + adv, idesc_b, 1);
                    }
                    umma_commit(&empty[s]);                           // frees the stage when these MMAs retire
                    if (kc == NKC - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
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
                unsigned char* stg = smem + size_t(s) * S::STAGE_BYTES;
                float4* xs = reinterpret_cast<float4*>(stg);
#pragma unroll 4
                for (int i = 0; i < S::XH_BYTES / 16 / AF_SPLITTERS; ++i) {
                    const int idx = wt + i * AF_SPLITTERS;                 // 16-byte chunk of the swizzled f32 tile
                    const float4 v = xs[idx];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    xs[idx] = h;
                    // logical position of this chunk: row = sample, lc = logical 16-byte chunk (columns 4 lc .. 4 lc + 3)
                    const int row = idx >> 3, lc = (idx & 7) ^ (row & 7);
                    // bf16 tile: 64-byte rows, 16-byte chunk (lc / 2) lives at chunk ((lc / 2) ^ ((row / 2) & 3)), 8 bytes per 4 columns
                    const int boff = row * 64 + ((((lc >> 1) ^ ((row >> 1) & 3)) << 4) | ((lc & 1) << 3));
                    *reinterpret_cast<uint2*>(stg + S::XLB_OFF + boff) = make_uint2(pack_bf16(v.x - h.x, v.y - h.y), pack_bf16(v.z - h.z, v.w - h.w));
                    *reinterpret_cast<uint2*>(stg + S::XB_OFF + boff) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
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

// ======================================================================================
// 2-SM variant (cta_group::2): a CTA pair (cluster of 2) computes a 256-sample tile with
// ONE tcgen05.mma.cta_group::2 per K step (M = 256).  Each CTA holds its own 128 sample
// rows (A) and only HALF of the W chunk (B: N/2 rows), so the per-CTA L2->SM traffic for W
// and its shared-memory footprint are halved -> three pipeline stages fit instead of two.
// Barriers: x_full / empty / acc_full are per CTA (empty and acc_full are signalled in both
// CTAs by a multicast tcgen05.commit); w_full, split_done and acc_empty live in the leader
// CTA (rank 0) and are signalled remotely by the peer.
// ======================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ncluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address of a CTA pair -> rank 0
__device__ __forceinline__ void tma_load_2d_to_leader_bar(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    // data lands in THIS CTA's shared memory, the transaction bytes are counted on the LEADER CTA's mbarrier
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(uint16_t(3))
                 : "memory");
}

template <int ND>
struct Affine2Smem {
    static constexpr int STAGES = 3;
    static constexpr int X_BYTES = AF_TILE_M * AF_KC * 4;        // 16 KB (this CTA's 128 rows)
    static constexpr int W_BYTES = (ND / 2) * AF_KC * 4;         // this CTA's half of the W chunk (16 KB at ND = 256)
    static constexpr int STAGE_BYTES = 2 * X_BYTES + 2 * W_BYTES;
    static constexpr int OUT_BYTES = 32 * AF_KC * 4;
    static constexpr int OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = OUT_OFF + AF_EPI_WARPS * 2 * OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr uint32_t ACC_COLS = ND;
    static constexpr uint32_t TMEM_COLS = 2 * ND < 32 ? 32 : 2 * ND;
};

template <int ND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AF_THREADS, 1)
affine2_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                    const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_y,
                    const float* __restrict__ bias, float* __restrict__ ladj, float ladj_const, int64_t N) {
    using S = Affine2Smem<ND>;
    constexpr int NKC = ND / AF_KC;
    constexpr int NH = ND / 2;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);  // local : this CTA's x chunk landed
    uint64_t* w_full = x_full + S::STAGES;                              // leader: both W halves landed
    uint64_t* split = w_full + S::STAGES;                               // leader: xh/xl written in both CTAs (4 warps)
    uint64_t* empty = split + S::STAGES;                                // local : stage consumed (multicast commit)
    uint64_t* acc_full = empty + S::STAGES;                             // local [2]: tile accumulated (multicast commit)
    uint64_t* acc_empty = acc_full + 2;                                 // leader [2]: both epilogues drained (8 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t n_ptiles = (N + 2 * AF_TILE_M - 1) / (2 * AF_TILE_M);   // 256-sample pair tiles

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&w_full[s], 1);
            mbar_init(&split[s], 2 * AF_SPLITTERS / 32);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 2 * AF_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, S::TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();          // the peer's barriers exist before anything remote is signalled
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t pt = cluster_id_x(); pt < n_ptiles; pt += ncluster_id_x()) {
                const int row0 = int(pt * 2 * AF_TILE_M + rank * AF_TILE_M);
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % S::STAGES;
                    if (it >= uint32_t(S::STAGES)) mbar_wait(&empty[s], ((it / S::STAGES) - 1) & 1);
                    unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                    mbar_expect_tx(&x_full[s], S::X_BYTES);
                    tma_load_2d(st, &map_x, kc * AF_KC, row0, &x_full[s]);
                    if (rank == 0) mbar_expect_tx(&w_full[s], 4 * S::W_BYTES);     // Wh + Wl halves of both CTAs
                    tma_load_2d_to_leader_bar(st + 2 * S::X_BYTES, &map_wh, kc * AF_KC, int(rank) * NH, &w_full[s]);
                    tma_load_2d_to_leader_bar(st + 2 * S::X_BYTES + S::W_BYTES, &map_wl, kc * AF_KC, int(rank) * NH, &w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only; one lane issues for the pair) =====
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * AF_TILE_M, ND);
            uint32_t it = 0, tcount = 0;
            for (int64_t pt = cluster_id_x(); pt < n_ptiles; pt += ncluster_id_x(), ++tcount) {
                const uint32_t buf = tcount & 1;
                if (tcount >= 2) mbar_wait(&acc_empty[buf], ((tcount >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + buf * S::ACC_COLS;
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % S::STAGES;
                    const uint32_t ph = (it / S::STAGES) & 1;
                    mbar_wait(&w_full[s], ph);
                    mbar_wait(&split[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        unsigned char* st = smem + size_t(s) * S::STAGE_BYTES;
                        const uint64_t dxh = make_desc_sw128(st), dxl = make_desc_sw128(st + S::X_BYTES);
                        const uint64_t dwh = make_desc_sw128(st + 2 * S::X_BYTES);
                        const uint64_t dwl = make_desc_sw128(st + 2 * S::X_BYTES + S::W_BYTES);
#pragma unroll
                        for (int j = 0; j < AF_KC / 8; ++j) {
                            const uint64_t adv = uint64_t((j * 32) >> 4);
                            umma2_tf32(acc, dxh + adv, dwh + adv, idesc, (kc | j) != 0);
                            umma2_tf32(acc, dxl + adv, dwh + adv, idesc, 1);
                            umma2_tf32(acc, dxh + adv, dwl + adv, idesc, 1);
                        }
                        umma2_commit_both(&empty[s]);
                        if (kc == NKC - 1) umma2_commit_both(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 2 + AF_SPLITTERS / 32) {
        // ===== splitters (both CTAs): local x chunk -> xh (in place) and xl; signal the leader =====
        const int wt = threadIdx.x - 64;
        uint32_t it = 0;
        for (int64_t pt = cluster_id_x(); pt < n_ptiles; pt += ncluster_id_x()) {
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % S::STAGES;
                mbar_wait(&x_full[s], (it / S::STAGES) & 1);
                float4* xs = reinterpret_cast<float4*>(smem + size_t(s) * S::STAGE_BYTES);
                float4* xl = reinterpret_cast<float4*>(smem + size_t(s) * S::STAGE_BYTES + S::X_BYTES);
#pragma unroll 4
                for (int i = 0; i < S::X_BYTES / 16 / AF_SPLITTERS; ++i) {
                    const float4 v = xs[wt + i * AF_SPLITTERS];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    xs[wt + i * AF_SPLITTERS] = h;
                    xl[wt + i * AF_SPLITTERS] = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&split[s], 0);
            }
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows of the accumulator =====
        const int quarter = warp & 3;
        unsigned char* stage_out = smem + S::OUT_OFF + size_t(warp - 2 - AF_SPLITTERS / 32) * 2 * S::OUT_BYTES;
        uint32_t tcount = 0, nbox = 0;
        for (int64_t pt = cluster_id_x(); pt < n_ptiles; pt += ncluster_id_x(), ++tcount) {
            const uint32_t buf = tcount & 1;
            mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row0 = pt * 2 * AF_TILE_M + rank * AF_TILE_M + quarter * 32;
            const uint32_t taddr = tmem_base + buf * S::ACC_COLS + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < ND / 32; ++c, ++nbox) {
                float v[32];
                tmem_ld32(taddr + uint32_t(c * 32), v);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float4* box = reinterpret_cast<float4*>(stage_out + (nbox & 1) * S::OUT_BYTES);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + j * 4);
                    box[lane * 8 + (j ^ (lane & 7))] =
                        make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (row0 < N) tma_store_2d(&map_y, box, c * 32, int(row0));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (ladj != nullptr && row0 + lane < N) __stcs(ladj + row0 + lane, ladj_const);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&acc_empty[buf], 0);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();          // nobody leaves while the peer may still signal or read this CTA
    if (warp == 1) tmem_dealloc2(tmem_base, S::TMEM_COLS);
}

}  // namespace

namespace {
// row-major [rows][cols] bf16 matrix, box [box_rows][32 cols] (64-byte rows), 64-byte swizzle
bool make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
uint16_t to_bf16(double v) {   // round-to-nearest-even
    const float f = float(v);
    uint32_t b;
    std::memcpy(&b, &f, 4);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return uint16_t(b >> 16);
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
                 std::vector<float>& wl, std::vector<float>& bias, std::vector<uint16_t>& wb) {
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
    // bf16 operands of the correction products (affine_bf_kernel): bf16(W) | bf16(W - tf32(W))
    wb.resize(2 * W.size());
    for (size_t i = 0; i < W.size(); ++i) {
        wb[i] = to_bf16(W[i]);
        wb[W.size() + i] = to_bf16(W[i] - double(wh[i]));
    }
}

// d_w: device buffer holding Wh | Wl | bias (2 D^2 + D floats) | bf16(W) | bf16(W - Wh) (2 D^2 bf16)
cudaError_t launch_affine(int D, const float* d_w, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                          int sm_count, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    // cta_group::2 pairs measured slower than independent CTAs in round 1 (profiles/README.md): opt-in
    static const bool two_sm = getenv("ENF_AFFINE_2SM") != nullptr;
    static const int kc_env = getenv("ENF_AFFINE_KC") ? atoi(getenv("ENF_AFFINE_KC")) : 0;
    const bool use2 = two_sm && D >= 128;
    const uint32_t kc = use2 ? 32u : (kc_env == 16 || kc_env == 32) ? uint32_t(kc_env) : 32u;
    CUtensorMap mx, mh, ml, my;
    const uint32_t wbox = use2 ? uint32_t(D / 2) : uint32_t(D);
    if (!make_map(&mx, x, uint64_t(N), uint64_t(D), AF_TILE_M, kc) || !make_map(&mh, d_w, uint64_t(D), uint64_t(D), wbox, kc) ||
        !make_map(&ml, d_w + size_t(D) * D, uint64_t(D), uint64_t(D), wbox, kc) ||
        !make_map(&my, y, uint64_t(N), uint64_t(D), 32))
        return cudaErrorInvalidValue;
    const float* bias = d_w + 2 * size_t(D) * D;
    float* lf = static_cast<float*>(ladj);
    const float lc = float(ladj_const);
    const int64_t tiles = (N + AF_TILE_M - 1) / AF_TILE_M;
    const unsigned grid = unsigned(tiles < sm_count ? tiles : sm_count);
    cudaError_t e = cudaSuccess;
    static const bool bf_corr = getenv("ENF_AFFINE_BF16") != nullptr;
    if (bf_corr && !use2 && kc == 32) {
        const uint16_t* wbf = reinterpret_cast<const uint16_t*>(d_w + 2 * size_t(D) * D + D);
        CUtensorMap mhb, mlb;
        if (!make_map_bf16(&mhb, wbf, uint64_t(D), uint64_t(D), uint32_t(D)) ||
            !make_map_bf16(&mlb, wbf + size_t(D) * D, uint64_t(D), uint64_t(D), uint32_t(D)))
            return cudaErrorInvalidValue;
#define ENF_AFFINE_BF_LAUNCH(ND)                                                                                       \
    {                                                                                                                  \
        const int smem = AffineBfSmem<ND>::TOTAL;                                                                      \
        static bool set[64] = {};                                                                                      \
        int dev = 0;                                                                                                   \
        cudaGetDevice(&dev);                                                                                           \
        if (!set[dev & 63]) {                                                                                          \
            e = cudaFuncSetAttribute(affine_bf_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);         \
            if (e != cudaSuccess) return e;                                                                            \
            set[dev & 63] = true;                                                                                      \
        }                                                                                                              \
        affine_bf_kernel<ND><<<grid, AF_THREADS, smem, st>>>(mx, mh, mhb, mlb, my, bias, lf, lc, N);                   \
    }
        if (D == 256) ENF_AFFINE_BF_LAUNCH(256)
        else if (D == 128) ENF_AFFINE_BF_LAUNCH(128)
        else ENF_AFFINE_BF_LAUNCH(64)
#undef ENF_AFFINE_BF_LAUNCH
        return cudaGetLastError();
    }
#define ENF_AFFINE_LAUNCH_KC(ND, KC)                                                                                   \
    {                                                                                                                  \
        const int smem = AffineSmem<ND, KC>::TOTAL;                                                                    \
        static bool set[64] = {};                                                                                      \
        int dev = 0;                                                                                                   \
        cudaGetDevice(&dev);                                                                                           \
        if (!set[dev & 63]) {                                                                                          \
            e = cudaFuncSetAttribute(affine_gemm_kernel<ND, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   \
            if (e != cudaSuccess) return e;                                                                            \
            set[dev & 63] = true;                                                                                      \
        }                                                                                                              \
        affine_gemm_kernel<ND, KC><<<grid, AF_THREADS, smem, st>>>(mx, mh, ml, my, bias, lf, lc, N);                   \
    }
#define ENF_AFFINE_LAUNCH(ND)                                                                                          \
    {                                                                                                                  \
        if (kc == 16) ENF_AFFINE_LAUNCH_KC(ND, 16) else ENF_AFFINE_LAUNCH_KC(ND, 32)                                   \
    }
    if (use2) {
        const int64_t ptiles = (N + 2 * AF_TILE_M - 1) / (2 * AF_TILE_M);
        const int max_pairs = sm_count / 2;
        const unsigned grid2 = 2u * unsigned(ptiles < max_pairs ? ptiles : max_pairs);
#define ENF_AFFINE2_LAUNCH(ND)                                                                                         \
    {                                                                                                                  \
        const int smem = Affine2Smem<ND>::TOTAL;                                                                       \
        static bool set[64] = {};                                                                                      \
        int dev = 0;                                                                                                   \
        cudaGetDevice(&dev);                                                                                           \
        if (!set[dev & 63]) {                                                                                          \
            e = cudaFuncSetAttribute(affine2_gemm_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
            if (e != cudaSuccess) return e;                                                                            \
            set[dev & 63] = true;                                                                                      \
        }                                                                                                              \
        affine2_gemm_kernel<ND><<<grid2, AF_THREADS, smem, st>>>(mx, mh, ml, my, bias, lf, lc, N);                     \
    }
        if (D == 256) ENF_AFFINE2_LAUNCH(256)
        else ENF_AFFINE2_LAUNCH(128)
#undef ENF_AFFINE2_LAUNCH
        return cudaGetLastError();
    }
    if (D == 256) ENF_AFFINE_LAUNCH(256)
    else if (D == 128) ENF_AFFINE_LAUNCH(128)
    else if (D == 64) ENF_AFFINE_LAUNCH(64)
    else return cudaErrorInvalidValue;
#undef ENF_AFFINE_LAUNCH
#undef ENF_AFFINE_LAUNCH_KC
    return cudaGetLastError();
}

}  // namespace enf
