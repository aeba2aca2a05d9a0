// F4 of SURVEY §2.3 in the form BASELINE.json's north star names: deep Householder stacks at large D applied in
// compact-WY form as tcgen05 / TMEM tensor-core GEMMs.
//
// A chain of HouseholderTrafo and ScaleShiftTrafo ops is "diagonal + low rank":  with the pre-scaled vectors
// v' = v sqrt(2 / v.v) every reflection (src/householder_trafo.jl:4-11) is I - v' v'^T, and the whole chain folds into
//        y = alpha . x  -  U (W^T x)  +  c ,        U, W : D x Kt  (Kt = number of reflections),
// on the host in float64 (wy_fold: one column of U and W per reflection, ScaleShift scales alpha, U and c).  Per sample
// that is 4 D Kt flops instead of the 2 D^2 of the dense fold y = W x + c (enf_affine.cu): at D = 256, Kt = 64 half
// the tensor work, and the dense kernel is tensor-bound.  Chains with Kt > 64 (or Kt > D/2) keep the dense kernel.
//
// Per 128-sample tile, TWO chained GEMMs on the 5th-generation tensor cores (kind::tf32, 3xTF32 for Float32 accuracy):
//   GEMM1   T[128 x 64]  = X[128 x D] . W[D x 64]       A = sample tile from shared memory (the column-major D x N sample
//                                                        matrix IS the K-major operand), B = W^T chunks, D = TMEM
//   hand-off T -> (Thi, Tlo): four warps read the accumulator (tcgen05.ld), split it into tf32 high / low parts and write
//                             them back to TENSOR MEMORY (tcgen05.st) -- GEMM2 takes its A operand from TMEM, so the
//                             128 x 64 intermediate never touches shared memory
//   GEMM2   V[128 x D]   = T[128 x 64] . U^T[64 x D]    A = Thi / Tlo in TMEM, B = U (resident in shared memory), D = TMEM
//   epilogue y = alpha . x - V + c: TMEM hands a lane one sample, global memory wants a lane to own columns: -V is
//            transposed through a 4 KB shared-memory box per warp, the sample tile is re-read from L2 (it was streamed
//            through the ring and is gone from shared memory) with coalesced loads, y is stored with coalesced stores.
// GEMM1 and GEMM2 are issued by two different warps, so that neither waits behind the other's barriers; T is double
// buffered, so GEMM1 runs up to two tiles ahead of GEMM2 and the hand-off / epilogue overlap tensor work.
//
// Shared memory at D = 256: 3-stage ring of 32-column chunks (x, xl, Wh, Wl: 48 KB per stage, 96 KB of TMA loads in
// flight) + two 32 KB buffers through which the four pieces of U (hi / lo x two k-halves) stream per tile + 4 staging
// boxes = 224 KB (a resident U would leave room for only 72 KB of ring: measured, the ring then starves).  TMEM: two T buffers of (main | correction) accumulators (2 x 128 columns; the hand-off
// rewrites a buffer in place as Thi | Tlo) | V (D columns) = 512.
// Warp roles (352 threads): warp 0 TMA producer (x, W), warp 1 TMEM owner + MMA issuer, warps 2-5 xh/xl split of every
// chunk, warps 6-9 hand-off + epilogue (one TMEM lane quarter each), warp 10 TMA producer of the U pieces.
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

constexpr int WY_TILE_M = 128;
constexpr int WY_KC = 32;            // columns per ring chunk: one 128-byte swizzle atom per row
constexpr int WY_KT = 64;            // reflections (padded with zero columns)
constexpr int WY_STAGES = 3;
constexpr int WY_SPLITTERS = 128;
constexpr int WY_THREADS = 64 + WY_SPLITTERS + 128 + 64;   // + warp 10: producer of the U pieces, warp 11: GEMM2 issuer
constexpr int WY_EPI_WARPS = 4;

template <int ND>
struct WySmem {
    static constexpr int X_BYTES = WY_TILE_M * WY_KC * 4;          // 16 KB
    static constexpr int W_BYTES = WY_KT * WY_KC * 4;              // 8 KB
    static constexpr int STAGE_BYTES = 2 * X_BYTES + 2 * W_BYTES;  // xh | xl | Wh | Wl = 48 KB
    static constexpr int U_CHUNK_BYTES = ND * 32 * 4;              // one piece of U: [ND rows x 32 k], 128B swizzle (32 KB at ND = 256)
    static constexpr int U_BYTES = 2 * U_CHUNK_BYTES;              // two piece buffers; the four pieces (Ul k0, Ul k1, Uh k0, Uh k1) of a
    static constexpr int RING_OFF = U_BYTES;                       // tile stream through them
    static constexpr int OUT_BYTES = 32 * 32 * 4;
    static constexpr int OUT_OFF = RING_OFF + WY_STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = OUT_OFF + WY_EPI_WARPS * OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    // tensor memory: two T buffers of (main | correction) accumulators, 64 columns each, then V.  The hand-off rewrites a
    // buffer in place as (Thi | Tlo).
    static constexpr uint32_t T_COL = 0, T_BUF_COLS = 128, TC_OFF = 64, V_COL = 256;
    static constexpr uint32_t TMEM_COLS = 512;
};

// D[tmem] (+)= A[tmem] . B[smem]   (A operand from tensor memory: lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int ND>
__global__ void __launch_bounds__(WY_THREADS, 1)
wy_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
               const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_uh,
               const __grid_constant__ CUtensorMap map_ul,
               const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ alpha,
               const float* __restrict__ cvec, float* __restrict__ ladj, float ladj_const, int64_t N) {
    using S = WySmem<ND>;
    constexpr int NKC = ND / WY_KC;                    // ring chunks per tile
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // TMA landed                     (1 + tx)
    uint64_t* split = full + WY_STAGES;                                // xh / xl written                (4 warps)
    uint64_t* empty = split + WY_STAGES;                               // MMAs of the stage retired      (tcgen05.commit)
    uint64_t* u_full = empty + WY_STAGES;                              // [2] a piece of U landed        (1 + tx)
    uint64_t* u_empty = u_full + 2;                                    // [2] its MMAs retired           (tcgen05.commit)
    uint64_t* t_full = u_empty + 2;                                    // [2] GEMM1 of a tile retired    (tcgen05.commit)
    uint64_t* t_split = t_full + 2;                                    // [2] Thi / Tlo written          (4 warps)
    uint64_t* t_free = t_split + 2;                                    // [2] GEMM2 has read Thi / Tlo   (tcgen05.commit)
    uint64_t* v_full = t_free + 2;                                     // GEMM2 of a tile retired        (tcgen05.commit)
    uint64_t* v_empty = v_full + 1;                                    // epilogue has drained V         (4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (N + WY_TILE_M - 1) / WY_TILE_M;
    const int my_tiles = blockIdx.x < n_tiles ? int((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WY_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], WY_SPLITTERS / 32);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&u_full[b], 1);
            mbar_init(&u_empty[b], 1);
            mbar_init(&t_full[b], 1);
            mbar_init(&t_split[b], WY_EPI_WARPS);
            mbar_init(&t_free[b], 1);
        }
        mbar_init(v_full, 1);
        mbar_init(v_empty, WY_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: x and W chunks =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int64_t tile = int64_t(blockIdx.x) + int64_t(i) * gridDim.x;
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % WY_STAGES;
                    if (it >= uint32_t(WY_STAGES)) mbar_wait(&empty[s], ((it / WY_STAGES) - 1) & 1);
                    unsigned char* st = smem + S::RING_OFF + size_t(s) * S::STAGE_BYTES;
                    mbar_expect_tx(&full[s], S::X_BYTES + 2 * S::W_BYTES);
                    tma_load_2d(st, &map_x, kc * WY_KC, int(tile * WY_TILE_M), &full[s]);              // x chunk  [128 x 32]
                    tma_load_2d(st + 2 * S::X_BYTES, &map_wh, kc * WY_KC, 0, &full[s]);                // Wh chunk [64 x 32]
                    tma_load_2d(st + 2 * S::X_BYTES + S::W_BYTES, &map_wl, kc * WY_KC, 0, &full[s]);   // Wl chunk
                }
            }
        }
    } else if (warp == 1) {
        // ===== GEMM1 issuer: T(i) = X(i) W into a (main | correction) accumulator pair =====
        constexpr uint32_t idesc1 = make_idesc_tf32(WY_TILE_M, WY_KT);
        uint32_t it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const uint32_t buf = uint32_t(i) & 1u;
            if (i >= 2) mbar_wait(&t_free[buf], uint32_t((i >> 1) - 1) & 1u);          // GEMM2 of tile i-2 has read this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tM = tmem_base + S::T_COL + buf * S::T_BUF_COLS, tC = tM + S::TC_OFF;
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % WY_STAGES;
                const uint32_t ph = (it / WY_STAGES) & 1;
                mbar_wait(&full[s], ph);
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    unsigned char* st = smem + S::RING_OFF + size_t(s) * S::STAGE_BYTES;
                    const uint64_t dxh = make_desc_kmajor<WY_KC>(st), dxl = make_desc_kmajor<WY_KC>(st + S::X_BYTES);
                    const uint64_t dwh = make_desc_kmajor<WY_KC>(st + 2 * S::X_BYTES);
                    const uint64_t dwl = make_desc_kmajor<WY_KC>(st + 2 * S::X_BYTES + S::W_BYTES);
#pragma unroll
                    for (int j = 0; j < WY_KC / 8; ++j) {          // UMMA K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        // The tensor core truncates when it adds into the f32 accumulator, one ulp of the ACCUMULATOR per MMA
                        // whatever the size of the addend.  The two correction products (2^-11 of the main one) get an
                        // accumulator of their own: the main one then sees 32 instead of 96 truncating additions per tile.
                        umma_tf32(tM, dxh + adv, dwh + adv, idesc1, (kc | j) != 0);
                        umma_tf32(tC, dxl + adv, dwh + adv, idesc1, (kc | j) != 0);
                        umma_tf32(tC, dxh + adv, dwl + adv, idesc1, 1);
                    }
                    umma_commit(&empty[s]);
                    if (kc == NKC - 1) umma_commit(&t_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 11) {
        // ===== GEMM2 issuer: V(j) = T(j) U^T, A = Thi | Tlo in tensor memory, B = the streamed pieces of U =====
        constexpr uint32_t idesc2 = make_idesc_tf32(WY_TILE_M, ND);
        for (int j = 0; j < my_tiles; ++j) {
            const uint32_t buf = uint32_t(j) & 1u;
            mbar_wait(&t_split[buf], uint32_t(j >> 1) & 1u);                           // Thi / Tlo of tile j are in TMEM
            if (j >= 1) mbar_wait(v_empty, uint32_t(j - 1) & 1u);                      // epilogue of tile j-1 has drained V
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tV = tmem_base + S::V_COL;
            const uint32_t tHI = tmem_base + S::T_COL + buf * S::T_BUF_COLS, tLO = tHI + S::TC_OFF;
            // pieces in the order Ul k0, Ul k1 (corrections Thi . Ul first: while the accumulator is still 2^-11 small their
            // truncation errors are too), then Uh k0, Uh k1 (correction Tlo . Uh, then the main product Thi . Uh)
#pragma unroll 1
            for (int pc = 0; pc < 4; ++pc) {
                const int q = 4 * j + pc, b = q & 1;
                mbar_wait(&u_full[b], uint32_t(q >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t du = make_desc_kmajor<32>(smem + b * S::U_CHUNK_BYTES);
                    const uint32_t k0 = uint32_t(pc & 1) * 32u;                        // first T column of this piece
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t adv = uint64_t((kk * 32) >> 4);
                        if (pc < 2) {
                            umma_tf32_ts(tV, tHI + k0 + uint32_t(kk * 8), du + adv, idesc2, (pc | kk) != 0);
                        } else {
                            umma_tf32_ts(tV, tLO + k0 + uint32_t(kk * 8), du + adv, idesc2, 1);
                            umma_tf32_ts(tV, tHI + k0 + uint32_t(kk * 8), du + adv, idesc2, 1);
                        }
                    }
                    umma_commit(&u_empty[b]);
                    if (pc == 3) {
                        umma_commit(v_full);
                        umma_commit(&t_free[buf]);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 10) {
        // ===== producer of the U pieces: per tile Ul k0, Ul k1, Uh k0, Uh k1 through two buffers =====
        if (lane == 0) {
            for (int q = 0; q < 4 * my_tiles; ++q) {
                const int b = q & 1, pc = q & 3;
                if (q >= 2) mbar_wait(&u_empty[b], uint32_t((q >> 1) - 1) & 1u);
                mbar_expect_tx(&u_full[b], S::U_CHUNK_BYTES);
                tma_load_2d(smem + b * S::U_CHUNK_BYTES, pc < 2 ? &map_ul : &map_uh, (pc & 1) * 32, 0, &u_full[b]);
            }
        }
    } else if (warp < 2 + WY_SPLITTERS / 32) {
        // ===== splitters: x -> xh (round-to-nearest tf32, in place) and xl = x - xh (second buffer) =====
        const int wt = threadIdx.x - 64;
        uint32_t it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % WY_STAGES;
                mbar_wait(&full[s], (it / WY_STAGES) & 1);
                float4* xs = reinterpret_cast<float4*>(smem + S::RING_OFF + size_t(s) * S::STAGE_BYTES);
                float4* xl = reinterpret_cast<float4*>(smem + S::RING_OFF + size_t(s) * S::STAGE_BYTES + S::X_BYTES);
#pragma unroll
                for (int q = 0; q < S::X_BYTES / 16 / WY_SPLITTERS; ++q) {
                    const float4 v = xs[wt + q * WY_SPLITTERS];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    xs[wt + q * WY_SPLITTERS] = h;
                    xl[wt + q * WY_SPLITTERS] = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else {
        // ===== warps 6-9: hand-off (T -> Thi | Tlo, in place in tensor memory) and epilogue (y = alpha x + c - V), whichever
        // is ready first: the hand-off of tile j+1 must not wait behind the epilogue of tile j, nor the other way round =====
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int ew = warp - 2 - WY_SPLITTERS / 32;
        float4* box = reinterpret_cast<float4*>(smem + S::OUT_OFF + size_t(ew) * S::OUT_BYTES);
        const uint32_t lane_off = uint32_t(quarter * 32) << 16;
        const int rq = lane >> 3, cq = lane & 7;                       // epilogue phase 2: rows rq + 4 i, 16-byte chunk cq
        int jh = 0, je = 0;                                            // next tile to hand off / to drain
        while (je < my_tiles) {
            int what = 0;                                              // 1: hand-off, 2: epilogue
            if (lane == 0) {
                for (;;) {
                    if (jh < my_tiles && jh <= je + 1 && mbar_try_wait(&t_full[jh & 1], uint32_t(jh >> 1) & 1u)) { what = 1; break; }
                    if (je < jh && mbar_try_wait(v_full, uint32_t(je) & 1u)) { what = 2; break; }
                }
            }
            what = __shfl_sync(0xffffffffu, what, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (what == 1) {
                // ---- hand-off: this warp's 32 rows of T(jh): main + correction, split into tf32 high / low parts
                const uint32_t buf = uint32_t(jh) & 1u;
#pragma unroll 1
                for (int h = 0; h < WY_KT / 32; ++h) {
                    const uint32_t tM = tmem_base + S::T_COL + buf * S::T_BUF_COLS + uint32_t(h * 32) + lane_off, tC = tM + S::TC_OFF;
                    float v[32], w[32];
                    tmem_ld32(tM, v);
                    tmem_ld32(tC, w);
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float t = v[e] + w[e];
                        const float hh = tf32_hi(t);
                        hi[e] = __float_as_uint(hh);
                        lo[e] = __float_as_uint(tf32_hi(t - hh));
                    }
                    tmem_st32(tM, hi);                                 // in place: (main | correction) -> (Thi | Tlo)
                    tmem_st32(tC, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_split[buf]);
                ++jh;
                continue;
            }
            // ---- epilogue of tile je.  TMEM hands a lane one SAMPLE (32 columns of it); global memory wants a lane to own
            // COLUMNS.  A 4 KB shared-memory box does the transpose: -V goes in by rows, then every lane adds alpha . x + c to
            // the 16-byte chunks of ITS column group (8 lanes cover one 128-byte row segment: coalesced re-read of the tile
            // from L2 and coalesced stores of y, 4 rows per instruction).
            const int64_t tile = int64_t(blockIdx.x) + int64_t(je) * gridDim.x;
            const int64_t row0 = tile * WY_TILE_M + quarter * 32;      // this warp's 32 samples
            const float* xbase = x + (row0 + rq) * int64_t(ND) + cq * 4;
            float* ybase = y + (row0 + rq) * int64_t(ND) + cq * 4;
            auto load_x = [&](int c, float4 (&xr)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    xr[i] = (row0 + rq + 4 * i < N) ? __ldcg(reinterpret_cast<const float4*>(xbase + int64_t(4 * i) * ND + c * 32))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            constexpr int PF = 3;                                      // boxes of x in flight (L2 latency >> time per box)
            float4 xr[PF][8];
#pragma unroll
            for (int c = 0; c < PF; ++c) load_x(c, xr[c]);
#pragma unroll
            for (int c = 0; c < ND / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + S::V_COL + uint32_t(c * 32) + lane_off, v);
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(alpha + c * 32 + cq * 4));
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cvec + c * 32 + cq * 4));
                // phase 1 (lane = sample row): -V into the box, 16-byte chunk q of row r at chunk (q ^ (r & 7)): conflict-free both ways
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    box[lane * 8 + (q ^ (lane & 7))] = make_float4(-v[4 * q], -v[4 * q + 1], -v[4 * q + 2], -v[4 * q + 3]);
                __syncwarp();
                // phase 2 (lane = column group)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rq + 4 * i;
                    const float4 b = box[r * 8 + (cq ^ (r & 7))];
                    const float4 xv = xr[c % PF][i];
                    const float4 o = make_float4(fmaf(a4.x, xv.x, c4.x) + b.x, fmaf(a4.y, xv.y, c4.y) + b.y,
                                                 fmaf(a4.z, xv.z, c4.z) + b.z, fmaf(a4.w, xv.w, c4.w) + b.w);
                    if (row0 + r < N) __stcs(reinterpret_cast<float4*>(ybase + int64_t(4 * i) * ND + c * 32), o);
                }
                if (c + PF < ND / 32) load_x(c + PF, xr[c % PF]);
                __syncwarp();
            }
            if (ladj != nullptr && row0 + lane < N) __stcs(ladj + row0 + lane, ladj_const);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(v_empty);
            ++je;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

void split_tf32(double v, float& h, float& l) {
    const float f = float(v);
    uint32_t bits;
    std::memcpy(&bits, &f, 4);
    bits = (bits + 0x1000u) & 0xFFFFE000u;          // round to tf32 (10-bit mantissa)
    std::memcpy(&h, &bits, 4);
    const float r = float(v - double(h));
    std::memcpy(&bits, &r, 4);
    bits = (bits + 0x1000u) & 0xFFFFE000u;          // remainder rounded to tf32 as well (operand truncation becomes exact)
    std::memcpy(&l, &bits, 4);
}

}  // namespace

// total number of reflections of a Householder/ScaleShift-only chain if the compact-WY kernel applies, else 0
int wy_rank(int dtype, int D, const ChainDesc& d) {
    if (dtype != 0 || !(D == 128 || D == 256)) return 0;
    int n_refl = 0;
    for (int o = 0; o < d.n_ops; ++o) {
        if (d.ops[o].kind != OP_HH && d.ops[o].kind != OP_SS) return 0;
        if (d.ops[o].kind == OP_HH) n_refl += d.ops[o].K;
    }
    // below 8 reflections the SIMT kernel is HBM-bound already; above D/4 + the dense fold needs no more tensor work
    return (n_refl >= 8 && n_refl <= WY_KT && 2 * n_refl <= D) ? n_refl : 0;
}

size_t wy_buffer_floats(int D) { return size_t(4) * WY_KT * D + 2 * size_t(D); }

// Fold the chain into y = alpha . x - U (W^T x) + c in float64 (one column of U, W per reflection) and lay the operands
// out for the kernel: Wt hi | Wt lo ([64][D]) | U hi | U lo ([D][64]) | alpha [D] | c [D].
void wy_fold(int D, int n_ops, const int* kinds, const int* Ks, const double* const* params, std::vector<float>& out) {
    std::vector<double> alpha(D, 1.0), c(D, 0.0), U, W;     // U, W: column-major D x kt
    int kt = 0;
    std::vector<double> vp(D), t;
    for (int o = 0; o < n_ops; ++o) {
        const double* p = params[o];
        if (kinds[o] == OP_SS) {
            for (int i = 0; i < D; ++i) {
                alpha[i] *= p[i];
                c[i] = c[i] * p[i] + p[D + i];
                for (int k = 0; k < kt; ++k) U[size_t(k) * D + i] *= p[i];
            }
            continue;
        }
        for (int r = 0; r < Ks[o]; ++r) {
            const double* v = p + size_t(r) * D;
            double n = 0.0;
            for (int i = 0; i < D; ++i) n += v[i] * v[i];
            const double sc = std::sqrt(2.0 / n);
            for (int i = 0; i < D; ++i) vp[i] = v[i] * sc;
            // U <- [U - v'(v'^T U), v'],  W <- [W, alpha . v'],  c <- c - v'(v'^T c)
            for (int k = 0; k < kt; ++k) {
                double a = 0.0;
                for (int i = 0; i < D; ++i) a += vp[i] * U[size_t(k) * D + i];
                for (int i = 0; i < D; ++i) U[size_t(k) * D + i] -= vp[i] * a;
            }
            double a = 0.0;
            for (int i = 0; i < D; ++i) a += vp[i] * c[i];
            for (int i = 0; i < D; ++i) c[i] -= vp[i] * a;
            U.resize(size_t(kt + 1) * D);
            W.resize(size_t(kt + 1) * D);
            for (int i = 0; i < D; ++i) {
                U[size_t(kt) * D + i] = vp[i];
                W[size_t(kt) * D + i] = alpha[i] * vp[i];
            }
            ++kt;
        }
    }
    out.assign(wy_buffer_floats(D), 0.f);
    float* wth = out.data();
    float* wtl = wth + size_t(WY_KT) * D;
    float* uh = wtl + size_t(WY_KT) * D;
    float* ul = uh + size_t(D) * WY_KT;
    float* al = ul + size_t(D) * WY_KT;
    float* cc = al + D;
    for (int k = 0; k < kt && k < WY_KT; ++k)
        for (int i = 0; i < D; ++i) {
            split_tf32(W[size_t(k) * D + i], wth[size_t(k) * D + i], wtl[size_t(k) * D + i]);
            split_tf32(U[size_t(k) * D + i], uh[size_t(i) * WY_KT + k], ul[size_t(i) * WY_KT + k]);
        }
    for (int i = 0; i < D; ++i) {
        al[i] = float(alpha[i]);
        cc[i] = float(c[i]);
    }
}

cudaError_t launch_wy(int D, const float* d_wy, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                      int sm_count, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const float* wth = d_wy;
    const float* wtl = wth + size_t(WY_KT) * D;
    const float* uh = wtl + size_t(WY_KT) * D;
    const float* ul = uh + size_t(D) * WY_KT;
    const float* al = ul + size_t(D) * WY_KT;
    const float* cc = al + D;
    CUtensorMap mx, mwh, mwl, muh, mul;
    if (!make_map(&mx, x, uint64_t(N), uint64_t(D), WY_TILE_M, WY_KC) ||
        !make_map(&mwh, wth, uint64_t(WY_KT), uint64_t(D), WY_KT, WY_KC) ||
        !make_map(&mwl, wtl, uint64_t(WY_KT), uint64_t(D), WY_KT, WY_KC) ||
        !make_map(&muh, uh, uint64_t(D), uint64_t(WY_KT), uint32_t(D), 32) ||
        !make_map(&mul, ul, uint64_t(D), uint64_t(WY_KT), uint32_t(D), 32))
        return cudaErrorInvalidValue;
    const int64_t tiles = (N + WY_TILE_M - 1) / WY_TILE_M;
    const unsigned grid = unsigned(tiles < sm_count ? tiles : sm_count);
    const float lc = float(ladj_const);
    float* lf = static_cast<float*>(ladj);
    const float* xf = static_cast<const float*>(x);
    float* yf = static_cast<float*>(y);
    cudaError_t e = cudaSuccess;
#define ENF_WY_LAUNCH(ND)                                                                                              \
    {                                                                                                                  \
        const int smem = WySmem<ND>::TOTAL;                                                                            \
        static bool set[64] = {};                                                                                      \
        int dev = 0;                                                                                                   \
        cudaGetDevice(&dev);                                                                                           \
        if (!set[dev & 63]) {                                                                                          \
            e = cudaFuncSetAttribute(wy_gemm_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);           \
            if (e != cudaSuccess) return e;                                                                            \
            set[dev & 63] = true;                                                                                      \
        }                                                                                                              \
        wy_gemm_kernel<ND><<<grid, WY_THREADS, smem, st>>>(mx, mwh, mwl, muh, mul, xf, yf, al, cc, lf, lc, N);         \
    }
    if (D == 256) ENF_WY_LAUNCH(256)
    else if (D == 128) ENF_WY_LAUNCH(128)
    else return cudaErrorInvalidValue;
#undef ENF_WY_LAUNCH
    return cudaGetLastError();
}

}  // namespace enf
