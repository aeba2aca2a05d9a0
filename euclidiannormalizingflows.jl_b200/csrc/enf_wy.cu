// F4 of SURVEY §2.3 in the form BASELINE.json's north star names: deep Householder stacks at large D applied in
// compact-WY form as tcgen05 / TMEM tensor-core GEMMs.
//
// A chain of HouseholderTrafo and ScaleShiftTrafo ops is "diagonal + low rank":  with the pre-scaled vectors
// v' = v sqrt(2 / v.v) every reflection (src/householder_trafo.jl:4-11) is I - v' v'^T, and the whole chain folds into
//        y = alpha . (x + U' (W^T x)) + c ,        U', W : D x Kt  (Kt = number of reflections),
// on the host in float64 (wy_fold: one column of U' and W per reflection; U' = -U / alpha row-wise).  Per sample that is
// 4 D Kt flops instead of the 2 D^2 of the dense fold y = W x + c (enf_affine.cu): at D = 256, Kt = 64 half the tensor
// work, and the dense kernel is tensor-bound.  Chains with Kt > 64 (or Kt > D/2, or a zero scale) keep the dense kernel.
//
// Per 128-sample tile the sample tile LIVES IN TENSOR MEMORY: the accumulator V (128 lanes x D columns) is initialised
// with x itself, serves as the A operand of the first GEMM, receives the low-rank update from the second GEMM, and is
// drained once as y.  Nothing is read twice from HBM or L2, and only the remainder xl = x - tf32(x) of the 3xTF32 split
// ever sits in shared memory as an operand.
//   split    four warps (one TMEM lane quarter each, thread = sample row) read a 32-column chunk of x from the TMA ring,
//            store it to V with tcgen05.st, and overwrite it IN PLACE with xl = tf32(x - trunc_tf32(x)) (the tensor core
//            reads the upper 19 bits of a 32-bit container, so V itself is the truncated high part)
//   GEMM1    T[128 x 64] = X W:  (main | correction) += V[:, k..k+8] (A from TMEM) . [Wh | Wl]  (one N = 128 MMA),
//                                 correction += xl (A from smem) . Wh
//   hand-off T -> (Thi, Tlo): eight warps read the accumulator pair, split the sum into tf32 high / low parts and write
//            them back to tensor memory in place; GEMM2 takes its A operand from there
//   GEMM2    V += Thi . U'h + Tlo . U'h + Thi . U'l   (A from TMEM, B = 16 KB pieces of U' streamed through four buffers),
//            first for the columns [0, D/2), then for [D/2, D): the epilogue drains the first half while the second runs
//   epilogue y = alpha . V + c: TMEM hands a lane one sample, global memory wants a lane to own columns: V is transposed
//            through a 4 KB shared-memory box per warp and stored with coalesced 16-byte stores.  A drained 32-column
//            chunk of V is handed back to the splitters at once, so the next tile's x flows in behind the epilogue.
// GEMM1 and GEMM2 are issued by two different warps; T is double buffered (GEMM1 of the next tile starts while the second
// half of GEMM2 still reads T).
//
// Shared memory: 4-stage ring of 32-column chunks (x -> xl 16 KB | Wh 8 KB | Wl 8 KB) = 128 KB, four 16 KB buffers for
// the pieces of U', eight 4 KB staging boxes = 224 KB.  TMEM: two T buffers of (main | correction) = 2 x 128 columns,
// then V (D columns).
// Warp roles (512 threads): warp 0 TMA producer (x, W), warp 1 TMEM owner + GEMM1 issuer, warps 2-5 splitters, warps 6-13
// hand-off + epilogue (two per TMEM lane quarter), warp 14 TMA producer of the U' pieces, warp 15 GEMM2 issuer.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
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
constexpr int WY_STAGES = 4;
constexpr int WY_UBUFS = 4;
constexpr int WY_SPLIT_WARPS = 4;
constexpr int WY_EPI_WARPS = 8;
#ifndef ENF_WY_ISSUERS
#define ENF_WY_ISSUERS 2     // threads issuing the MMAs of each GEMM (on alternating chunks / pieces)
#endif
constexpr int WY_ISSUERS = ENF_WY_ISSUERS;
constexpr int WY_THREADS = 32 * (2 + WY_SPLIT_WARPS + WY_EPI_WARPS + 2 + 2 * (WY_ISSUERS - 1));

template <int ND>
struct WySmem {
    static constexpr int HALF = ND / 2;                             // output columns per GEMM2 pass
    static constexpr int X_BYTES = WY_TILE_M * WY_KC * 4;           // 16 KB
    static constexpr int W_BYTES = WY_KT * WY_KC * 4;               // 8 KB
    static constexpr int U_PIECE_BYTES = HALF * 32 * 4;             // one piece of U': [HALF rows x 32 k], 128B swizzle
    static constexpr int XRING_OFF = WY_UBUFS * U_PIECE_BYTES;      // x chunks: a slot is free again once the splitters have read it
    static constexpr int WRING_OFF = XRING_OFF + WY_STAGES * X_BYTES;   // Wh | Wl chunks: free once GEMM1 of the chunk has retired
    static constexpr int OUT_BYTES = 32 * 32 * 4;
    static constexpr int OUT_OFF = WRING_OFF + WY_STAGES * 2 * W_BYTES;
    static constexpr int AC_OFF = OUT_OFF + WY_EPI_WARPS * OUT_BYTES;   // alpha | c (2 ND floats)
    static constexpr int BAR_OFF = AC_OFF + 2 * ND * 4;
    static constexpr int TOTAL = BAR_OFF + 512;                     // the dynamic window starts 1024-byte aligned (checked)
    // tensor memory: T (main | correction accumulators, rewritten in place as Thi | Tlo by the hand-off), a window of
    // WY_STAGES 32-column chunks of xl (the A operand of the correction product), V
    static constexpr uint32_t T_COL = 0, TC_OFF = 64, XL_COL = 128, V_COL = 256;
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

#ifdef ENF_WY_TRACE   // tuning aid (tools/build_variant.sh): clock64() of the hand-shakes of CTA 0 for tiles 100-103
__device__ long long wy_trace_buf[7 * 4 * 16];
#define WY_T(role, j, ev)                                                                \
    if (blockIdx.x == 0 && lane == 0 && (j) >= 100 && (j) < 104) wy_trace_buf[((role) * 4 + ((j) - 100)) * 16 + (ev)] = clock64();
#else
#define WY_T(role, j, ev)
#endif

template <int ND>
__global__ void __launch_bounds__(WY_THREADS, 1)
wy_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
               const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_uh,
               const __grid_constant__ CUtensorMap map_ul, const __grid_constant__ CUtensorMap map_y,
               const float* __restrict__ alpha,
               const float* __restrict__ cvec, float* __restrict__ ladj, float ladj_const, int64_t N) {
    using S = WySmem<ND>;
    constexpr int NKC = ND / WY_KC;                    // ring chunks per tile = 32-column chunks of V
    constexpr int HALF = S::HALF;
    extern __shared__ __align__(1024) unsigned char smem[];       // no static shared memory in this kernel: offset 0 of the window
    if (smem_u32(smem) & 1023u) __trap();
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF); // x chunk landed                 (1 + tx)
    uint64_t* w_full = x_full + WY_STAGES;                             // Wh | Wl chunk landed           (1 + tx)
    uint64_t* split = w_full + WY_STAGES;                              // x in V, xl in the TMEM window  (4 warps)
    uint64_t* empty = split + WY_STAGES;                               // MMAs of the chunk retired      (tcgen05.commit)
    uint64_t* u_full = empty + WY_STAGES;                              // a piece of U' landed           (1 + tx)
    uint64_t* u_empty = u_full + WY_UBUFS;                             // its MMAs retired               (tcgen05.commit)
    uint64_t* t_full = u_empty + WY_UBUFS;                             // GEMM1 of a tile retired        (tcgen05.commit)
    uint64_t* t_split = t_full + 1;                                    // Thi / Tlo written              (8 warps)
    uint64_t* t_free = t_split + 1;                                    // GEMM2 has read Thi / Tlo       (tcgen05.commit)
    uint64_t* v_full = t_free + 1;                                     // [2] a column half of V is final (tcgen05.commit)
    uint64_t* v_free = v_full + 2;                                     // [NKC] a chunk of V is drained  (4 warps)
    uint64_t* g1_go = v_free + NKC;                                    // the zero-initialising MMAs of a tile have retired (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g1_go + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (N + WY_TILE_M - 1) / WY_TILE_M;
    const int my_tiles = blockIdx.x < n_tiles ? int((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WY_STAGES; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&w_full[s], 1);
            mbar_init(&split[s], WY_SPLIT_WARPS);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < WY_UBUFS; ++b) {
            mbar_init(&u_full[b], 1);
            mbar_init(&u_empty[b], 1);
        }
        mbar_init(t_full, WY_ISSUERS);
        mbar_init(g1_go, 1);
        mbar_init(t_split, WY_EPI_WARPS);
        mbar_init(t_free, WY_ISSUERS);
        mbar_init(&v_full[0], WY_ISSUERS);
        mbar_init(&v_full[1], WY_ISSUERS);
        for (int c = 0; c < NKC; ++c) mbar_init(&v_free[c], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    float* s_ac = reinterpret_cast<float*>(smem + S::AC_OFF);
    for (int i = threadIdx.x; i < ND; i += WY_THREADS) {
        s_ac[i] = alpha[i];
        s_ac[ND + i] = cvec[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tV = tmem_base + S::V_COL;

    if (warp == 0) {
        // ===== TMA producer of the x chunks: runs up to WY_STAGES chunks ahead of the splitters =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int64_t tile = int64_t(blockIdx.x) + int64_t(i) * gridDim.x;
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    const int s = it % WY_STAGES;
                    if (it >= uint32_t(WY_STAGES)) mbar_wait(&split[s], ((it / WY_STAGES) - 1) & 1);
                    WY_T(4, i, kc)
                    mbar_expect_tx(&x_full[s], S::X_BYTES);
                    tma_load_2d_hint(smem + S::XRING_OFF + s * S::X_BYTES, &map_x, kc * WY_KC, int(tile * WY_TILE_M), &x_full[s], L2_EVICT_FIRST);
                }
            }
        }
    } else if (warp == 1 || (WY_ISSUERS == 2 && warp == 16)) {
        // ===== GEMM1 issuer: T(i) = X(i) W into the (main | correction) accumulator pair.  ONE thread runs the whole loop: the
        // tensor core's instruction queue is shallow (tools/mma_rate.cu: every cycle the issuing thread spends on barriers,
        // re-convergence or commits between two MMAs is a cycle the tensor pipe idles), so the loop carries nothing warp-wide,
        // and TWO such threads (warps 1 and 16) take alternating chunks: while one sits in its waits and commits, the other
        // one's MMAs keep the pipe fed.  The only order that matters is that the zero-initialising MMA of a tile comes first:
        // the second thread starts a tile once the first one's chunk 0 has retired (g1_go) =====
        constexpr uint32_t idesc_n128 = make_idesc_tf32(WY_TILE_M, 2 * WY_KT);
        constexpr uint32_t idesc_n64 = make_idesc_tf32(WY_TILE_M, WY_KT);
        const uint32_t tM = tmem_base + S::T_COL, tC = tM + S::TC_OFF;
        const int me = warp == 1 ? 0 : 1;
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                if (i >= 1) mbar_wait(t_free, uint32_t(i - 1) & 1u);                   // GEMM2 of tile i-1 has read Thi | Tlo
                if (me == 1) mbar_wait(g1_go, uint32_t(i) & 1u);
                for (int kc = 0; kc < NKC; ++kc, ++it) {
                    if (WY_ISSUERS == 2 && (kc & 1) != me) continue;
                    const int s = it % WY_STAGES;
                    mbar_wait(&w_full[s], (it / WY_STAGES) & 1);
                    mbar_wait(&split[s], (it / WY_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    WY_T(0, i, kc)
                    const uint64_t dw = make_desc_kmajor<WY_KC>(smem + S::WRING_OFF + s * 2 * S::W_BYTES);   // rows 0-63 Wh, 64-127 Wl
                    const uint32_t tX = tV + uint32_t(kc * WY_KC), tXL = tmem_base + S::XL_COL + uint32_t(s * WY_KC);
#pragma unroll
                    for (int j = 0; j < WY_KC / 8; ++j) {          // UMMA K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        // (main | correction) (+)= trunc(x) . [Wh | Wl]: A = the chunk's columns of V.  The tensor core truncates
                        // when it adds into the f32 accumulator, one ulp of the ACCUMULATOR per MMA whatever the size of the
                        // addend: the 2^-11-small correction products have an accumulator of their own.
                        umma_tf32_ts(tM, tX + uint32_t(j * 8), dw + adv, idesc_n128, (kc | j) != 0);
                        umma_tf32_ts(tC, tXL + uint32_t(j * 8), dw + adv, idesc_n64, 1);       // correction += xl . Wh
                    }
                    umma_commit(&empty[s]);
                    if (WY_ISSUERS == 2 && kc == 0) umma_commit(g1_go);
                    if (kc >= NKC - WY_ISSUERS) umma_commit(t_full);                   // this thread's last chunk of the tile
                }
            }
        }
    } else if (warp == 15 || (WY_ISSUERS == 2 && warp == 17)) {
        // ===== GEMM2 issuer (one thread, see GEMM1): V(j) += T(j) U'^T, A = Thi | Tlo in tensor memory, B = the pieces of U' =====
        constexpr uint32_t idesc2 = make_idesc_tf32(WY_TILE_M, HALF);
        const uint32_t tHI = tmem_base + S::T_COL, tLO = tHI + S::TC_OFF;
        const int me = warp == 15 ? 0 : 1;                                             // two threads on alternating pieces (see GEMM1)
        if (lane == 0) {
            uint32_t q = 0;
            for (int j = 0; j < my_tiles; ++j) {
                mbar_wait(t_split, uint32_t(j) & 1u);                                  // Thi / Tlo of tile j are in TMEM
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                WY_T(3, j, 0)
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    // pieces of a half: U'l k0, U'l k1 (correction Thi . U'l), U'h k0, U'h k1 (Tlo . U'h and the main product)
                    const uint32_t tD = tV + uint32_t(h * HALF);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc, ++q) {
                        if (WY_ISSUERS == 2 && (pc & 1) != me) continue;
                        const uint32_t b = q % WY_UBUFS;
                        mbar_wait(&u_full[b], (q / WY_UBUFS) & 1u);
                        WY_T(3, j, 1 + 4 * h + pc)
                        const uint64_t du = make_desc_kmajor<32>(smem + b * S::U_PIECE_BYTES);
                        const uint32_t k0 = uint32_t(pc & 1) * 32u;                    // first T column of this piece
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t adv = uint64_t((kk * 32) >> 4);
                            if (pc >= 2) umma_tf32_ts(tD, tLO + k0 + uint32_t(kk * 8), du + adv, idesc2, 1);
                            umma_tf32_ts(tD, tHI + k0 + uint32_t(kk * 8), du + adv, idesc2, 1);
                        }
                        umma_commit(&u_empty[b]);
                    }
                    umma_commit(&v_full[h]);
                }
                umma_commit(t_free);
            }
        }
    } else if (warp == 14) {
        // ===== TMA producer of the W chunks and of the U' pieces (both come from L2): one thread polls both queues, neither
        // may wait behind the other.  Per tile and column half the pieces are U'l k0, U'l k1, U'h k0, U'h k1 =====
        if (lane == 0) {
            const uint32_t n_w = uint32_t(my_tiles) * NKC, n_u = uint32_t(my_tiles) * 8u;
            uint32_t wq = 0, uq = 0;
            while (wq < n_w || uq < n_u) {
                if (wq < n_w) {
                    const uint32_t s = wq % WY_STAGES;
                    if (wq < uint32_t(WY_STAGES) || mbar_test(&empty[s], ((wq / WY_STAGES) - 1) & 1u)) {
                        const int kc = int(wq % NKC);
                        unsigned char* dst = smem + S::WRING_OFF + s * 2 * S::W_BYTES;
                        mbar_expect_tx(&w_full[s], 2 * S::W_BYTES);
                        tma_load_2d_hint(dst, &map_wh, kc * WY_KC, 0, &w_full[s], L2_EVICT_LAST);              // Wh chunk [64 x 32]
                        tma_load_2d_hint(dst + S::W_BYTES, &map_wl, kc * WY_KC, 0, &w_full[s], L2_EVICT_LAST);
                        ++wq;
                    }
                }
                if (uq < n_u) {
                    const uint32_t b = uq % WY_UBUFS, pc = uq & 3u, h = (uq >> 2) & 1u;
                    if (uq < uint32_t(WY_UBUFS) || mbar_test(&u_empty[b], ((uq / WY_UBUFS) - 1) & 1u)) {
                        WY_T(5, int(uq >> 3), int(uq & 7u))
                        mbar_expect_tx(&u_full[b], S::U_PIECE_BYTES);
                        tma_load_2d_hint(smem + b * S::U_PIECE_BYTES, pc < 2 ? &map_ul : &map_uh, int(pc & 1u) * 32, int(h) * HALF, &u_full[b],
                                         L2_EVICT_LAST);
                        ++uq;
                    }
                }
            }
        }
    } else if (warp < 2 + WY_SPLIT_WARPS) {
        // ===== splitters (thread = sample row of this warp's TMEM lane quarter): x -> V, xl = tf32(x - trunc(x)) in place =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = uint32_t(quarter * 32) << 16;
        uint32_t it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            for (int kc = 0; kc < NKC; ++kc, ++it) {
                const int s = it % WY_STAGES;
                mbar_wait(&x_full[s], (it / WY_STAGES) & 1);
                if (warp == 2) { WY_T(1, i, kc) }
                if (it >= uint32_t(WY_STAGES)) mbar_wait(&empty[s], ((it / WY_STAGES) - 1) & 1);   // the xl window slot is free
                if (i >= 1) mbar_wait(&v_free[kc], uint32_t(i - 1) & 1u);      // the epilogue of tile i-1 has drained these columns
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // 16-byte chunk q of row r sits at chunk q ^ (r & 7) of the row (TMA 128-byte swizzle): conflict-free LDS.128
                const uint4* xs = reinterpret_cast<const uint4*>(smem + S::XRING_OFF + s * S::X_BYTES) + row * 8;
                uint32_t xv[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 v = xs[q ^ (row & 7)];
                    xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
                }
                tmem_st32(tV + uint32_t(kc * WY_KC) + lane_off, xv);
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    xv[e] = __float_as_uint(tf32_hi(__uint_as_float(xv[e]) - __uint_as_float(xv[e] & 0xFFFFE000u)));
                tmem_st32(tmem_base + S::XL_COL + uint32_t(s * WY_KC) + lane_off, xv);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
                if (warp == 2) { WY_T(1, i, 8 + kc) }
            }
        }
    } else {
        // ===== warps 6-13: per tile the hand-off (T -> Thi | Tlo, in place in tensor memory), then the epilogue (y = alpha V + c).
        // Two warps share a TMEM lane quarter: member p takes the T columns [32 p, 32 p + 32) and the V chunks c = p (mod 2) =====
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int ew = warp - 2 - WY_SPLIT_WARPS;
        const int p = ew >> 2;
        float4* box = reinterpret_cast<float4*>(smem + S::OUT_OFF + size_t(ew) * S::OUT_BYTES);
        const uint32_t lane_off = uint32_t(quarter * 32) << 16;
        for (int j = 0; j < my_tiles; ++j) {
            {
                // ---- hand-off: this warp's 32 rows x 32 columns of T(j): main + correction, split into tf32 high / low parts
                mbar_wait(t_full, uint32_t(j) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 6) { WY_T(2, j, 0) }
                const uint32_t tM = tmem_base + S::T_COL + uint32_t(p * 32) + lane_off, tC = tM + S::TC_OFF;
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
                tmem_st32(tM, hi);                                     // in place: (main | correction) -> (Thi | Tlo)
                tmem_st32(tC, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(t_split);
                if (warp == 6) { WY_T(2, j, 1) }
            }
            // ---- epilogue of tile j.  TMEM hands a lane one SAMPLE (32 columns of it): y = alpha V + c goes row by row into a
            // 4 KB staging box in the tensor map's 128-byte swizzle (conflict-free STS.128), and one bulk tensor store per box
            // writes full 128-byte row segments (rows beyond N are clipped by the tensor map).
            const int64_t tile = int64_t(blockIdx.x) + int64_t(j) * gridDim.x;
            const int64_t row0 = tile * WY_TILE_M + quarter * 32;      // this warp's 32 samples
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                mbar_wait(&v_full[h], uint32_t(j) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 6) { WY_T(2, j, 2 + 4 * h) }
#pragma unroll 1
                for (int cc = 0; cc < NKC / 4; ++cc) {
                    const int c = h * (NKC / 2) + 2 * cc + p;
                    float v[32];
                    tmem_ld32(tV + uint32_t(c * 32) + lane_off, v);
                    if (warp == 6) { WY_T(2, j, 3 + 4 * h + cc) }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (warp == 6 && cc == 0) { WY_T(6, j, h * 8 + 0) }
                    if (lane == 0) {
                        mbar_arrive(&v_free[c]);                       // the splitters may refill these columns
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the box
                    }
                    __syncwarp();
                    if (warp == 6 && cc == 0) { WY_T(6, j, h * 8 + 1) }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 a4 = *reinterpret_cast<const float4*>(s_ac + c * 32 + q * 4);          // broadcast LDS.128
                        const float4 c4 = *reinterpret_cast<const float4*>(s_ac + ND + c * 32 + q * 4);
                        box[lane * 8 + (q ^ (lane & 7))] = make_float4(fmaf(a4.x, v[4 * q], c4.x), fmaf(a4.y, v[4 * q + 1], c4.y),
                                                                       fmaf(a4.z, v[4 * q + 2], c4.z), fmaf(a4.w, v[4 * q + 3], c4.w));
                    }
                    if (warp == 6 && cc == 0) { WY_T(6, j, h * 8 + 2) }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (warp == 6 && cc == 0) { WY_T(6, j, h * 8 + 3) }
                    if (lane == 0) {
                        tma_store_2d(&map_y, box, c * 32, int(row0));
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (warp == 6 && cc == 0) { WY_T(6, j, h * 8 + 4) }
                }
            }
            if (p == 0 && ladj != nullptr && row0 + lane < N) __stcs(ladj + row0 + lane, ladj_const);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    // D = 128 is instantiated and correct, but measured level with the dense fold there (0.72 vs 0.75 of HBM peak: a 64 KB
    // tile does not amortise the per-tile hand-shakes); D = 256: 0.73 vs 0.50
    static const bool wy128 = getenv("ENF_WY_D128") != nullptr;
    if (dtype != 0 || !(D == 256 || (D == 128 && wy128))) return 0;
    int n_refl = 0;
    for (int o = 0; o < d.n_ops; ++o) {
        if (d.ops[o].kind != OP_HH && d.ops[o].kind != OP_SS) return 0;
        if (d.ops[o].kind == OP_HH) n_refl += d.ops[o].K;
    }
    // below 8 reflections the SIMT kernel is HBM-bound already; above D/4 + the dense fold needs no more tensor work
    return (n_refl >= 8 && n_refl <= WY_KT && 2 * n_refl <= D) ? n_refl : 0;
}

size_t wy_buffer_floats(int D) { return size_t(4) * WY_KT * D + 2 * size_t(D); }

// Fold the chain into y = alpha . x - U (W^T x) + c = alpha . (x + U' (W^T x)) + c in float64 (one column of U, W per
// reflection; U' = -U / alpha row by row) and lay the operands out for the kernel:
// Wt hi | Wt lo ([64][D]) | U' hi | U' lo ([D][64]) | alpha [D] | c [D].  False if a scale is zero (or so small that U' is
// not finite in Float32): the caller then uses the dense fold.
bool wy_fold(int D, int n_ops, const int* kinds, const int* Ks, const double* const* params, std::vector<float>& out) {
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
    bool ok = true;
    for (int k = 0; k < kt && k < WY_KT; ++k)
        for (int i = 0; i < D; ++i) {
            split_tf32(W[size_t(k) * D + i], wth[size_t(k) * D + i], wtl[size_t(k) * D + i]);
            const double up = -U[size_t(k) * D + i] / alpha[i];
            if (!(std::fabs(up) < 1e30)) ok = false;
            split_tf32(ok ? up : 0.0, uh[size_t(i) * WY_KT + k], ul[size_t(i) * WY_KT + k]);
        }
    for (int i = 0; i < D; ++i) {
        al[i] = float(alpha[i]);
        cc[i] = float(c[i]);
    }
    return ok;
}

cudaError_t launch_wy(int D, const float* d_wy, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                      int sm_count, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return cudaErrorInvalidValue;
    if (N <= 0) return cudaSuccess;
    const float* wth = d_wy;
    const float* wtl = wth + size_t(WY_KT) * D;
    const float* uh = wtl + size_t(WY_KT) * D;
    const float* ul = uh + size_t(D) * WY_KT;
    const float* al = ul + size_t(D) * WY_KT;
    const float* cc = al + D;
    CUtensorMap mx, mwh, mwl, muh, mul, my;
    if (!make_map(&mx, x, uint64_t(N), uint64_t(D), WY_TILE_M, WY_KC) || !make_map(&my, y, uint64_t(N), uint64_t(D), 32, 32) ||
        !make_map(&mwh, wth, uint64_t(WY_KT), uint64_t(D), WY_KT, WY_KC) ||
        !make_map(&mwl, wtl, uint64_t(WY_KT), uint64_t(D), WY_KT, WY_KC) ||
        !make_map(&muh, uh, uint64_t(D), uint64_t(WY_KT), uint32_t(D / 2), 32) ||
        !make_map(&mul, ul, uint64_t(D), uint64_t(WY_KT), uint32_t(D / 2), 32))
        return cudaErrorInvalidValue;
    const int64_t tiles = (N + WY_TILE_M - 1) / WY_TILE_M;
    const unsigned grid = unsigned(tiles < sm_count ? tiles : sm_count);
    const float lc = float(ladj_const);
    float* lf = static_cast<float*>(ladj);
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
        wy_gemm_kernel<ND><<<grid, WY_THREADS, smem, st>>>(mx, mwh, mwl, muh, mul, my, al, cc, lf, lc, N);             \
    }
    if (D == 256) ENF_WY_LAUNCH(256)
    else if (D == 128) ENF_WY_LAUNCH(128)
    else return cudaErrorInvalidValue;
#undef ENF_WY_LAUNCH
#ifdef ENF_WY_TRACE
    {
        static long long h[7 * 4 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, wy_trace_buf, sizeof(h));
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        const char* names[7] = {"gemm1 issue  ", "split (w2)   ", "epilogue (w6)", "gemm2 issue  ", "x producer   ", "u producer   ", "epi detail   "};
        for (int j = 0; j < 4; ++j)
            for (int r = 0; r < 7; ++r) {
                fprintf(stderr, "tile %d %s", 100 + j, names[r]);
                for (int e = 0; e < 16; ++e) fprintf(stderr, " %6lld", h[(r * 4 + j) * 16 + e] ? h[(r * 4 + j) * 16 + e] - t0 : -1);
                fprintf(stderr, "\n");
            }
    }
#endif
    return cudaGetLastError();
}

}  // namespace enf
