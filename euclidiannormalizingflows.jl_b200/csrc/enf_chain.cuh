// Fused trafo-chain kernels (sm_100a): one pass over a D x N column-major sample
// matrix applies a whole chain (CenterStretch/Contract, Johnson/Inv, ScaleShift,
// Householder stacks), optionally with the per-sample ladj, or with the
// whitening loss and the raw parameter-gradient sums.
//
// Data layout / thread mapping ("lane groups").  A sample is a contiguous column
// of D elements.  It is split into 16-byte vectors (VE = 4 floats / 2 doubles);
// a group of G = 2^LG adjacent lanes owns one sample, lane g of the group owns
// vectors g, g+G, ... (CH of them).  A warp therefore reads 32 consecutive
// 16-byte vectors per load instruction: every HBM access is a fully coalesced
// 128-bit LDG/STG and the per-sample registers are filled without a transpose.
// Row-wise parameters become per-lane constants (the lane's rows never change),
// Householder dot products and the per-sample ladj are finished with log2(G)
// xor-shuffles inside the group.  D that is not a multiple of VE, or unaligned
// pointers, use the same mapping with masked scalar accesses (MODE_SCALAR);
// D < VE (D dividing VE) packs VE/D samples into one vector (MODE_PACK, or
// MODE_PACKU with scalar accesses when a pointer is not 16-byte aligned).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "enf_math.cuh"

namespace enf {

constexpr int MAX_OPS = 24;
constexpr int MAX_TARGET_K = 8;
#ifndef ENF_NT
#define ENF_NT 256
#endif
constexpr int NT = ENF_NT;  // threads per CTA of the chain kernels

enum : int { OP_CS = 0, OP_CC = 1, OP_JO = 2, OP_JI = 3, OP_SS = 4, OP_HH = 5 };
enum : int { MODE_VEC = 0, MODE_SCALAR = 1, MODE_PACK = 2, MODE_PACKU = 3 };  // PACKU: pack layout, unaligned pointers

struct DevOp {
    int kind;  // OP_*
    int K;     // reflections (OP_HH)
    int coff;  // offset of this op's constants (elements) in the constants block
    int roff;  // first per-row raw-sum slot of this op
    int soff;  // first scalar raw-sum slot (OP_HH: sum_j p'_j q'_j per reflection)
    int save;  // index of the saved-input tile (elementwise ops), -1 for OP_HH
};

struct ChainDesc {
    int n_ops;
    int D;           // rows
    int Dp;          // padded rows (G*CH*VE >= D, or VE in MODE_PACK); constants are padded with neutral values
    int n_consts;    // elements in the constants block
    int n_save;      // saved-input tiles the gradient kernel needs
    int n_rowslots;  // raw-sum slots holding one value per row
    int n_scalars;   // raw-sum slots holding one value per chain
    // Objective of the loss / gradient kernels.  0: the whitening loss of src/optimize_whitening.jl:7-15 (target =
    // standard normal).  1: the negative ELBO of examples/nf_variational_1d.jl:29-41 with an element-wise Gaussian-mixture
    // target log-density (the example's my_ll, :25-27): component k has log weight tlw[k] = log(w_k / (sigma_k sqrt(2 pi))),
    // mean tmu[k], inverse width tis[k].
    int target_kind;
    int target_K;
    double tlw[MAX_TARGET_K], tmu[MAX_TARGET_K], tis[MAX_TARGET_K];
    DevOp ops[MAX_OPS];
};

__host__ __device__ constexpr int n_consts_of(int kind, int K) {
    return (kind == OP_CS || kind == OP_CC) ? 9 : kind == OP_JO ? 5 : kind == OP_JI ? 6 : kind == OP_SS ? 2 : K;   // HH: the pre-scaled v'_k
}
__host__ __device__ constexpr int n_rowslots_of(int kind, int K) {
    return (kind == OP_CS || kind == OP_CC) ? 3 : (kind == OP_JO || kind == OP_JI) ? 4 : kind == OP_SS ? 2 : K;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------ 16-byte accessors
template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int VE = 4; };
template <> struct Vec<double> { static constexpr int VE = 2; };

__device__ __forceinline__ void ld16_stream(const float* p, float (&o)[4]) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void ld16_stream(const double* p, double (&o)[2]) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    o[0] = t.x; o[1] = t.y;
}
__device__ __forceinline__ void st16_stream(float* p, const float (&o)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
}
__device__ __forceinline__ void st16_stream(double* p, const double (&o)[2]) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(o[0], o[1]));
}
__device__ __forceinline__ void ld16_shared(const float* p, float (&o)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void ld16_shared(const double* p, double (&o)[2]) {
    double2 t = *reinterpret_cast<const double2*>(p);
    o[0] = t.x; o[1] = t.y;
}
__device__ __forceinline__ void st16_shared(float* p, const float (&o)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void st16_shared(double* p, const double (&o)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
}
// 16-byte load from a 32-bit shared-window address (keeps the ring reads LDS.128 with immediate offsets;
// through a generic pointer the compiler emits 64-bit generic LD.E.128)
__device__ __forceinline__ void lds16(uint32_t a, float (&o)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(a));
}
__device__ __forceinline__ void lds16(uint32_t a, double (&o)[2]) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(a));
}
__device__ __forceinline__ void sts16(uint32_t a, const float (&o)[4]) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, const double (&o)[2]) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(o[0]), "d"(o[1]) : "memory");
}
// 16-byte read-modify-write of a per-thread accumulator vector in shared memory
template <typename T, int VE>
__device__ __forceinline__ void acc16_shared(uint32_t p, const T (&a)[VE]) {   // p: 32-bit shared-window address
    T cur[VE];
    lds16(p, cur);
#if ENF_F32X2
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int e = 0; e < VE; e += 2) {
            const float2 s2 = add2(make_float2(cur[e], cur[e + 1]), make_float2(a[e], a[e + 1]));
            cur[e] = s2.x;
            cur[e + 1] = s2.y;
        }
        sts16(p, cur);
        return;
    }
#endif
#pragma unroll
    for (int e = 0; e < VE; ++e) cur[e] += a[e];
    sts16(p, cur);
}

// ------------------------------------------------------------------ configuration
template <typename T_, int LG_, int CH_, int MODE_, int PD_, int SPT_>
struct Cfg {
    using T = T_;
    static constexpr int LG = LG_;
    static constexpr int G = 1 << LG_;
    static constexpr int CH = CH_;
    static constexpr int MODE = MODE_;
    static constexpr int PD = PD_;  // rows per sample in MODE_PACK (divides VE), else 0
    static constexpr int SPT = SPT_;
    static constexpr int VE = Vec<T_>::VE;
    static constexpr bool PACKED = (MODE_ == MODE_PACK || MODE_ == MODE_PACKU);
    static constexpr int LN = PACKED ? (Vec<T_>::VE / (PD_ > 0 ? PD_ : 1)) : 1;  // samples per tile row of a thread
    static constexpr int PDD = PD_ > 0 ? PD_ : 1;
    static constexpr int SB = NT / G;  // samples (packed: vectors) per CTA step
    static constexpr int DP = PACKED ? Vec<T_>::VE : (1 << LG_) * CH_ * Vec<T_>::VE;  // padded rows == ChainDesc::Dp
    static_assert(!PACKED || (LG_ == 0 && CH_ == 1 && PD_ > 0), "pack mode is one vector per thread");
    // ladj / mask slot of element e of a vector
    static __host__ __device__ constexpr int slot(int e) { return PACKED ? e / PDD : 0; }
};

template <class C> struct Tile {
    typename C::T v[C::SPT][C::CH][C::VE];
};

// first sample (MODE_PACK: first vector) of row u of a tile, for this thread
// W = true: warp-private tiles (the forward kernels): `tile` counts tiles of SPT * (32 / G) items owned by ONE warp
template <class C, bool W = false>
__device__ __forceinline__ int64_t tile_item(int64_t tile, int u) {
    if (W) return (tile * C::SPT + u) * (32 / C::G) + ((threadIdx.x & 31) >> C::LG);
    return (tile * C::SPT + u) * C::SB + (threadIdx.x >> C::LG);
}

// Loads tile `tile` of this thread.  nv[u] = number of valid samples in tile row u
// (0/1 in the lane-group modes, 0..LN in the packed modes); invalid elements read 0.
template <class C, bool W = false>
__device__ __forceinline__ void load_tile(const typename C::T* x, int64_t N, int D, int64_t tile,
                                          Tile<C>& t, int (&nv)[C::SPT]) {
    using T = typename C::T;
    const int g = threadIdx.x & (C::G - 1);
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        const int64_t s = tile_item<C, W>(tile, u);
        if (C::PACKED) {
            const int64_t left = N - s * C::LN;
            nv[u] = left <= 0 ? 0 : (left >= C::LN ? C::LN : int(left));
            if (C::MODE == MODE_PACK && nv[u] == C::LN) ld16_stream(x + s * C::VE, t.v[u][0]);
            else {
#pragma unroll
                for (int e = 0; e < C::VE; ++e)
                    t.v[u][0][e] = (C::slot(e) < nv[u]) ? __ldcs(x + s * C::VE + e) : T(0);
            }
        } else {
            nv[u] = s < N ? 1 : 0;
#pragma unroll
            for (int q = 0; q < C::CH; ++q) {
                const int row0 = (q * C::G + g) * C::VE;
                if (C::MODE == MODE_VEC) {
                    if (nv[u] && row0 < D) ld16_stream(x + s * D + row0, t.v[u][q]);
                    else {
#pragma unroll
                        for (int e = 0; e < C::VE; ++e) t.v[u][q][e] = T(0);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < C::VE; ++e)
                        t.v[u][q][e] = (nv[u] && row0 + e < D) ? __ldcs(x + s * D + row0 + e) : T(0);
                }
            }
        }
    }
}

template <class C, bool W = false>
__device__ __forceinline__ void store_tile(typename C::T* y, int D, int64_t tile, const Tile<C>& t,
                                           const int (&nv)[C::SPT]) {
    const int g = threadIdx.x & (C::G - 1);
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        if (nv[u] == 0) continue;
        const int64_t s = tile_item<C, W>(tile, u);
        if (C::PACKED) {
            if (C::MODE == MODE_PACK && nv[u] == C::LN) st16_stream(y + s * C::VE, t.v[u][0]);
            else {
#pragma unroll
                for (int e = 0; e < C::VE; ++e)
                    if (C::slot(e) < nv[u]) __stcs(y + s * C::VE + e, t.v[u][0][e]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < C::CH; ++q) {
                const int row0 = (q * C::G + g) * C::VE;
                if (C::MODE == MODE_VEC) {
                    if (row0 < D) st16_stream(y + s * D + row0, t.v[u][q]);
                } else {
#pragma unroll
                    for (int e = 0; e < C::VE; ++e)
                        if (row0 + e < D) __stcs(y + s * D + row0 + e, t.v[u][q][e]);
                }
            }
        }
    }
}

// xor-shuffle sum over the G lanes of a group
template <class C, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int off = C::G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// xor-shuffle sum over the G lanes of a group of a PAIR of values (Float32: one packed add per step)
template <class C, typename T>
__device__ __forceinline__ void group_sum2(T& a, T& b) {
#if ENF_F32X2
    if constexpr (sizeof(T) == 4) {
        float2 v = make_float2(a, b);
#pragma unroll
        for (int off = C::G / 2; off > 0; off >>= 1)
            v = add2(v, make_float2(__shfl_xor_sync(0xffffffffu, v.x, off), __shfl_xor_sync(0xffffffffu, v.y, off)));
        a = v.x;
        b = v.y;
        return;
    }
#endif
    a = group_sum<C>(a);
    b = group_sum<C>(b);
}

// offset of this lane's constants for vector q inside one length-Dp constant array
template <class C>
__device__ __forceinline__ int const_off(int q) {
    return C::PACKED ? 0 : (q * C::G + (threadIdx.x & (C::G - 1))) * C::VE;
}

// ------------------------------------------------------------------ forward ops on a tile
// Which of an op's constant arrays the forward pass needs, in the order the
// *_fwd_v functions take them (slot numbers: see enf_math.cuh).
__host__ __device__ constexpr int n_fwd_consts(int kind) { return kind == OP_CS ? 6 : kind == OP_SS ? 2 : 4; }
__host__ __device__ constexpr int fwd_const_slot(int kind, int j) {
    return kind == OP_CS ? (j == 0 ? 0 : j == 1 ? 6 : j == 2 ? 2 : j == 3 ? 3 : j == 4 ? 7 : 8) : j;
}
constexpr int MAX_FWD_CONSTS = 6;

#ifndef ENF_HH_UNROLL
#define ENF_HH_UNROLL 1
#endif
constexpr int HH_UNROLL = ENF_HH_UNROLL;   // unroll factor of the reflection loop of the interpretive kernel

// one Householder reflection y = x - (v'.x) v' with the pre-scaled v' = v sqrt(2/v.v)
// (src/householder_trafo.jl:4-11): partial dot products per lane, xor-shuffles inside the group
template <class C>
__device__ __forceinline__ void hh_reflect(Tile<C>& t, const typename C::T (&vk)[C::CH][C::VE]) {
    using T = typename C::T;
    constexpr int VE = C::VE;
    if (C::PACKED) {
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int p = 0; p < C::LN; ++p) {
                T d = T(0);
#pragma unroll
                for (int e = 0; e < C::PDD; ++e) d = Prim<T>::fma_(vk[0][p * C::PDD + e], t.v[u][0][p * C::PDD + e], d);
#pragma unroll
                for (int e = 0; e < C::PDD; ++e)
                    t.v[u][0][p * C::PDD + e] = Prim<T>::fma_(-d, vk[0][p * C::PDD + e], t.v[u][0][p * C::PDD + e]);
            }
    } else {
#if ENF_F32X2
        if constexpr (sizeof(T) == 4) {
            // FFMA2: two rows per instruction for the dot product and for the update
            float d[C::SPT];
#pragma unroll
            for (int u = 0; u < C::SPT; ++u) {
                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < C::CH; ++q) {
                    a = fma2(make_float2(vk[q][0], vk[q][1]), make_float2(t.v[u][q][0], t.v[u][q][1]), a);
                    a = fma2(make_float2(vk[q][2], vk[q][3]), make_float2(t.v[u][q][2], t.v[u][q][3]), a);
                }
                d[u] = a.x + a.y;
            }
            if constexpr (C::SPT % 2 == 0) {
#pragma unroll
                for (int u = 0; u < C::SPT; u += 2) group_sum2<C>(d[u], d[u + 1]);   // one packed add per shuffle step and sample pair
            } else {
#pragma unroll
                for (int u = 0; u < C::SPT; ++u) d[u] = group_sum<C>(d[u]);
            }
#pragma unroll
            for (int u = 0; u < C::SPT; ++u) {
                const float2 nd = make_float2(-d[u], -d[u]);
#pragma unroll
                for (int q = 0; q < C::CH; ++q) {
                    const float2 lo = fma2(nd, make_float2(vk[q][0], vk[q][1]), make_float2(t.v[u][q][0], t.v[u][q][1]));
                    const float2 hi = fma2(nd, make_float2(vk[q][2], vk[q][3]), make_float2(t.v[u][q][2], t.v[u][q][3]));
                    t.v[u][q][0] = lo.x; t.v[u][q][1] = lo.y; t.v[u][q][2] = hi.x; t.v[u][q][3] = hi.y;
                }
            }
            return;
        }
#endif
        T d[C::SPT];
#pragma unroll
        for (int u = 0; u < C::SPT; ++u) {
            d[u] = T(0);
#pragma unroll
            for (int q = 0; q < C::CH; ++q)
#pragma unroll
                for (int e = 0; e < VE; ++e) d[u] = Prim<T>::fma_(vk[q][e], t.v[u][q][e], d[u]);
        }
#pragma unroll
        for (int u = 0; u < C::SPT; ++u) d[u] = group_sum<C>(d[u]);
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int q = 0; q < C::CH; ++q)
#pragma unroll
                for (int e = 0; e < VE; ++e) t.v[u][q][e] = Prim<T>::fma_(-d[u], vk[q][e], t.v[u][q][e]);
    }
}

// elementwise trafo KIND on vector q of every sample of the tile; k[j] = j-th forward constant of the lane's rows.
// DEFER: Jacobian factors are multiplied into pn/pd (one pair per sample) instead of being logged per vector.
template <class C, int KIND, bool LADJ, bool SAFE, bool DEFER>
__device__ __forceinline__ void elem_fwd_q(Tile<C>& t, int q, const typename C::T (&k)[MAX_FWD_CONSTS][C::VE],
                                           typename C::T (&l)[C::SPT][C::LN], bool& bad, typename C::T (&pn)[C::SPT],
                                           typename C::T (&pd)[C::SPT]) {
    using T = typename C::T;
    constexpr int VE = C::VE;
    constexpr int GR = C::PACKED ? C::PDD : VE;   // consecutive elements that belong to one sample
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        if (KIND == OP_SS) {
#if ENF_F32X2
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int e = 0; e < VE; e += 2) {
                    const float2 y = fma2(make_float2(t.v[u][q][e], t.v[u][q][e + 1]), make_float2(k[0][e], k[0][e + 1]),
                                          make_float2(k[1][e], k[1][e + 1]));
                    t.v[u][q][e] = y.x;
                    t.v[u][q][e + 1] = y.y;
                }
            } else
#endif
            {
#pragma unroll
                for (int e = 0; e < VE; ++e) t.v[u][q][e] = Prim<T>::fma_(t.v[u][q][e], k[0][e], k[1][e]);
            }
        } else {
#pragma unroll
            for (int p = 0; p < VE / GR; ++p) {
                T* v = &t.v[u][q][p * GR];
                T& ll = l[u][C::slot(p * GR)];
                const int o = p * GR;
                if (KIND == OP_CS)
                    cs_fwd_v<T, GR, LADJ, SAFE, DEFER>(v, k[0] + o, k[1] + o, k[2] + o, k[3] + o, k[4] + o, k[5] + o, ll, bad, &pn[u], &pd[u]);
                else if (KIND == OP_CC) cc_fwd_v<T, GR, LADJ, SAFE, DEFER>(v, k[0] + o, k[1] + o, k[2] + o, k[3] + o, ll, bad, &pn[u]);
                else if (KIND == OP_JO) jo_fwd_v<T, GR, LADJ, SAFE, DEFER>(v, k[0] + o, k[1] + o, k[2] + o, k[3] + o, ll, bad, &pn[u]);
                else ji_fwd_v<T, GR, LADJ, SAFE, DEFER>(v, k[0] + o, k[1] + o, k[2] + o, k[3] + o, ll, bad, &pn[u]);
            }
        }
    }
}

template <class C, int KIND, bool LADJ, bool SAFE>
__device__ __forceinline__ void elem_fwd_from(const typename C::T* cb, Tile<C>& t, typename C::T (&l)[C::SPT][C::LN],
                                              bool& bad) {
    using T = typename C::T;
    using P = Prim<T>;
    // one log per op and lane-sample (product over all CH vectors of the lane) on the fast path
    constexpr bool DEFER = LADJ && !SAFE && !C::PACKED && C::CH > 1 && KIND != OP_SS;
    T pn[C::SPT], pd[C::SPT];
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) pn[u] = pd[u] = T(1);
#pragma unroll
    for (int q = 0; q < C::CH; ++q) {
        T k[MAX_FWD_CONSTS][C::VE];
#pragma unroll
        for (int j = 0; j < n_fwd_consts(KIND); ++j) ld16_shared(cb + fwd_const_slot(KIND, j) * C::DP + const_off<C>(q), k[j]);
        elem_fwd_q<C, KIND, LADJ, SAFE, DEFER>(t, q, k, l, bad, pn, pd);
    }
    if (DEFER) {
#pragma unroll
        for (int u = 0; u < C::SPT; ++u) {
            const T L = P::lg(KIND == OP_CS ? pn[u] * P::rcp(pd[u]) : pn[u]);
            bad = bad || !(P::abs_(L) < P::LG_SAFE);
            l[u][0] += (KIND == OP_JO && sizeof(T) == 8) ? T(-0.5) * L : L;
        }
    }
}

// interpretive dispatch: the op list is data (ChainDesc), constants come from shared memory
template <class C, bool LADJ, bool SAFE>
__device__ __forceinline__ void apply_op_fwd(const DevOp& op, const typename C::T* s_c, Tile<C>& t,
                                             typename C::T (&l)[C::SPT][C::LN], bool& bad) {
    using T = typename C::T;
    const T* cb = s_c + op.coff;
    switch (op.kind) {
        case OP_HH:
#pragma unroll HH_UNROLL
            for (int k = 0; k < op.K; ++k) {
                T vk[C::CH][C::VE];
#pragma unroll
                for (int q = 0; q < C::CH; ++q) ld16_shared(cb + k * C::DP + const_off<C>(q), vk[q]);
                hh_reflect<C>(t, vk);
            }
            break;
        case OP_SS: elem_fwd_from<C, OP_SS, LADJ, SAFE>(cb, t, l, bad); break;
        case OP_CS: elem_fwd_from<C, OP_CS, LADJ, SAFE>(cb, t, l, bad); break;
        case OP_CC: elem_fwd_from<C, OP_CC, LADJ, SAFE>(cb, t, l, bad); break;
        case OP_JO: elem_fwd_from<C, OP_JO, LADJ, SAFE>(cb, t, l, bad); break;
        default: elem_fwd_from<C, OP_JI, LADJ, SAFE>(cb, t, l, bad); break;
    }
}

template <class C>
__device__ __forceinline__ void stage_constants(const ChainDesc& desc, const typename C::T* consts, typename C::T* s_c) {
    for (int i = threadIdx.x; i < desc.n_consts; i += blockDim.x) s_c[i] = consts[i];
    __syncthreads();
}

template <class C>
__device__ __forceinline__ int64_t num_tiles(int64_t N) {
    const int64_t items = C::PACKED ? ((N + C::LN - 1) / C::LN) : N;
    const int64_t per_tile = int64_t(C::SB) * C::SPT;
    return (items + per_tile - 1) / per_tile;
}

template <typename T, int LN>
__device__ __forceinline__ void store_ladj_vec16(T* dst, const T (&v)[LN]) {
    if constexpr (sizeof(T) == 4 && LN == 4) __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
    else if constexpr (sizeof(T) == 8 && LN == 2) __stcs(reinterpret_cast<double2*>(dst), make_double2(v[0], v[1]));
}
template <typename T, int LN>
__device__ __forceinline__ void store_ladj_vec8(T* dst, const T (&v)[LN]) {
    if constexpr (sizeof(T) == 4 && LN == 2) __stcs(reinterpret_cast<float2*>(dst), make_float2(v[0], v[1]));
}

// per-sample ladj: finish the sum over the lanes of a group, convert from lg units, add the row constants
template <class C, bool W = false>
__device__ __forceinline__ void store_ladj(typename C::T* ladj, int64_t tile, typename C::T (&l)[C::SPT][C::LN],
                                           const int (&nv)[C::SPT], typename C::T ladj_const) {
    using T = typename C::T;
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        const int64_t s = tile_item<C, W>(tile, u);
        if (C::PACKED) {
            T vals[C::LN];
#pragma unroll
            for (int p = 0; p < C::LN; ++p) vals[p] = Prim<T>::fma_(l[u][p], Prim<T>::LGU, ladj_const);
            T* dst = ladj + s * C::LN;
            // the LN values of one vector are consecutive: one 16- or 8-byte store when the row is aligned
            if (C::LN * sizeof(T) == 16 && nv[u] == C::LN && (reinterpret_cast<uintptr_t>(ladj) & 15u) == 0) {
                store_ladj_vec16(dst, vals);
            } else if (C::LN * sizeof(T) == 8 && sizeof(T) == 4 && nv[u] == C::LN && (reinterpret_cast<uintptr_t>(ladj) & 7u) == 0) {
                store_ladj_vec8(dst, vals);
            } else {
#pragma unroll
                for (int p = 0; p < C::LN; ++p)
                    if (p < nv[u]) __stcs(dst + p, vals[p]);
            }
        } else {
            const T tot = group_sum<C>(l[u][0]);
            if (nv[u] && (threadIdx.x & (C::G - 1)) == 0) __stcs(ladj + s, Prim<T>::fma_(tot, Prim<T>::LGU, ladj_const));
        }
    }
}

// ------------------------------------------------------------------ TMA input ring (MODE_VEC)
// A CTA's tile (SPT*SB consecutive samples) is one contiguous block of global
// memory: one elected thread brings it into shared memory with a single bulk
// async copy (cp.async.bulk, the 1-D form of TMA) that signals an mbarrier, RING
// tiles ahead of the math.  Every thread then picks its own 16-byte vectors out of
// the staged tile with LDS.128, so no warp ever waits on HBM latency and no
// registers are tied up by loads in flight.
#ifndef ENF_RING
#define ENF_RING 2   // 2 x 32 KB stages per CTA (8 vectors per thread), 3 CTAs per SM
#endif
constexpr int RING = ENF_RING;

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef ENF_MBAR_SUSPEND_NS
#define ENF_MBAR_SUSPEND_NS 100000   // let the hardware park a waiting warp instead of spinning through issue slots
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(uint32_t(ENF_MBAR_SUSPEND_NS))
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <class C>
struct Ring {
    using T = typename C::T;
    static constexpr bool ON = (C::MODE == MODE_VEC);
    static constexpr int TILE_SAMPLES = C::SPT * C::SB;
    static constexpr size_t STAGE_BYTES = ON ? size_t(TILE_SAMPLES) * C::DP * sizeof(T) : 0;   // D <= DP
    static constexpr size_t BYTES = ON ? RING * STAGE_BYTES + 128 : 0;                          // + barriers
    };

// issue the bulk copy of tile `tile` into ring slot `slot` (one thread)
template <class C>
__device__ __forceinline__ void ring_issue(const typename C::T* x, int64_t N, int D, int64_t tile, unsigned char* stage0,
                                           uint64_t* bars, int slot) {
    using T = typename C::T;
    const int64_t first = tile * Ring<C>::TILE_SAMPLES;
    int64_t n = N - first;
    if (n > Ring<C>::TILE_SAMPLES) n = Ring<C>::TILE_SAMPLES;
    const uint32_t bytes = uint32_t(n) * uint32_t(D) * uint32_t(sizeof(T));
    mbar_expect_tx(&bars[slot], bytes);
    bulk_g2s(stage0 + size_t(slot) * Ring<C>::STAGE_BYTES, x + first * D, bytes, &bars[slot]);
}

// this thread's vectors of the staged tile -> registers
template <class C>
__device__ __forceinline__ void ring_read(const unsigned char* stage, int64_t N, int D, int64_t tile, Tile<C>& t,
                                          int (&nv)[C::SPT]) {
    using T = typename C::T;
    const T* sp = reinterpret_cast<const T*>(stage);
    const int g = threadIdx.x & (C::G - 1);
    const int s_in = threadIdx.x >> C::LG;
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        nv[u] = tile_item<C>(tile, u) < N ? 1 : 0;
#pragma unroll
        for (int q = 0; q < C::CH; ++q) {
            const int row0 = (q * C::G + g) * C::VE;
            if (nv[u] && row0 < D) ld16_shared(sp + (u * C::SB + s_in) * D + row0, t.v[u][q]);
            else {
#pragma unroll
                for (int e = 0; e < C::VE; ++e) t.v[u][q][e] = T(0);
            }
        }
    }
}

// Fast paths for a tile whose samples all exist (every tile but the last): no per-sample predicates, one
// base address per tile, immediate offsets per vector.
template <class C>
__device__ __forceinline__ void ring_read_full(uint32_t stage, int D, Tile<C>& t) {
    using T = typename C::T;
    const int g = threadIdx.x & (C::G - 1);
    const int s_in = threadIdx.x >> C::LG;
    const uint32_t base = stage + uint32_t((s_in * D + g * C::VE) * int(sizeof(T)));
    const uint32_t ustride = uint32_t(C::SB * D * int(sizeof(T)));
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
#pragma unroll
        for (int q = 0; q < C::CH; ++q) {
            if ((q * C::G + g) * C::VE < D) lds16(base + u * ustride + uint32_t(q * C::G * C::VE * int(sizeof(T))), t.v[u][q]);
            else {
#pragma unroll
                for (int e = 0; e < C::VE; ++e) t.v[u][q][e] = T(0);
            }
        }
    }
}

template <class C>
__device__ __forceinline__ void store_tile_full(typename C::T* y, int D, int64_t tile, const Tile<C>& t) {
    using T = typename C::T;
    const int g = threadIdx.x & (C::G - 1);
    T* yb = y + tile_item<C>(tile, 0) * D + g * C::VE;
    const int ustride = C::SB * D;
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
#pragma unroll
        for (int q = 0; q < C::CH; ++q)
            if ((q * C::G + g) * C::VE < D) st16_stream(yb + u * ustride + q * C::G * C::VE, t.v[u][q]);
    }
}

template <class C>
__device__ __forceinline__ void store_ladj_full(typename C::T* ladj, int64_t tile, typename C::T (&l)[C::SPT][C::LN],
                                                typename C::T ladj_const) {
    using T = typename C::T;
    T* lb = ladj + tile_item<C>(tile, 0);
#pragma unroll
    for (int u = 0; u < C::SPT; ++u) {
        const T tot = group_sum<C>(l[u][0]);
        if ((threadIdx.x & (C::G - 1)) == 0) __stcs(lb + u * C::SB, Prim<T>::fma_(tot, Prim<T>::LGU, ladj_const));
    }
}

// ------------------------------------------------------------------ forward (+ ladj) kernels
// Grid-stride loop over tiles shared by the interpretive and the static kernel;
// `apply(t, l)` runs the chain on one register tile.  `ring_smem` is 128-byte
// aligned dynamic shared memory of Ring<C>::BYTES bytes (unused unless MODE_VEC).
template <class C, bool LADJ, class Apply>
__device__ __forceinline__ void fwd_tile_loop(const typename C::T* x, typename C::T* y, typename C::T* ladj, int64_t N,
                                              int D, typename C::T ladj_const, unsigned char* ring_smem, Apply&& apply) {
    using T = typename C::T;
    const int64_t nt = num_tiles<C>(N);
    unsigned char* stage0 = ring_smem;
    const uint32_t stage0_u32 = smem_u32(ring_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring_smem + RING * Ring<C>::STAGE_BYTES);   // TMA landed
    uint64_t* empty = full + RING;                                                            // all warps have read
    if (Ring<C>::ON) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < RING; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], NT / 32);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < RING; ++i) {
                const int64_t tl = int64_t(blockIdx.x) + int64_t(i) * gridDim.x;
                if (tl < nt) ring_issue<C>(x, N, D, tl, stage0, full, i);
            }
        }
    }
    int k = 0;
    for (int64_t tile = blockIdx.x; tile < nt; tile += gridDim.x, ++k) {
        Tile<C> t;
        int nv[C::SPT];
        T l[C::SPT][C::LN];
        // MODE_VEC: every tile but the last is complete -> predicate-free loads and stores
        const bool full_tile = Ring<C>::ON && (tile + 1) * Ring<C>::TILE_SAMPLES <= N;
        if (Ring<C>::ON) {
            const int slot = k % RING;
            const uint32_t parity = uint32_t(k / RING) & 1u;
            // thread 0 refills the slot every warp finished reading one iteration ago (tile k-1 -> tile k-1+RING):
            // it only ever waits for warps that are more than a whole tile behind
            if (threadIdx.x == 0 && k >= 1) {
                const int64_t nxt = tile + int64_t(RING - 1) * gridDim.x;
                if (nxt < nt) {
                    const int ps = (k - 1) % RING;
                    while (!mbar_try_wait(&empty[ps], uint32_t((k - 1) / RING) & 1u)) {}
                    ring_issue<C>(x, N, D, nxt, stage0, full, ps);
                }
            }
            while (!mbar_try_wait(&full[slot], parity)) {}
            if (full_tile) {
                ring_read_full<C>(stage0_u32 + uint32_t(slot) * uint32_t(Ring<C>::STAGE_BYTES), D, t);
#pragma unroll
                for (int u = 0; u < C::SPT; ++u) nv[u] = 1;
            } else {
                ring_read<C>(stage0 + size_t(slot) * Ring<C>::STAGE_BYTES, N, D, tile, t, nv);
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);
        } else {
            load_tile<C>(x, N, D, tile, t, nv);
        }
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int p = 0; p < C::LN; ++p) l[u][p] = T(0);
        bool bad = false;
        apply(std::false_type{}, t, l, bad);
        if (LADJ && __any_sync(0xffffffffu, bad)) {
            // a Jacobian-factor product left the float range somewhere in this warp: redo the tile
            // with per-element logs (x is still intact: the outputs have not been stored yet)
            load_tile<C>(x, N, D, tile, t, nv);
#pragma unroll
            for (int u = 0; u < C::SPT; ++u)
#pragma unroll
                for (int p = 0; p < C::LN; ++p) l[u][p] = T(0);
            apply(std::true_type{}, t, l, bad);
        }
        if (full_tile) {
            store_tile_full<C>(y, D, tile, t);
            if (LADJ) store_ladj_full<C>(ladj, tile, l, ladj_const);
        } else {
            store_tile<C>(y, D, tile, t, nv);
            if (LADJ) store_ladj<C>(ladj, tile, l, nv, ladj_const);
        }
    }
}


template <class C>
__device__ __forceinline__ unsigned char* ring_base(unsigned char* after_consts) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(after_consts) + 127) & ~uintptr_t(127));
}

#ifndef ENF_GRAD_SAFE_ONLY
#define ENF_GRAD_SAFE_ONLY 0
#endif
template <class C> using FwdRing = Ring<C>;
#define ENF_FWD_LOOP fwd_tile_loop

// F1/F2 of SURVEY §2.3: (f::Trafo)(x) and with_logabsdet_jacobian(f, x) for a whole chain.
#ifndef ENF_FWD_MIN_CTAS
#define ENF_FWD_MIN_CTAS 3
#endif
template <class C, bool LADJ>
__global__ void __launch_bounds__(NT, ENF_FWD_MIN_CTAS) chain_fwd_kernel(const __grid_constant__ ChainDesc desc,
                                                                         const typename C::T* __restrict__ consts,
                                                                         const typename C::T* x, typename C::T* y,
                                                                         typename C::T* ladj, int64_t N,
                                                                         typename C::T ladj_const) {
    using T = typename C::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* s_c = reinterpret_cast<T*>(smem_raw);
    stage_constants<C>(desc, consts, s_c);
    ENF_FWD_LOOP<C, LADJ>(x, y, ladj, N, desc.D, ladj_const, ring_base<C>(smem_raw + size_t(desc.n_consts) * sizeof(T)),
                           [&](auto safe, Tile<C>& t, T (&l)[C::SPT][C::LN], bool& bad) {
                               constexpr bool SAFE = decltype(safe)::value;
                               for (int o = 0; o < desc.n_ops; ++o) apply_op_fwd<C, LADJ, SAFE>(desc.ops[o], s_c, t, l, bad);
                           });
}

// backward of one elementwise op on vector q of a tile: xt = the op's input, yt = its output,
// gt <- input cotangent, ra[k][e] += m[u] * raw integrand k  (FULL: every sample of the tile exists, no mask)
template <class C, int KIND, bool FULL>
__device__ __forceinline__ void bwd_elem(const Tile<C>& xt, const Tile<C>& yt, Tile<C>& gt, int q,
                                         const typename C::T (&m)[C::SPT][C::LN],
                                         const typename C::T (&c0)[C::VE], const typename C::T (&c1)[C::VE],
                                         const typename C::T (&c2)[C::VE], const typename C::T (&c3)[C::VE],
                                         const typename C::T (&c4)[C::VE], const typename C::T (&c5)[C::VE],
                                         typename C::T (&ra)[4][C::VE]) {
    using T = typename C::T;
    constexpr int NR = n_rowslots_of(KIND, 0);
#if ENF_F32X2
    if constexpr (sizeof(T) == 4) {
        // two rows per FP32 instruction: the same templates instantiated for the pair type F2
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int e = 0; e < C::VE; e += 2) {
                auto pr = [&](const float (&a)[C::VE]) { return F2(a[e], a[e + 1]); };
                F2 r[4] = {F2(0.f), F2(0.f), F2(0.f), F2(0.f)};
                const F2 xin(xt.v[u][q][e], xt.v[u][q][e + 1]), yout(yt.v[u][q][e], yt.v[u][q][e + 1]);
                const F2 G(gt.v[u][q][e], gt.v[u][q][e + 1]);
                F2 gx;
                if (KIND == OP_SS) {
                    r[0] = G * xin;
                    r[1] = G;
                    gx = G * pr(c0);
                } else if (KIND == OP_CS) {
                    gx = cs_bwd<F2>(xin, yout, G, pr(c0), pr(c1), pr(c2), pr(c3), pr(c4), pr(c5), r);
                } else if (KIND == OP_CC) {
                    gx = cc_bwd<F2>(xin, yout, G, pr(c0), pr(c1), pr(c2), pr(c3), pr(c4), pr(c5), r);
                } else if (KIND == OP_JO) {
                    gx = jo_bwd<F2>(xin, yout, G, pr(c0), pr(c1), pr(c4), r);
                } else {
                    gx = ji_bwd<F2>(xin, yout, G, pr(c0), pr(c1), pr(c2), pr(c3), pr(c4), pr(c5), r);
                }
                gt.v[u][q][e] = gx.v.x;
                gt.v[u][q][e + 1] = gx.v.y;
#pragma unroll
                for (int k = 0; k < NR; ++k) {
                    float2 acc;
                    if (u == 0) acc = FULL ? r[k].v : mul2(make_float2(m[u][C::slot(e)], m[u][C::slot(e + 1)]), r[k].v);   // ra starts here
                    else if (FULL) acc = add2(r[k].v, make_float2(ra[k][e], ra[k][e + 1]));
                    else acc = fma2(make_float2(m[u][C::slot(e)], m[u][C::slot(e + 1)]), r[k].v, make_float2(ra[k][e], ra[k][e + 1]));
                    ra[k][e] = acc.x;
                    ra[k][e + 1] = acc.y;
                }
            }
        return;
    }
#endif
#pragma unroll
    for (int u = 0; u < C::SPT; ++u)
#pragma unroll
        for (int e = 0; e < C::VE; ++e) {
            T r[4] = {T(0), T(0), T(0), T(0)};
            const T xin = xt.v[u][q][e], yout = yt.v[u][q][e], G = gt.v[u][q][e];
            T gx;
            if (KIND == OP_SS) {
                r[0] = G * xin;
                r[1] = G;
                gx = G * c0[e];
            } else if (KIND == OP_CS) {
                gx = cs_bwd<T>(xin, yout, G, c0[e], c1[e], c2[e], c3[e], c4[e], c5[e], r);
            } else if (KIND == OP_CC) {
                gx = cc_bwd<T>(xin, yout, G, c0[e], c1[e], c2[e], c3[e], c4[e], c5[e], r);
            } else if (KIND == OP_JO) {
                gx = jo_bwd<T>(xin, yout, G, c0[e], c1[e], c4[e], r);
            } else {
                gx = ji_bwd<T>(xin, yout, G, c0[e], c1[e], c2[e], c3[e], c4[e], c5[e], r);
            }
            gt.v[u][q][e] = gx;
#pragma unroll
            for (int k = 0; k < NR; ++k)
                ra[k][e] = u == 0 ? (FULL ? r[k] : m[u][C::slot(e)] * r[k])
                                  : (FULL ? ra[k][e] + r[k] : Prim<T>::fma_(m[u][C::slot(e)], r[k], ra[k][e]));
        }
}

// ... and the tile's contribution to the op's per-thread raw sums (the number of row slots is a compile-time property of
// the op kind here: no zero-initialised spare slots, no predicated updates)
template <class C, int KIND, bool FULL>
__device__ __forceinline__ void bwd_elem_acc(const Tile<C>& xt, const Tile<C>& yt, Tile<C>& gt, int q,
                                             const typename C::T (&m)[C::SPT][C::LN],
                                             const typename C::T (&c0)[C::VE], const typename C::T (&c1)[C::VE],
                                             const typename C::T (&c2)[C::VE], const typename C::T (&c3)[C::VE],
                                             const typename C::T (&c4)[C::VE], const typename C::T (&c5)[C::VE], uint32_t acc0) {
    using T = typename C::T;
    constexpr int NR = n_rowslots_of(KIND, 0);
    T ra[4][C::VE];                      // written by the first sample of the tile (bwd_elem), slots >= NR never touched
    bwd_elem<C, KIND, FULL>(xt, yt, gt, q, m, c0, c1, c2, c3, c4, c5, ra);
#pragma unroll
    for (int k = 0; k < NR; ++k) acc16_shared<T, C::VE>(acc0 + uint32_t(k * C::CH) * uint32_t(NT * 16), ra[k]);
}

// -log p(z) and its derivative for the Gaussian-mixture target of the ELBO objective (log-sum-exp over the components).
// Deliberately out of line: the whitening loss never calls it, and it must not cost the hot path any registers.
template <typename T> struct ValGrad { T val, grad; };
template <typename T>
__device__ __noinline__ ValGrad<T> target_eval(const ChainDesc& d, T z) {
    T e[MAX_TARGET_K];
    T mx = T(-1e30);
    for (int k = 0; k < d.target_K; ++k) {
        const T t = (z - T(d.tmu[k])) * T(d.tis[k]);
        e[k] = T(d.tlw[k]) - T(0.5) * t * t;
        mx = e[k] > mx ? e[k] : mx;
    }
    T s = T(0), gsum = T(0);
    for (int k = 0; k < d.target_K; ++k) {
        const T p = sizeof(T) == 4 ? T(expf(float(e[k] - mx))) : T(exp(double(e[k] - mx)));
        s += p;
        gsum += p * (z - T(d.tmu[k])) * T(d.tis[k]) * T(d.tis[k]);
    }
    ValGrad<T> r;
    r.val = -(mx + (sizeof(T) == 4 ? T(logf(float(s))) : T(log(double(s)))));
    r.grad = gsum / s;
    return r;
}

// ------------------------------------------------------------------ loss / gradient kernel
// F3 of SURVEY §2.3: mvnormal_negll_trafo and the reverse pass of
// mvnormal_negll_trafograd (src/optimize_whitening.jl:7-22) in one pass over x.
// Per CTA it emits `n_raw` float64 partial sums:
//   [n_rowslots][Dp] per-row raw sums | [n_scalars] | sum_j 1/2 |y_j|^2 | sum_j ladj_j (variable part)
// which reduce_partials_kernel adds up in a fixed order (bitwise reproducible).
//
// Shared memory: constants | saved op inputs [n_save][TV][NT] x 16 B | accumulators [n_rowslots][CH][NT] x 16 B
// | scalar accumulators [n_scalars][NT].  Every per-thread datum is a 16-byte vector at [..][tid]:
// LDS.128 / STS.128, conflict-free.
template <typename T>
struct GradSmem {
    const T* c;      // constants
    uint32_t c32;    //   ... as a 32-bit shared-window address
    uint32_t save;   // 32-bit shared-window addresses of THIS thread's 16-byte slot [..][tid] (immediate offsets per vector):
    uint32_t acc;    //   saved inputs of the elementwise ops / per-row raw-sum accumulators
    T* sc;           // per-thread scalar accumulators
};

// Forward pass of one tile (saving the input of every elementwise op when GRAD).  FULL: every sample of the tile
// exists (MODE_VEC): predicate-free loads, no masks.  Returns the range flag of the fast ladj path (!SAFE).
template <class C, bool GRAD, bool FULL, bool SAFE>
__device__ __forceinline__ bool grad_fwd_tile(const ChainDesc& desc, const GradSmem<typename C::T>& sm, const typename C::T* x,
                                              int64_t N, int64_t tile, Tile<C>& zt, typename C::T (&l)[C::SPT][C::LN],
                                              typename C::T (&m)[C::SPT][C::LN]) {
    using T = typename C::T;
    constexpr int VE = C::VE;
    constexpr int TV = C::SPT * C::CH;
    const int tid = threadIdx.x;
    const int D = desc.D;
    if constexpr (FULL) {
        const int g = tid & (C::G - 1);
        const T* xb = x + tile_item<C>(tile, 0) * D + g * VE;
        const int ustride = C::SB * D;
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int q = 0; q < C::CH; ++q) {
                if ((q * C::G + g) * VE < D) ld16_stream(xb + u * ustride + q * C::G * VE, zt.v[u][q]);
                else {
#pragma unroll
                    for (int e = 0; e < VE; ++e) zt.v[u][q][e] = T(0);
                }
            }
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int p = 0; p < C::LN; ++p) l[u][p] = T(0);
    } else {
        int nv[C::SPT];
        load_tile<C>(x, N, D, tile, zt, nv);
#pragma unroll
        for (int u = 0; u < C::SPT; ++u)
#pragma unroll
            for (int p = 0; p < C::LN; ++p) {
                m[u][p] = p < nv[u] ? T(1) : T(0);
                l[u][p] = T(0);
            }
    }
    bool bad = false;
    for (int o = 0; o < desc.n_ops; ++o) {
        const DevOp op = desc.ops[o];
        if (GRAD && op.save >= 0) {
            const uint32_t sv = sm.save + uint32_t(op.save) * uint32_t(TV * NT * 16);
#pragma unroll
            for (int u = 0; u < C::SPT; ++u)
#pragma unroll
                for (int q = 0; q < C::CH; ++q) sts16(sv + uint32_t((u * C::CH + q) * NT * 16), zt.v[u][q]);
        }
        apply_op_fwd<C, true, SAFE>(op, sm.c, zt, l, bad);
    }
    return bad;
}

// Reverse pass of one tile: gt = N dL/d(activation), seeded by the caller (y for the whitening loss,
// src/optimize_whitening.jl:12; -dlog p/dz for the ELBO objective); zt holds the chain's output on entry and is walked
// back to the input.
template <class C, bool FULL>
__device__ __forceinline__ void grad_bwd_tile(const ChainDesc& desc, const GradSmem<typename C::T>& sm, Tile<C>& zt, Tile<C>& gt,
                                              const typename C::T (&m)[C::SPT][C::LN]) {
    using T = typename C::T;
    using P = Prim<T>;
    constexpr int VE = C::VE;
    constexpr int TV = C::SPT * C::CH;
    constexpr int Dp = C::DP;
    const int tid = threadIdx.x;
    const int g = tid & (C::G - 1);
    for (int o = desc.n_ops - 1; o >= 0; --o) {
        const DevOp op = desc.ops[o];
        const uint32_t cb = sm.c32 + uint32_t(op.coff) * uint32_t(sizeof(T));   // this op's constants
        if (op.kind == OP_HH) {
            // reverse sweep with recomputation (src/householder_trafo.jl:88-114)
            for (int k = op.K - 1; k >= 0; --k) {
                T vk[C::CH][VE];
#pragma unroll
                for (int q = 0; q < C::CH; ++q) lds16(cb + uint32_t((k * Dp + const_off<C>(q)) * int(sizeof(T))), vk[q]);
                const uint32_t acc = sm.acc + uint32_t(op.roff + k) * uint32_t(C::CH * NT * 16);
                T* sc = sm.sc + size_t(op.soff + k) * NT + tid;
                T a1[C::CH][VE];
                T a2 = T(0);
#pragma unroll
                for (int q = 0; q < C::CH; ++q)
#pragma unroll
                    for (int e = 0; e < VE; ++e) a1[q][e] = T(0);
                if (C::PACKED) {
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u)
#pragma unroll
                        for (int p = 0; p < C::LN; ++p) {
                            T po = T(0), qd = T(0);
#pragma unroll
                            for (int e = 0; e < C::PD; ++e) {
                                po = P::fma_(vk[0][p * C::PD + e], zt.v[u][0][p * C::PD + e], po);
                                qd = P::fma_(vk[0][p * C::PD + e], gt.v[u][0][p * C::PD + e], qd);
                            }
                            const T pm = -po * m[u][p], qm = qd * m[u][p];
#pragma unroll
                            for (int e = 0; e < C::PD; ++e) {
                                const int i = p * C::PD + e;
                                zt.v[u][0][i] = P::fma_(-po, vk[0][i], zt.v[u][0][i]);
                                a1[0][i] = P::fma_(pm, gt.v[u][0][i], P::fma_(qm, zt.v[u][0][i], a1[0][i]));
                                gt.v[u][0][i] = P::fma_(-qd, vk[0][i], gt.v[u][0][i]);
                            }
                            a2 = P::fma_(pm, qd, a2);
                        }
                } else if constexpr (ENF_F32X2 && sizeof(T) == 4) {
                    // two rows per FFMA2 (same arithmetic as the scalar branch below)
                    float po[C::SPT], qd[C::SPT];
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) {
                        float2 ap = make_float2(0.f, 0.f), aq = make_float2(0.f, 0.f);
#pragma unroll
                        for (int q = 0; q < C::CH; ++q)
#pragma unroll
                            for (int e = 0; e < VE; e += 2) {
                                const float2 v2 = make_float2(vk[q][e], vk[q][e + 1]);
                                ap = fma2(v2, make_float2(zt.v[u][q][e], zt.v[u][q][e + 1]), ap);
                                aq = fma2(v2, make_float2(gt.v[u][q][e], gt.v[u][q][e + 1]), aq);
                            }
                        const float2 pq = add2(make_float2(ap.x, aq.x), make_float2(ap.y, aq.y));
                        po[u] = pq.x;
                        qd[u] = pq.y;
                    }
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) group_sum2<C>(po[u], qd[u]);
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) {
                        const float pm = FULL ? -po[u] : -po[u] * m[u][0], qm = FULL ? qd[u] : qd[u] * m[u][0];
                        const float2 npo = make_float2(-po[u], -po[u]), nqd = make_float2(-qd[u], -qd[u]);
                        const float2 pm2 = make_float2(pm, pm), qm2 = make_float2(qm, qm);
#pragma unroll
                        for (int q = 0; q < C::CH; ++q)
#pragma unroll
                            for (int e = 0; e < VE; e += 2) {
                                const float2 v2 = make_float2(vk[q][e], vk[q][e + 1]);
                                const float2 g2 = make_float2(gt.v[u][q][e], gt.v[u][q][e + 1]);
                                const float2 z2 = fma2(npo, v2, make_float2(zt.v[u][q][e], zt.v[u][q][e + 1]));  // reflection input
                                const float2 a2v = fma2(pm2, g2, fma2(qm2, z2, make_float2(a1[q][e], a1[q][e + 1])));
                                const float2 gn = fma2(nqd, v2, g2);
                                zt.v[u][q][e] = z2.x; zt.v[u][q][e + 1] = z2.y;
                                a1[q][e] = a2v.x; a1[q][e + 1] = a2v.y;
                                gt.v[u][q][e] = gn.x; gt.v[u][q][e + 1] = gn.y;
                            }
                        a2 = P::fma_(pm, qd[u], a2);
                    }
                } else {
                    T po[C::SPT], qd[C::SPT];
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) {
                        po[u] = T(0);
                        qd[u] = T(0);
#pragma unroll
                        for (int q = 0; q < C::CH; ++q)
#pragma unroll
                            for (int e = 0; e < VE; ++e) {
                                po[u] = P::fma_(vk[q][e], zt.v[u][q][e], po[u]);
                                qd[u] = P::fma_(vk[q][e], gt.v[u][q][e], qd[u]);
                            }
                    }
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) {
                        po[u] = group_sum<C>(po[u]);
                        qd[u] = group_sum<C>(qd[u]);
                    }
#pragma unroll
                    for (int u = 0; u < C::SPT; ++u) {
                        const T pm = FULL ? -po[u] : -po[u] * m[u][0], qm = FULL ? qd[u] : qd[u] * m[u][0];
#pragma unroll
                        for (int q = 0; q < C::CH; ++q)
#pragma unroll
                            for (int e = 0; e < VE; ++e) {
                                zt.v[u][q][e] = P::fma_(-po[u], vk[q][e], zt.v[u][q][e]);  // reflection input
                                a1[q][e] = P::fma_(pm, gt.v[u][q][e], P::fma_(qm, zt.v[u][q][e], a1[q][e]));
                                gt.v[u][q][e] = P::fma_(-qd[u], vk[q][e], gt.v[u][q][e]);
                            }
                        a2 = P::fma_(pm, qd[u], a2);
                    }
                }
#pragma unroll
                for (int q = 0; q < C::CH; ++q) acc16_shared<T, VE>(acc + uint32_t(q * NT * 16), a1[q]);
                if (g == 0) *sc += a2;
            }
            continue;
        }
        // elementwise op: zt holds its output; reload its input, differentiate
        Tile<C> xt;
        {
            const uint32_t sv = sm.save + uint32_t(op.save) * uint32_t(TV * NT * 16);
#pragma unroll
            for (int u = 0; u < C::SPT; ++u)
#pragma unroll
                for (int q = 0; q < C::CH; ++q) lds16(sv + uint32_t((u * C::CH + q) * NT * 16), xt.v[u][q]);
        }
#pragma unroll
        for (int q = 0; q < C::CH; ++q) {
            const int co = const_off<C>(q);
            T c0[VE], c1[VE], c2[VE], c3[VE], c4[VE], c5[VE];
            lds16(cb + uint32_t((0 * Dp + co) * int(sizeof(T))), c0);
            lds16(cb + uint32_t((1 * Dp + co) * int(sizeof(T))), c1);
            if (op.kind != OP_SS) {
                lds16(cb + uint32_t((2 * Dp + co) * int(sizeof(T))), c2);
                lds16(cb + uint32_t((3 * Dp + co) * int(sizeof(T))), c3);
                lds16(cb + uint32_t((4 * Dp + co) * int(sizeof(T))), c4);
            }
            if (op.kind == OP_CS || op.kind == OP_CC || op.kind == OP_JI) lds16(cb + uint32_t((5 * Dp + co) * int(sizeof(T))), c5);
            const uint32_t acc0 = sm.acc + uint32_t(op.roff * C::CH + q) * uint32_t(NT * 16);   // slot k of vector q: + k CH NT 16
            switch (op.kind) {
                case OP_SS: bwd_elem_acc<C, OP_SS, FULL>(xt, zt, gt, q, m, c0, c1, c2, c3, c4, c5, acc0); break;
                case OP_CS: bwd_elem_acc<C, OP_CS, FULL>(xt, zt, gt, q, m, c0, c1, c2, c3, c4, c5, acc0); break;
                case OP_CC: bwd_elem_acc<C, OP_CC, FULL>(xt, zt, gt, q, m, c0, c1, c2, c3, c4, c5, acc0); break;
                case OP_JO: bwd_elem_acc<C, OP_JO, FULL>(xt, zt, gt, q, m, c0, c1, c2, c3, c4, c5, acc0); break;
                default: bwd_elem_acc<C, OP_JI, FULL>(xt, zt, gt, q, m, c0, c1, c2, c3, c4, c5, acc0); break;
            }
        }
        zt = xt;   // the input of this op is the output of the previous one
    }
}

#ifndef ENF_GRAD_FLUSH_TILES
#define ENF_GRAD_FLUSH_TILES 64   // Float32 per-thread accumulation depth: 64 tiles x SPT samples, then folded into float64
#endif
constexpr int GRAD_FLUSH_TILES = ENF_GRAD_FLUSH_TILES;
#ifndef ENF_GRAD_MIN_CTAS
#define ENF_GRAD_MIN_CTAS 2
#endif
template <class C, bool GRAD>
__global__ void __launch_bounds__(NT, ENF_GRAD_MIN_CTAS) chain_grad_kernel(const __grid_constant__ ChainDesc desc,
                                                        const typename C::T* __restrict__ consts,
                                                        const typename C::T* __restrict__ x, int64_t N,
                                                        double* __restrict__ partials) {
    using T = typename C::T;
    using P = Prim<T>;
    constexpr int VE = C::VE;
    constexpr int TV = C::SPT * C::CH;  // 16-byte vectors per thread per tile
    constexpr bool HOTFULL = (C::MODE == MODE_VEC);   // the hot instantiation has no masks: the (one) ragged tile takes the masked path
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* s_c = reinterpret_cast<T*>(smem_raw);
    const int n_consts_al = (desc.n_consts + 3) & ~3;
    const int tid = threadIdx.x;
    T* s_save = s_c + n_consts_al;
    T* s_acc = s_save + (GRAD ? size_t(desc.n_save) * TV * NT * VE : 0);
    GradSmem<T> sm;
    sm.c = s_c;
    sm.c32 = smem_u32(s_c);
    sm.save = smem_u32(s_save) + uint32_t(tid) * 16u;
    sm.acc = smem_u32(s_acc) + uint32_t(tid) * 16u;
    sm.sc = s_acc + (GRAD ? size_t(desc.n_rowslots) * C::CH * NT * VE : 0);
    // Programmatic dependent launch (the device-side fit loop chains gradient kernel -> update kernel -> gradient kernel ...
    // with the programmatic-serialization launch attribute): do what does not depend on the previous kernel, wait for it
    // before touching what it wrote (the constants), and let the next kernel of the chain be scheduled.  Both instructions
    // are no-ops for ordinary launches.
    if (GRAD) {
        const T zero[VE] = {};
        for (int i = 0; i < desc.n_rowslots * C::CH; ++i) st16_shared(s_acc + (size_t(i) * NT + tid) * VE, zero);
        for (int i = 0; i < desc.n_scalars; ++i) sm.sc[size_t(i) * NT + tid] = T(0);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // only now: the update kernel launched behind this one may read the parameters before ITS wait, which is safe once
    // the previous update kernel (the one this kernel has just waited for) is complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    stage_constants<C>(desc, consts, s_c);
    constexpr int Dp = C::DP;
    T loss_y = T(0), loss_l = T(0);
    double loss_y64 = 0.0, loss_l64 = 0.0;
    const int64_t nt = num_tiles<C>(N);
    const int64_t n_items = C::PACKED ? ((N + C::LN - 1) / C::LN) : N;
    const int n_raw = desc.n_rowslots * Dp + desc.n_scalars + 2;
    double* out = partials + size_t(blockIdx.x) * n_raw;
    // The per-thread accumulators are Float32 for Float32 chains.  To keep their depth bounded however many tiles a CTA
    // visits (the grid is capped at a few CTAs per SM), they are folded into the float64 partial sums of this CTA - in
    // a fixed order, so the result stays bitwise reproducible - every GRAD_FLUSH_TILES tiles and cleared.
    bool first_flush = true;
    auto flush = [&](bool first) {
        __syncthreads();
        if (GRAD) {
            for (int pi = tid; pi < desc.n_rowslots * Dp; pi += NT) {
                const int rs = pi / Dp, row = pi - rs * Dp;
                const int vecidx = row / VE, e = row - vecidx * VE;
                int q, gg;
                if (C::PACKED) { q = 0; gg = 0; }
                else { q = vecidx >> C::LG; gg = vecidx & (C::G - 1); }
                const T* acc = s_acc + (size_t(rs) * C::CH + q) * NT * VE + e;
                double sacc = 0.0;
                for (int j = gg; j < NT; j += C::G) sacc += double(acc[size_t(j) * VE]);
                out[pi] = first ? sacc : out[pi] + sacc;
            }
            for (int k = tid; k < desc.n_scalars; k += NT) {
                const T* sc = sm.sc + size_t(k) * NT;
                double sacc = 0.0;
                for (int j = 0; j < NT; j += C::G) sacc += double(sc[j]);
                double* o = out + desc.n_rowslots * Dp + k;
                *o = first ? sacc : *o + sacc;
            }
            __syncthreads();
            const T zero[VE] = {};
            for (int i = 0; i < desc.n_rowslots * C::CH; ++i) st16_shared(s_acc + (size_t(i) * NT + tid) * VE, zero);
            for (int i = 0; i < desc.n_scalars; ++i) sm.sc[size_t(i) * NT + tid] = T(0);
        } else if (first) {
            for (int pi = tid; pi < desc.n_rowslots * Dp + desc.n_scalars; pi += NT) out[pi] = 0.0;
        }
        loss_y64 += double(loss_y);
        loss_l64 += double(loss_l);
        loss_y = T(0);
        loss_l = T(0);
    };
    int visited = 0;
    for (int64_t tile = blockIdx.x; tile < nt; tile += gridDim.x, ++visited) {
        if (visited != 0 && (visited % GRAD_FLUSH_TILES) == 0) {
            flush(first_flush);
            first_flush = false;
        }
        Tile<C> zt;
        T l[C::SPT][C::LN];
        T m[C::SPT][C::LN];  // 1 for real samples, 0 for the padding of the last tile (unused on full tiles)
        const bool full = HOTFULL && (tile + 1) * (int64_t(C::SB) * C::SPT) <= n_items;
        // forward with the fast ladj (one log per op and lane-sample); the whole forward again with per-element
        // logs (and masks) if a factor product left the float range; the ragged last tile is always masked
        bool use_mask;
        if (full) {
            use_mask = __any_sync(0xffffffffu, grad_fwd_tile<C, GRAD, HOTFULL, false>(desc, sm, x, N, tile, zt, l, m));
            if (use_mask) grad_fwd_tile<C, GRAD, false, true>(desc, sm, x, N, tile, zt, l, m);
        } else if (HOTFULL) {
            grad_fwd_tile<C, GRAD, false, true>(desc, sm, x, N, tile, zt, l, m);
            use_mask = true;
        } else {
            // layouts without a mask-free path (scalar / packed accesses): fast masked forward first
            if (__any_sync(0xffffffffu, grad_fwd_tile<C, GRAD, false, false>(desc, sm, x, N, tile, zt, l, m)))
                grad_fwd_tile<C, GRAD, false, true>(desc, sm, x, N, tile, zt, l, m);
            use_mask = true;
        }
        Tile<C> gt;
        // The ELBO objective is compiled into the scalar-access layouts only (enf_elbo_grad always dispatches to them): its
        // batches are a few hundred draws, and the hot vectorised instantiations must not pay registers for it.
        constexpr bool HAS_TARGET = (C::MODE == MODE_SCALAR || C::MODE == MODE_PACKU);
        if (HAS_TARGET && desc.target_kind != 0) {
            // loss term -log p(z) per element, cotangent seed -dlog p/dz (cold path, out of line)
            const int g_ = tid & (C::G - 1);
#pragma unroll
            for (int u = 0; u < C::SPT; ++u) {
#pragma unroll
                for (int q = 0; q < C::CH; ++q)
#pragma unroll
                    for (int e = 0; e < VE; ++e) {
                        const bool row_ok = C::PACKED ? true : ((q * C::G + g_) * VE + e < desc.D);
                        const T w = use_mask ? m[u][C::slot(e)] : T(1);
                        const ValGrad<T> r = target_eval<T>(desc, zt.v[u][q][e]);
                        loss_y = P::fma_(row_ok ? w : T(0), r.val, loss_y);
                        gt.v[u][q][e] = row_ok ? r.grad : T(0);
                    }
#pragma unroll
                for (int p = 0; p < C::LN; ++p) loss_l = use_mask ? P::fma_(m[u][p], l[u][p], loss_l) : loss_l + l[u][p];
            }
        } else if (!use_mask) {
#pragma unroll
            for (int u = 0; u < C::SPT; ++u) {
                T sy = T(0);
#pragma unroll
                for (int q = 0; q < C::CH; ++q)
#pragma unroll
                    for (int e = 0; e < VE; ++e) {
                        sy = P::fma_(zt.v[u][q][e], zt.v[u][q][e], sy);
                        gt.v[u][q][e] = zt.v[u][q][e];
                    }
                loss_y = P::fma_(T(0.5), sy, loss_y);
#pragma unroll
                for (int p = 0; p < C::LN; ++p) loss_l += l[u][p];
            }
        } else {
#pragma unroll
            for (int u = 0; u < C::SPT; ++u) {
                T sy = T(0);
#pragma unroll
                for (int q = 0; q < C::CH; ++q)
#pragma unroll
                    for (int e = 0; e < VE; ++e) {
                        sy = P::fma_(m[u][C::slot(e)] * zt.v[u][q][e], zt.v[u][q][e], sy);
                        gt.v[u][q][e] = zt.v[u][q][e];
                    }
                loss_y = P::fma_(T(0.5), sy, loss_y);
#pragma unroll
                for (int p = 0; p < C::LN; ++p) loss_l = P::fma_(m[u][p], l[u][p], loss_l);
            }
        }
        if constexpr (!GRAD) continue;
        if (!use_mask) grad_bwd_tile<C, HOTFULL>(desc, sm, zt, gt, m);
        else grad_bwd_tile<C, false>(desc, sm, zt, gt, m);
    }
    flush(first_flush);
    // loss partials: warp shuffle, then one double per warp through shared memory
    __shared__ double s_loss[2][NT / 32];
    double ly = loss_y64, ll = loss_l64 * double(Prim<T>::LGU);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ly += __shfl_xor_sync(0xffffffffu, ly, off);
        ll += __shfl_xor_sync(0xffffffffu, ll, off);
    }
    if ((tid & 31) == 0) { s_loss[0][tid >> 5] = ly; s_loss[1][tid >> 5] = ll; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < NT / 32; ++w) { a += s_loss[0][w]; b += s_loss[1][w]; }
        out[n_raw - 2] = a;
        out[n_raw - 1] = b;
    }
}

// sums[i] (+)= sum_b partials[b][i], b ascending: deterministic
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int n_blocks, int n_raw,
                                       double* __restrict__ sums, int accumulate);

}  // namespace enf
