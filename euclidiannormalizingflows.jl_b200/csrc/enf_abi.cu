// C ABI of libenf_b200.so (include/enf_b200.h): contexts, device buffers, chain
// construction (derived per-row constants), kernel dispatch, mapping of the raw
// device sums to (negll, parameter gradients), the host-buffer pipeline and the
// NCCL group used for sharded gradient steps.
//
// There is no CPU fallback anywhere in this file: every compute entry point
// launches the sm_100a kernels of enf_chain.cuh or fails with an error code.
#include "enf_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <sys/mman.h>
#include <nccl.h>  // types only; the library is dlopen'ed (see NcclApi)

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "enf_chain.cuh"
#include "enf_launch.h"

namespace enf {
void affine_fold(int D, int n_ops, const int* kinds, const int* Ks, const double* const* params, std::vector<float>& wh,
                 std::vector<float>& wl, std::vector<float>& bias);
}

namespace enf {
bool wy_fold(int D, int n_ops, const int* kinds, const int* Ks, const double* const* params, std::vector<float>& out);
}

using namespace enf;

// ------------------------------------------------------------------ errors
namespace {
thread_local std::string t_last_error;

int fail(enf_ctx* ctx, int code, const char* fmt, ...);

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr double LOG2E_D = 1.4426950408889634073599246810019;
constexpr int HOST_SLOTS = 3;
}  // namespace

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct enf_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t slot_stream[HOST_SLOTS] = {nullptr, nullptr, nullptr};
    void* slot_buf[HOST_SLOTS] = {nullptr, nullptr, nullptr};
    size_t slot_bytes = 0;
    cudaEvent_t ev = nullptr;
    cudaEvent_t timing[ENF_N_EVENTS] = {};
    int64_t launches = 0;
    std::string last_error;
    // pinned host buffers placed on the GPU's NUMA node (enf_host_alloc): base -> mapped bytes
    std::map<void*, size_t> numa_allocs;
    int numa_node = -2;             // -2: not looked up yet, -1: unknown
    // NCCL group
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory all-reduce (enf_p2p.cu): every rank's buffer mapped through CUDA IPC; off -> ncclAllReduce
    bool p2p = false;
    P2PDesc p2p_desc = {};
};

namespace {

int fail(enf_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_last_error = buf;
    if (ctx) ctx->last_error = buf;
    return code;
}

#define CU(ctx, call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(ctx, ENF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

std::mutex g_nccl_mu;
NcclApi g_nccl;

int load_nccl(enf_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return ENF_OK;
    // ENF_NCCL_LIB: a specific libnccl.so.2.  The dynamic linker keeps ONE library per soname and process, so a process
    // that will also load another NCCL user later (e.g. PyTorch with its bundled, newer NCCL) must load that copy first.
    void* h = nullptr;
    if (const char* p = getenv("ENF_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(ctx, ENF_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString)
        return fail(ctx, ENF_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return ENF_OK;
}

#define NC(ctx, call)                                                                              \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess)                                                                    \
            return fail(ctx, ENF_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__));    \
    } while (0)

struct HostOp {
    int kind;
    int K;
    size_t poff;     // offset of this op's params in the packed parameter vector
    size_t nparams;  // number of parameters (fields * D, or D * K)
};

inline size_t params_of(int kind, int K, int D) {
    switch (kind) {
        case OP_CS: case OP_CC: return size_t(3) * D;
        case OP_JO: case OP_JI: return size_t(4) * D;
        case OP_SS: return size_t(2) * D;
        default: return size_t(D) * K;
    }
}

}  // namespace

struct enf_chain {
    enf_ctx* ctx = nullptr;
    int dtype = ENF_F32;
    int D = 0;
    std::vector<HostOp> ops;
    std::vector<double> params;  // packed, float64 copy of what the caller passed
    size_t n_params = 0;
    Plan plan;
    ChainDesc desc;
    double ladj_const_ss = 0.0;     // sum_i log|a_i| over ScaleShift ops
    double ladj_const_other = 0.0;  // Johnson row constants
    // device / staging
    void* d_consts = nullptr;
    void* h_consts = nullptr;  // pinned
    cudaEvent_t consts_copied = nullptr;
    bool consts_pending = false;
    int n_raw = 0;
    int max_blocks = 0;
    double* d_partials = nullptr;
    double* d_sums = nullptr;  // n_raw + 1 (last: N_local, for the group all-reduce)
    double* h_sums = nullptr;  // pinned, n_raw + 1
    // Householder/ScaleShift-only chains at large D: folded affine map for the tensor-core kernel (enf_affine.cu)
    bool affine = false;
    bool wy = false;            // ... of which those with few enough reflections run in compact-WY form (enf_wy.cu)
    bool wy_valid = false;      // ... unless the current parameters have a zero scale
    float* d_wy = nullptr;      // Wt hi | Wt lo | U hi | U lo | alpha | c
    float* d_affine = nullptr;  // Wh | Wl | bias
    // ... and their loss/gradient from the batch's second moments (enf_moments.cu): the raw sums are
    // [[S, m], [m^T, N]] ((D+1)^2 doubles) instead of per-op sums
    bool moments = false;
    bool affine_dirty = false;
    double* d_params64 = nullptr;  // packed float64 parameters (device-side chain rule)
    double* d_mom_out = nullptr;   // [negll, grads...]
    double* d_mom_part = nullptr;  // per-CTA partial [negll, grads...] of the chain-rule kernel
    double* d_mom_all = nullptr;   // per-batch moment matrices of the device-side fit loop (kept between calls)
    size_t mom_all_cap = 0;        // ... capacity in doubles
    double* h_mom_out = nullptr;   // pinned
};

namespace {

size_t elem_size(int dtype) { return dtype == ENF_F32 ? 4 : 8; }

template <typename T>
void put(void* base, size_t i, double v) { static_cast<T*>(base)[i] = T(v); }

// Fill the constants block (layout: for each op, n_consts_of(kind) arrays of
// length Dp) from the float64 parameters.  Padding rows get neutral values that
// map 0 -> 0 with zero ladj, so they never need masking inside the kernels.
int derive_constants(enf_chain* ch) {
    enf_ctx* ctx = ch->ctx;
    const int D = ch->D, Dp = ch->desc.Dp;
    const bool packed = ch->plan.packed;
    if (ch->consts_pending) {
        CU(ctx, cudaEventSynchronize(ch->consts_copied));
        ch->consts_pending = false;
    }
    auto set = [&](size_t idx, double v) {
        if (ch->dtype == ENF_F32) put<float>(ch->h_consts, idx, v);
        else put<double>(ch->h_consts, idx, v);
    };
    // unit of the device log / exp: Float32 kernels use lg2 / ex2 (enf_math.cuh Prim<float>)
    const double lgu = ch->dtype == ENF_F32 ? 0.69314718055994530942 : 1.0;
    const double exu = ch->dtype == ENF_F32 ? LOG2E_D : 1.0;
    ch->ladj_const_ss = 0.0;
    ch->ladj_const_other = 0.0;
    for (size_t o = 0; o < ch->ops.size(); ++o) {
        const HostOp& op = ch->ops[o];
        const DevOp& dop = ch->desc.ops[o];
        const double* p = ch->params.data() + op.poff;
        for (int r = 0; r < Dp; ++r) {
            const bool real = packed || r < D;
            const int i = packed ? r % D : r;
            const size_t base = size_t(dop.coff) + r;
            switch (op.kind) {
                case OP_CS:
                case OP_CC: {
                    const double a = real ? p[i] : 0.0, b = real ? p[D + i] : 1.0, c = real ? p[2 * D + i] : 0.0;
                    const double A = std::exp(b * a);
                    set(base + 0 * size_t(Dp), -b * LOG2E_D);
                    set(base + 1 * size_t(Dp), A);
                    set(base + 2 * size_t(Dp), lgu / b);
                    set(base + 3 * size_t(Dp), c);
                    set(base + 4 * size_t(Dp), a);
                    set(base + 5 * size_t(Dp), b);
                    set(base + 6 * size_t(Dp), 0.5 * A);
                    set(base + 7 * size_t(Dp), 2.0 / A);
                    set(base + 8 * size_t(Dp), (1.0 + A * A) / A);
                    break;
                }
                case OP_JO: {
                    const double gm = real ? p[i] : 0.0, dl = real ? p[D + i] : 1.0, xi = real ? p[2 * D + i] : 0.0,
                                 lm = real ? p[3 * D + i] : 1.0;
                    set(base + 0 * size_t(Dp), 1.0 / lm);
                    set(base + 1 * size_t(Dp), -xi / lm);
                    set(base + 2 * size_t(Dp), gm);
                    set(base + 3 * size_t(Dp), dl * lgu);
                    set(base + 4 * size_t(Dp), dl);
                    if (r < D) ch->ladj_const_other += std::log(std::fabs(dl / lm));
                    break;
                }
                case OP_JI: {
                    const double gm = real ? p[i] : 0.0, dl = real ? p[D + i] : 1.0, xi = real ? p[2 * D + i] : 0.0,
                                 lm = real ? p[3 * D + i] : 1.0;
                    set(base + 0 * size_t(Dp), exu / dl);
                    set(base + 1 * size_t(Dp), -gm * exu / dl);
                    set(base + 2 * size_t(Dp), lm);
                    set(base + 3 * size_t(Dp), xi);
                    set(base + 4 * size_t(Dp), 1.0 / dl);
                    set(base + 5 * size_t(Dp), 1.0 / lm);
                    if (r < D) ch->ladj_const_other += std::log(std::fabs(lm / dl));
                    break;
                }
                case OP_SS: {
                    const double a = real ? p[i] : 1.0, b = real ? p[D + i] : 0.0;
                    set(base + 0 * size_t(Dp), a);
                    set(base + 1 * size_t(Dp), b);
                    if (r < D) ch->ladj_const_ss += std::log(std::fabs(a));
                    break;
                }
                default: {  // OP_HH: v'_k = v_k sqrt(2 / v_k.v_k), filled below (needs all rows at once)
                    break;
                }
            }
        }
        if (op.kind == OP_HH) {
            // v'_k = v_k sqrt(2 / v_k.v_k): y = x - (v'.x) v' (src/householder_trafo.jl:4-11 without the division)
            for (int k = 0; k < op.K; ++k) {
                const double* v = p + size_t(k) * D;
                double n = 0.0;
                for (int j = 0; j < D; ++j) n += v[j] * v[j];
                const double sc = std::sqrt(2.0 / n);
                for (int r = 0; r < Dp; ++r) {
                    const bool real = packed || r < D;
                    const int i = packed ? r % D : r;
                    set(size_t(dop.coff) + size_t(k) * Dp + r, real ? v[i] * sc : 0.0);
                }
            }
        }
    }
    CU(ctx, cudaMemcpyAsync(ch->d_consts, ch->h_consts, size_t(ch->desc.n_consts) * elem_size(ch->dtype),
                            cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaEventRecord(ch->consts_copied, ctx->stream));
    ch->consts_pending = true;
    ch->affine_dirty = ch->affine;   // W = fold of the chain is rebuilt lazily, by the first forward call that needs it
    if (ch->moments) {
        // float64 parameters for the device-side chain rule (pageable source is staged before the call returns)
        CU(ctx, cudaMemcpyAsync(ch->d_params64, ch->params.data(), ch->n_params * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<double> norms;   // v.v of every reflection, in application order
        for (const HostOp& op : ch->ops)
            if (op.kind == OP_HH)
                for (int k = 0; k < op.K; ++k) {
                    const double* v = ch->params.data() + op.poff + size_t(k) * D;
                    double n = 0.0;
                    for (int i = 0; i < D; ++i) n += v[i] * v[i];
                    norms.push_back(n);
                }
        CU(ctx, cudaMemcpyAsync(ch->d_params64 + ch->n_params, norms.data(), norms.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    return ENF_OK;
}

// Householder/ScaleShift chains on the tensor-core forward path: fold the chain into y = W x + c (host, float64)
int ensure_affine(enf_chain* ch) {
    enf_ctx* ctx = ch->ctx;
    const int D = ch->D;
    if (ch->affine && ch->affine_dirty) {
        std::vector<int> kinds, Ks;
        std::vector<const double*> pp;
        for (const HostOp& op : ch->ops) {
            kinds.push_back(op.kind);
            Ks.push_back(op.K);
            pp.push_back(ch->params.data() + op.poff);
        }
        if (ch->wy) {
            std::vector<float> buf;
            ch->wy_valid = wy_fold(D, int(kinds.size()), kinds.data(), Ks.data(), pp.data(), buf);
            CU(ctx, cudaMemcpyAsync(ch->d_wy, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
        std::vector<float> wh, wl, bias;
        affine_fold(D, int(kinds.size()), kinds.data(), Ks.data(), pp.data(), wh, wl, bias);
        const size_t n2 = size_t(D) * D;
        // pageable source: the copies are staged before cudaMemcpyAsync returns, so the vectors may die here
        CU(ctx, cudaMemcpyAsync(ch->d_affine, wh.data(), n2 * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ch->d_affine + n2, wl.data(), n2 * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ch->d_affine + 2 * n2, bias.data(), size_t(D) * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        ch->affine_dirty = false;
    }
    return ENF_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ENF_MOMENTS_HOST=1: map second moments to gradients with the float64 host code (finish_moments) instead of the
// cluster kernel - the cross-check of the two implementations
bool host_chain_rule() {
    static const bool v = getenv("ENF_MOMENTS_HOST") != nullptr;
    return v;
}

int pick_mode(const enf_chain* ch, const void* x, const void* y) {
    const int VE = ch->dtype == ENF_F32 ? 4 : 2;
    const bool al = aligned16(x) && (y == nullptr || aligned16(y));
    if (ch->plan.packed) return al ? MODE_PACK : MODE_PACKU;
    return (al && ch->D % VE == 0) ? MODE_VEC : MODE_SCALAR;
}

// Householder/ScaleShift-only chains: second moments [[S, m], [m^T, N]] -> (negll, grads), O(K D^2) in float64.
// With x^ = [x; 1], every intermediate of the chain is x_i = B_i x^ (B_i: D x (D+1)), the loss cotangent of
// the output is y/N, and for every op  sum_j g_ij x_(i-1)j^T = Z_i B_(i-1)^T  with  Z_n = B_n S^/N,
// Z_(i-1) = A_i^T Z_i.  The reverse sweep recovers B_(i-1) from B_i exactly like the reference recovers a
// reflection's input from its output (src/householder_trafo.jl:88-103) and applies the closed forms of
// SURVEY §8a (pullback_v: :22-40) to the D x D moment matrix instead of to the D x N batch.
void finish_moments(const enf_chain* ch, const double* sums, int64_t N, int flags, double* negll, std::vector<double>* grads) {
    const int D = ch->D, D1 = D + 1;
    const double Nd = double(N);
    std::vector<double> B(size_t(D) * D1, 0.0), Z(size_t(D) * D1), w(D1), tB(D1), tZ(D1);
    for (int k = 0; k < D; ++k) {
        B[size_t(k) * D1 + k] = 1.0;
        for (int c = 0; c < D1; ++c) Z[size_t(k) * D1 + c] = sums[size_t(k) * D1 + c] / Nd;
    }
    for (int c = 0; c < D1; ++c) w[c] = sums[size_t(D) * D1 + c] / Nd;
    // t = v^T M, then M -= s v t^T   (M: D x D1 row-major)
    auto reflect = [&](std::vector<double>& M, const double* v, double s, std::vector<double>& t) {
        std::fill(t.begin(), t.end(), 0.0);
        for (int k = 0; k < D; ++k) {
            const double vk = v[k];
            const double* row = M.data() + size_t(k) * D1;
            for (int c = 0; c < D1; ++c) t[c] += vk * row[c];
        }
        for (int k = 0; k < D; ++k) {
            const double f = s * v[k];
            double* row = M.data() + size_t(k) * D1;
            for (int c = 0; c < D1; ++c) row[c] -= f * t[c];
        }
    };
    auto norm2 = [&](const double* v) { double n = 0.0; for (int k = 0; k < D; ++k) n += v[k] * v[k]; return n; };
    // forward: B_n and Z_n = B_n S^/N (the chain applied to the columns of S^/N; w carries the homogeneous weight)
    for (size_t o = 0; o < ch->ops.size(); ++o) {
        const HostOp& op = ch->ops[o];
        const double* p = ch->params.data() + op.poff;
        if (op.kind == OP_SS) {
            for (int k = 0; k < D; ++k) {
                const double a = p[k], b = p[D + k];
                double* rb = B.data() + size_t(k) * D1;
                double* rz = Z.data() + size_t(k) * D1;
                for (int c = 0; c < D1; ++c) { rb[c] *= a; rz[c] = a * rz[c] + b * w[c]; }
                rb[D] += b;
            }
        } else {
            for (int r = 0; r < op.K; ++r) {
                const double* v = p + size_t(r) * D;
                const double s = 2.0 / norm2(v);
                reflect(B, v, s, tB);
                reflect(Z, v, s, tZ);
            }
        }
    }
    double sum_y = 0.0;   // sum_j |y_j|^2 / 2 = N/2 <Z_n, B_n>
    for (size_t i = 0; i < B.size(); ++i) sum_y += Z[i] * B[i];
    sum_y *= 0.5 * Nd;
    const double lconst = ch->ladj_const_other + ((flags & ENF_NEGLL_ZYGOTE_PRIMAL) ? 0.0 : ch->ladj_const_ss);
    if (negll) *negll = (sum_y + 0.5 * LOG2PI * Nd * D - Nd * lconst) / Nd;
    if (!grads) return;
    grads->assign(ch->n_params, 0.0);
    for (size_t oo = ch->ops.size(); oo-- > 0;) {
        const HostOp& op = ch->ops[oo];
        const double* p = ch->params.data() + op.poff;
        double* g = grads->data() + op.poff;
        if (op.kind == OP_SS) {
            for (int k = 0; k < D; ++k) {
                const double a = p[k], b = p[D + k], ia = 1.0 / a;
                double* rb = B.data() + size_t(k) * D1;
                double* rz = Z.data() + size_t(k) * D1;
                g[D + k] = rz[D];
                rb[D] -= b;
                double acc = 0.0;
                for (int c = 0; c < D1; ++c) { rb[c] *= ia; acc += rz[c] * rb[c]; rz[c] *= a; }
                g[k] = acc - ia;
            }
        } else {
            for (int r = op.K; r-- > 0;) {
                const double* v = p + size_t(r) * D;
                const double n = norm2(v), s = 2.0 / n;
                reflect(B, v, s, tB);                       // B: output -> input of this reflection; v^T B_in = -tB
                std::fill(tZ.begin(), tZ.end(), 0.0);
                for (int k = 0; k < D; ++k) {
                    const double vk = v[k];
                    const double* rz = Z.data() + size_t(k) * D1;
                    for (int c = 0; c < D1; ++c) tZ[c] += vk * rz[c];
                }
                double vcv = 0.0;
                for (int c = 0; c < D1; ++c) vcv -= tZ[c] * tB[c];
                for (int k = 0; k < D; ++k) {
                    const double* rb = B.data() + size_t(k) * D1;
                    double* rz = Z.data() + size_t(k) * D1;
                    double cv = 0.0, ctv = 0.0;
                    for (int c = 0; c < D1; ++c) { cv -= rz[c] * tB[c]; ctv += rb[c] * tZ[c]; }
                    g[size_t(r) * D + k] = -s * (cv + ctv) + (4.0 / (n * n)) * vcv * v[k];
                    const double f = s * v[k];
                    for (int c = 0; c < D1; ++c) rz[c] -= f * tZ[c];   // Z: output -> input cotangent moments
                }
            }
        }
    }
}

// raw device sums -> (negll, grads).  Transcribes tests/device_model.py *_finish.
// elbo: the objective is the negative ELBO of examples/nf_variational_1d.jl:29-41 (sum_y then holds sum -log p(z)):
// nELBO = -[(sum log p(z) + sum ladj) / N - (log 2 pi + 1)/2 * D]
void finish(const enf_chain* ch, const double* sums, int64_t N, int flags, double* negll, std::vector<double>* grads, bool elbo = false) {
    if (ch->moments) return finish_moments(ch, sums, N, flags, negll, grads);
    const int D = ch->D, Dp = ch->desc.Dp;
    const bool packed = ch->plan.packed;
    const double Nd = double(N);
    const double LB = -1.0;
    const int n_raw = ch->n_raw;
    const double sum_y = sums[n_raw - 2], sum_l = sums[n_raw - 1];
    const double lconst = ch->ladj_const_other + ((flags & ENF_NEGLL_ZYGOTE_PRIMAL) ? 0.0 : ch->ladj_const_ss);
    if (negll) *negll = (sum_y + 0.5 * (LOG2PI + (elbo ? 1.0 : 0.0)) * Nd * D - (sum_l + Nd * lconst)) / Nd;
    if (!grads) return;
    grads->assign(ch->n_params, 0.0);
    std::vector<double> R(size_t(4) * D);
    for (size_t o = 0; o < ch->ops.size(); ++o) {
        const HostOp& op = ch->ops[o];
        const DevOp& dop = ch->desc.ops[o];
        const double* p = ch->params.data() + op.poff;
        double* g = grads->data() + op.poff;
        auto row_sum = [&](int slot, int i) {  // raw per-row sum `slot` of this op for row i
            const double* base = sums + size_t(dop.roff + slot) * Dp;
            if (!packed) return base[i];
            double s = 0.0;
            for (int r = i; r < Dp; r += D) s += base[r];
            return s;
        };
        if (op.kind == OP_HH) {
            for (int k = 0; k < op.K; ++k) {
                const double* v = p + size_t(k) * D;
                double n = 0.0;
                for (int j = 0; j < D; ++j) n += v[j] * v[j];
                const double acc2 = sums[size_t(ch->desc.n_rowslots) * Dp + dop.soff + k];
                for (int i = 0; i < D; ++i)
                    g[size_t(k) * D + i] = (-std::sqrt(2.0 / n) * row_sum(k, i) + (2.0 / n) * acc2 * v[i]) / Nd;
            }
            continue;
        }
        for (int i = 0; i < D; ++i) {
            const double r0 = row_sum(0, i), r1 = row_sum(1, i);
            switch (op.kind) {
                case OP_CS: {      // raw sums: dc, -da, -b db (enf_math.cuh: cs_bwd)
                    const double r2 = row_sum(2, i);
                    g[i] = -r1 / Nd; g[D + i] = -r2 / (p[D + i] * Nd); g[2 * D + i] = r0 / Nd;
                    break;
                }
                case OP_CC: {      // raw sums: -dc, da, b db (enf_math.cuh: cc_bwd)
                    const double r2 = row_sum(2, i);
                    g[i] = r1 / Nd; g[D + i] = r2 / (p[D + i] * Nd); g[2 * D + i] = -r0 / Nd;
                    break;
                }
                case OP_JO: {
                    const double r2 = row_sum(2, i), r3 = row_sum(3, i);
                    const double gm = p[i], dl = p[D + i], lm = p[3 * D + i];
                    g[i] = r0 / Nd;
                    g[D + i] = ((r1 - gm * r0) / dl + LB * Nd / dl) / Nd;     // sum G asinh z = sum G (y - gamma)/delta
                    g[2 * D + i] = (-r2 / lm) / Nd;
                    g[3 * D + i] = (-(r3 + LB * Nd) / lm) / Nd;
                    break;
                }
                case OP_JI: {
                    const double r2 = row_sum(2, i), r3 = row_sum(3, i);
                    const double dl = p[D + i], xi = p[2 * D + i], lm = p[3 * D + i];
                    g[i] = (-r0 / dl) / Nd;
                    g[D + i] = (-(r1 + LB * Nd) / dl) / Nd;
                    g[2 * D + i] = r2 / Nd;
                    g[3 * D + i] = ((r3 - xi * r2) / lm + LB * Nd / lm) / Nd;  // sum G sinh s = sum G (y - xi)/lambda
                    break;
                }
                default: {  // OP_SS
                    const double a = p[i];
                    g[i] = (r0 + LB * Nd / a) / Nd;
                    g[D + i] = r1 / Nd;
                    break;
                }
            }
        }
    }
}

void export_grads(const enf_chain* ch, const std::vector<double>& g, void* out) {
    if (ch->dtype == ENF_F32) {
        float* o = static_cast<float*>(out);
        for (size_t i = 0; i < g.size(); ++i) o[i] = float(g[i]);
    } else {
        std::memcpy(out, g.data(), g.size() * sizeof(double));
    }
}

int run_partial(enf_chain* ch, const void* x, int64_t N, bool grad, const ChainDesc* desc_override = nullptr) {
    enf_ctx* ctx = ch->ctx;
    if (N < 0) return fail(ctx, ENF_ERR_INVALID, "N must be >= 0");
    if (ch->moments) {
        if (N > 0 && !aligned16(x))
            return fail(ctx, ENF_ERR_INVALID, "loss/gradient of Householder/ScaleShift chains at D=%d needs a 16-byte aligned sample matrix", ch->D);
        CU(ctx, launch_moments(ch->D, x, N, ch->d_partials, ch->d_sums, ctx->sm_count, ctx->stream));
        ctx->launches += 2;
        return ENF_OK;
    }
    KernelSet ks;
    // a target density (ELBO objective) is compiled into the scalar-access layouts only
    const int mode = desc_override ? (ch->plan.packed ? MODE_PACKU : MODE_SCALAR) : pick_mode(ch, x, nullptr);
    if (!select_kernels(ch->dtype, ch->plan, mode, ks))
        return fail(ctx, ENF_ERR_INVALID, "no kernel variant for dtype=%d D=%d", ch->dtype, ch->D);
    const size_t smem = grad_smem_bytes(ch->dtype, ch->desc, ks, grad);
    if (smem > 227 * 1024)
        return fail(ctx, ENF_ERR_INVALID,
                    "chain needs %zu bytes of shared memory per CTA for the fused %s kernel (limit 232448)", smem,
                    grad ? "gradient" : "loss");
    int blocks = 0;
    CU(ctx, launch_grad(ch->dtype, ks, desc_override ? *desc_override : ch->desc, ch->d_consts, x, N, grad, ch->d_partials,
                        ch->max_blocks, &blocks, ctx->sm_count, ctx->stream));
    CU(ctx, launch_reduce(ch->d_partials, blocks, ch->n_raw, ch->d_sums, false, ctx->stream));
    ctx->launches += 2;
    return ENF_OK;
}

// second-moment chains: d_sums (possibly all-reduced) -> negll, grads through the device-side chain rule
int moments_finish_device(enf_chain* ch, int flags, double* negll, void* grads_host) {
    enf_ctx* ctx = ch->ctx;
    std::vector<int> kinds, Ks, poffs;
    for (const HostOp& op : ch->ops) {
        kinds.push_back(op.kind);
        Ks.push_back(op.K);
        poffs.push_back(int(op.poff));
    }
    const double lconst = ch->ladj_const_other + ((flags & ENF_NEGLL_ZYGOTE_PRIMAL) ? 0.0 : ch->ladj_const_ss);
    CU(ctx, launch_moments_chainrule(ch->D, int(kinds.size()), kinds.data(), Ks.data(), poffs.data(), int(ch->n_params), ch->d_params64,
                                     ch->d_params64 + ch->n_params, ch->d_sums, lconst, nullptr, ch->d_mom_part, ch->d_mom_out, ctx->stream));
    ctx->launches += 2;
    CU(ctx, cudaMemcpyAsync(ch->h_mom_out, ch->d_mom_out, (ch->n_params + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (negll) *negll = ch->h_mom_out[0];
    if (grads_host) {
        std::vector<double> g(ch->h_mom_out + 1, ch->h_mom_out + 1 + ch->n_params);
        export_grads(ch, g, grads_host);
    }
    return ENF_OK;
}

}  // namespace

// ------------------------------------------------------------------ context
extern "C" int enf_version(void) { return 100; }

extern "C" const char* enf_last_error(const enf_ctx* ctx) {
    if (ctx && !ctx->last_error.empty()) return ctx->last_error.c_str();
    return t_last_error.c_str();
}

extern "C" int enf_device_count(int* n) {
    if (!n) return fail(nullptr, ENF_ERR_INVALID, "n is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(nullptr, ENF_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *n = c;
    return ENF_OK;
}

extern "C" int enf_init(int device, enf_ctx** out) {
    if (!out) return fail(nullptr, ENF_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, ENF_ERR_CUDA, "no usable CUDA device (%s); libenf_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, ENF_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, ENF_ERR_CUDA, "device %d is sm_%d%d; libenf_b200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    enf_ctx* ctx = new enf_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    CU(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < HOST_SLOTS; ++i) CU(ctx, cudaStreamCreateWithFlags(&ctx->slot_stream[i], cudaStreamNonBlocking));
    CU(ctx, cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming));
    for (int i = 0; i < ENF_N_EVENTS; ++i) CU(ctx, cudaEventCreate(&ctx->timing[i]));
    *out = ctx;
    return ENF_OK;
}

static void p2p_teardown(enf_ctx* ctx);

extern "C" int enf_destroy(enf_ctx* ctx) {
    if (!ctx) return ENF_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    p2p_teardown(ctx);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < HOST_SLOTS; ++i) {
        if (ctx->slot_stream[i]) { cudaStreamSynchronize(ctx->slot_stream[i]); cudaStreamDestroy(ctx->slot_stream[i]); }
        if (ctx->slot_buf[i]) cudaFree(ctx->slot_buf[i]);
    }
    if (ctx->ev) cudaEventDestroy(ctx->ev);
    for (int i = 0; i < ENF_N_EVENTS; ++i)
        if (ctx->timing[i]) cudaEventDestroy(ctx->timing[i]);
    for (auto& kv : ctx->numa_allocs) {                 // host buffers the caller never freed
        cudaHostUnregister(kv.first);
        munmap(kv.first, kv.second);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ENF_OK;
}

extern "C" int enf_sync(enf_ctx* ctx) {
    if (!ctx) return fail(nullptr, ENF_ERR_INVALID, "ctx is NULL");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ENF_OK;
}

extern "C" int enf_event_record(enf_ctx* ctx, int slot) {
    if (!ctx) return fail(nullptr, ENF_ERR_INVALID, "ctx is NULL");
    if (slot < 0 || slot >= ENF_N_EVENTS) return fail(ctx, ENF_ERR_INVALID, "event slot %d out of range", slot);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventRecord(ctx->timing[slot], ctx->stream));
    return ENF_OK;
}

extern "C" int enf_event_elapsed_ms(enf_ctx* ctx, int a, int b, float* ms) {
    if (!ctx || !ms) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if (a < 0 || a >= ENF_N_EVENTS || b < 0 || b >= ENF_N_EVENTS) return fail(ctx, ENF_ERR_INVALID, "event slot out of range");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventSynchronize(ctx->timing[b]));
    CU(ctx, cudaEventElapsedTime(ms, ctx->timing[a], ctx->timing[b]));
    return ENF_OK;
}

extern "C" int enf_launch_count(const enf_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return fail(nullptr, ENF_ERR_INVALID, "NULL argument");
    *n = ctx->launches;
    return ENF_OK;
}

// ------------------------------------------------------------------ memory
extern "C" int enf_alloc(enf_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    *dptr = nullptr;
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return fail(ctx, ENF_ERR_NOMEM, "cudaMalloc(%zu) failed: out of device memory", bytes);
    }
    CU(ctx, e);
    return ENF_OK;
}

extern "C" int enf_free(enf_ctx* ctx, void* dptr) {
    if (!ctx) return fail(ctx, ENF_ERR_INVALID, "ctx is NULL");
    if (!dptr) return ENF_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaFree(dptr));
    return ENF_OK;
}

static std::mutex g_host_mu;      // guards enf_ctx::numa_allocs / numa_node (host buffers may be managed from several threads)

// NUMA node the GPU hangs off (sysfs), or -1
static int gpu_numa_node(int device) {
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; ++c) *c = char(tolower(*c));
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// the CPUs of a NUMA node that this thread may run on ("0-31,64-95" in sysfs)
static bool node_cpus(int node, const cpu_set_t& allowed, cpu_set_t& out) {
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(&out);
    int n = 0, lo = 0, hi = 0;
    while (fscanf(f, "%d", &lo) == 1) {
        hi = lo;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &hi) != 1) break;
            c = fgetc(f);
        }
        for (int i = lo; i <= hi && i < CPU_SETSIZE; ++i)
            if (CPU_ISSET(i, &allowed)) { CPU_SET(i, &out); ++n; }
        if (c != ',') break;
    }
    fclose(f);
    return n > 0;
}

// Pinned host memory for the host-matrix entry points.  With several GPUs per box every rank streams ~90 GB/s through
// its buffers; if they all sit on the NUMA node the processes happened to start on, the ranks of the other socket pull
// everything across the inter-socket link and one node's DRAM carries all of it (round 1: 0.18 weak-scaling efficiency
// of the end-to-end leg at 8 GPUs).  So the pages are placed on the node the GPU hangs off: the calling thread is moved
// onto that node's CPUs while it first-touches an anonymous mapping (local allocation policy; needs no CAP_SYS_NICE,
// unlike mbind), and the mapping is then pinned with cudaHostRegister.  Falls back to cudaMallocHost when the topology
// is not visible or anything fails (ENF_NO_NUMA=1 forces that).
extern "C" int enf_host_alloc(enf_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    *hptr = nullptr;
    std::unique_lock<std::mutex> lk(g_host_mu);
    if (ctx->numa_node == -2)      // ENF_NUMA_NODE=<n> overrides the sysfs lookup (boxes that hide the topology; tests)
        ctx->numa_node = getenv("ENF_NO_NUMA") ? -1 : getenv("ENF_NUMA_NODE") ? atoi(getenv("ENF_NUMA_NODE")) : gpu_numa_node(ctx->device);
    const int node = ctx->numa_node;
    lk.unlock();
    cpu_set_t old_set, node_set;
    if (bytes >= (size_t(1) << 20) && node >= 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0 &&
        node_cpus(node, old_set, node_set)) {
        const size_t map_bytes = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
        void* p = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, map_bytes, MADV_HUGEPAGE);
            const bool moved = sched_setaffinity(0, sizeof node_set, &node_set) == 0;
            std::memset(p, 0, map_bytes);                                  // first touch on the GPU's node
            if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
            if (moved && cudaHostRegister(p, map_bytes, cudaHostRegisterDefault) == cudaSuccess) {
                lk.lock();
                ctx->numa_allocs[p] = map_bytes;
                *hptr = p;
                return ENF_OK;
            }
            cudaGetLastError();
            munmap(p, map_bytes);
        }
    }
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, ENF_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return ENF_OK;
}

extern "C" int enf_host_free(enf_ctx* ctx, void* hptr) {
    if (!ctx) return fail(ctx, ENF_ERR_INVALID, "ctx is NULL");
    if (!hptr) return ENF_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    size_t mapped = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = ctx->numa_allocs.find(hptr);
        if (it != ctx->numa_allocs.end()) {
            mapped = it->second;
            ctx->numa_allocs.erase(it);
        }
    }
    if (mapped) {
        CU(ctx, cudaHostUnregister(hptr));
        munmap(hptr, mapped);
        return ENF_OK;
    }
    CU(ctx, cudaFreeHost(hptr));
    return ENF_OK;
}

extern "C" int enf_h2d(enf_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return ENF_OK;
}

extern "C" int enf_d2h(enf_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ENF_OK;
}

extern "C" int enf_memset(enf_ctx* ctx, void* dst, int value, size_t bytes) {
    if (!ctx || (bytes && !dst)) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return ENF_OK;
}

extern "C" int enf_fill_normal(enf_ctx* ctx, int dtype, void* x, int D, int64_t N, int64_t col0, uint64_t seed) {
    if (!ctx || !x) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if (dtype != ENF_F32 && dtype != ENF_F64) return fail(ctx, ENF_ERR_INVALID, "bad dtype %d", dtype);
    if (D < 1 || N < 0) return fail(ctx, ENF_ERR_INVALID, "bad shape D=%d N=%lld", D, (long long)N);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_fill_normal(dtype, x, D, N, col0, seed, ctx->stream));
    ctx->launches += 1;
    return ENF_OK;
}

extern "C" int enf_convert(enf_ctx* ctx, int dst_dtype, void* dst, int src_dtype, const void* src, int64_t n) {
    if (!ctx || (n > 0 && (!dst || !src))) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if ((dst_dtype != ENF_F32 && dst_dtype != ENF_F64) || (src_dtype != ENF_F32 && src_dtype != ENF_F64))
        return fail(ctx, ENF_ERR_INVALID, "bad dtype %d <- %d", dst_dtype, src_dtype);
    if (n < 0) return fail(ctx, ENF_ERR_INVALID, "bad element count %lld", (long long)n);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_convert(dst_dtype, dst, src_dtype, src, n, ctx->stream));
    ctx->launches += 1;
    return ENF_OK;
}

// ------------------------------------------------------------------ chains
// Parameter domain of the kernels.  CenterStretch / CenterContract are evaluated with w = e^{-b|x|} and ln2/b, i.e. for
// b > 0: the reference itself writes exp(abs(b*x)) and sign(x) (src/center_stretch.jl:6-7), which is the b > 0 branch of
// the inverse; b <= 0 is rejected instead of silently returning another branch.  Johnson needs delta, lambda != 0
// (src/johnson_trafo.jl:29-42 divides by both), ScaleShift a != 0 (log|a|, 1/a), Householder v != 0.
static int validate_params(enf_chain* ch) {
    const int D = ch->D;
    for (size_t o = 0; o < ch->ops.size(); ++o) {
        const HostOp& op = ch->ops[o];
        const double* p = ch->params.data() + op.poff;
        for (size_t i = 0; i < op.nparams; ++i)
            if (!std::isfinite(p[i])) return fail(ch->ctx, ENF_ERR_INVALID, "op %zu: parameter %zu is not finite", o, i);
        switch (op.kind) {
            case OP_CS: case OP_CC:
                for (int i = 0; i < D; ++i)
                    if (!(p[D + i] > 0.0))
                        return fail(ch->ctx, ENF_ERR_INVALID, "op %zu: CenterStretch/CenterContract need b > 0 (b[%d] = %g)", o, i, p[D + i]);
                break;
            case OP_JO: case OP_JI:
                for (int i = 0; i < D; ++i)
                    if (p[D + i] == 0.0 || p[3 * D + i] == 0.0)
                        return fail(ch->ctx, ENF_ERR_INVALID, "op %zu: JohnsonTrafo needs delta != 0 and lambda != 0 (row %d)", o, i);
                break;
            case OP_SS:
                for (int i = 0; i < D; ++i)
                    if (p[i] == 0.0) return fail(ch->ctx, ENF_ERR_INVALID, "op %zu: ScaleShiftTrafo needs a != 0 (row %d)", o, i);
                break;
            default:
                for (int k = 0; k < op.K; ++k) {
                    double n = 0.0;
                    for (int i = 0; i < D; ++i) n += p[size_t(k) * D + i] * p[size_t(k) * D + i];
                    if (!(n > 0.0)) return fail(ch->ctx, ENF_ERR_INVALID, "op %zu: Householder vector %d is zero", o, k);
                }
        }
    }
    return ENF_OK;
}

extern "C" int enf_chain_create(enf_ctx* ctx, int dtype, int D, int n_ops, const enf_op* ops, enf_chain** out) {
    if (!ctx || !ops || !out) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (dtype != ENF_F32 && dtype != ENF_F64) return fail(ctx, ENF_ERR_INVALID, "bad dtype %d", dtype);
    if (D < 1) return fail(ctx, ENF_ERR_INVALID, "D must be >= 1 (got %d)", D);
    if (n_ops < 1 || n_ops > MAX_OPS) return fail(ctx, ENF_ERR_INVALID, "n_ops must be in [1,%d] (got %d)", MAX_OPS, n_ops);
    CU(ctx, cudaSetDevice(ctx->device));
    enf_chain* ch = new enf_chain();
    ch->ctx = ctx;
    ch->dtype = dtype;
    ch->D = D;
    ChainDesc& d = ch->desc;
    // Lay the op list out for a plan.  The three-vectors-per-lane plans (make_plan) triple the per-thread accumulators of
    // the gradient kernel against the one-vector layout the power-of-two plan would give it: if those no longer fit the
    // shared memory, the chain takes the power-of-two plan.
    for (int attempt = 0; attempt < 2; ++attempt) {
    if (!make_plan(dtype, D, ch->plan, attempt == 0)) {
        delete ch;
        return fail(ctx, ENF_ERR_INVALID, "D=%d is not supported (max %d rows for this dtype)", D,
                    dtype == ENF_F32 ? 1024 : 512);
    }
    ch->ops.clear();
    std::memset(&d, 0, sizeof d);
    d.n_ops = n_ops;
    d.D = D;
    d.Dp = ch->plan.Dp;
    size_t poff = 0;
    int coff = 0, roff = 0, soff = 0, save = 0;
    for (int o = 0; o < n_ops; ++o) {
        const int kind = ops[o].kind;
        int K = ops[o].K;
        if (kind < OP_CS || kind > OP_HH) { delete ch; return fail(ctx, ENF_ERR_INVALID, "op %d: bad kind %d", o, kind); }
        if (kind == OP_HH) {
            if (K < 1) { delete ch; return fail(ctx, ENF_ERR_INVALID, "op %d: Householder needs K >= 1 (got %d)", o, K); }
        } else K = 0;
        if (!ops[o].params) { delete ch; return fail(ctx, ENF_ERR_INVALID, "op %d: params is NULL", o); }
        HostOp h{kind, K, poff, params_of(kind, K, D)};
        ch->ops.push_back(h);
        DevOp& dop = d.ops[o];
        dop.kind = kind;
        dop.K = K;
        dop.coff = coff;
        dop.roff = roff;
        dop.soff = kind == OP_HH ? soff : 0;
        dop.save = kind == OP_HH ? -1 : save;
        coff += n_consts_of(kind, K) * d.Dp;
        roff += n_rowslots_of(kind, K);
        if (kind == OP_HH) soff += K; else save += 1;
        poff += h.nparams;
    }
    ch->n_params = poff;
    d.n_consts = coff;
    d.n_rowslots = roff;
    d.n_scalars = soff;
    d.n_save = save;
    ch->n_raw = d.n_rowslots * d.Dp + d.n_scalars + 2;
    if (attempt == 0 && ch->plan.CH == 3) {
        KernelSet ks;
        if (!select_kernels(dtype, ch->plan, MODE_VEC, ks) || grad_smem_bytes(dtype, d, ks, true) > 227 * 1024 ||
            fwd_smem_bytes(dtype, d) > 200 * 1024)
            continue;
    }
    break;
    }
    ch->affine = affine_supported(dtype, D, d);
    ch->wy = ch->affine && wy_rank(dtype, D, d) > 0;
    ch->moments = ch->affine && moments_supported(dtype, D) && getenv("ENF_NO_MOMENTS") == nullptr;
    if (ch->moments) ch->n_raw = (D + 1) * (D + 1);
    if (fwd_smem_bytes(dtype, d) > 200 * 1024) {
        delete ch;
        return fail(ctx, ENF_ERR_INVALID, "chain constants (%zu bytes) exceed the shared-memory budget", fwd_smem_bytes(dtype, d));
    }
    // gather params (float64 host copy)
    ch->params.resize(ch->n_params);
    for (int o = 0; o < n_ops; ++o) {
        const HostOp& h = ch->ops[o];
        for (size_t i = 0; i < h.nparams; ++i)
            ch->params[h.poff + i] = dtype == ENF_F32 ? double(static_cast<const float*>(ops[o].params)[i])
                                                       : static_cast<const double*>(ops[o].params)[i];
    }
    const size_t cbytes = size_t(d.n_consts) * elem_size(dtype);
    ch->max_blocks = 4 * ctx->sm_count;
    cudaError_t e;
    if ((e = cudaMalloc(&ch->d_consts, cbytes)) != cudaSuccess ||
        (e = cudaMallocHost(&ch->h_consts, cbytes)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_partials),
                        ch->moments ? moments_partial_bytes(D, ctx->sm_count) : size_t(ch->max_blocks) * ch->n_raw * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_sums), size_t(ch->n_raw + 1) * sizeof(double))) != cudaSuccess ||
        (e = cudaMallocHost(reinterpret_cast<void**>(&ch->h_sums), size_t(ch->n_raw + 1) * sizeof(double))) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ch->consts_copied, cudaEventDisableTiming)) != cudaSuccess) {
        enf_chain_destroy(ch);
        return fail(ctx, ENF_ERR_CUDA, "chain allocation failed: %s", cudaGetErrorString(e));
    }
    if (ch->moments &&
        ((e = cudaMalloc(reinterpret_cast<void**>(&ch->d_params64), (ch->n_params + size_t(d.n_scalars) + 1) * sizeof(double))) != cudaSuccess ||
         (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_mom_out), (ch->n_params + 1) * sizeof(double))) != cudaSuccess ||
         (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_mom_part), moments_chainrule_part_bytes(D, int(ch->n_params)))) != cudaSuccess ||
         (e = cudaMallocHost(reinterpret_cast<void**>(&ch->h_mom_out), (ch->n_params + 1) * sizeof(double))) != cudaSuccess)) {
        enf_chain_destroy(ch);
        return fail(ctx, ENF_ERR_CUDA, "chain allocation failed: %s", cudaGetErrorString(e));
    }
    if (ch->affine && (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_affine), (2 * size_t(D) * D + D) * sizeof(float))) != cudaSuccess) {
        enf_chain_destroy(ch);
        return fail(ctx, ENF_ERR_CUDA, "chain allocation failed: %s", cudaGetErrorString(e));
    }
    if (ch->wy && (e = cudaMalloc(reinterpret_cast<void**>(&ch->d_wy), wy_buffer_floats(D) * sizeof(float))) != cudaSuccess) {
        enf_chain_destroy(ch);
        return fail(ctx, ENF_ERR_CUDA, "chain allocation failed: %s", cudaGetErrorString(e));
    }
    int rc = validate_params(ch);
    if (rc == ENF_OK) rc = derive_constants(ch);
    if (rc != ENF_OK) { enf_chain_destroy(ch); return rc; }
    *out = ch;
    return ENF_OK;
}

extern "C" int enf_chain_set_params(enf_chain* ch, const void* packed) {
    if (!ch || !packed) return fail(ch ? ch->ctx : nullptr, ENF_ERR_INVALID, "NULL argument");
    CU(ch->ctx, cudaSetDevice(ch->ctx->device));
    std::vector<double> old = ch->params;
    for (size_t i = 0; i < ch->n_params; ++i)
        ch->params[i] = ch->dtype == ENF_F32 ? double(static_cast<const float*>(packed)[i])
                                              : static_cast<const double*>(packed)[i];
    int rc = validate_params(ch);
    if (rc != ENF_OK) { ch->params = old; return rc; }     // the chain keeps its previous parameters
    return derive_constants(ch);
}

extern "C" int enf_chain_num_params(const enf_chain* ch, int64_t* n) {
    if (!ch || !n) return fail(nullptr, ENF_ERR_INVALID, "NULL argument");
    *n = int64_t(ch->n_params);
    return ENF_OK;
}

extern "C" int enf_chain_destroy(enf_chain* ch) {
    if (!ch) return ENF_OK;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    for (int i = 0; i < HOST_SLOTS; ++i) cudaStreamSynchronize(ch->ctx->slot_stream[i]);
    if (ch->d_consts) cudaFree(ch->d_consts);
    if (ch->h_consts) cudaFreeHost(ch->h_consts);
    if (ch->d_partials) cudaFree(ch->d_partials);
    if (ch->d_sums) cudaFree(ch->d_sums);
    if (ch->h_sums) cudaFreeHost(ch->h_sums);
    if (ch->d_affine) cudaFree(ch->d_affine);
    if (ch->d_wy) cudaFree(ch->d_wy);
    if (ch->d_params64) cudaFree(ch->d_params64);
    if (ch->d_mom_out) cudaFree(ch->d_mom_out);
    if (ch->d_mom_part) cudaFree(ch->d_mom_part);
    if (ch->d_mom_all) cudaFree(ch->d_mom_all);
    if (ch->h_mom_out) cudaFreeHost(ch->h_mom_out);
    if (ch->consts_copied) cudaEventDestroy(ch->consts_copied);
    delete ch;
    return ENF_OK;
}

extern "C" int enf_chain_describe(const enf_chain* ch, char* buf, size_t buflen) {
    if (!ch || !buf || buflen == 0) return fail(nullptr, ENF_ERR_INVALID, "NULL argument");
    KernelSet ks;
    const int mode = ch->plan.packed ? MODE_PACK : ((ch->D % (ch->dtype == ENF_F32 ? 4 : 2)) == 0 ? MODE_VEC : MODE_SCALAR);
    select_kernels(ch->dtype, ch->plan, mode, ks);
    snprintf(buf, buflen,
             "dtype=%s D=%d Dp=%d layout=%s lanes_per_sample=%d vectors_per_lane=%d ops=%d consts=%d "
             "fwd_smem=%zu grad_smem=%zu n_raw=%d forward=%s",
             ch->dtype == ENF_F32 ? "f32" : "f64", ch->D, ch->desc.Dp, ch->plan.packed ? "packed" : "lane-group",
             1 << ch->plan.LG, ch->plan.CH, ch->desc.n_ops, ch->desc.n_consts, fwd_smem_bytes(ch->dtype, ch->desc),
             grad_smem_bytes(ch->dtype, ch->desc, ks, true), ch->n_raw,
             !ch->affine ? "simt" : (ch->wy && getenv("ENF_NO_WY") == nullptr) ? "tcgen05-compact-wy" : "tcgen05-dense-fold");
    return ENF_OK;
}

// ------------------------------------------------------------------ forward
static int forward_impl(enf_chain* ch, const void* x, int64_t N, void* y, void* ladj, bool want_ladj, cudaStream_t st) {
    enf_ctx* ctx = ch->ctx;
    if (N < 0) return fail(ctx, ENF_ERR_INVALID, "N must be >= 0");
    if (N == 0) return ENF_OK;
    if (!x || !y || (want_ladj && !ladj)) return fail(ctx, ENF_ERR_INVALID, "NULL device pointer");
    KernelSet ks;
    const int mode = pick_mode(ch, x, y);
    const double lc = ch->ladj_const_other + ch->ladj_const_ss;
    static const bool no_affine = getenv("ENF_NO_AFFINE") != nullptr;
    if (ch->affine && mode == MODE_VEC && !no_affine) {
        // Householder / ScaleShift stack at large D: one tcgen05 GEMM per tile (enf_affine.cu)
        int rca = ensure_affine(ch);
        if (rca != ENF_OK) return rca;
        const bool no_wy = getenv("ENF_NO_WY") != nullptr;             // cross-check: dense fold instead of compact WY
        if (ch->wy && ch->wy_valid && !no_wy)
            CU(ctx, launch_wy(ch->D, ch->d_wy, x, y, want_ladj ? ladj : nullptr, N, lc, ctx->sm_count, st));
        else
            CU(ctx, launch_affine(ch->D, ch->d_affine, x, y, want_ladj ? ladj : nullptr, N, lc, ctx->sm_count, st));
        ctx->launches += 1;
        return ENF_OK;
    }
    if (!select_kernels(ch->dtype, ch->plan, mode, ks))
        return fail(ctx, ENF_ERR_INVALID, "no kernel variant for dtype=%d D=%d", ch->dtype, ch->D);
    CU(ctx, launch_fwd(ch->dtype, ks, ch->desc, ch->d_consts, x, y, want_ladj ? ladj : nullptr, N, lc, ctx->sm_count, st));
    ctx->launches += 1;
    return ENF_OK;
}

extern "C" int enf_forward(enf_chain* ch, const void* x, int64_t N, void* y) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    CU(ch->ctx, cudaSetDevice(ch->ctx->device));
    return forward_impl(ch, x, N, y, nullptr, false, ch->ctx->stream);
}

extern "C" int enf_forward_ladj(enf_chain* ch, const void* x, int64_t N, void* y, void* ladj) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    CU(ch->ctx, cudaSetDevice(ch->ctx->device));
    return forward_impl(ch, x, N, y, ladj, true, ch->ctx->stream);
}

extern "C" int enf_forward_ladj_host(enf_chain* ch, const void* x_host, int64_t N, void* y_host, void* ladj_host) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (N < 0) return fail(ctx, ENF_ERR_INVALID, "N must be >= 0");
    if (N == 0) return ENF_OK;
    if (!x_host || !y_host) return fail(ctx, ENF_ERR_INVALID, "NULL host pointer");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t es = elem_size(ch->dtype);
    const size_t col_bytes = size_t(ch->D) * es;
    // chunk: ~32 MiB of samples, a multiple of 1024 columns so every chunk stays 16-byte aligned
    static const size_t chunk_mb = getenv("ENF_HOST_CHUNK_MB") ? size_t(atoi(getenv("ENF_HOST_CHUNK_MB"))) : 32;   // tuning aid
    int64_t chunk = int64_t(((chunk_mb ? chunk_mb : 32) << 20) / col_bytes);
    chunk = (chunk / 1024) * 1024;
    if (chunk < 1024) chunk = 1024;
    if (chunk > N) chunk = N;
    const size_t x_bytes = size_t(chunk) * col_bytes, l_bytes = ((size_t(chunk) * es + 255) / 256) * 256;
    const size_t need = 2 * ((x_bytes + 255) / 256 * 256) + l_bytes;
    if (ctx->slot_bytes < need) {
        for (int i = 0; i < HOST_SLOTS; ++i) {
            CU(ctx, cudaStreamSynchronize(ctx->slot_stream[i]));
            if (ctx->slot_buf[i]) { CU(ctx, cudaFree(ctx->slot_buf[i])); ctx->slot_buf[i] = nullptr; }
        }
        ctx->slot_bytes = 0;
        for (int i = 0; i < HOST_SLOTS; ++i) CU(ctx, cudaMalloc(&ctx->slot_buf[i], need));
        ctx->slot_bytes = need;
    }
    // the slot streams must see the chain's constants (uploaded on ctx->stream)
    CU(ctx, cudaEventRecord(ctx->ev, ctx->stream));
    for (int i = 0; i < HOST_SLOTS; ++i) CU(ctx, cudaStreamWaitEvent(ctx->slot_stream[i], ctx->ev, 0));
    const size_t xb_al = (x_bytes + 255) / 256 * 256;
    const bool copy_only = getenv("ENF_HOST_COPY_ONLY") != nullptr;
    int64_t done = 0;
    for (int it = 0; done < N; ++it) {
        const int s = it % HOST_SLOTS;
        const int64_t n = (N - done < chunk) ? (N - done) : chunk;
        char* base = static_cast<char*>(ctx->slot_buf[s]);
        void* dx = base;
        void* dy = base + xb_al;
        void* dl = base + 2 * xb_al;
        cudaStream_t st = ctx->slot_stream[s];
        CU(ctx, cudaMemcpyAsync(dx, static_cast<const char*>(x_host) + size_t(done) * col_bytes, size_t(n) * col_bytes,
                                cudaMemcpyHostToDevice, st));
        // ENF_HOST_COPY_ONLY=1 (diagnostic): the same pipeline without the kernel = the copy ceiling of the platform
        int rc = copy_only ? ENF_OK : forward_impl(ch, dx, n, dy, dl, ladj_host != nullptr, st);
        if (rc != ENF_OK) return rc;
        CU(ctx, cudaMemcpyAsync(static_cast<char*>(y_host) + size_t(done) * col_bytes, dy, size_t(n) * col_bytes,
                                cudaMemcpyDeviceToHost, st));
        if (ladj_host)
            CU(ctx, cudaMemcpyAsync(static_cast<char*>(ladj_host) + size_t(done) * es, dl, size_t(n) * es,
                                    cudaMemcpyDeviceToHost, st));
        done += n;
    }
    for (int i = 0; i < HOST_SLOTS; ++i) CU(ctx, cudaStreamSynchronize(ctx->slot_stream[i]));
    return ENF_OK;
}

// ------------------------------------------------------------------ loss / gradient
extern "C" int enf_negll(enf_chain* ch, const void* x, int64_t N, double* negll) {
    if (!ch || !negll) return fail(ch ? ch->ctx : nullptr, ENF_ERR_INVALID, "NULL argument");
    enf_ctx* ctx = ch->ctx;
    if (N < 1) return fail(ctx, ENF_ERR_INVALID, "N must be >= 1");
    if (!x) return fail(ctx, ENF_ERR_INVALID, "NULL device pointer");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = run_partial(ch, x, N, false);
    if (rc != ENF_OK) return rc;
    if (ch->moments && !host_chain_rule()) return moments_finish_device(ch, 0, negll, nullptr);
    CU(ctx, cudaMemcpyAsync(ch->h_sums, ch->d_sums, size_t(ch->n_raw) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    finish(ch, ch->h_sums, N, 0, negll, nullptr);
    return ENF_OK;
}

extern "C" int enf_negll_grad_partial(enf_chain* ch, const void* x, int64_t N_local, double** sums_dev, int64_t* n) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (N_local > 0 && !x) return fail(ctx, ENF_ERR_INVALID, "NULL device pointer");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = run_partial(ch, x, N_local, true);
    if (rc != ENF_OK) return rc;
    if (sums_dev) *sums_dev = ch->d_sums;
    if (n) *n = ch->n_raw;
    return ENF_OK;
}

extern "C" int enf_negll_grad_finish(enf_chain* ch, const double* sums_host, int64_t N_global, int flags,
                                     double* negll, void* grads_host) {
    if (!ch || !sums_host) return fail(ch ? ch->ctx : nullptr, ENF_ERR_INVALID, "NULL argument");
    if (N_global < 1) return fail(ch->ctx, ENF_ERR_INVALID, "N_global must be >= 1");
    std::vector<double> g;
    finish(ch, sums_host, N_global, flags, negll, grads_host ? &g : nullptr);
    if (grads_host) export_grads(ch, g, grads_host);
    return ENF_OK;
}

extern "C" int enf_negll_grad(enf_chain* ch, const void* x, int64_t N, int flags, double* negll, void* grads_host) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (N < 1) return fail(ctx, ENF_ERR_INVALID, "N must be >= 1");
    int rc = enf_negll_grad_partial(ch, x, N, nullptr, nullptr);
    if (rc != ENF_OK) return rc;
    if (ch->moments && !host_chain_rule()) return moments_finish_device(ch, flags, negll, grads_host);
    CU(ctx, cudaMemcpyAsync(ch->h_sums, ch->d_sums, size_t(ch->n_raw) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return enf_negll_grad_finish(ch, ch->h_sums, N, flags, negll, grads_host);
}

// ------------------------------------------------------------------ ELBO objective, JohnsonSU operations
// nELBO(trafo, xi) and its parameter gradient (examples/nf_variational_1d.jl:29-47) for xi of shape D x N (samples are
// columns, as everywhere in this library; the example passes a (2 batchsize) x 1 matrix and swaps the roles of rows and
// columns, :32-34 -- the shim transposes).  The target log-density is applied element-wise like the example's my_ll.(z).
extern "C" int enf_elbo_grad(enf_chain* ch, const enf_target* target, const void* xi_dev, int64_t N, int flags, double* nelbo,
                             void* grads_host) {
    if (!ch || !target || !nelbo) return fail(ch ? ch->ctx : nullptr, ENF_ERR_INVALID, "NULL argument");
    enf_ctx* ctx = ch->ctx;
    if (N < 1 || !xi_dev) return fail(ctx, ENF_ERR_INVALID, "N must be >= 1 and xi_dev non-NULL");
    if (target->kind != ENF_TARGET_GAUSS_MIXTURE) return fail(ctx, ENF_ERR_INVALID, "unknown target kind %d", target->kind);
    if (target->K < 1 || target->K > MAX_TARGET_K || !target->weights || !target->means || !target->sigmas)
        return fail(ctx, ENF_ERR_INVALID, "Gaussian-mixture target needs 1..%d components", MAX_TARGET_K);
    if (ch->moments) return fail(ctx, ENF_ERR_INVALID, "the ELBO objective is not available for second-moment (Householder/ScaleShift-only, D=%d) chains", ch->D);
    CU(ctx, cudaSetDevice(ctx->device));
    ChainDesc d = ch->desc;
    d.target_kind = 1;
    d.target_K = target->K;
    for (int k = 0; k < target->K; ++k) {
        const double w = target->weights[k], sg = target->sigmas[k];
        if (!(w > 0.0) || !(sg > 0.0) || !std::isfinite(target->means[k]))
            return fail(ctx, ENF_ERR_INVALID, "mixture component %d: weight and sigma must be positive, the mean finite", k);
        d.tlw[k] = std::log(w / sg) - 0.5 * LOG2PI;
        d.tmu[k] = target->means[k];
        d.tis[k] = 1.0 / sg;
    }
    int rc = run_partial(ch, xi_dev, N, grads_host != nullptr, &d);
    if (rc != ENF_OK) return rc;
    CU(ctx, cudaMemcpyAsync(ch->h_sums, ch->d_sums, size_t(ch->n_raw) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<double> g;
    finish(ch, ch->h_sums, N, flags, nelbo, grads_host ? &g : nullptr, true);
    if (grads_host) export_grads(ch, g, grads_host);
    return ENF_OK;
}

// Element-wise JohnsonSU operations (src/johnson_trafo.jl:109-129): out[i] = op(JohnsonSU(gamma, delta, xi, lambda), x[i]).
extern "C" int enf_johnsonsu(enf_ctx* ctx, int dtype, int op, const double* params4, const void* x_dev, int64_t N, void* out_dev) {
    if (!ctx || !params4) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if (dtype != ENF_F32 && dtype != ENF_F64) return fail(ctx, ENF_ERR_INVALID, "bad dtype %d", dtype);
    if (op < ENF_JSU_PDF || op > ENF_JSU_QUANTILE) return fail(ctx, ENF_ERR_INVALID, "bad JohnsonSU operation %d", op);
    if (N < 0 || (N > 0 && (!x_dev || !out_dev))) return fail(ctx, ENF_ERR_INVALID, "bad N / NULL device pointer");
    if (params4[1] == 0.0 || params4[3] == 0.0 || !std::isfinite(params4[0] + params4[1] + params4[2] + params4[3]))
        return fail(ctx, ENF_ERR_INVALID, "JohnsonSU needs finite parameters with delta != 0 and lambda != 0");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_johnsonsu(dtype, op, x_dev, out_dev, N, params4, ctx->sm_count, ctx->stream));
    if (N > 0) ctx->launches += 1;
    return ENF_OK;
}

// ------------------------------------------------------------------ NCCL group
extern "C" int enf_group_unique_id(void* id_out) {
    if (!id_out) return fail(nullptr, ENF_ERR_INVALID, "id_out is NULL");
    int rc = load_nccl(nullptr);
    if (rc != ENF_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == ENF_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(nullptr, g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return ENF_OK;
}

// Map every rank's exchange buffer into this process (CUDA IPC handles travel through one ncclAllReduce(sum) of a
// zero-padded int32 table).  Any failure just leaves ctx->p2p off: the group calls then use ncclAllReduce.
static int p2p_setup(enf_ctx* ctx) {
    ctx->p2p = false;
    const int R = ctx->nranks;
    if (R < 2 || R > P2P_MAX_RANKS || getenv("ENF_NO_P2P") != nullptr) return ENF_OK;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* local = nullptr;
    int* d_tab = nullptr;
    std::vector<int> tab(size_t(R) * 17, 0);           // per rank: 16 ints of handle + 1 "ok" flag
    bool ok = cudaMalloc(&local, p2p_bytes(R)) == cudaSuccess && cudaMemset(local, 0, p2p_bytes(R)) == cudaSuccess;
    cudaIpcMemHandle_t h;
    ok = ok && cudaIpcGetMemHandle(&h, local) == cudaSuccess;
    if (ok) {
        std::memcpy(&tab[size_t(ctx->rank) * 17], &h, 64);
        tab[size_t(ctx->rank) * 17 + 16] = 1;
    }
    cudaGetLastError();
    // the table exchange is collective: every rank takes part even if its own allocation failed
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&d_tab), tab.size() * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, g_nccl.AllReduce(d_tab, d_tab, tab.size(), ncclInt32, ncclSum, ctx->comm, ctx->stream));
    CU(ctx, cudaMemcpyAsync(tab.data(), d_tab, tab.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_tab);
    for (int r = 0; r < R; ++r) ok = ok && tab[size_t(r) * 17 + 16] == 1;
    P2PDesc d = {};
    d.nranks = R;
    d.rank = ctx->rank;
    int opened = 0;
    for (int r = 0; ok && r < R; ++r) {
        if (r == ctx->rank) { d.peer[r] = local; continue; }
        cudaIpcMemHandle_t hr;
        std::memcpy(&hr, &tab[size_t(r) * 17], 64);
        if (cudaIpcOpenMemHandle(&d.peer[r], hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
        ++opened;
    }
    // all ranks must agree: one more tiny all-reduce of the outcome
    int good = ok ? 1 : 0, *d_good = nullptr;
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&d_good), sizeof(int)));
    CU(ctx, cudaMemcpyAsync(d_good, &good, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(ctx, g_nccl.AllReduce(d_good, d_good, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&good, d_good, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_good);
    cudaGetLastError();
    if (good == R) {
        ctx->p2p = true;
        ctx->p2p_desc = d;
    } else {
        for (int r = 0; r < R; ++r)
            if (r != ctx->rank && d.peer[r]) cudaIpcCloseMemHandle(d.peer[r]);
        if (local) cudaFree(local);
        cudaGetLastError();
    }
    (void)opened;
    return ENF_OK;
}

static void p2p_teardown(enf_ctx* ctx) {
    if (!ctx->p2p) return;
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) continue;
        if (ctx->p2p_desc.peer[r]) cudaIpcCloseMemHandle(ctx->p2p_desc.peer[r]);
    }
    cudaFree(ctx->p2p_desc.peer[ctx->rank]);
    ctx->p2p = false;
    ctx->p2p_desc = P2PDesc{};
    cudaGetLastError();
}

// sum `n` doubles over the group in place: peer-memory kernel when available and small enough, else ncclAllReduce
static int group_allreduce(enf_ctx* ctx, double* d_vals, size_t n) {
    if (ctx->p2p && n <= size_t(P2P_SLOT)) {
        CU(ctx, launch_p2p_allreduce(ctx->p2p_desc, d_vals, int(n), ctx->stream));
        return ENF_OK;
    }
    NC(ctx, g_nccl.AllReduce(d_vals, d_vals, n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return ENF_OK;
}

// did a peer fail to show up in one of the peer-memory all-reduces since the last check? (call after a stream sync)
static int p2p_check(enf_ctx* ctx) {
    if (!ctx->p2p) return ENF_OK;
    int err = 0;
    CU(ctx, cudaMemcpy(&err, static_cast<unsigned char*>(ctx->p2p_desc.peer[ctx->rank]) + P2P_ERR_OFF, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        CU(ctx, cudaMemset(static_cast<unsigned char*>(ctx->p2p_desc.peer[ctx->rank]) + P2P_ERR_OFF, 0, sizeof(int)));
        return fail(ctx, ENF_ERR_NCCL, "peer-memory all-reduce timed out waiting for a rank of the group");
    }
    return ENF_OK;
}

extern "C" int enf_group_init(enf_ctx* ctx, int nranks, int rank, const void* id_bytes) {
    if (!ctx || !id_bytes) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, ENF_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    if (ctx->comm) return fail(ctx, ENF_ERR_INVALID, "group already initialised");
    int rc = load_nccl(ctx);
    if (rc != ENF_OK) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    NC(ctx, g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->nranks = nranks;
    ctx->rank = rank;
    return p2p_setup(ctx);
}

extern "C" int enf_group_destroy(enf_ctx* ctx) {
    if (!ctx) return fail(nullptr, ENF_ERR_INVALID, "ctx is NULL");
    if (ctx->comm) {
        CU(ctx, cudaSetDevice(ctx->device));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        p2p_teardown(ctx);
        NC(ctx, g_nccl.CommDestroy(ctx->comm));
        ctx->comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return ENF_OK;
}

// all-reduce (sum over the group) of the raw sums enf_negll_grad_partial left on the device, plus the local sample
// count: the exchange step of a sharded gradient step on its own (asynchronous on the context stream)
extern "C" int enf_group_allreduce_sums(enf_chain* ch, int64_t N_local) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (!ctx->comm) return fail(ctx, ENF_ERR_INVALID, "enf_group_init has not been called on this context");
    CU(ctx, cudaSetDevice(ctx->device));
    ch->h_sums[ch->n_raw] = double(N_local);
    CU(ctx, cudaMemcpyAsync(ch->d_sums + ch->n_raw, ch->h_sums + ch->n_raw, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return group_allreduce(ctx, ch->d_sums, size_t(ch->n_raw + 1));
}

extern "C" int enf_negll_grad_group(enf_chain* ch, const void* x, int64_t N_local, int flags, double* negll,
                                    void* grads_host) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (!ctx->comm) return fail(ctx, ENF_ERR_INVALID, "enf_group_init has not been called on this context");
    int rc = enf_negll_grad_partial(ch, x, N_local, nullptr, nullptr);
    if (rc != ENF_OK) return rc;
    rc = enf_group_allreduce_sums(ch, N_local);      // N_local is appended so one all-reduce also yields the global batch size
    if (rc != ENF_OK) return rc;
    if (ch->moments && !host_chain_rule()) {                 // N_global = S^[D][D]
        rc = moments_finish_device(ch, flags, negll, grads_host);        // synchronises the stream
        if (rc != ENF_OK) return rc;
        return p2p_check(ctx);                                           // a rank that never arrived fails every rank
    }
    CU(ctx, cudaMemcpyAsync(ch->h_sums, ch->d_sums, size_t(ch->n_raw + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (std::isnan(ch->h_sums[ch->n_raw])) {      // poisoned by the peer-memory all-reduce: some rank never arrived (sticky on all ranks)
        p2p_check(ctx);                           // clears the flag
        return fail(ctx, ENF_ERR_NCCL, "peer-memory all-reduce timed out waiting for a rank of the group");
    }
    const int64_t N_global = int64_t(std::llround(ch->h_sums[ch->n_raw]));
    return enf_negll_grad_finish(ch, ch->h_sums, N_global, flags, negll, grads_host);
}

// Device-side fit loop for second-moment (Householder/ScaleShift) chains: the batches are contiguous, unshuffled and
// the same in every epoch (src/optimize_whitening.jl:31-38), and the loss depends on a batch only through its moment
// matrix, so ONE pass over the data computes [[S, m], [m^T, N]] of every batch; after that a step is two launches
// (chain-rule cluster kernel + optimizer kernel) whose cost does not depend on the number of samples.
static int optimize_whitening_moments(enf_chain* ch, const void* x, const std::vector<int64_t>& starts, const std::vector<int64_t>& counts,
                                      int64_t nepochs, double eta, double epsilon, int flags, int use_group, int fresh_state,
                                      double* state_inout, void* params_out, double* history_out) {
    enf_ctx* ctx = ch->ctx;
    const size_t P = ch->n_params, stride = size_t(ch->n_raw) + 1;
    const int64_t nb = int64_t(counts.size());
    const int64_t n_steps = nb * nepochs;
    for (int64_t b = 0; b < nb; ++b)
        if (counts[size_t(b)] > 0 && (!aligned16(x) || (starts[size_t(b)] * ch->D * 4) % 16 != 0))
            return fail(ctx, ENF_ERR_INVALID, "second-moment chains need 16-byte aligned batches (D=%d, batch %lld starts at column %lld)",
                        ch->D, static_cast<long long>(b), static_cast<long long>(starts[size_t(b)]));
    std::vector<int> kinds, Ks, poffs;
    for (const HostOp& op : ch->ops) {
        kinds.push_back(op.kind);
        Ks.push_back(op.K);
        poffs.push_back(int(op.poff));
    }
    double *d_all = nullptr, *d_state = nullptr, *d_hist = nullptr, *d_lc = nullptr;
    long long* d_step = nullptr;
    auto cleanup = [&]() {
        if (d_state) cudaFree(d_state);
        if (d_hist) cudaFree(d_hist);
        if (d_lc) cudaFree(d_lc);
        if (d_step) cudaFree(d_step);
    };
#define CUF(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            cleanup();                                                                                   \
            return fail(ctx, ENF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)
    std::vector<double> st0(P);
    for (size_t i = 0; i < P; ++i) st0[i] = fresh_state ? epsilon : state_inout[i];
    // ladj row constants, one device slot per op (ScaleShift: sum log|a|; ENF_NEGLL_ZYGOTE_PRIMAL: dropped)
    std::vector<double> lc0(ch->ops.size(), 0.0);
    for (size_t o = 0; o < ch->ops.size(); ++o)
        if (ch->ops[o].kind == OP_SS && !(flags & ENF_NEGLL_ZYGOTE_PRIMAL))
            for (int i = 0; i < ch->D; ++i) lc0[o] += std::log(std::fabs(ch->params[ch->ops[o].poff + size_t(i)]));
    if (ch->mom_all_cap < size_t(nb) * stride) {   // tens of MB: allocate once per chain, not once per call
        if (ch->d_mom_all) cudaFree(ch->d_mom_all);
        ch->d_mom_all = nullptr;
        ch->mom_all_cap = 0;
        CUF(cudaMalloc(reinterpret_cast<void**>(&ch->d_mom_all), size_t(nb) * stride * sizeof(double)));
        ch->mom_all_cap = size_t(nb) * stride;
    }
    d_all = ch->d_mom_all;
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_state), P * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_hist), size_t(n_steps ? n_steps : 1) * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_lc), lc0.size() * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_step), sizeof(long long)));
    CUF(cudaMemcpyAsync(d_state, st0.data(), P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUF(cudaMemcpyAsync(d_lc, lc0.data(), lc0.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUF(cudaMemsetAsync(d_step, 0, sizeof(long long), ctx->stream));
    CUF(cudaStreamSynchronize(ctx->stream));
    // the one pass over the data
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t start = starts[size_t(b)], nbt = counts[size_t(b)];
        if (nbt == 0) {      // this rank holds no column of the batch (sharded batches): its moments are zero
            CUF(cudaMemsetAsync(d_all + size_t(b) * stride, 0, stride * sizeof(double), ctx->stream));
            continue;
        }
        const char* xb = static_cast<const char*>(x) + size_t(start) * size_t(ch->D) * 4;
        CUF(launch_moments(ch->D, xb, nbt, ch->d_partials, d_all + size_t(b) * stride, ctx->sm_count, ctx->stream));
        ctx->launches += 2;
    }
    if (use_group) {
        ncclResult_t r = g_nccl.AllReduce(d_all, d_all, size_t(nb) * stride, ncclDouble, ncclSum, ctx->comm, ctx->stream);
        if (r != ncclSuccess) {
            cleanup();
            return fail(ctx, ENF_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
        }
    }
    double* d_par = ch->d_params64;          // parameters + v.v array: uploaded by derive_constants, updated in place
    double* d_norms = ch->d_params64 + P;
    auto enqueue_epoch = [&]() -> int {
        for (int64_t b = 0; b < nb; ++b) {
            CUF(launch_moments_chainrule(ch->D, int(kinds.size()), kinds.data(), Ks.data(), poffs.data(), int(P), d_par, d_norms,
                                         d_all + size_t(b) * stride, 0.0, d_lc, ch->d_mom_part, ch->d_mom_out, ctx->stream));
            CUF(launch_moments_update(ch->D, int(kinds.size()), kinds.data(), Ks.data(), poffs.data(), ch->d_mom_out, d_par,
                                      d_norms, d_state, eta, epsilon, flags, d_lc, d_hist, d_step, ctx->stream));
            ctx->launches += 3;
        }
        return ENF_OK;
    };
    int64_t ep = 0;
    if (nepochs >= 1) {          // first epoch directly: sets the kernels' shared-memory attributes (not capturable)
        int rc = enqueue_epoch();
        if (rc != ENF_OK) return rc;
        ep = 1;
    }
    static const bool no_graph = getenv("ENF_NO_GRAPH") != nullptr;
    if (nepochs - ep >= 2 && !no_graph && nb <= 4096) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            int rc = enqueue_epoch();
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc != ENF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            ok = ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        }
        if (ok) {
            for (; ep < nepochs; ++ep)
                if (cudaGraphLaunch(exec, ctx->stream) != cudaSuccess) { ok = false; break; }
        } else {
            cudaGetLastError();
        }
        if (exec) { cudaStreamSynchronize(ctx->stream); cudaGraphExecDestroy(exec); }
        if (graph) cudaGraphDestroy(graph);
    }
    for (; ep < nepochs; ++ep) {
        int rc = enqueue_epoch();
        if (rc != ENF_OK) return rc;
    }
    std::vector<double> pfin(P);
    CUF(cudaMemcpyAsync(pfin.data(), d_par, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUF(cudaMemcpyAsync(state_inout, d_state, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_steps) CUF(cudaMemcpyAsync(history_out, d_hist, size_t(n_steps) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUF(cudaStreamSynchronize(ctx->stream));
#undef CUF
    cleanup();
    ch->params = pfin;
    export_grads(ch, pfin, params_out);
    int rcv = validate_params(ch);
    if (rcv != ENF_OK) return rcv;
    return derive_constants(ch);
}

// ------------------------------------------------------------------ device-side fit loop
// every rank of a group must run the same number of steps with the same number of collectives: compare the batch
// count over the ranks before anything is launched (max and -min in one all-reduce) and fail on EVERY rank otherwise
static int group_check_same(enf_ctx* ctx, int64_t value, const char* what) {
    long long h[2] = {static_cast<long long>(value), -static_cast<long long>(value)};
    long long* d = nullptr;
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&d), sizeof h));
    CU(ctx, cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllReduce(d, d, 2, ncclInt64, ncclMax, ctx->comm, ctx->stream);
    if (r != ncclSuccess) { cudaFree(d); return fail(ctx, ENF_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); }
    CU(ctx, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    if (h[0] != -h[1])
        return fail(ctx, ENF_ERR_INVALID, "the ranks of the group disagree on %s (%lld ... %lld); pass explicit per-batch column "
                    "counts (enf_optimize_whitening_batches)", what, -h[1], h[0]);
    return ENF_OK;
}

extern "C" int enf_optimize_whitening_batches(enf_chain* ch, const void* x, int64_t n_batches, const int64_t* local_counts,
                                              int64_t nepochs, double eta, double epsilon, int flags, int use_group, int fresh_state,
                                              double* state_inout, void* params_out, double* history_out, int64_t* n_steps_out) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (!x || !local_counts || !state_inout || !params_out || !history_out) return fail(ctx, ENF_ERR_INVALID, "NULL argument");
    if (n_batches < 1 || nepochs < 0) return fail(ctx, ENF_ERR_INVALID, "bad number of batches / epochs");
    if (use_group && !ctx->comm) return fail(ctx, ENF_ERR_INVALID, "enf_group_init has not been called on this context");
    CU(ctx, cudaSetDevice(ctx->device));
    const int64_t nb = n_batches;
    std::vector<int64_t> starts(static_cast<size_t>(nb)), cnts(local_counts, local_counts + nb);
    int64_t N = 0;
    for (int64_t b = 0; b < nb; ++b) {
        if (cnts[size_t(b)] < 0 || (cnts[size_t(b)] == 0 && !use_group))
            return fail(ctx, ENF_ERR_INVALID, "batch %lld has %lld columns", static_cast<long long>(b), static_cast<long long>(cnts[size_t(b)]));
        starts[size_t(b)] = N;
        N += cnts[size_t(b)];
    }
    if (use_group) {
        int rcg = group_check_same(ctx, nb, "the number of batches");
        if (rcg != ENF_OK) return rcg;
    }
    const int64_t n_steps = nb * nepochs;
    if (n_steps_out) *n_steps_out = n_steps;
    if (ch->moments)
        return optimize_whitening_moments(ch, x, starts, cnts, nepochs, eta, epsilon, flags, use_group, fresh_state,
                                          state_inout, params_out, history_out);
    const size_t P = ch->n_params;
    const size_t es = elem_size(ch->dtype);

    FitDesc fd;
    std::memset(&fd, 0, sizeof fd);
    fd.n_ops = int(ch->ops.size());
    fd.D = ch->D;
    fd.Dp = ch->desc.Dp;
    fd.packed = ch->plan.packed ? 1 : 0;
    fd.n_rowslots = ch->desc.n_rowslots;
    fd.n_raw = ch->n_raw;
    for (int o = 0; o < fd.n_ops; ++o) {
        fd.ops[o].kind = ch->ops[o].kind;
        fd.ops[o].K = ch->ops[o].K;
        fd.ops[o].poff = int(ch->ops[o].poff);
        fd.ops[o].coff = ch->desc.ops[o].coff;
        fd.ops[o].roff = ch->desc.ops[o].roff;
        fd.ops[o].soff = ch->desc.ops[o].soff;
    }
    KernelSet ks_probe;
    if (!select_kernels(ch->dtype, ch->plan, ch->plan.packed ? MODE_PACKU : MODE_SCALAR, ks_probe))
        return fail(ctx, ENF_ERR_INVALID, "no kernel variant for dtype=%d D=%d", ch->dtype, ch->D);
    if (grad_smem_bytes(ch->dtype, ch->desc, ks_probe, true) > 227 * 1024)
        return fail(ctx, ENF_ERR_INVALID, "chain too large for the fused gradient kernel");

    double *d_params = nullptr, *d_state = nullptr, *d_lconst = nullptr, *d_hist = nullptr, *d_counts = nullptr;
    long long* d_step = nullptr;
    std::vector<double> counts(static_cast<size_t>(nb), 0.0);
    std::vector<double> st0(P, 0.0);
    for (int64_t b = 0; b < nb; ++b) counts[size_t(b)] = double(cnts[size_t(b)]);
    for (size_t i = 0; i < P; ++i) st0[i] = fresh_state ? epsilon : state_inout[i];
    auto cleanup = [&]() {
        if (d_params) cudaFree(d_params);
        if (d_state) cudaFree(d_state);
        if (d_lconst) cudaFree(d_lconst);
        if (d_hist) cudaFree(d_hist);
        if (d_counts) cudaFree(d_counts);
        if (d_step) cudaFree(d_step);
    };
#define CUF(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            cleanup();                                                                                   \
            return fail(ctx, ENF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_params), P * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_state), P * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_lconst), 2 * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_hist), size_t(n_steps ? n_steps : 1) * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_counts), size_t(nb) * sizeof(double)));
    CUF(cudaMalloc(reinterpret_cast<void**>(&d_step), sizeof(long long)));
    CUF(cudaMemcpyAsync(d_params, ch->params.data(), P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUF(cudaMemcpyAsync(d_state, st0.data(), P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUF(cudaMemcpyAsync(d_counts, counts.data(), size_t(nb) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUF(cudaStreamSynchronize(ctx->stream));   // the pageable sources above may now go away / be reused
    if (ch->consts_pending) {
        CUF(cudaEventSynchronize(ch->consts_copied));
        ch->consts_pending = false;
    }
    CUF(cudaMemsetAsync(d_step, 0, sizeof(long long), ctx->stream));
    CUF(launch_fit_derive(ch->dtype, fd, d_params, ch->d_consts, d_lconst, ctx->stream));   // constants + ladj row constants of step 0
    // one epoch = nb steps of (derive, grad, reduce, count, [all-reduce], update); identical every epoch
    auto enqueue_epoch = [&]() -> int {
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t start = starts[size_t(b)], nbt = cnts[size_t(b)];
            const char* xb = static_cast<const char*>(x) + size_t(start) * size_t(ch->D) * es;
            KernelSet ks;
            select_kernels(ch->dtype, ch->plan, pick_mode(ch, xb, nullptr), ks);
            int blocks = 0;
            // gradient kernel -> update kernel -> gradient kernel ... as programmatic dependent launches: each kernel is
            // scheduled while its predecessor still runs and waits on the device (griddepcontrol.wait), which takes the
            // launch latency out of a step that is two short kernels (ENF_NO_PDL=1: plain stream order)
            static const bool pdl = getenv("ENF_NO_PDL") == nullptr;
            const bool pdl_u = pdl && !(use_group && !(ctx->p2p && size_t(ch->n_raw + 1) <= size_t(P2P_SLOT)));
            CUF(launch_grad(ch->dtype, ks, ch->desc, ch->d_consts, xb, nbt, true, ch->d_partials, ch->max_blocks, &blocks,
                            ctx->sm_count, ctx->stream, pdl_u));
            if (use_group && ctx->p2p && size_t(ch->n_raw + 1) <= size_t(P2P_SLOT)) {
                // sharded batch, small payload: the update kernel sums the partials, exchanges the sums with the other ranks
                // through NVLink peer memory and applies the step - still two launches per step
                CUF(launch_fit_update(ch->dtype, fd, ch->d_sums, ch->d_partials, blocks, double(nbt), d_lconst, d_params, d_state,
                                      eta, epsilon, flags, d_hist, d_step, ch->d_consts, ctx->stream, &ctx->p2p_desc,
                                      pdl_u && blocks < ctx->sm_count));
                ctx->launches += 2;
            } else if (use_group) {
                CUF(launch_reduce(ch->d_partials, blocks, ch->n_raw, ch->d_sums, false, ctx->stream));
                CUF(cudaMemcpyAsync(ch->d_sums + ch->n_raw, d_counts + b, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                if (group_allreduce(ctx, ch->d_sums, size_t(ch->n_raw + 1)) != ENF_OK) {
                    cleanup();
                    return ENF_ERR_NCCL;
                }
                CUF(launch_fit_update(ch->dtype, fd, ch->d_sums, nullptr, 0, 0.0, d_lconst, d_params, d_state, eta, epsilon, flags,
                                      d_hist, d_step, ch->d_consts, ctx->stream));
                ctx->launches += 3;
            } else {
                CUF(launch_fit_update(ch->dtype, fd, ch->d_sums, ch->d_partials, blocks, double(nbt), d_lconst, d_params, d_state,
                                      eta, epsilon, flags, d_hist, d_step, ch->d_consts, ctx->stream, nullptr,
                                      pdl_u && blocks < ctx->sm_count));
                ctx->launches += 2;
            }
        }
        return ENF_OK;
    };
    int64_t ep = 0;
    if (nepochs >= 1) {          // first epoch directly: it also sets the kernels' shared-memory attributes (not capturable)
        int rc = enqueue_epoch();
        if (rc != ENF_OK) return rc;
        ep = 1;
    }
    static const bool no_graph = getenv("ENF_NO_GRAPH") != nullptr;
    if (nepochs - ep >= 2 && !no_graph && nb <= 4096) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            int rc = enqueue_epoch();
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc != ENF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            ok = ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        }
        if (ok) {
            for (; ep < nepochs; ++ep) {
                cudaError_t le = cudaGraphLaunch(exec, ctx->stream);
                if (le != cudaSuccess) { ok = false; break; }
            }
        } else {
            cudaGetLastError();   // capture unsupported here: fall through to direct launches
        }
        if (exec) { cudaStreamSynchronize(ctx->stream); cudaGraphExecDestroy(exec); }
        if (graph) cudaGraphDestroy(graph);
    }
    for (; ep < nepochs; ++ep) {
        int rc = enqueue_epoch();
        if (rc != ENF_OK) return rc;
    }
    std::vector<double> pfin(P);
    CUF(cudaMemcpyAsync(pfin.data(), d_params, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUF(cudaMemcpyAsync(state_inout, d_state, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_steps) CUF(cudaMemcpyAsync(history_out, d_hist, size_t(n_steps) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUF(cudaStreamSynchronize(ctx->stream));
#undef CUF
    cleanup();
    if (use_group) {
        int prc = p2p_check(ctx);
        if (prc != ENF_OK) return prc;
    }
    ch->params = pfin;
    export_grads(ch, pfin, params_out);        // same packed layout / dtype conversion as gradients
    int rcv = validate_params(ch);             // the optimizer may have stepped a parameter out of the trafo's domain
    if (rcv != ENF_OK) return rcv;
    return derive_constants(ch);
}

// src/optimize_whitening.jl:31-32: batchsize = round(Int, length(smpls) / nbatches) (ties to even), contiguous column
// ranges in fixed order, the last one possibly short
extern "C" int enf_optimize_whitening(enf_chain* ch, const void* x, int64_t N, int64_t nbatches, int64_t nepochs,
                                      double eta, double epsilon, int flags, int use_group, int fresh_state,
                                      double* state_inout, void* params_out, double* history_out, int64_t* n_steps_out) {
    if (!ch) return fail(nullptr, ENF_ERR_INVALID, "chain is NULL");
    enf_ctx* ctx = ch->ctx;
    if (N < 1 || nbatches < 1 || nepochs < 0) return fail(ctx, ENF_ERR_INVALID, "bad N / nbatches / nepochs");
    const int64_t batchsize = int64_t(std::nearbyint(double(N) / double(nbatches)));
    if (batchsize < 1) return fail(ctx, ENF_ERR_INVALID, "nbatches exceeds the number of samples");
    const int64_t nb = (N + batchsize - 1) / batchsize;
    std::vector<int64_t> counts(static_cast<size_t>(nb));
    for (int64_t b = 0; b < nb; ++b) counts[size_t(b)] = std::min(batchsize, N - b * batchsize);
    return enf_optimize_whitening_batches(ch, x, nb, counts.data(), nepochs, eta, epsilon, flags, use_group, fresh_state,
                                          state_inout, params_out, history_out, n_steps_out);
}
