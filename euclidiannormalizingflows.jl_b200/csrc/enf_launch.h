// Host-side interface between the C ABI (enf_abi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace enf {

struct ChainDesc;

// How a D-row sample is mapped onto lanes (see enf_chain.cuh).
struct Plan {
    bool packed = false;  // D < VE, VE % D == 0: several samples per 16-byte vector
    int PD = 0;           // rows per sample when packed
    int LG = 0;           // log2(lanes per sample)
    int CH = 1;           // 16-byte vectors per lane per sample
    int gLG = 0, gCH = 1; // the same for the loss / gradient kernels
    int Dp = 0;           // padded row count the constants block is laid out for
};

struct KernelSet {
    const void* fwd = nullptr;
    const void* fwd_ladj = nullptr;
    const void* grad = nullptr;
    const void* negll = nullptr;
    int fwd_items_per_tile = 0;
    int grad_items_per_tile = 0;
    int grad_tile_elems = 0;
    // one vector per thread per tile: the variant for batches too small to give every SM a tile of the regular kernel
    // (packed modes only: the D = 1 fit of BASELINE configs[1] runs 1e5-sample batches = 25 regular tiles on 148 SMs)
    const void* grad_small = nullptr;
    const void* negll_small = nullptr;
    int grad_small_items_per_tile = 0;
    int grad_small_tile_elems = 0;
    int LN = 1;  // samples per item (packed modes)
    int G = 1, CH = 1, VE = 4;
    size_t fwd_ring_bytes = 0;  // TMA input ring of the forward kernels (MODE_VEC)
    int fwd_threads = 256;      // CTA size of the forward kernels
};

bool make_plan(int dtype, int D, Plan& plan, bool allow_three = true);
bool select_kernels(int dtype, const Plan& plan, int mode, KernelSet& k);
size_t fwd_smem_bytes(int dtype, const ChainDesc& d);
size_t grad_smem_bytes(int dtype, const ChainDesc& d, const KernelSet& k, bool grad);

cudaError_t launch_fwd(int dtype, const KernelSet& k, const ChainDesc& desc, const void* consts, const void* x,
                       void* y, void* ladj, int64_t N, double ladj_const, int sm_count, cudaStream_t st);
cudaError_t launch_grad(int dtype, const KernelSet& k, const ChainDesc& desc, const void* consts, const void* x,
                        int64_t N, bool grad, double* partials, int max_blocks, int* blocks_used, int sm_count,
                        cudaStream_t st, bool pdl = false);
cudaError_t launch_reduce(const double* partials, int n_blocks, int n_raw, double* sums, bool accumulate,
                          cudaStream_t st);

// affine chains on the tensor cores (enf_affine.cu)
// peer-memory all-reduce (enf_p2p.cu): layout of a rank's buffer and the descriptor the kernel takes by value
constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_SLOT = 8192;                         // doubles per rank and parity
constexpr size_t P2P_SEQ_OFF = 0, P2P_ERR_OFF = 8, P2P_FLAG_OFF = 256, P2P_DATA_OFF = 1024;
constexpr long long P2P_SPIN_LIMIT = 400000000;        // ~2 minutes of waiting for a peer (rank skew: first-call set-up, host jitter) before giving up
// data area: [parity][source rank][P2P_SLOT] entries of 16 bytes: every double travels as two 8-byte words
// {32 data bits, 32-bit sequence number} (enf_p2p.cuh)
inline size_t p2p_bytes(int nranks) { return P2P_DATA_OFF + size_t(2) * nranks * P2P_SLOT * 16; }
struct P2PDesc {
    void* peer[P2P_MAX_RANKS];                         // peer[r]: rank r's buffer as mapped in this process
    int nranks, rank;
};
cudaError_t launch_p2p_allreduce(const P2PDesc& d, double* sums, int n, cudaStream_t st);

bool affine_supported(int dtype, int D, const ChainDesc& d);
// compact-WY form on the tensor cores (enf_wy.cu): y = alpha . x - U (W^T x) + c, two chained tcgen05 GEMMs per tile
int wy_rank(int dtype, int D, const ChainDesc& d);      // number of reflections if the WY kernel applies, else 0
size_t wy_buffer_floats(int D);
cudaError_t launch_wy(int D, const float* d_wy, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                      int sm_count, cudaStream_t st);
// second moments [[S, m], [m^T, N]] of a D x N batch on tensor cores (enf_moments.cu); d_part: scratch of
// moments_partial_bytes() bytes, d_sums: (D+1)^2 + 1 doubles
bool moments_supported(int dtype, int D);
size_t moments_partial_bytes(int D, int sm_count);
cudaError_t launch_moments(int D, const void* x, int64_t N, void* d_part, double* d_sums, int sm_count, cudaStream_t st);
size_t moments_chainrule_part_bytes(int D, int n_params);
cudaError_t launch_moments_chainrule(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs, int n_params,
                                     const double* d_params, const double* d_norms, const double* d_sums, double lconst,
                                     const double* d_lconst, double* d_part, double* d_out, cudaStream_t st);
// device-side optimizer step on the chain-rule kernel's output (ADAGrad + Householder column normalisation)
cudaError_t launch_moments_update(int D, int n_ops, const int* kinds, const int* Ks, const int* poffs, const double* d_out,
                                  double* d_params, double* d_norms, double* d_state, double eta, double eps, int flags,
                                  double* d_lconst, double* d_history, long long* d_step, cudaStream_t st);
cudaError_t launch_affine(int D, const float* d_w, const void* x, void* y, void* ladj, int64_t N, double ladj_const,
                          int sm_count, cudaStream_t st);

// device-side optimize_whitening loop (enf_fit.cu)
struct FitOp {
    int kind, K;
    int poff;   // offset of the op's parameters in the packed parameter vector
    int coff;   // offset of its constants in the constants block
    int roff;   // first per-row raw-sum slot
    int soff;   // first scalar raw-sum slot
};
struct FitDesc {
    int n_ops, D, Dp, packed;
    int n_rowslots, n_raw;
    FitOp ops[24];
};
cudaError_t launch_fit_derive(int dtype, const FitDesc& fd, const double* params, void* consts, double* lconst,
                              cudaStream_t st);
cudaError_t launch_fit_update(int dtype, const FitDesc& fd, double* sums, const double* partials, int n_blocks, double count,
                              double* lconst, double* params, double* state, double eta, double eps, int flags,
                              double* history, long long* step_ctr, void* consts, cudaStream_t st, const P2PDesc* p2p = nullptr,
                              bool pdl = false);

// JohnsonSU distribution operations (enf_johnsonsu.cu); `op` values are the ABI's enf_johnsonsu_op
enum : int { JSU_PDF = 0, JSU_LOGPDF = 1, JSU_CDF = 2, JSU_LOGCDF = 3, JSU_CCDF = 4, JSU_LOGCCDF = 5, JSU_QUANTILE = 6 };
cudaError_t launch_johnsonsu(int dtype, int op, const void* x, void* out, int64_t N, const double* params4, int sm_count,
                             cudaStream_t st);

// synthetic data (enf_fill.cu)
cudaError_t launch_fill_normal(int dtype, void* x, int D, int64_t N, int64_t col0, uint64_t seed, cudaStream_t st);
cudaError_t launch_convert(int dst_dtype, void* dst, int src_dtype, const void* src, int64_t n, cudaStream_t st);

}  // namespace enf
