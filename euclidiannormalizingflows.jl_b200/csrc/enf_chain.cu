// Instantiation table and launchers of the fused chain kernels (enf_chain.cuh).
#include "enf_chain.cuh"
#include "enf_launch.h"

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace enf {

// Fixed-order sum of the CTA partials (bitwise reproducible, independent of the launch).  Eight lanes share one output:
// lane j adds the partials of the CTAs b = j (mod 8) in increasing order, four independent loads in flight at a time,
// and the eight subtotals are combined by a fixed xor tree -- a single thread walking all CTAs was a chain of ~300
// dependent L2 round trips (34 us for the C5 chain's 296 x 423 partials).
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int n_blocks, int n_raw,
                                       double* __restrict__ sums, int accumulate) {
    const int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3), j = threadIdx.x & 7;
    const bool live = i < n_raw;
    double s = 0.0;
    if (live) {
        int b = j;
        for (; b + 24 < n_blocks; b += 32) {
            const double p0 = partials[size_t(b) * n_raw + i], p1 = partials[size_t(b + 8) * n_raw + i];
            const double p2 = partials[size_t(b + 16) * n_raw + i], p3 = partials[size_t(b + 24) * n_raw + i];
            s += p0;
            s += p1;
            s += p2;
            s += p3;
        }
        for (; b < n_blocks; b += 8) s += partials[size_t(b) * n_raw + i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && j == 0) sums[i] = accumulate ? sums[i] + s : s;
}

bool select_f32_vec(const Plan& p, KernelSet& k);
bool select_f32_scalar(const Plan& p, KernelSet& k);
bool select_f64_vec(const Plan& p, KernelSet& k);
bool select_f64_scalar(const Plan& p, KernelSet& k);
bool select_pack(int dtype, int PD, int mode, KernelSet& k);

bool select_kernels(int dtype, const Plan& plan, int mode, KernelSet& k) {
    if (plan.packed) {
        if (mode != MODE_PACK && mode != MODE_PACKU) return false;
        return select_pack(dtype, plan.PD, mode, k);
    }
    if (mode != MODE_VEC && mode != MODE_SCALAR) return false;
    if (dtype == 0) return mode == MODE_VEC ? select_f32_vec(plan, k) : select_f32_scalar(plan, k);
    return mode == MODE_VEC ? select_f64_vec(plan, k) : select_f64_scalar(plan, k);
}

bool make_plan(int dtype, int D, Plan& plan, bool allow_three) {
    const int VE = dtype == 0 ? 4 : 2;
    plan = Plan{};
    if (D < 1) return false;
    if (D < VE && VE % D == 0) {
        plan.packed = true;
        plan.PD = D;
        plan.LG = 0;
        plan.CH = 1;
        plan.gLG = 0;
        plan.gCH = 1;
        plan.Dp = VE;
        return true;
    }
    const int nvec = (D + VE - 1) / VE;
    int LG = 0;
    while ((1 << LG) < nvec && LG < 5) ++LG;
    // two vectors per lane when the sample has at least two: half the shuffles per Householder
    // reflection and per ladj, still sector-coalesced (measured fastest, profiles/)
    if (LG > 0 && nvec <= 32) --LG;
    // tuning override: ENF_PLAN_LG=<log2 lanes per sample> (the dispatcher derives vectors per lane from it)
    if (const char* e = getenv("ENF_PLAN_LG")) {
        const int v = atoi(e);
        if ((v == 0 && nvec == 4) || (v == 2 && nvec == 4)) LG = v;   // only (0,4) and (2,1) are instantiated
    }
    int CH = (nvec + (1 << LG) - 1) >> LG;
    int CHp = 1;
    while (CHp < CH) CHp <<= 1;
    if (CHp > 8) return false;
    plan.packed = false;
    plan.PD = 0;
    plan.LG = LG;
    plan.CH = CHp;
    plan.Dp = (1 << LG) * CHp * VE;
    // gradient kernels: as many lanes per sample as possible (same padded row count)
    plan.gLG = LG;
    plan.gCH = CHp;
    while (plan.gCH > 1 && plan.gLG < 5) { ++plan.gLG; plan.gCH >>= 1; }
    // Three vectors per lane where that pads fewer rows than the power-of-two plan (nvec = 3, 5-6, 9-12, 17-24, ...:
    // D = 24 ran with a quarter of its lanes idle, D = 20 with three eighths).  Forward and gradient kernels share the
    // padded row count, so both take the (lanes, 3) layout.
    if (allow_three && !getenv("ENF_NO_CH3"))
        for (int lg = 0; lg <= 5; ++lg) {
            const int cap = 3 << lg;
            if (cap >= nvec && cap * VE < plan.Dp) {
                plan.LG = plan.gLG = lg;
                plan.CH = plan.gCH = 3;
                plan.Dp = cap * VE;
                break;
            }
        }
    return true;
}

// ---- occupancy / attribute cache -------------------------------------------------
namespace {
std::mutex g_mu;
// both caches are per device: function attributes live in the device's context
std::map<std::tuple<int, const void*, size_t>, int> g_occ;   // (device, kernel, smem) -> CTAs per SM
std::map<std::pair<int, const void*>, size_t> g_smem_set;     // (device, kernel) -> max dynamic smem opted in

cudaError_t prepare_kernel(const void* fn, size_t smem, int& ctas_per_sm, int threads = NT) {
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    if (smem > 48 * 1024) {
        auto it = g_smem_set.find({dev, fn});
        if (it == g_smem_set.end() || it->second < smem) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            g_smem_set[{dev, fn}] = smem;
        }
    }
    auto key = std::make_tuple(dev, fn, smem);
    auto it = g_occ.find(key);
    if (it == g_occ.end()) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem);
        if (e != cudaSuccess) return e;
        if (nb < 1) return cudaErrorLaunchOutOfResources;
        it = g_occ.emplace(key, nb).first;
    }
    ctas_per_sm = it->second;
    return cudaSuccess;
}
}  // namespace

size_t fwd_smem_bytes(int dtype, const ChainDesc& d) { return size_t(d.n_consts) * (dtype == 0 ? 4 : 8); }

size_t grad_smem_bytes(int dtype, const ChainDesc& d, const KernelSet& k, bool grad) {
    const size_t es = dtype == 0 ? 4 : 8;
    size_t n = size_t((d.n_consts + 3) & ~3);
    if (grad) {
        n += size_t(d.n_save) * k.grad_tile_elems * NT;
        n += size_t(d.n_rowslots) * k.CH * k.VE * NT;
        n += size_t(d.n_scalars) * NT;
    }
    return n * es;
}

cudaError_t launch_fwd(int dtype, const KernelSet& k, const ChainDesc& desc, const void* consts, const void* x,
                       void* y, void* ladj, int64_t N, double ladj_const, int sm_count, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const void* fn = ladj ? k.fwd_ladj : k.fwd;
    // constants | pad to 128 B | TMA ring (MODE_VEC)
    const size_t smem = fwd_smem_bytes(dtype, desc) + (k.fwd_ring_bytes ? k.fwd_ring_bytes + 128 : 0);
    int per_sm = 0;
    cudaError_t e = prepare_kernel(fn, smem, per_sm, k.fwd_threads);
    if (e != cudaSuccess) return e;
    const int64_t items = (N + k.LN - 1) / k.LN;
    const int64_t tiles = (items + k.fwd_items_per_tile - 1) / k.fwd_items_per_tile;
    const int64_t cap = int64_t(per_sm) * sm_count;
    const unsigned grid = unsigned(tiles < cap ? tiles : cap);
    float lc32 = float(ladj_const);
    double lc64 = ladj_const;
    void* args[] = {const_cast<ChainDesc*>(&desc), &consts, &x, &y, &ladj, &N,
                    dtype == 0 ? static_cast<void*>(&lc32) : static_cast<void*>(&lc64)};
    return cudaLaunchKernel(fn, dim3(grid), dim3(k.fwd_threads), args, smem, st);
}

cudaError_t launch_grad(int dtype, const KernelSet& k, const ChainDesc& desc, const void* consts, const void* x,
                        int64_t N, bool grad, double* partials, int max_blocks, int* blocks_used, int sm_count,
                        cudaStream_t st, bool pdl) {
    const int64_t items = (N + k.LN - 1) / k.LN;
    KernelSet ks = k;
    static const bool no_small = getenv("ENF_NO_SMALL_GRAD") != nullptr;
    if (k.grad_small && !no_small && (items + k.grad_items_per_tile - 1) / k.grad_items_per_tile < sm_count) {
        // fewer regular tiles than SMs: the one-vector-per-thread variant spreads the batch over more of them
        ks.grad = k.grad_small;
        ks.negll = k.negll_small;
        ks.grad_items_per_tile = k.grad_small_items_per_tile;
        ks.grad_tile_elems = k.grad_small_tile_elems;
    }
    const void* fn = grad ? ks.grad : ks.negll;
    const size_t smem = grad_smem_bytes(dtype, desc, ks, grad);
    int per_sm = 0;
    cudaError_t e = prepare_kernel(fn, smem, per_sm);
    if (e != cudaSuccess) return e;
    int64_t tiles = (items + ks.grad_items_per_tile - 1) / ks.grad_items_per_tile;
    if (tiles < 1) tiles = 1;
    int64_t cap = int64_t(per_sm) * sm_count;
    if (cap > max_blocks) cap = max_blocks;
    const unsigned grid = unsigned(tiles < cap ? tiles : cap);
    *blocks_used = int(grid);
    void* args[] = {const_cast<ChainDesc*>(&desc), &consts, &x, &N, &partials};
    // (only for grids smaller than the GPU: a step of two short kernels is latency-bound and gains ~2 %; with a full-size
    // grid the early-resident CTAs of the next kernel cost more than the hidden launch latency - measured -1.4 % on C5)
    if (!pdl || int(grid) >= sm_count) return cudaLaunchKernel(fn, dim3(grid), dim3(NT), args, smem, st);
    // programmatic dependent launch: may start before the previous kernel of the stream has finished (chain_grad_kernel
    // waits for it with griddepcontrol.wait before it reads the constants)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

cudaError_t launch_reduce(const double* partials, int n_blocks, int n_raw, double* sums, bool accumulate,
                          cudaStream_t st) {
    const int threads = 256;                       // 32 outputs per CTA, 8 lanes each
    const int grid = (n_raw + 31) / 32;
    reduce_partials_kernel<<<grid, threads, 0, st>>>(partials, n_blocks, n_raw, sums, accumulate ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace enf
