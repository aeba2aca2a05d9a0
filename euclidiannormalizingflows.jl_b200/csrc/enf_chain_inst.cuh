// Kernel-set construction shared by the instantiation translation units
// (enf_chain_inst_*.cu); split so the variants compile in parallel.
#pragma once
#include "enf_chain.cuh"
#include "enf_launch.h"

namespace enf {
namespace {

#ifndef ENF_FWD_VECS
#define ENF_FWD_VECS 8   // 16-byte vectors per thread per tile in the forward kernels
#endif
#ifndef ENF_FWD_VECS_DIRECT
#define ENF_FWD_VECS_DIRECT 4   // 8 measured 0.58 instead of 0.74 of the HBM peak at D = 1
#endif

template <typename T, int LG, int CH, int MODE, int PD>
void fill_fwd(KernelSet& k) {
    // the staged (MODE_VEC) tiles carry ENF_FWD_VECS vectors per thread; modes that load straight into registers use
    // ENF_FWD_VECS_DIRECT
    constexpr int V = MODE == MODE_VEC ? ENF_FWD_VECS : ENF_FWD_VECS_DIRECT;
    using CF = Cfg<T, LG, CH, MODE, PD, (CH >= V ? 1 : V / CH)>;
    k.fwd = reinterpret_cast<const void*>(&chain_fwd_kernel<CF, false>);
    k.fwd_ladj = reinterpret_cast<const void*>(&chain_fwd_kernel<CF, true>);
    k.fwd_items_per_tile = CF::SB * CF::SPT;
    k.fwd_ring_bytes = FwdRing<CF>::BYTES;
    k.fwd_threads = NT;
    k.LN = CF::LN;
}

template <typename T, int LG, int CH, int MODE, int PD>
void fill_grad(KernelSet& k) {
#ifndef ENF_GRAD_VECS
#define ENF_GRAD_VECS 4   // 16-byte vectors per thread per tile in the gradient kernels (measured 15 % faster than 2)
#endif
    using CG = Cfg<T, LG, CH, MODE, PD, (CH >= ENF_GRAD_VECS ? 1 : ENF_GRAD_VECS / CH)>;
    k.grad = reinterpret_cast<const void*>(&chain_grad_kernel<CG, true>);
    k.negll = reinterpret_cast<const void*>(&chain_grad_kernel<CG, false>);
    k.grad_items_per_tile = CG::SB * CG::SPT;
    k.grad_tile_elems = CG::SPT * CG::CH * CG::VE;
    k.G = CG::G;
    k.CH = CH;
    k.VE = CG::VE;
}

template <typename T, int LG, int CH, int MODE, int PD>
void fill_grad_small(KernelSet& k) {
    using CG = Cfg<T, LG, CH, MODE, PD, 1>;
    k.grad_small = reinterpret_cast<const void*>(&chain_grad_kernel<CG, true>);
    k.negll_small = reinterpret_cast<const void*>(&chain_grad_kernel<CG, false>);
    k.grad_small_items_per_tile = CG::SB * CG::SPT;
    k.grad_small_tile_elems = CG::SPT * CG::CH * CG::VE;
}

// (log2 lanes per sample, vectors per lane) pairs make_plan() can produce for the forward kernels
// (two vectors per lane where possible) plus (2,1) and (0,4) for the ENF_PLAN_LG tuning override ...
template <typename T, int MODE>
bool select_group_fwd(int LG, int CH, KernelSet& k) {
#define ENF_CASE(lg, ch) if (LG == lg && CH == ch) { fill_fwd<T, lg, ch, MODE, 0>(k); return true; }
    ENF_CASE(0, 1) ENF_CASE(0, 2) ENF_CASE(1, 2) ENF_CASE(2, 2) ENF_CASE(3, 2) ENF_CASE(4, 2)
    ENF_CASE(5, 2) ENF_CASE(5, 4) ENF_CASE(5, 8) ENF_CASE(2, 1) ENF_CASE(0, 4)
    // three vectors per lane: D = 12, 24, 48, 96, ... (and the sizes just below) without a quarter of the lanes idle
    ENF_CASE(0, 3) ENF_CASE(1, 3) ENF_CASE(2, 3) ENF_CASE(3, 3) ENF_CASE(4, 3) ENF_CASE(5, 3)
#undef ENF_CASE
    return false;
}

// ... and for the gradient kernels (one vector per lane where possible: their per-thread
// accumulators and saved activations scale with the vectors a lane owns)
template <typename T, int MODE>
bool select_group_grad(int LG, int CH, KernelSet& k) {
#define ENF_CASE(lg, ch) if (LG == lg && CH == ch) { fill_grad<T, lg, ch, MODE, 0>(k); return true; }
    ENF_CASE(0, 1) ENF_CASE(1, 1) ENF_CASE(2, 1) ENF_CASE(3, 1) ENF_CASE(4, 1) ENF_CASE(5, 1)
    ENF_CASE(5, 2) ENF_CASE(5, 4) ENF_CASE(5, 8)
    ENF_CASE(0, 3) ENF_CASE(1, 3) ENF_CASE(2, 3) ENF_CASE(3, 3) ENF_CASE(4, 3) ENF_CASE(5, 3)
#undef ENF_CASE
    return false;
}

template <typename T, int MODE>
bool select_group(const Plan& p, KernelSet& k) {
    return select_group_fwd<T, MODE>(p.LG, p.CH, k) && select_group_grad<T, MODE>(p.gLG, p.gCH, k);
}

}  // namespace

}  // namespace enf
