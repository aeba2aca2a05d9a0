// Registry of chains compiled as static op sequences (enf_chain.cuh: StaticStages).
//
// A chain whose (dtype, D, op kinds, Householder K) matches an entry runs a kernel
// in which the op list is a template parameter pack: no per-op dispatch, unrolled
// Householder loops, per-row constants held in registers.  Any other chain -- and
// any call whose pointers are not 16-byte aligned -- runs the interpretive kernel.
// The entries are the chain shapes of the reference's examples and of the
// benchmark configurations (BASELINE.json configs).
#include <cstdlib>

#include "enf_chain.cuh"
#include "enf_launch.h"

namespace enf {
namespace {

constexpr int code(int kind, int K = 0) { return kind | (K << 8); }

template <class C, int... CODES>
void fill(StaticKernel& k) {
    k.fwd = reinterpret_cast<const void*>(&chain_fwd_static_kernel<C, false, CODES...>);
    k.fwd_ladj = reinterpret_cast<const void*>(&chain_fwd_static_kernel<C, true, CODES...>);
    k.items_per_tile = C::SB * C::SPT;
    k.LN = C::LN;
    k.Dp = C::DP;
    k.ring_bytes = FwdRing<C>::BYTES;   // + the constants block, added by the launcher
    k.threads = NT;
}

bool same(const ChainDesc& d, std::initializer_list<int> codes) {
    if (d.n_ops != int(codes.size())) return false;
    int i = 0;
    for (int c : codes) {
        if (d.ops[i].kind != (c & 0xff) || d.ops[i].K != (c >> 8)) return false;
        ++i;
    }
    return true;
}

}  // namespace

// variant: tuning knob (ENF_STATIC_VARIANT), 0 = default mapping
bool select_static(int dtype, const ChainDesc& d, int mode, StaticKernel& k) {
    // measured slower than the interpretive kernel in round 1 (register-resident constants cost
    // occupancy, profiles/README.md): opt-in for experiments only
    static const bool disabled = getenv("ENF_STATIC") == nullptr;
    static const int variant = getenv("ENF_STATIC_VARIANT") ? atoi(getenv("ENF_STATIC_VARIANT")) : 0;
    if (disabled || mode != MODE_VEC) return false;
    if (dtype == 0 && d.D == 16) {
        // C3: CenterStretch ∘ JohnsonTrafo ∘ HouseholderTrafo(16x4) and its inverse
        if (same(d, {code(OP_HH, 4), code(OP_JO), code(OP_CS)})) {
            if (variant == 1) fill<Cfg<float, 2, 1, MODE_VEC, 0, 4>, code(OP_HH, 4), code(OP_JO), code(OP_CS)>(k);
            else if (variant == 2) fill<Cfg<float, 0, 4, MODE_VEC, 0, 1>, code(OP_HH, 4), code(OP_JO), code(OP_CS)>(k);
            else if (variant == 3) fill<Cfg<float, 1, 2, MODE_VEC, 0, 1>, code(OP_HH, 4), code(OP_JO), code(OP_CS)>(k);
            else if (variant == 4) fill<Cfg<float, 2, 1, MODE_VEC, 0, 2>, code(OP_HH, 4), code(OP_JO), code(OP_CS)>(k);
            else fill<Cfg<float, 1, 2, MODE_VEC, 0, 2>, code(OP_HH, 4), code(OP_JO), code(OP_CS)>(k);
            return k.Dp == d.Dp;
        }
        if (same(d, {code(OP_CC), code(OP_JI), code(OP_HH, 4)})) {
            fill<Cfg<float, 1, 2, MODE_VEC, 0, 2>, code(OP_CC), code(OP_JI), code(OP_HH, 4)>(k);
            return k.Dp == d.Dp;
        }
    }
    if (dtype == 0 && d.D == 32) {
        // C5 fit chain: ScaleShift ∘ Householder(32x4) ∘ JohnsonTrafo ∘ CenterContract
        if (same(d, {code(OP_CC), code(OP_JO), code(OP_HH, 4), code(OP_SS)})) {
            fill<Cfg<float, 2, 2, MODE_VEC, 0, 2>, code(OP_CC), code(OP_JO), code(OP_HH, 4), code(OP_SS)>(k);
            return k.Dp == d.Dp;
        }
    }
    return false;
}

}  // namespace enf
