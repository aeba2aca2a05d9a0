// Kernel-set construction shared by the instantiation translation units
// (enf_chain_inst_*.cu); split so the variants compile in parallel.
#pragma once
#include "enf_chain.cuh"
#include "enf_launch.h"

namespace enf {
namespace {

template <typename T, int LG, int CH, int MODE, int PD>
KernelSet make_set() {
    using CF = Cfg<T, LG, CH, MODE, PD, (CH >= 4 ? 1 : 4 / CH)>;   // forward: 16 B x 4 in flight per thread
    using CG = Cfg<T, LG, CH, MODE, PD, (CH >= 2 ? 1 : 2 / CH)>;   // gradient: two register tiles live
    KernelSet k;
    k.fwd = reinterpret_cast<const void*>(&chain_fwd_kernel<CF, false>);
    k.fwd_ladj = reinterpret_cast<const void*>(&chain_fwd_kernel<CF, true>);
    k.grad = reinterpret_cast<const void*>(&chain_grad_kernel<CG, true>);
    k.negll = reinterpret_cast<const void*>(&chain_grad_kernel<CG, false>);
    k.fwd_items_per_tile = CF::SB * CF::SPT;
    k.grad_items_per_tile = CG::SB * CG::SPT;
    k.grad_tile_elems = CG::SPT * CG::CH * CG::VE;
    k.LN = CF::LN;
    k.G = CF::G;
    k.CH = CH;
    k.VE = CF::VE;
    return k;
}

template <typename T, int MODE>
bool select_group(int LG, int CH, KernelSet& k) {
    if (CH == 1) {
        switch (LG) {
            case 0: k = make_set<T, 0, 1, MODE, 0>(); return true;
            case 1: k = make_set<T, 1, 1, MODE, 0>(); return true;
            case 2: k = make_set<T, 2, 1, MODE, 0>(); return true;
            case 3: k = make_set<T, 3, 1, MODE, 0>(); return true;
            case 4: k = make_set<T, 4, 1, MODE, 0>(); return true;
            case 5: k = make_set<T, 5, 1, MODE, 0>(); return true;
        }
        return false;
    }
    if (LG != 5) return false;
    switch (CH) {
        case 2: k = make_set<T, 5, 2, MODE, 0>(); return true;
        case 4: k = make_set<T, 5, 4, MODE, 0>(); return true;
        case 8: k = make_set<T, 5, 8, MODE, 0>(); return true;
    }
    return false;
}

}  // namespace

}  // namespace enf
