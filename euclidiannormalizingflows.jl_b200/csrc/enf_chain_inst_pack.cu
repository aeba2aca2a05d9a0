// Instantiates the packed-sample chain kernels (D < VE).
#include "enf_chain_inst.cuh"
namespace enf {
bool select_pack(int dtype, int PD, int mode, KernelSet& k) {
    if (dtype == 0) {
        if (PD == 1) { k = mode == MODE_PACK ? make_set<float, 0, 1, MODE_PACK, 1>() : make_set<float, 0, 1, MODE_PACKU, 1>(); return true; }
        if (PD == 2) { k = mode == MODE_PACK ? make_set<float, 0, 1, MODE_PACK, 2>() : make_set<float, 0, 1, MODE_PACKU, 2>(); return true; }
        return false;
    }
    if (PD == 1) { k = mode == MODE_PACK ? make_set<double, 0, 1, MODE_PACK, 1>() : make_set<double, 0, 1, MODE_PACKU, 1>(); return true; }
    return false;
}
}  // namespace enf
