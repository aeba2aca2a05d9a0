// Instantiates the packed-sample chain kernels (D < VE).
#include "enf_chain_inst.cuh"
namespace enf {
namespace {
template <typename T, int MODE, int PD>
void fill(KernelSet& k) {
    fill_fwd<T, 0, 1, MODE, PD>(k);
    fill_grad<T, 0, 1, MODE, PD>(k);
    fill_grad_small<T, 0, 1, MODE, PD>(k);
}
}  // namespace
bool select_pack(int dtype, int PD, int mode, KernelSet& k) {
    if (dtype == 0) {
        if (PD == 1) { mode == MODE_PACK ? fill<float, MODE_PACK, 1>(k) : fill<float, MODE_PACKU, 1>(k); return true; }
        if (PD == 2) { mode == MODE_PACK ? fill<float, MODE_PACK, 2>(k) : fill<float, MODE_PACKU, 2>(k); return true; }
        return false;
    }
    if (PD == 1) { mode == MODE_PACK ? fill<double, MODE_PACK, 1>(k) : fill<double, MODE_PACKU, 1>(k); return true; }
    return false;
}
}  // namespace enf
