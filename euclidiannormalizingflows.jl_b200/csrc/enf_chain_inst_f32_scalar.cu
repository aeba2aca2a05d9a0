// Instantiates the lane-group chain kernels for float, MODE_SCALAR.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f32_scalar(int LG, int CH, KernelSet& k) { return select_group<float, MODE_SCALAR>(LG, CH, k); }
}  // namespace enf
