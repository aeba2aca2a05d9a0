// Instantiates the lane-group chain kernels for float, MODE_VEC.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f32_vec(int LG, int CH, KernelSet& k) { return select_group<float, MODE_VEC>(LG, CH, k); }
}  // namespace enf
