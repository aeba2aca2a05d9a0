// Instantiates the lane-group chain kernels for float, MODE_VEC.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f32_vec(const Plan& p, KernelSet& k) { return select_group<float, MODE_VEC>(p, k); }
}  // namespace enf
