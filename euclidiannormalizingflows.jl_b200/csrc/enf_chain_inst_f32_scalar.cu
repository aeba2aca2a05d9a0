// Instantiates the lane-group chain kernels for float, MODE_SCALAR.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f32_scalar(const Plan& p, KernelSet& k) { return select_group<float, MODE_SCALAR>(p, k); }
}  // namespace enf
