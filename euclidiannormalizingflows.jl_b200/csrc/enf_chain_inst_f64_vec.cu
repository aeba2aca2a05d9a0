// Instantiates the lane-group chain kernels for double, MODE_VEC.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f64_vec(const Plan& p, KernelSet& k) { return select_group<double, MODE_VEC>(p, k); }
}  // namespace enf
