// Instantiates the lane-group chain kernels for double, MODE_SCALAR.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f64_scalar(const Plan& p, KernelSet& k) { return select_group<double, MODE_SCALAR>(p, k); }
}  // namespace enf
