// Instantiates the lane-group chain kernels for double, MODE_SCALAR.
#include "enf_chain_inst.cuh"
namespace enf {
bool select_f64_scalar(int LG, int CH, KernelSet& k) { return select_group<double, MODE_SCALAR>(LG, CH, k); }
}  // namespace enf
