// Peer-memory all-reduce of the raw sums of a sharded gradient step.
//
// The exchange of a step is a few KB of float64 sums: pure latency.  Instead of a library collective, every rank
// keeps a small buffer that all the other ranks of the node map through CUDA IPC (NVLink / NVSwitch peer memory).
// One kernel per rank does the whole all-reduce: it STORES its sums into its slot of every peer's buffer as 8-byte
// words that carry the exchange's sequence number next to 32 data bits each, polls its own buffer until the words of
// every other rank show that number, and adds the slots in rank order - every rank obtains the bitwise identical
// result with one NVLink store latency and no fence (enf_p2p.cuh).  The sequence counter lives on the device, so the
// kernel can sit inside a captured CUDA graph (enf_optimize_whitening).  Data slots are double-buffered by sequence
// parity: a rank can only be one all-reduce ahead of the slowest one.
#include <cuda_runtime.h>

#include <cstdint>

#include "enf_launch.h"
#include "enf_p2p.cuh"

namespace enf {
namespace {

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const __grid_constant__ P2PDesc d, double* __restrict__ sums, int n) {
    p2p_allreduce_block(d, sums, n);
}

}  // namespace

cudaError_t launch_p2p_allreduce(const P2PDesc& d, double* sums, int n, cudaStream_t st) {
    if (n > P2P_SLOT) return cudaErrorInvalidValue;
    p2p_allreduce_kernel<<<1, 256, 0, st>>>(d, sums, n);
    return cudaGetLastError();
}

}  // namespace enf
