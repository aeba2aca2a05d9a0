// Device side of the peer-memory all-reduce (see enf_p2p.cu): callable from any one-CTA kernel of 256 threads.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "enf_launch.h"

namespace enf {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// all threads of the (single) CTA call this; sums[0..n) is replaced by the sum over the ranks
__device__ __forceinline__ void p2p_allreduce_block(const P2PDesc& d, double* __restrict__ sums, int n) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    const int tid = threadIdx.x, R = d.nranks;
    unsigned char* local = static_cast<unsigned char*>(d.peer[d.rank]);
    if (tid == 0) {
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(local + P2P_SEQ_OFF);
        s_seq = *ctr + 1;
        *ctr = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const size_t par_off = P2P_DATA_OFF + size_t(seq & 1) * size_t(R) * P2P_SLOT * sizeof(double);
    // publish: my sums -> slot [rank] of every rank's buffer (own buffer included)
    for (int r = 0; r < R; ++r) {
        double* dst = reinterpret_cast<double*>(static_cast<unsigned char*>(d.peer[r]) + par_off) + size_t(d.rank) * P2P_SLOT;
        for (int i = tid; i < n; i += blockDim.x) dst[i] = sums[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < R)
        st_release_sys(reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(d.peer[tid]) + P2P_FLAG_OFF) + d.rank, seq);
    // gather: wait for every rank's sequence number in MY buffer (bounded: a missing rank must not hang the GPU)
    if (tid < R) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(local + P2P_FLAG_OFF) + tid;
        long long spins = 0;
        while (ld_acquire_sys(flag) < seq) {
            if (++spins > P2P_SPIN_LIMIT) {
                // a rank never arrived: raise the sticky error flag in EVERY rank's buffer, so that the ranks that do
                // complete this exchange later fail the same way at their next p2p_check (enf_abi.cu)
                for (int r = 0; r < R; ++r)
                    *reinterpret_cast<volatile int*>(static_cast<unsigned char*>(d.peer[r]) + P2P_ERR_OFF) = 1;
                __threadfence_system();
                s_timeout = 1;
                break;
            }
            __nanosleep(spins < 4096 ? 32 : 256);
        }
    }
    __syncthreads();
    const double* mine = reinterpret_cast<const double*>(local + par_off);
    for (int i = tid; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < R; ++r) a += mine[size_t(r) * P2P_SLOT + i];
        sums[i] = a;
    }
    __syncthreads();
    // a rank did not show up (here, or in an earlier exchange of any rank: the error flag is sticky and raised in every
    // rank's buffer): poison the last value (the sample count of the batch) so that the host notices with the copy it
    // makes anyway
    if (tid == 0 && (s_timeout || *reinterpret_cast<volatile int*>(local + P2P_ERR_OFF) != 0))
        sums[n - 1] = __longlong_as_double(0x7FF8000000000000LL);
}

}  // namespace
}  // namespace enf
