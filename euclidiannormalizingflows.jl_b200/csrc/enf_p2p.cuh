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

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// all threads of the (single) CTA call this; sums[0..n) is replaced by the sum over the ranks.
//
// Flag-in-data exchange: every double is stored into the peers' buffers as two 8-byte words, each carrying 32 data bits
// and the low 32 bits of this exchange's sequence number.  An aligned 8-byte access is single-copy atomic, so a reader
// that sees the sequence number in a word has its data bits too: no fence, no separate flag, no second round trip -
// the exchange costs one NVLink store latency plus the skew of the ranks (the first version published the data, fenced
// at system scope, released a flag and polled it: ~9 us of the D = 1 fit step on 8 GPUs).  Slots are double-buffered by
// sequence parity: a rank can be at most one exchange ahead of the slowest one.
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
    const unsigned long long tag = seq % 0xFFFFFFFFull + 1;      // 1 .. 2^32 - 1, never 0: the buffers start zeroed
    const size_t par_off = P2P_DATA_OFF + size_t(seq & 1) * size_t(R) * P2P_SLOT * 16;
    // publish: my sums -> slot [rank] of every other rank's buffer
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(sums[i]));
        const unsigned long long w0 = (bits << 32) | (tag & 0xFFFFFFFFull), w1 = (bits & 0xFFFFFFFF00000000ull) | (tag & 0xFFFFFFFFull);
        for (int r = 0; r < R; ++r) {
            if (r == d.rank) continue;
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(d.peer[r]) + par_off) +
                                      (size_t(d.rank) * P2P_SLOT + i) * 2;
            st_relaxed_sys(dst, w0);
            st_relaxed_sys(dst + 1, w1);
        }
    }
    // gather: the other ranks' words arrive in MY buffer; add in rank order (bitwise identical on every rank).  The wait is
    // bounded: a missing rank must not hang the GPU
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(local + par_off);
    bool timed_out = false;
    for (int i = tid; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < R; ++r) {
            if (r == d.rank) { a += sums[i]; continue; }
            const unsigned long long* src = mine + (size_t(r) * P2P_SLOT + i) * 2;
            unsigned long long w0 = ld_relaxed_sys(src), w1 = ld_relaxed_sys(src + 1);
            long long spins = 0;
            while (!timed_out && ((w0 & 0xFFFFFFFFull) != (tag & 0xFFFFFFFFull) || (w1 & 0xFFFFFFFFull) != (tag & 0xFFFFFFFFull))) {
                if (++spins > P2P_SPIN_LIMIT) { timed_out = true; break; }
                __nanosleep(spins < 4096 ? 20 : 256);
                w0 = ld_relaxed_sys(src);
                w1 = ld_relaxed_sys(src + 1);
            }
            a += __longlong_as_double(static_cast<long long>((w1 & 0xFFFFFFFF00000000ull) | (w0 >> 32)));
        }
        sums[i] = a;
    }
    if (timed_out) {
        // a rank never arrived: raise the sticky error flag in EVERY rank's buffer, so that the ranks that do complete this
        // exchange later fail the same way at their next p2p_check (enf_abi.cu)
        for (int r = 0; r < R; ++r)
            *reinterpret_cast<volatile int*>(static_cast<unsigned char*>(d.peer[r]) + P2P_ERR_OFF) = 1;
        __threadfence_system();
        s_timeout = 1;
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
