// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, M = 128, K = 8) by N and by where the A operand lives
// (shared memory descriptor vs tensor memory).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -I euclidiannormalizingflows.jl_b200/csrc -I include tools/mma_rate.cu -o build/mma_rate -lcuda
#include <cstdio>
#include "enf_tc.cuh"
using namespace enf;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 SS, 1 TS, 2 alternating TS(N) + SS(N/2) (the GEMM1 pattern)
template <int N, int MODE>
__global__ void k(long long* out, int n_mma) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = slot;
    if (warp == 0) {
        const uint64_t da = make_desc_kmajor<32>(smem), db = make_desc_kmajor<32>(smem + 16384);
        constexpr uint32_t idN = make_idesc_tf32(128, N), idH = make_idesc_tf32(128, N / 2 < 8 ? 8 : N / 2);
        long long t0 = clock64();
        if (lane == 0) {
            for (int i = 0; i < n_mma; ++i) {
                const uint64_t adv = uint64_t(((i & 3) * 32) >> 4);
                if (MODE == 0) umma_tf32(t, da + adv, db + adv, idN, 1);
                else if (MODE == 1) umma_ts(t, t + 256 + (i & 3) * 8, db + adv, idN, 1);
                else { umma_ts(t, t + 256 + (i & 3) * 8, db + adv, idN, 1); umma_tf32(t + 128, da + adv, db + adv, idH, 1); }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    else mbar_wait(&bar, 0);      // the other warps (if any) poll the completion barrier the way the kernels' consumer warps do
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(t, 512);
}

// the kernels' issue pattern: per chunk two waits on (completed) mbarriers, a fence, 8 MMAs (TS N=128 + TS N=64), a commit.
// out[2 + i] = clock after issuing MMA i of the first 64
__global__ void k_chunked(long long* out, int n_chunks, int n_wait, int flags = 7) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, done[2], ready;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done[0], 1); mbar_init(&done[1], 1); mbar_init(&ready, 1); mbar_arrive(&ready);
                            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = slot;
    if (warp == 0) {
        const uint64_t db = make_desc_kmajor<32>(smem + 16384);
        constexpr uint32_t id128 = make_idesc_tf32(128, 128), id64 = make_idesc_tf32(128, 64);
        long long t0 = clock64();
        if (!(flags & 4)) {        // the whole loop in one thread: no per-chunk divergence / reconvergence
            if (lane == 0)
                for (int c = 0; c < n_chunks; ++c) {
                    for (int w = 0; w < n_wait; ++w) mbar_wait(&ready, 0);
                    if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t adv = uint64_t((j * 32) >> 4);
                        umma_ts(t, t + 256 + j * 8, db + adv, id128, 1);
                        umma_ts(t + 64, t + 128 + j * 8, db + adv, id64, 1);
                    }
                    if (flags & 1) umma_commit(&done[c & 1]);
                }
            __syncwarp();
        } else
        for (int c = 0; c < n_chunks; ++c) {
            for (int w = 0; w < n_wait; ++w) mbar_wait(&ready, 0);          // completed long ago
            if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                for (int j = 0; j < 4; ++j) {
                    const uint64_t adv = uint64_t((j * 32) >> 4);
                    umma_ts(t, t + 256 + j * 8, db + adv, id128, 1);
                    if (c < 8) out[2 + c * 8 + 2 * j] = clock64() - t0;
                    umma_ts(t + 64, t + 128 + j * 8, db + adv, id64, 1);
                    if (c < 8) out[2 + c * 8 + 2 * j + 1] = clock64() - t0;
                }
                if (flags & 1) umma_commit(&done[c & 1]);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0) { out[0] = n_chunks; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(t, 512);
}

template <int N, int MODE>
void run(const char* name, long long* d, int threads = 32) {
    const int n = 512;
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        k<N, MODE><<<1, threads, 64 * 1024>>>(d, n);
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    }
    printf("%-28s %2d warps N=%3d: issue %7.1f clk/iter, complete %7.1f clk/iter   (%s)\n", name, threads / 32, N, double(h[0]) / n, double(h[1]) / n,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d;
    cudaMalloc(&d, 1024);
    run<256, 0>("SS (A smem)", d); run<128, 0>("SS (A smem)", d); run<64, 0>("SS (A smem)", d);
    run<256, 1>("TS (A tmem)", d); run<128, 1>("TS (A tmem)", d); run<64, 1>("TS (A tmem)", d);
    run<128, 2>("TS N + SS N/2 (GEMM1 pattern)", d); run<256, 2>("TS N + SS N/2", d);
    run<128, 1>("TS, 15 warps polling", d, 512); run<128, 0>("SS, 15 warps polling", d, 512); run<128, 2>("TS N + SS N/2, 15 polling", d, 512);
    run<128, 1>("TS, 7 warps polling", d, 256);
    cudaFuncSetAttribute(k_chunked, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int fl = 0; fl < 8; ++fl) {
        long long h[66];
        for (int rep = 0; rep < 2; ++rep) { k_chunked<<<1, 32, 64 * 1024>>>(d, 64, 1, fl); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); }
        printf("chunked, 1 wait, flags commit=%d fence=%d per-chunk-branch=%d: %.1f clk per chunk\n", fl & 1, (fl >> 1) & 1, (fl >> 2) & 1, double(h[1]) / 64);
    }
    for (int nw = 0; nw <= 2; ++nw) {
        long long h[66];
        for (int rep = 0; rep < 2; ++rep) { k_chunked<<<1, 32, 64 * 1024>>>(d, 64, nw, 7); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); }
        printf("chunked (8 MMAs: TS128+TS64 x4, commit), %d waits per chunk: %.1f clk per chunk (floor 440)\n  issue clocks:", nw, double(h[1]) / 64);
        for (int i = 0; i < 32; ++i) printf(" %lld", h[2 + i]);
        printf("\n");
    }
    return 0;
}
