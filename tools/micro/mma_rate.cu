// Micro-benchmark: issue rate of tcgen05.mma (bf16, M=128, N in {64,128,256}, K=16, SS mode, 128B-swizzled K-major tiles)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../flatland-marl_b200/csrc/policy -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

template <int N>
__global__ void __launch_bounds__(128, 1) k_rate(long long *out, int iters, int distinct, int commit_every) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t bars2[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 * 6) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (warp == 0) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; i++) mbar_init(&bars2[i], 1); mbar_init_fence(); }
        __syncwarp();
        tmem_alloc(&slot, 512);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 2 * 16384;
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t st = distinct ? (uint32_t)(i & 1) * 16384u : 0u;
            const uint64_t ad = desc_sw128(a0 + st), bd = desc_sw128(b0 + st * 2);
            for (int k = 0; k < 4; k++) mma_bf16(tmem + (i & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
            if (commit_every && (i % commit_every) == commit_every - 1) mma_commit(&bars2[i & 7]);
        }
        mma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N> void run(long long *d, int ctas) {
    cudaFuncSetAttribute(k_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 6 + 2048);
    for (int distinct = 0; distinct < 2; distinct++) {
        const int commit_every = distinct;     // second run: a tcgen05.commit after every 4 MMAs (one k-block), nobody waits on it
        long long h[2];
        const int iters = 2000;
        k_rate<N><<<ctas, 128, 16384 * 6 + 2048>>>(d, iters, distinct, commit_every);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%3d ctas=%3d distinct stages + commit per k-block=%d: %s  issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", N, ctas, distinct, cudaGetErrorString(e),
               (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), 128 * N / 256);
    }
}

int main() {
    long long *d;
    cudaMalloc(&d, 64);
    run<64>(d, 1); run<128>(d, 1); run<256>(d, 1);
    run<128>(d, 148); run<256>(d, 148);
    return 0;
}
