// Micro-benchmark: delivery rate of TMA tile loads (bf16, 64-column x R-row boxes, SWIZZLE_128B) into one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../flatland-marl_b200/csrc/policy -o tma_rate tma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) k_tma(const __grid_constant__ CUtensorMap tm, long long *out, int nbox, int box_bytes, int box_rows, int rounds) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) mbar_init(&bar[i], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        long long stamps[16];
        for (int r = 0; r < rounds; r++) {
            for (int i = 0; i < nbox; i++) {
                mbar_expect_tx(&bar[i], box_bytes);
                const int tile = (blockIdx.x * rounds + r) * nbox + i;   // distinct data for every box
                tma_load_2d(smem_u32(smem + (size_t)i * box_bytes), &tm, (tile & 3) * 64, (tile >> 2) * box_rows, &bar[i]);
            }
            for (int i = 0; i < nbox; i++) {
                mbar_wait(&bar[i], r & 1);
                if (r == rounds - 1) stamps[i] = clock64() - t0;
            }
            if (r == 0) t0 = clock64() - (clock64() - t0);   // keep t0
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            for (int i = 0; i < nbox; i++) out[1 + i] = stamps[i];
        }
    }
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const size_t rows = 1 << 20, cols = 256;      // 512 MB? no: 1M rows x 256 cols x 2 B = 512 MB
    void *d;
    cudaMalloc(&d, rows * cols * 2);
    cudaMemset(d, 0, rows * cols * 2);
    long long *out;
    cudaMalloc(&out, 256);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int box_rows : {128, 256}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {cols, rows};
        cuuint64_t strides[1] = {cols * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int box_bytes = box_rows * 128;
        for (int ctas : {1, 148}) {
            for (int nbox : {1, 2, 4, 8}) {
                if (nbox * box_bytes > 190 * 1024) continue;
                const int rounds = 20;
                long long h[17];
                k_tma<<<ctas, 128, 200 * 1024>>>(tm, out, nbox, box_bytes, box_rows, rounds);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                printf("box %3d rows (%2d KB) ctas=%3d boxes in flight=%d: %s  %.0f cycles per round -> %.1f B/clk/SM; last round completions:", box_rows,
                       box_bytes / 1024, ctas, nbox, cudaGetErrorString(e), (double)h[0] / rounds, (double)nbox * box_bytes * rounds / (double)h[0]);
                for (int i = 0; i < nbox; i++) printf(" %lld", h[1 + i] - (i ? 0 : 0));
                printf("\n");
            }
        }
    }
    return 0;
}
