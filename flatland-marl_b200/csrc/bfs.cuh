// bfs.cuh — k_bfs: DistanceMap._compute (distance_map.py:57-160), reset-time only.
#pragma once
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// k_bfs: DistanceMap (distance_map.py:57-160) as a level-synchronous pull BFS over (cell, orientation)
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(1024) k_bfs(FlBatch b, const int32_t *__restrict__ env_ids) {
    const int H = (int)b.H, W = (int)b.W, HW = H * W, ns = (int)b.n_slots;
    const int e = env_ids ? env_ids[blockIdx.x / ns] : (int)(blockIdx.x / ns), s = blockIdx.x % ns;   // env_ids: only the listed environments
    extern __shared__ __align__(16) unsigned char smraw[];
    uint16_t *out = b.dist + (size_t)e * b.dist_stride + (size_t)s * HW * 4;
    uint16_t *dd = SMEM ? reinterpret_cast<uint16_t *>(smraw) : out;
    const uint16_t *__restrict__ g = b.grid + (size_t)e * b.grid_stride;
    const int tr = b.slot_rc[((size_t)e * ns + s) * 2], tc = b.slot_rc[((size_t)e * ns + s) * 2 + 1];
    for (int k = threadIdx.x; k < HW * 4; k += blockDim.x) dd[k] = FL_DIST_INF;
    __syncthreads();
    if (tr >= 0) {
        if (threadIdx.x < 4) dd[(tr * W + tc) * 4 + threadIdx.x] = 0;
        __syncthreads();
        for (int level = 1; level < 0xFFFF; level++) {
            int changed = 0;
            for (int cell = threadIdx.x; cell < HW; cell += blockDim.x) {
                const unsigned gc = __ldg(g + cell);
                if (!gc) continue;
                const int r = cell / W, c = cell - r * W;
                for (int o = 0; o < 4; o++) {
                    if (dd[cell * 4 + o] != FL_DIST_INF) continue;
                    const int nb = nibble(gc, o);
                    for (int m = 0; m < 4; m++) {
                        if (!tbit(nb, m)) continue;
                        const int rr = r + d_row(m), cc = c + d_col(m);
                        if (rr < 0 || cc < 0 || rr >= H || cc >= W) continue;
                        // (compute-sanitizer racecheck flags this line: a thread may read a neighbour's distance while its
                        // owner sets it to `level`.  Either value it can see — infinity before, `level` after — differs from
                        // level - 1, so the comparison has one outcome; 16-bit aligned stores are not torn.)
                        if (dd[(rr * W + cc) * 4 + m] == level - 1) { dd[cell * 4 + o] = (uint16_t)level; changed = 1; break; }
                    }
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k < HW * 4; k += blockDim.x) out[k] = dd[k];
    }
}


}  // namespace
