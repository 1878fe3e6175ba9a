// wire.cuh — compact device-to-host format of one observation step, and its expansion on the host.
//
// fl_step_observe_host ships 2,438 bytes per agent over PCIe (every observation tensor as f32 / i32, the layout the policy
// reads).  Most of those bytes carry no information: about 40 % of the 31 tree nodes are the constant "no branch here"
// vector (twelve times -1), adjacency / evaluation orders are small integers stored as i32, 70 of the 83 attribute entries
// are 0/1 flags.  k_pack rewrites one environment range into a byte stream of ~1.1 KB per agent — bit-exact, every float
// that is not a flag or a -1 travels as its own 32 bits — straight into pinned host memory (the kernel's stores go over
// PCIe; nothing is staged in device memory, no size has to be known to the host beforehand), and expand_env() rebuilds
// the f32 / i32 tensors in the caller's host buffers, one environment per task of a small thread pool.
//
// Stream layout of a chunk (uint32 words): table[n_env] (word offset of the environment's block inside the chunk | dones
// ["__all__"] << 31), then the blocks in the order the CTAs finished.  A block is N agent records back to back:
//   w[0]      node mask, bit n = tree node n differs from the all -1 vector
//   w[1..3]   attribute entries 0..69 as bits (feature_parser.cpp:19-77: one-hot codes and flags, exactly 0.0 or 1.0)
//   w[4..16]  attribute entries 70..82 (floats)
//   w[17]     dist_target (float)      w[18] reward (int32)      w[19] dones[i]
//   w[20..57] 151 bytes: adjacency[30][3], node_order[31], edge_order[30] as int8 (values -2..30);  w[58..63] zero
//   then 12 floats per set bit of the node mask, in node order; the record is padded to a multiple of 8 words, so that
//   every record starts on a 32-byte boundary and every store instruction of the pack kernel covers whole sectors.
#pragma once
#include "common.cuh"

#include <cstring>
#include <vector>

namespace {

constexpr int WIRE_FIXED_WORDS = 64;
constexpr int WIRE_MAX_WORDS = (WIRE_FIXED_WORDS + FL_MAX_NODES * FL_NODE_F + 7) & ~7;   // 440
__host__ __device__ inline int wire_record_words(unsigned mask) { return (WIRE_FIXED_WORDS + FL_NODE_F * __builtin_popcount(mask) + 7) & ~7; }

struct WireSrc {   // device pointers of the environment range being packed (already advanced to its first environment)
    const float *attr, *forest, *dist_target;
    const int32_t *adjacency, *node_order, *edge_order, *rewards;
    const uint8_t *dones;
};

constexpr int PACK_THREADS = 128;

__global__ void __launch_bounds__(PACK_THREADS)
k_pack(WireSrc src, int N, int n_env, uint32_t *__restrict__ wire, uint32_t *__restrict__ cursor) {
    const int el = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ uint32_t pk_smem[];
    const int Np = (N + 3) & ~3;                       // keeps the staging areas 16-byte aligned
    uint32_t *s_mask = pk_smem, *s_off = pk_smem + Np, *stage = pk_smem + 2 * Np + 4 + warp * WIRE_MAX_WORDS;
    uint32_t *s_base = pk_smem + 2 * Np;
    const size_t a0 = (size_t)el * N;
    // pass 1: which nodes are not the all -1 vector
    for (int i = warp; i < N; i += PACK_THREADS / 32) {
        bool real = false;
        if (lane < FL_MAX_NODES) {
            const float4 *p = reinterpret_cast<const float4 *>(src.forest + (a0 + i) * (FL_MAX_NODES * FL_NODE_F) + lane * FL_NODE_F);
            const float4 x = p[0], y = p[1], z = p[2];
            real = x.x != -1.f || x.y != -1.f || x.z != -1.f || x.w != -1.f || y.x != -1.f || y.y != -1.f || y.z != -1.f ||
                   y.w != -1.f || z.x != -1.f || z.y != -1.f || z.z != -1.f || z.w != -1.f;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, real);
        if (lane == 0) s_mask[i] = m;
    }
    __syncthreads();
    if (warp == 0) {                                   // offsets of the agent records inside the block
        unsigned run = 0;
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            const unsigned len = i < N ? (unsigned)((WIRE_FIXED_WORDS + FL_NODE_F * __popc(s_mask[i]) + 7) & ~7) : 0u;
            unsigned x = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
            if (i < N) s_off[i] = run + x - len;
            run += __shfl_sync(0xFFFFFFFFu, x, 31);
        }
        if (lane == 0) {
            const unsigned base = (unsigned)((n_env + 7) & ~7) + atomicAdd(cursor, run);
            s_base[0] = base;
            wire[el] = base | ((uint32_t)(src.dones[(size_t)el * (N + 1) + N] != 0) << 31);
        }
    }
    __syncthreads();
    const unsigned base = s_base[0];
    // pass 2: one warp per agent stages the record in shared memory and streams it out with coalesced stores
    for (int i = warp; i < N; i += PACK_THREADS / 32) {
        const size_t ea = a0 + i;
        const unsigned mask = s_mask[i];
        const float *at = src.attr + ea * FL_ATTR_F;
        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, at[lane] != 0.0f), b1 = __ballot_sync(0xFFFFFFFFu, at[32 + lane] != 0.0f);
        const unsigned b2 = __ballot_sync(0xFFFFFFFFu, lane < 6 && at[64 + lane] != 0.0f);
        if (lane == 0) { stage[0] = mask; stage[1] = b0; stage[2] = b1; stage[3] = b2; }
        if (lane < 13) stage[4 + lane] = __float_as_uint(at[70 + lane]);
        if (lane == 13) stage[17] = __float_as_uint(src.dist_target[ea]);
        if (lane == 14) stage[18] = (uint32_t)src.rewards[ea];
        if (lane == 15) stage[19] = src.dones[(size_t)el * (N + 1) + i];
        uint8_t *sb = reinterpret_cast<uint8_t *>(stage + 20);
        const int32_t *adj = src.adjacency + ea * ((FL_MAX_NODES - 1) * 3);
        for (int k = lane; k < 90; k += 32) sb[k] = (uint8_t)(int8_t)adj[k];
        if (lane < FL_MAX_NODES) sb[90 + lane] = (uint8_t)(int8_t)src.node_order[ea * FL_MAX_NODES + lane];
        if (lane < FL_MAX_NODES - 1) sb[121 + lane] = (uint8_t)(int8_t)src.edge_order[ea * (FL_MAX_NODES - 1) + lane];
        if (lane == 31) sb[151] = 0;
        if (lane >= 26) stage[58 + lane - 26] = 0u;
        if (lane < FL_MAX_NODES && ((mask >> lane) & 1u)) {
            const float4 *p = reinterpret_cast<const float4 *>(src.forest + ea * (FL_MAX_NODES * FL_NODE_F) + lane * FL_NODE_F);
            float4 *q = reinterpret_cast<float4 *>(stage + WIRE_FIXED_WORDS + FL_NODE_F * __popc(mask & ((1u << lane) - 1u)));
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
        }
        __syncwarp();
        const int len = (WIRE_FIXED_WORDS + FL_NODE_F * __popc(mask) + 7) & ~7;   // the padding words carry stale bytes of the stage
        uint4 *out = reinterpret_cast<uint4 *>(wire + base + s_off[i]);
        const uint4 *st4 = reinterpret_cast<const uint4 *>(stage);
        for (int k = lane; k < len / 4; k += 32) out[k] = st4[k];                  // 16 bytes per lane, 512 contiguous bytes per store
        __syncwarp();
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct WireDst {   // host pointers of the WHOLE batch (the reference-facing layout of FlObsBuffers)
    float *attr, *forest, *dist_target;
    int32_t *adjacency, *node_order, *edge_order, *rewards;
    uint8_t *valid_actions, *dones;
};

// Rebuilds environment `e_global` (local index el of the chunk whose stream starts at `wire`) in the host tensors; returns
// the words consumed.  Portable version.
inline uint64_t expand_env_scalar(const uint32_t *wire, int el, long long e_global, int N, const WireDst &d) {
    const uint32_t tw = wire[el];
    const uint32_t *p = wire + (tw & 0x7FFFFFFFu);
    const uint32_t *const p0 = p;
    if (d.dones) d.dones[(size_t)e_global * (N + 1) + N] = (uint8_t)(tw >> 31);
    for (int i = 0; i < N; i++) {
        const size_t ea = (size_t)e_global * N + i;
        const uint32_t mask = p[0];
        if (d.attr) {
            float *a = d.attr + ea * FL_ATTR_F;
            for (int k = 0; k < 32; k++) a[k] = (float)((p[1] >> k) & 1u);
            for (int k = 0; k < 32; k++) a[32 + k] = (float)((p[2] >> k) & 1u);
            for (int k = 0; k < 6; k++) a[64 + k] = (float)((p[3] >> k) & 1u);
            std::memcpy(a + 70, p + 4, 13 * sizeof(float));
        }
        if (d.valid_actions) for (int k = 0; k < 5; k++) d.valid_actions[ea * 5 + k] = (uint8_t)((p[3] >> (1 + k)) & 1u);
        if (d.dist_target) std::memcpy(d.dist_target + ea, p + 17, 4);
        if (d.rewards) d.rewards[ea] = (int32_t)p[18];
        if (d.dones) d.dones[(size_t)e_global * (N + 1) + i] = (uint8_t)p[19];
        const int8_t *sb = reinterpret_cast<const int8_t *>(p + 20);
        if (d.adjacency) { int32_t *q = d.adjacency + ea * ((FL_MAX_NODES - 1) * 3); for (int k = 0; k < 90; k++) q[k] = sb[k]; }
        if (d.node_order) { int32_t *q = d.node_order + ea * FL_MAX_NODES; for (int k = 0; k < FL_MAX_NODES; k++) q[k] = sb[90 + k]; }
        if (d.edge_order) { int32_t *q = d.edge_order + ea * (FL_MAX_NODES - 1); for (int k = 0; k < FL_MAX_NODES - 1; k++) q[k] = sb[121 + k]; }
        const uint32_t *nodes = p + WIRE_FIXED_WORDS;
        if (d.forest) {
            float *f = d.forest + ea * (FL_MAX_NODES * FL_NODE_F);
            for (int n = 0; n < FL_MAX_NODES; n++) {
                if ((mask >> n) & 1u) { std::memcpy(f + n * FL_NODE_F, nodes, FL_NODE_F * sizeof(float)); nodes += FL_NODE_F; }
                else for (int k = 0; k < FL_NODE_F; k++) f[n * FL_NODE_F + k] = -1.0f;
            }
        }
        p += wire_record_words(mask);
    }
    return (uint64_t)(p - p0);
}

#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
}  // namespace
#include <immintrin.h>
namespace {
struct WireBitLut {   // byte -> eight floats 0.0 / 1.0
    alignas(32) float f[256][8];
    WireBitLut() { for (int b = 0; b < 256; b++) for (int k = 0; k < 8; k++) f[b][k] = (float)((b >> k) & 1); }
};
inline const WireBitLut &wire_bit_lut() { static const WireBitLut lut; return lut; }

__attribute__((target("avx2")))
inline void wire_widen(int32_t *q, const int8_t *src, int n) {   // int8 -> int32, eight at a time
    int k = 0;
    for (; k + 8 <= n; k += 8)
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(q + k), _mm256_cvtepi8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(src + k))));
    for (; k < n; k++) q[k] = src[k];
}

// dst <- src with non-temporal 16-byte stores where dst is aligned (head and tail: ordinary stores); both 4-byte aligned
__attribute__((target("avx2")))
inline void wire_stream_copy(void *dst, const void *src, size_t bytes) {
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head > bytes) head = bytes;
    std::memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    for (; bytes >= 16; bytes -= 16, d += 16, s += 16)
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), _mm_loadu_si128(reinterpret_cast<const __m128i *>(s)));
    std::memcpy(d, s, bytes);
}

// The same with AVX2: flags through a byte -> 8 floats table, int8 -> int32 eight at a time, and non-temporal stores so that
// 125 MB per step do not first pull their cache lines in: nt_f for the forest rows (61 % of the bytes; every 48-byte node
// slot is 16-byte aligned), nt for the other tensors as well (staged per environment, see below).
__attribute__((target("avx2")))
inline uint64_t expand_env_avx2(const uint32_t *wire, int el, long long e_global, int N, const WireDst &d, bool nt_f, bool nt) {
    const WireBitLut &lut = wire_bit_lut();
    const uint32_t tw = wire[el];
    const uint32_t *p = wire + (tw & 0x7FFFFFFFu);
    const uint32_t *const p0 = p;
    if (d.dones) d.dones[(size_t)e_global * (N + 1) + N] = (uint8_t)(tw >> 31);
    const __m128 minus1 = _mm_set1_ps(-1.0f);
    const bool nt_forest = nt_f && d.forest && (reinterpret_cast<uintptr_t>(d.forest) & 15) == 0;
    // with non-temporal stores the small tensors of the environment (attributes, adjacency, orders: rows that are not multiples
    // of a cache line) are assembled in a thread-local block first and leave as ONE sequential stream per tensor, so that the
    // write-combining buffers always see whole lines; the forest rows are 16-byte aligned and stream directly
    static thread_local std::vector<int32_t> scratch;
    float *s_attr = nullptr;
    int32_t *s_adj = nullptr, *s_no = nullptr, *s_eo = nullptr;
    if (nt) {
        const size_t need = (size_t)N * (FL_ATTR_F + 90 + FL_MAX_NODES + FL_MAX_NODES - 1) + 64;
        if (scratch.size() < need) scratch.resize(need);
        s_attr = reinterpret_cast<float *>(scratch.data());
        s_adj = scratch.data() + (size_t)N * FL_ATTR_F + 8;
        s_no = s_adj + (size_t)N * 90 + 8;
        s_eo = s_no + (size_t)N * FL_MAX_NODES + 8;
    }
    for (int i = 0; i < N; i++) {
        const size_t ea = (size_t)e_global * N + i;
        const uint32_t mask = p[0];
        if (d.attr) {
            float *a = nt ? s_attr + (size_t)i * FL_ATTR_F : d.attr + ea * FL_ATTR_F;
            const uint8_t *bits = reinterpret_cast<const uint8_t *>(p + 1);          // bytes 0..8: entries 0..71 (70, 71 are zero bits)
            for (int k = 0; k < 9; k++) _mm256_storeu_ps(a + 8 * k, _mm256_load_ps(lut.f[bits[k]]));
            std::memcpy(a + 70, p + 4, 13 * sizeof(float));
        }
        if (d.valid_actions) for (int k = 0; k < 5; k++) d.valid_actions[ea * 5 + k] = (uint8_t)((p[3] >> (1 + k)) & 1u);
        if (d.dist_target) std::memcpy(d.dist_target + ea, p + 17, 4);
        if (d.rewards) d.rewards[ea] = (int32_t)p[18];
        if (d.dones) d.dones[(size_t)e_global * (N + 1) + i] = (uint8_t)p[19];
        const int8_t *sb = reinterpret_cast<const int8_t *>(p + 20);
        if (nt) {
            wire_widen(s_adj + (size_t)i * 90, sb, 90);
            wire_widen(s_no + (size_t)i * FL_MAX_NODES, sb + 90, FL_MAX_NODES);
            wire_widen(s_eo + (size_t)i * (FL_MAX_NODES - 1), sb + 121, FL_MAX_NODES - 1);
        } else {
            if (d.adjacency) wire_widen(d.adjacency + ea * ((FL_MAX_NODES - 1) * 3), sb, 90);
            if (d.node_order) wire_widen(d.node_order + ea * FL_MAX_NODES, sb + 90, FL_MAX_NODES);
            if (d.edge_order) wire_widen(d.edge_order + ea * (FL_MAX_NODES - 1), sb + 121, FL_MAX_NODES - 1);
        }
        const float *nodes = reinterpret_cast<const float *>(p + WIRE_FIXED_WORDS);
        if (d.forest) {
            float *f = d.forest + ea * (FL_MAX_NODES * FL_NODE_F);
            for (int n = 0; n < FL_MAX_NODES; n++, f += FL_NODE_F) {
                __m128 x = minus1, y = minus1, z = minus1;
                if ((mask >> n) & 1u) { x = _mm_loadu_ps(nodes); y = _mm_loadu_ps(nodes + 4); z = _mm_loadu_ps(nodes + 8); nodes += FL_NODE_F; }
                if (nt_forest) { _mm_stream_ps(f, x); _mm_stream_ps(f + 4, y); _mm_stream_ps(f + 8, z); }
                else { _mm_storeu_ps(f, x); _mm_storeu_ps(f + 4, y); _mm_storeu_ps(f + 8, z); }
            }
        }
        p += wire_record_words(mask);
    }
    if (nt_forest && !nt) _mm_sfence();
    if (nt) {
        const size_t a0 = (size_t)e_global * N;
        if (d.attr) wire_stream_copy(d.attr + a0 * FL_ATTR_F, s_attr, (size_t)N * FL_ATTR_F * 4);
        if (d.adjacency) wire_stream_copy(d.adjacency + a0 * 90, s_adj, (size_t)N * 90 * 4);
        if (d.node_order) wire_stream_copy(d.node_order + a0 * FL_MAX_NODES, s_no, (size_t)N * FL_MAX_NODES * 4);
        if (d.edge_order) wire_stream_copy(d.edge_order + a0 * (FL_MAX_NODES - 1), s_eo, (size_t)N * (FL_MAX_NODES - 1) * 4);
        _mm_sfence();
    }
    return (uint64_t)(p - p0);
}
#define FL_WIRE_HAVE_AVX2 1
#endif

inline uint64_t expand_env(const uint32_t *wire, int el, long long e_global, int N, const WireDst &d, int mode) {
#ifdef FL_WIRE_HAVE_AVX2
    if (mode > 0 && __builtin_cpu_supports("avx2")) return expand_env_avx2(wire, el, e_global, N, d, mode > 1, mode > 2);
#endif
    return expand_env_scalar(wire, el, e_global, N, d);
}

}  // namespace
