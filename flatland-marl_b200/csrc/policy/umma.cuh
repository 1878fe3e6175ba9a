// umma.cuh — thin PTX wrappers for the sm_100a pieces the policy kernels use: mbarriers, cp.async,
// tcgen05 (TMEM allocation, single-thread MMA issue with shared-memory descriptors, commit, TMEM loads).
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors" (K-major operands, 128-byte swizzle).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded spin: a protocol error traps (visible as a launch failure) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- cp.async (16-byte, L2 only) and the generic -> async proxy fence ----------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// The mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
// (.noinc: the arrival counts towards the barrier's expected count), so a producer never blocks on its own loads.
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Loads that are issued where they are written (the compiler may not sink them to their first use): the producers fetch
// the list entries of the NEXT tile before they queue the current tile's cp.async gathers, because a load queued behind
// two dozen gathers comes back only when they do.
__device__ __forceinline__ uint32_t ldg_u32_now(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u8_now(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- bulk stores shared -> global (TMA engine, 1-D): one thread moves a whole row of the output tile ----------
__device__ __forceinline__ void bulk_store(void *dst_global, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- TMA tile loads (2-D tensor map, 128-byte swizzle: the box lands in exactly the layout desc_sw128 describes) ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---- tcgen05 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major bf16 tile stored as rows of 128 bytes (64 elements) with the
// 128-byte swizzle (16-byte chunk c of row r lives at chunk c ^ (r & 7)); 8-row groups are 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows * 128 B   [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)       [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B             [61,64)
    return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> f32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same, B operand MN-major (its N index is the contiguous one): a [K rows][64 N-elements] tile with the 128-byte swizzle,
// i.e. exactly what a TMA box of 64 columns x K rows leaves in shared memory.
__host__ __device__ constexpr uint32_t idesc_bf16_bmn(int M, int N) { return idesc_bf16(M, N) | (1u << 16); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: 32 lanes (this warp's quarter) x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- math ---------------------------------------------------------------------------------------------------
// Branch-free forms for the epilogues (one warp per scheduler has no other warps to hide a divergent libm call):
// tanh.approx.f32 is one MUFU op (max relative error 2^-11, below the bf16 rounding of the stored result);
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
// Two tanh per MUFU operation (tanh.approx.f16x2).  The special-function unit retires 16 lanes per clock per SM, and the
// LSTM gates need four transcendentals per hidden unit: with scalar tanh.approx.f32 the leaf kernel alone is 4096 MUFU
// cycles per 128-node tile.  The f16 result carries the same ~2^-11 relative error as tanh.approx.f32.
__device__ __forceinline__ float2 tanh2_fast(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t *>(&h)));
    return __half22float2(*reinterpret_cast<const __half2 *>(&r));
}
__device__ __forceinline__ float2 sigmoid2_fast(float a, float b) {
    const float2 t = tanh2_fast(0.5f * a, 0.5f * b);
    return make_float2(fmaf(0.5f, t.x, 0.5f), fmaf(0.5f, t.y, 0.5f));
}
__device__ __forceinline__ float2 gelu2_erf(float x, float y) {
    const float ux = fminf(x * x, 49.0f), uy = fminf(y * y, 49.0f);
    float px = fmaf(ux, -3.51516783e-4f, 3.70056460e-2f), py = fmaf(uy, -3.51516783e-4f, 3.70056460e-2f);
    px = fmaf(ux, px, 7.97507884e-1f);
    py = fmaf(uy, py, 7.97507884e-1f);
    const float2 t = tanh2_fast(x * px, y * py);
    const float hx = 0.5f * x, hy = 0.5f * y;
    return make_float2(fmaf(hx, t.x, hx), fmaf(hy, t.y, hy));
}
// GELU (erf form, nn.GELU default) as 0.5 x (1 + tanh(x P(x^2))): P is a quadratic fitted to the erf form (max absolute
// deviation 2.5e-5 over all x; the textbook tanh form is off by 4.7e-4), x^2 clamped where tanh has saturated.
// 8 instructions, one MUFU — the erf-by-exp form (17 instructions, two MUFU) made the epilogue the pace of k_lin.
__device__ __forceinline__ float gelu_erf(float x) {
    const float u = fminf(x * x, 49.0f);
    float pl = fmaf(u, -3.51516783e-4f, 3.70056460e-2f);
    pl = fmaf(u, pl, 7.97507884e-1f);
    const float h = 0.5f * x;
    return fmaf(h, tanh_fast(x * pl), h);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

}  // namespace umma
