// Warp-level PTX primitives of the SIMT / mma.sync kernels (attention, row-wise): cp.async, ldmatrix, mma.sync.m16n8k16,
// named barriers, cluster barrier / DSMEM stores, MUFU approximations.
//
// Every primitive is ONE inline-PTX statement behind a function, so the kernels that use them contain no inline PTX of
// their own.  That is what lets tests/emu/ compile the very same kernel sources with g++ (-DDSHEG_EMU swaps this header's
// PTX bodies for host implementations that model the documented fragment layouts) and run them, thread for thread, on a
// CPU: index math, swizzles, fragment maps and barrier protocols are checked against the oracle without a GPU.  The
// product build never defines DSHEG_EMU; the emulation is test infrastructure and is not linked into the library.
#pragma once
#include "common.cuh"

#ifdef DSHEG_EMU
#include "emu_prims.h"   // tests/emu/emu_prims.h: same names, host bodies
#else

namespace dsheg {
namespace prims {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// named barrier `id` (1..15) over `nthreads` threads of the CTA
template <int NTHREADS> __device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory"); }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// packed fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, two IEEE fp32 operations per lane and issue slot)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// ---- thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without release semantics: signals only "this CTA has started" (no memory is published, so no fence is needed)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store (a, b) at the address `local_addr` of CTA `peer`'s shared memory
__device__ __forceinline__ void st_peer_f32x2(uint32_t local_addr, uint32_t peer, float a, float b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(peer));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}

}  // namespace prims
}  // namespace dsheg

// dynamic shared memory of the running CTA
#define DSHEG_DYN_SMEM(name, align) extern __shared__ __align__(align) uint8_t name[]

#endif  // DSHEG_EMU
