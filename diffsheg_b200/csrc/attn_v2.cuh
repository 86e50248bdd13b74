// Linear-attention core + StylizationBlock prologue, bf16 fast path (D = 512, 8 heads of 64, T <= 96).
//
// One CTA per sample, one WARP per head (8 warps).  Per head (reference transformer.py:122-128):
//   K' = softmax_t(K)   Q' = softmax_d(Q)   A = K'^T V  [64x64]   Y = Q' A  [T x 64]
// * K and V head tiles land in shared memory with cp.async (16-B chunks, XOR-swizzled so that
//   ldmatrix is conflict-free without padding); Q goes global -> registers directly as mma A fragments.
// * both contractions run on tensor cores with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate):
//   the 64x64 per-head products are far too small for a tcgen05/TMEM pipeline, and the kernel only has
//   to keep up with HBM (32 FLOP per byte).  A^T = V^T K' is computed so that its accumulator fragments
//   ARE the B-operand fragments of the second product (pure register permutation, no smem round trip).
// * softmax denominators are applied after the products (per column of A, per row of Y) in fp32.
// * Y (bf16) replaces the dead V tile in smem; after a CTA barrier each warp normalises full 512-wide rows
//   (LayerNorm over all heads), applies (1+scale), shift and SiLU (transformer.py:92-96) and writes z with
//   1 KB coalesced row stores.  The attention output never round-trips through HBM before the out-proj GEMM.
// Algorithmic HBM traffic: read q,k,v + write z = 4 * T * 512 * 2 bytes per sample.
#pragma once
#include "common.cuh"

namespace dsheg {
namespace av2 {

constexpr int TP = 96;                  // padded time rows (6 m-tiles of 16)
constexpr int HD = 64, NH = 8, D = 512;
constexpr int TILE_BYTES = TP * HD * 2;  // 12288
constexpr int SMEM_BYTES = NH * 2 * TILE_BYTES + NH * HD * 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w);
  return make_float2(__bfloat162float(v.x), __bfloat162float(v.y));
}
// byte offset of 16-B chunk `c` (0..7) of row `r` inside a swizzled [TP][64] bf16 tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// Q fragments of one 16-row m-tile straight from global memory (rows >= T read as zero)
__device__ __forceinline__ void load_q_tile(const bf16* qhead, int row0, int T, int g, int q, uint32_t (&qa)[4][4]) {
  const int r0 = row0 + g, r1 = row0 + g + 8;
  const uint32_t* p0 = reinterpret_cast<const uint32_t*>(qhead + (size_t)r0 * (3 * D));
  const uint32_t* p1 = reinterpret_cast<const uint32_t*>(qhead + (size_t)r1 * (3 * D));
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    qa[ks][0] = r0 < T ? __ldg(p0 + ks * 8 + q) : 0u;
    qa[ks][1] = r1 < T ? __ldg(p1 + ks * 8 + q) : 0u;
    qa[ks][2] = r0 < T ? __ldg(p0 + ks * 8 + 4 + q) : 0u;
    qa[ks][3] = r1 < T ? __ldg(p1 + ks * 8 + 4 + q) : 0u;
  }
}

__global__ void __launch_bounds__(256, 1)
attn_v2_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ z, int T, int B, const float* __restrict__ ln_g,
               const float* __restrict__ ln_b, const float* __restrict__ ss, int ss_ld) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int smp = blockIdx.x;
  const size_t row0 = (size_t)smp * T;
  uint8_t* Ks = sm + warp * 2 * TILE_BYTES;
  uint8_t* Vs = Ks + TILE_BYTES;
  float* inv = reinterpret_cast<float*>(sm + NH * 2 * TILE_BYTES) + warp * HD;
  const uint32_t ks_addr = smem_addr(Ks), vs_addr = smem_addr(Vs);
  const bf16* qhead = qkv + row0 * (3 * D) + warp * HD;
  const int n_mt = (T + 15) >> 4;  // 16-row tiles that contain valid frames

  // ---- 1. K and V head tiles -> smem (cp.async), padding rows of V zeroed
  for (int i = lane; i < T * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    const bf16* src = qhead + (size_t)r * (3 * D) + c * 8;
    cp_async16(ks_addr + swz(r, c), src + D);
    cp_async16(vs_addr + swz(r, c), src + 2 * D);
  }
  for (int i = lane; i < (n_mt * 16 - T) * 8; i += 32) {
    const int r = T + (i >> 3), c = i & 7;
    *reinterpret_cast<uint4*>(Vs + swz(r, c)) = make_uint4(0, 0, 0, 0);
  }
  // ---- 2. first Q m-tile (global -> registers) overlaps the cp.async latency
  uint32_t qa[4][4];
  load_q_tile(qhead, 0, T, g, q, qa);
  cp_async_wait_all();
  __syncwarp();

  // ---- 3. softmax over time, per K column: lane owns columns 2*lane, 2*lane+1 (one 32-bit word per row)
  {
    const int c = lane >> 2, w = lane & 3;
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int r = 0; r < T; ++r) {
      const float2 v = unpack2(*reinterpret_cast<const uint32_t*>(Ks + swz(r, c) + w * 4));
      m0 = fmaxf(m0, v.x);
      m1 = fmaxf(m1, v.y);
    }
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < T; ++r) {
      uint32_t* p = reinterpret_cast<uint32_t*>(Ks + swz(r, c) + w * 4);
      const float2 v = unpack2(*p);
      // round first so that the normaliser is the sum of exactly what the tensor core multiplies
      const __nv_bfloat162 e = __floats2bfloat162_rn(__expf(v.x - m0), __expf(v.y - m1));
      s0 += __bfloat162float(e.x);
      s1 += __bfloat162float(e.y);
      *p = *reinterpret_cast<const uint32_t*>(&e);
    }
    inv[2 * lane] = 1.f / s0;
    inv[2 * lane + 1] = 1.f / s1;
    for (int i = lane; i < (n_mt * 16 - T) * 8; i += 32) {
      const int r = T + (i >> 3), cc = i & 7;
      *reinterpret_cast<uint4*>(Ks + swz(r, cc)) = make_uint4(0, 0, 0, 0);
    }
  }
  __syncwarp();

  // ---- 4. A^T[l][d] = sum_t V[t][l] K'[t][d]; its accumulators become the B fragments bfr[kd][nl] of step 5
  uint32_t bfr[4][8][2];
  {
    const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {  // 16 l's per m-tile
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      for (int kt = 0; kt < n_mt; ++kt) {  // 16 frames per k-step
        uint32_t a[4];
        {
          const int r = kt * 16 + rr + ((mat >> 1) << 3), c = 2 * mt + (mat & 1);
          ldsm_x4_trans(vs_addr + swz(r, c), a[0], a[1], a[2], a[3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // two d n-tiles per ldmatrix.x4
          uint32_t b0, b1, b2, b3;
          const int r = kt * 16 + rr + ((mat & 1) << 3), c = 2 * np + (mat >> 1);
          ldsm_x4_trans(ks_addr + swz(r, c), b0, b1, b2, b3);
          mma_bf16(acc[2 * np], a, b0, b1);
          mma_bf16(acc[2 * np + 1], a, b2, b3);
        }
      }
      // normalise column d by 1/sum_t exp(K) and permute C^T fragments into B fragments (see header)
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
        const float i00 = inv[16 * kd + 2 * q], i01 = inv[16 * kd + 2 * q + 1];
        const float i10 = inv[16 * kd + 8 + 2 * q], i11 = inv[16 * kd + 8 + 2 * q + 1];
        bfr[kd][2 * mt][0] = pack2(acc[2 * kd][0] * i00, acc[2 * kd][1] * i01);
        bfr[kd][2 * mt][1] = pack2(acc[2 * kd + 1][0] * i10, acc[2 * kd + 1][1] * i11);
        bfr[kd][2 * mt + 1][0] = pack2(acc[2 * kd][2] * i00, acc[2 * kd][3] * i01);
        bfr[kd][2 * mt + 1][1] = pack2(acc[2 * kd + 1][2] * i10, acc[2 * kd + 1][3] * i11);
      }
    }
  }
  __syncwarp();  // every lane is done reading V: its tile now receives Y

  // ---- 5. Y[t][l] = softmax_d(Q)[t][:] . A ; bf16 Y -> the V tile
  for (int mt = 0; mt < n_mt; ++mt) {
    // row softmax numerators in registers: this lane holds 16 of the 64 d's of rows g and g+8; the quad holds all
    float f[4][8];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float2 a0 = unpack2(qa[ks][0]), a1 = unpack2(qa[ks][1]), a2 = unpack2(qa[ks][2]), a3 = unpack2(qa[ks][3]);
      f[ks][0] = a0.x; f[ks][1] = a0.y; f[ks][2] = a2.x; f[ks][3] = a2.y;   // row g
      f[ks][4] = a1.x; f[ks][5] = a1.y; f[ks][6] = a3.x; f[ks][7] = a3.y;   // row g + 8
      mx0 = fmaxf(mx0, fmaxf(fmaxf(a0.x, a0.y), fmaxf(a2.x, a2.y)));
      mx1 = fmaxf(mx1, fmaxf(fmaxf(a1.x, a1.y), fmaxf(a3.x, a3.y)));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sm0 = 0.f, sm1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const __nv_bfloat162 e0 = __floats2bfloat162_rn(__expf(f[ks][0] - mx0), __expf(f[ks][1] - mx0));
      const __nv_bfloat162 e2 = __floats2bfloat162_rn(__expf(f[ks][2] - mx0), __expf(f[ks][3] - mx0));
      const __nv_bfloat162 e1 = __floats2bfloat162_rn(__expf(f[ks][4] - mx1), __expf(f[ks][5] - mx1));
      const __nv_bfloat162 e3 = __floats2bfloat162_rn(__expf(f[ks][6] - mx1), __expf(f[ks][7] - mx1));
      sm0 += __bfloat162float(e0.x) + __bfloat162float(e0.y) + __bfloat162float(e2.x) + __bfloat162float(e2.y);
      sm1 += __bfloat162float(e1.x) + __bfloat162float(e1.y) + __bfloat162float(e3.x) + __bfloat162float(e3.y);
      pa[ks][0] = *reinterpret_cast<const uint32_t*>(&e0);
      pa[ks][1] = *reinterpret_cast<const uint32_t*>(&e1);
      pa[ks][2] = *reinterpret_cast<const uint32_t*>(&e2);
      pa[ks][3] = *reinterpret_cast<const uint32_t*>(&e3);
    }
    sm0 += __shfl_xor_sync(0xffffffffu, sm0, 1); sm0 += __shfl_xor_sync(0xffffffffu, sm0, 2);
    sm1 += __shfl_xor_sync(0xffffffffu, sm1, 1); sm1 += __shfl_xor_sync(0xffffffffu, sm1, 2);
    if (mt + 1 < n_mt) load_q_tile(qhead, (mt + 1) * 16, T, g, q, qa);  // prefetch under the MMAs
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_bf16(acc[nt], pa[ks], bfr[ks][nt][0], bfr[ks][nt][1]);
    const float r0 = 1.f / sm0, r1 = 1.f / sm1;
    const int ra = mt * 16 + g, rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(Vs + swz(ra, nt) + q * 4) = pack2(acc[nt][0] * r0, acc[nt][1] * r0);
      *reinterpret_cast<uint32_t*>(Vs + swz(rb, nt) + q * 4) = pack2(acc[nt][2] * r1, acc[nt][3] * r1);
    }
  }
  __syncthreads();  // all 8 heads of the sample are in smem

  // ---- 6. StylizationBlock prologue over full rows: LN(512) * (1 + scale) + shift, SiLU; one warp per row
  {
    const int hh = lane >> 2, c0 = (lane & 3) * 2;  // this lane covers head hh, 16-B chunks c0 and c0+1
    const uint8_t* Yh = sm + hh * 2 * TILE_BYTES + TILE_BYTES;
    const int col0 = hh * HD + c0 * 8;
    const float* sc = ss + (size_t)(smp % B) * ss_ld;
    float gg[16], bb[16], s1[16], s2[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      gg[e] = __ldg(ln_g + col0 + e); bb[e] = __ldg(ln_b + col0 + e);
      s1[e] = 1.f + __ldg(sc + col0 + e); s2[e] = __ldg(sc + D + col0 + e);
    }
    for (int t = warp; t < T; t += NH) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(Yh + swz(t, c0));
      const uint4 u1 = *reinterpret_cast<const uint4*>(Yh + swz(t, c0 + 1));
      const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      float v[16];
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float2 p2 = unpack2(w[e]); v[2 * e] = p2.x; v[2 * e + 1] = p2.y; s += p2.x + p2.y; }
      const float mean = warp_sum(s) * (1.f / D);
      float var = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) { const float dlt = v[e] - mean; var += dlt * dlt; }
      const float rstd = rsqrtf(warp_sum(var) * (1.f / D) + 1e-5f);
      uint32_t o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a0 = ((v[2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e]) * s1[2 * e] + s2[2 * e];
        const float a1 = ((v[2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1]) * s1[2 * e + 1] + s2[2 * e + 1];
        o[e] = pack2(silu_f(a0), silu_f(a1));
      }
      uint4* dst = reinterpret_cast<uint4*>(z + (row0 + t) * (size_t)D + col0);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
}

}  // namespace av2
}  // namespace dsheg
