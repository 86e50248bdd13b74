// Linear-attention core + StylizationBlock prologue for FP32 activations (the tf32 precision mode): D = 512-class widths with heads of 64.
//
// Same mathematics as the generic kernel (kernels.cuh attn_kernel; reference transformer.py:112-130 + :86-97), same structure -- one
// CTA per sample, heads one after the other, fp32 tiles in shared memory, Y through an fp32 row scratch, LayerNorm / modulate / SiLU
// over the full rows at the end -- but the two products of a head run on the tensor cores with mma.sync.m16n8k8 TF32 (fp32
// accumulation): A = K'^T V (64 x 64, contraction over the frames) and Y = Q' A (T x 64, contraction over 64).  The softmaxes, their
// sums and normalisations, and the LayerNorm are exact fp32; only the four product operands are rounded to TF32 (round-to-nearest,
// once, when they are written to shared memory) -- the precision contract of the tf32 mode (gemm_tf32.cuh).
// The SIMT kernel spends 2 shared-memory loads per FMA: 3.3 ms per layer launch at the headline batch; this one: 12 loads per 4 MMAs.
#pragma once
#include "kernels.cuh"

namespace dsheg {
namespace at32 {

constexpr int HD = 64;
constexpr int LDK = 72;   // row stride (floats) of the K', V and A tiles: fragment loads index [k][m] -> banks 8 q + g: conflict-free
constexpr int LDQ = 68;   // row stride of the Q' tile: fragment loads index [m][k] -> banks 4 g + q: conflict-free
constexpr int NTHREADS = 256;

inline size_t smem_bytes(int T) {
  const int Tp = (T + 15) / 16 * 16;
  return (size_t)(Tp * LDQ + 2 * Tp * LDK + HD * LDK + 4 * HD) * sizeof(float);
}

#ifdef DSHEG_EMU
// tests/emu models of the two PTX instructions (test infrastructure; the product build takes the #else branch):
//   cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits kept
//   mma.m16n8k8.row.col tf32 (g = lane / 4, q = lane % 4): A regs {(g, q), (g + 8, q), (g, q + 4), (g + 8, q + 4)},
//   B regs {(k = q, n = g), (k = q + 4, n = g)}, C/D {(g, 2q), (g, 2q + 1), (g + 8, 2q), (g + 8, 2q + 1)}; fp32 accumulation
inline float to_tf32(float x) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
  return __uint_as_float(u);
}
inline void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  struct Frag { uint32_t a[4], b[2]; } mine{{a[0], a[1], a[2], a[3]}, {b0, b1}};
  const int lane = emu::self().lane, g = lane >> 2, q = lane & 3;
  uint8_t(*slots)[64] = emu::warp_exchange(&mine, sizeof(Frag), "mma.sync tf32");
  auto frag = [&](int l) { Frag f; memcpy(&f, slots[l], sizeof(Frag)); return f; };
  for (int e = 0; e < 4; ++e) {
    const int row = g + 8 * (e >> 1), n = 2 * q + (e & 1);
    float acc = c[e];
    for (int k = 0; k < 8; ++k) {
      const Frag fa = frag(row % 8 * 4 + (k & 3)), fb = frag(n * 4 + (k & 3));
      const float av = __uint_as_float(fa.a[(row >> 3) + 2 * (k >> 2)] & 0xffffe000u), bv = __uint_as_float(fb.b[k >> 2] & 0xffffe000u);
      acc += av * bv;
    }
    c[e] = acc;
  }
}
#else
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#endif

// qkv: [n_samples * T, 3 D] fp32 (q | k | v), y32: [n_samples * T, D] fp32 scratch, z: [n_samples * T, D] fp32.
__global__ void __launch_bounds__(NTHREADS) attn_tf32_kernel(const float* __restrict__ qkv, float* __restrict__ y32, float* __restrict__ z, int T, int D,
                                                             int H, int B, const float* __restrict__ g, const float* __restrict__ b,
                                                             const float* __restrict__ ss, int ss_ld) {
#ifdef DSHEG_EMU
  float* sm = reinterpret_cast<float*>(emu::self().cta->smem);
#else
  extern __shared__ float sm[];
#endif
  const int Tp = (T + 15) / 16 * 16;          // frames padded to whole m-tiles (rows T .. Tp-1 are zero)
  float* Qs = sm;                              // [Tp][LDQ]
  float* Ks = Qs + Tp * LDQ;                   // [Tp][LDK]
  float* Vs = Ks + Tp * LDK;                   // [Tp][LDK]
  float* As = Vs + Tp * LDK;                   // [64][LDK]
  float* red = As + HD * LDK;                  // [4][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, q = lane & 3;      // mma fragment coordinates
  const int smp = blockIdx.x;
  const size_t row0 = (size_t)smp * T;
  const int col = tid & 63, part = tid >> 6;   // column softmax: 4 row partitions x 64 columns

  for (int h = 0; h < H; ++h) {
    // ---- stage the head's Q, K, V tiles (16-byte loads; V already rounded to TF32; padding rows zero)
    for (int e = tid; e < Tp * (HD / 4); e += NTHREADS) {
      const int t = e / (HD / 4), c4 = (e % (HD / 4)) * 4;
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv;
      if (t < T) {
        const float* p = qkv + (row0 + t) * (size_t)(3 * D) + h * HD + c4;
        qv = __ldg(reinterpret_cast<const float4*>(p));
        kv = __ldg(reinterpret_cast<const float4*>(p + D));
        vv = __ldg(reinterpret_cast<const float4*>(p + 2 * D));
      }
      *reinterpret_cast<float4*>(Qs + t * LDQ + c4) = qv;
      *reinterpret_cast<float4*>(Ks + t * LDK + c4) = kv;
      *reinterpret_cast<float4*>(Vs + t * LDK + c4) = make_float4(to_tf32(vv.x), to_tf32(vv.y), to_tf32(vv.z), to_tf32(vv.w));
    }
    __syncthreads();
    // ---- softmax over time for every K column (tr:123, dim=1), exact fp32; the normalised weights are stored as TF32
    {
      float m = -INFINITY;
      for (int t = part; t < T; t += 4) m = fmaxf(m, Ks[t * LDK + col]);
      red[part * HD + col] = m;
      __syncthreads();
      m = fmaxf(fmaxf(red[col], red[HD + col]), fmaxf(red[2 * HD + col], red[3 * HD + col]));
      __syncthreads();
      float s = 0.f;
      for (int t = part; t < T; t += 4) {
        const float e = __expf(Ks[t * LDK + col] - m);
        Ks[t * LDK + col] = e;
        s += e;
      }
      red[part * HD + col] = s;
      __syncthreads();
      const float inv = 1.f / (red[col] + red[HD + col] + red[2 * HD + col] + red[3 * HD + col]);
      for (int t = part; t < T; t += 4) Ks[t * LDK + col] = to_tf32(Ks[t * LDK + col] * inv);
    }
    // ---- softmax over the head dim for every Q row (tr:122, dim=-1): one warp per row
    for (int t = warp; t < T; t += NTHREADS / 32) {
      const float v0 = Qs[t * LDQ + lane], v1 = Qs[t * LDQ + lane + 32];
      const float mx = warp_max(fmaxf(v0, v1));
      const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx);
      const float inv = 1.f / warp_sum(e0 + e1);
      Qs[t * LDQ + lane] = to_tf32(e0 * inv);
      Qs[t * LDQ + lane + 32] = to_tf32(e1 * inv);
    }
    __syncthreads();
    // ---- A[d][l] = sum_t K'[t][d] V[t][l] on the tensor cores: warp -> d block (16 rows) x l half (32 columns = 4 n-tiles)
    {
      const int d0 = (warp & 3) * 16, l0 = (warp >> 2) * 32;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      for (int t0 = 0; t0 < Tp; t0 += 8) {
        uint32_t a[4];   // A operand = K'^T: a(m = d, k = t) = Ks[t][d]
        a[0] = __float_as_uint(Ks[(t0 + q) * LDK + d0 + gq]);
        a[1] = __float_as_uint(Ks[(t0 + q) * LDK + d0 + gq + 8]);
        a[2] = __float_as_uint(Ks[(t0 + q + 4) * LDK + d0 + gq]);
        a[3] = __float_as_uint(Ks[(t0 + q + 4) * LDK + d0 + gq + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {   // B operand (k = t, n = l) = Vs[t][l]
          const uint32_t b0 = __float_as_uint(Vs[(t0 + q) * LDK + l0 + 8 * nt + gq]);
          const uint32_t b1 = __float_as_uint(Vs[(t0 + q + 4) * LDK + l0 + 8 * nt + gq]);
          mma_tf32(acc[nt], a, b0, b1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {   // C fragment: rows d0 + gq (+ 8), columns l0 + 8 nt + 2 q (+ 1)
        *reinterpret_cast<float2*>(As + (d0 + gq) * LDK + l0 + 8 * nt + 2 * q) = make_float2(to_tf32(acc[nt][0]), to_tf32(acc[nt][1]));
        *reinterpret_cast<float2*>(As + (d0 + gq + 8) * LDK + l0 + 8 * nt + 2 * q) = make_float2(to_tf32(acc[nt][2]), to_tf32(acc[nt][3]));
      }
    }
    __syncthreads();
    // ---- Y[t][l] = sum_d Q'[t][d] A[d][l]: work items (m-tile of 16 frames, l half), round-robin over the 8 warps
    for (int item = warp; item < (Tp / 16) * 2; item += NTHREADS / 32) {
      const int t0 = (item >> 1) * 16, l0 = (item & 1) * 32;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int k0 = 0; k0 < HD; k0 += 8) {
        uint32_t a[4];   // A operand (m = t, k = d) = Qs[t][d]
        a[0] = __float_as_uint(Qs[(t0 + gq) * LDQ + k0 + q]);
        a[1] = __float_as_uint(Qs[(t0 + gq + 8) * LDQ + k0 + q]);
        a[2] = __float_as_uint(Qs[(t0 + gq) * LDQ + k0 + q + 4]);
        a[3] = __float_as_uint(Qs[(t0 + gq + 8) * LDQ + k0 + q + 4]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {   // B operand (k = d, n = l) = As[d][l]
          const uint32_t b0 = __float_as_uint(As[(k0 + q) * LDK + l0 + 8 * nt + gq]);
          const uint32_t b1 = __float_as_uint(As[(k0 + q + 4) * LDK + l0 + 8 * nt + gq]);
          mma_tf32(acc[nt], a, b0, b1);
        }
      }
      const int ta = t0 + gq, tb = ta + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = h * HD + l0 + 8 * nt + 2 * q;
        if (ta < T) *reinterpret_cast<float2*>(y32 + (row0 + ta) * (size_t)D + c) = make_float2(acc[nt][0], acc[nt][1]);
        if (tb < T) *reinterpret_cast<float2*>(y32 + (row0 + tb) * (size_t)D + c) = make_float2(acc[nt][2], acc[nt][3]);
      }
    }
    __syncthreads();
  }
  // ---- StylizationBlock prologue over the full rows (all heads done; this CTA's own y32 rows: L2-resident)
  const float* sc = ss + (size_t)(smp % B) * ss_ld;
  for (int t = warp; t < T; t += NTHREADS / 32)
    ln_mod_silu_row<float, float>(y32 + (row0 + t) * (size_t)D, z + (row0 + t) * (size_t)D, D, g, b, sc, sc + D, lane);
}

}  // namespace at32
}  // namespace dsheg
