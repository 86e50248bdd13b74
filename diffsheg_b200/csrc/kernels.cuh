// Memory-bound kernels of the denoiser: row statistics, LayerNorm-modulate-SiLU, the linear
// attention core, embedding MLPs, the hubert conv stack and the small glue around the GEMMs.
// Every kernel is templated on the activation type TA (float in fp32 mode, bf16 in bf16 mode);
// statistics, softmaxes and accumulations are always fp32.
#pragma once
#include "common.cuh"

namespace dsheg {

constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default (reference transformer.py:79,105,285)

// ---------------------------------------------------------------------------------------------
// feat_prep: one warp per hidden row.
//   uncond rows (row < n_uncond):  h[row] += nullc        (feat_proj(null_cond_emb) is a per-layer
//                                   constant: transformer.py:326-338)
//   cond rows:                     mu/rstd over the virtual concat (h | xf | hubert [| expr]) for
//                                   the LayerNorm(P) folded into the feat_proj GEMM (tr:284-289)
// Segment 0 is h itself (indexed by the absolute row); segments 1.. are indexed by row - n_uncond.
// ---------------------------------------------------------------------------------------------
template <typename TA>
__global__ void feat_prep_kernel(TA* h, int ldh, int D, int n_uncond, int n_rows, const float* nullc,
                                 Seg s1, Seg s2, Seg s3, int nseg_extra, float* mu, float* rstd) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  TA* hr = h + (size_t)row * ldh;
  if (row < n_uncond) {
    for (int c = lane; c < D; c += 32) AT<TA>::st(hr + c, AT<TA>::ld(hr + c) + nullc[c]);
    return;
  }
  const int cr = row - n_uncond;
  const Seg segs[3] = {s1, s2, s3};
  float sum = 0.f;
  int P = D;
  for (int c = lane; c < D; c += 32) sum += AT<TA>::ld(hr + c);
  for (int s = 0; s < nseg_extra; ++s) {
    const TA* p = reinterpret_cast<const TA*>(segs[s].ptr) + (size_t)cr * segs[s].ld;
    for (int c = lane; c < segs[s].k; c += 32) sum += AT<TA>::ld(p + c);
    P += segs[s].k;
  }
  const float mean = warp_sum(sum) / (float)P;
  float var = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = AT<TA>::ld(hr + c) - mean; var += v * v; }
  for (int s = 0; s < nseg_extra; ++s) {
    const TA* p = reinterpret_cast<const TA*>(segs[s].ptr) + (size_t)cr * segs[s].ld;
    for (int c = lane; c < segs[s].k; c += 32) { const float v = AT<TA>::ld(p + c) - mean; var += v * v; }
  }
  var = warp_sum(var) / (float)P;
  if (lane == 0) { mu[cr] = mean; rstd[cr] = rsqrtf(var + LN_EPS); }
}

// Row statistics of a plain [n_rows, D] matrix (LayerNorm folded into the QKV GEMM, tr:119-125).
template <typename TA>
__global__ void rowstats_kernel(const TA* x, int ld, int D, int n_rows, float* mu, float* rstd) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  const TA* xr = x + (size_t)row * ld;
  float sum = 0.f;
  for (int c = lane; c < D; c += 32) sum += AT<TA>::ld(xr + c);
  const float mean = warp_sum(sum) / (float)D;
  float var = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = AT<TA>::ld(xr + c) - mean; var += v * v; }
  var = warp_sum(var) / (float)D;
  if (lane == 0) { mu[row] = mean; rstd[row] = rsqrtf(var + LN_EPS); }
}

// ---------------------------------------------------------------------------------------------
// StylizationBlock prologue (tr:92-96): z = SiLU( LN(y; g, b) * (1 + scale[s]) + shift[s] ),
// sample s = (row / T) % B; scale = ss[s*ss_ld + 0..D), shift = ss[s*ss_ld + D..2D).
// One warp per row; D <= 32*LMS_MAXV.
// ---------------------------------------------------------------------------------------------
constexpr int LMS_MAXV = 16;

template <typename TIN, typename TA>
__device__ __forceinline__ void ln_mod_silu_row(const TIN* yr, TA* zr, int D, const float* g, const float* b,
                                                const float* scale, const float* shift, int lane) {
  float v[LMS_MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LMS_MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < D) ? AT<TIN>::ld(yr + c) : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < LMS_MAXV; ++i) {
    const int c = lane + 32 * i;
    const float dlt = (c < D) ? v[i] - mean : 0.f;
    var += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)D + LN_EPS);
#pragma unroll
  for (int i = 0; i < LMS_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      float t = (v[i] - mean) * rstd * g[c] + b[c];
      t = t * (1.f + scale[c]) + shift[c];
      AT<TA>::st(zr + c, silu_f(t));
    }
  }
}

template <typename TIN, typename TA>
__global__ void ln_mod_silu_kernel(const TIN* y, int ldy, TA* z, int ldz, int D, int n_rows, int T, int B,
                                   const float* g, const float* b, const float* ss, int ss_ld) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  const int s = (row / T) % B;
  const float* sc = ss + (size_t)s * ss_ld;
  ln_mod_silu_row<TIN, TA>(y + (size_t)row * ldy, z + (size_t)row * ldz, D, g, b, sc, sc + D, lane);
}

// ---------------------------------------------------------------------------------------------
// bf16 fast paths of the row-wise kernels: one warp per row, 16-byte (8 x bf16) accesses, the row is read
// ONCE and kept in registers for the second (variance / normalise) pass.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __bfloat162float(h2.x);
    f[2 * e + 1] = __bfloat162float(h2.y);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// feat_prep for bf16 (same contract as feat_prep_kernel).  D % 8 == 0, D <= 512, extra segments <= 256 wide,
// every segment row 16-byte aligned.
__global__ void feat_prep_bf16_kernel(bf16* h, int ldh, int D, int n_uncond, int n_rows, const float* nullc, Seg s1, Seg s2,
                                      Seg s3, int nseg_extra, float* mu, float* rstd) {
  DSHEG_PDL_ENTER();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  bf16* hr = h + (size_t)row * ldh;
  if (row < n_uncond) {
    for (int c = lane; c < D / 8; c += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(hr + c * 8), f);
      const float4 n0 = __ldg(reinterpret_cast<const float4*>(nullc + c * 8)), n1 = __ldg(reinterpret_cast<const float4*>(nullc + c * 8 + 4));
      f[0] += n0.x; f[1] += n0.y; f[2] += n0.z; f[3] += n0.w; f[4] += n1.x; f[5] += n1.y; f[6] += n1.z; f[7] += n1.w;
      *reinterpret_cast<uint4*>(hr + c * 8) = pack8(f);
    }
    return;
  }
  const int cr = row - n_uncond;
  // h: up to 2 chunks per lane (D <= 512); each extra segment: 1 chunk per lane (k <= 256)
  float vh[2][8], vx[3][8];
  float sum = 0.f;
  int P = D;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + 32 * i) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) vh[i][e] = 0.f;
    if (c < D) {
      unpack8(*reinterpret_cast<const uint4*>(hr + c), vh[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += vh[i][e];
    }
  }
  const Seg segs[3] = {s1, s2, s3};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
#pragma unroll
    for (int e = 0; e < 8; ++e) vx[s][e] = 0.f;
    if (s < nseg_extra) {
      const int c = lane * 8, k = segs[s].k;
      P += k;
      if (c < k) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(segs[s].ptr) + (size_t)cr * segs[s].ld + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { vx[s][e] = (c + e < k) ? f[e] : 0.f; sum += vx[s][e]; }
      }
    }
  }
  const float mean = warp_sum(sum) / (float)P;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if ((lane + 32 * i) * 8 < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float dlt = vh[i][e] - mean; var += dlt * dlt; }
    }
  }
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < nseg_extra) {
      const int c = lane * 8, k = segs[s].k;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float dlt = (c + e < k) ? vx[s][e] - mean : 0.f; var += dlt * dlt; }
    }
  }
  var = warp_sum(var) / (float)P;
  if (lane == 0) { mu[cr] = mean; rstd[cr] = rsqrtf(var + LN_EPS); }
}

// (sum, sum of squares) of the conditioning part (xf | hubert [| expr]) of the feat_proj input row: step-wide
// constant across the 8 layers of a net, combined with the residual-stream partials in the feat1 epilogue.
__global__ void cond_stats_bf16_kernel(Seg s1, Seg s2, Seg s3, int nseg, int n_rows, float2* cs) {
  DSHEG_PDL_ENTER();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  const Seg segs[3] = {s1, s2, s3};
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < nseg) {
      const int c = lane * 8, k = segs[s].k;
      if (c < k) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(segs[s].ptr) + (size_t)row * segs[s].ld + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) if (c + e < k) { sum += f[e]; sq = fmaf(f[e], f[e], sq); }
      }
    }
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if (lane == 0) cs[row] = make_float2(sum, sq);
}

// rowstats for bf16, D % 8 == 0, D <= 1024
__global__ void rowstats_bf16_kernel(const bf16* x, int ld, int D, int n_rows, float* mu, float* rstd) {
  DSHEG_PDL_ENTER();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  const bf16* xr = x + (size_t)row * ld;
  float v[4][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    if (c < D) {
      unpack8(*reinterpret_cast<const uint4*>(xr + c), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if ((lane + 32 * i) * 8 < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float dlt = v[i][e] - mean; var += dlt * dlt; }
    }
  }
  var = warp_sum(var) / (float)D;
  if (lane == 0) { mu[row] = mean; rstd[row] = rsqrtf(var + LN_EPS); }
}

// ln_mod_silu for bf16 in / bf16 out, D % 8 == 0, D <= 512
__global__ void ln_mod_silu_bf16_kernel(const bf16* y, int ldy, bf16* z, int ldz, int D, int n_rows, int T, int B, const float* g,
                                        const float* b, const float* ss, int ss_ld) {
  DSHEG_PDL_ENTER();
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n_rows) return;
  const bf16* yr = y + (size_t)row * ldy;
  bf16* zr = z + (size_t)row * ldz;
  const float* sc = ss + (size_t)((row / T) % B) * ss_ld;
  float v[2][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + 32 * i) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    if (c < D) {
      unpack8(*reinterpret_cast<const uint4*>(yr + c), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if ((lane + 32 * i) * 8 < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float dlt = v[i][e] - mean; var += dlt * dlt; }
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)D + LN_EPS);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < D) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c)), g1 = __ldg(reinterpret_cast<const float4*>(g + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(sc + c)), p1 = __ldg(reinterpret_cast<const float4*>(sc + c + 4));
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(sc + D + c)), q1 = __ldg(reinterpret_cast<const float4*>(sc + D + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float s1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      const float s2[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = ((v[i][e] - mean) * rstd * gg[e] + bb[e]) * (1.f + s1[e]) + s2[e];
        o[e] = silu_f(t);
      }
      *reinterpret_cast<uint4*>(zr + c) = pack8(o);
    }
  }
}

// ln_mod_silu, D == 512, bf16: one CTA (8 warps) per SAMPLE so that the per-column coefficients
//   t = ((v-mean)*rstd*g + b)*(1+scale) + shift = (v-mean)*rstd*G + Bc
// are folded once per lane and reused for the sample's T rows (the per-row variant spends 8x more load
// instructions on g/b/scale/shift than on data).  SiLU(x) = h + h*tanh(h), h = x/2 (MUFU.TANH).
__global__ void __launch_bounds__(256) ln_mod_silu_sample_bf16_kernel(const bf16* y, bf16* z, int T, int B, const float* g,
                                                                      const float* b, const float* ss, int ss_ld) {
  DSHEG_PDL_ENTER();
  constexpr int D = 512;
  const int smp = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const float* sc = ss + (size_t)(smp % B) * ss_ld;
  float G[16], Bc[16];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + 32 * i) * 8;
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g + c + e)), b4 = __ldg(reinterpret_cast<const float4*>(b + c + e));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc + c + e)), d4 = __ldg(reinterpret_cast<const float4*>(sc + D + c + e));
      G[i * 8 + e] = a.x * (1.f + c4.x); G[i * 8 + e + 1] = a.y * (1.f + c4.y);
      G[i * 8 + e + 2] = a.z * (1.f + c4.z); G[i * 8 + e + 3] = a.w * (1.f + c4.w);
      Bc[i * 8 + e] = fmaf(b4.x, 1.f + c4.x, d4.x); Bc[i * 8 + e + 1] = fmaf(b4.y, 1.f + c4.y, d4.y);
      Bc[i * 8 + e + 2] = fmaf(b4.z, 1.f + c4.z, d4.z); Bc[i * 8 + e + 3] = fmaf(b4.w, 1.f + c4.w, d4.w);
    }
  }
  const size_t row0 = (size_t)smp * T;
  // two rows per iteration: 4 independent 16-byte loads per lane in flight (the kernel is latency / MLP bound otherwise)
  for (int t = warp; t < T; t += 16) {
    const int t2 = t + 8;
    const bool has2 = t2 < T;
    const bf16* yr = y + (row0 + t) * D;
    const bf16* yr2 = y + (row0 + (has2 ? t2 : t)) * D;
    const uint4 a0 = *reinterpret_cast<const uint4*>(yr + lane * 8), a1 = *reinterpret_cast<const uint4*>(yr + 256 + lane * 8);
    const uint4 b0 = *reinterpret_cast<const uint4*>(yr2 + lane * 8), b1 = *reinterpret_cast<const uint4*>(yr2 + 256 + lane * 8);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr == 1 && !has2) break;
      float v[16];
      unpack8(rr == 0 ? a0 : b0, *reinterpret_cast<float(*)[8]>(&v[0]));
      unpack8(rr == 0 ? a1 : b1, *reinterpret_cast<float(*)[8]>(&v[8]));
      float s = 0.f, sq = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) { s += v[e]; sq = fmaf(v[e], v[e], sq); }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o2); sq += __shfl_xor_sync(0xffffffffu, sq, o2); }
      const float mean = s * (1.f / D);
      // FFN output rows are O(1) with |mean| << std: E[x^2] - mean^2 is safe in fp32 (clamped at 0)
      const float rstd = rsqrtf(fmaxf(sq * (1.f / D) - mean * mean, 0.f) + LN_EPS);
      const float nmr = -mean * rstd;
      float o[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float hh = 0.5f * fmaf(fmaf(v[e], rstd, nmr), G[e], Bc[e]);
        float th;
#ifdef DSHEG_EMU
        th = tanhf(hh);
#else
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hh));
#endif
        o[e] = fmaf(hh, th, hh);
      }
      bf16* zr = z + (row0 + (rr == 0 ? t : t2)) * D;
      *reinterpret_cast<uint4*>(zr + lane * 8) = pack8(*reinterpret_cast<float(*)[8]>(&o[0]));
      *reinterpret_cast<uint4*>(zr + 256 + lane * 8) = pack8(*reinterpret_cast<float(*)[8]>(&o[8]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Linear ("efficient") self-attention core (tr:122-128), one CTA per sample, heads in sequence:
//   Q' = softmax_d(Q)   K' = softmax_t(K)   A = K'^T V  [HD x HD]   Y = Q' A
// followed, in the same CTA, by the StylizationBlock prologue over the full D-wide row (needs all
// heads), so the attention output never leaves the chip before LN/modulate/SiLU ... except for the
// fp32 row scratch y32 (L2-resident: written and re-read by the same CTA).
// qkv: [n_samples*T, 3D] (q | k | v), z: [n_samples*T, D].
// ---------------------------------------------------------------------------------------------
template <typename TA, int HD>
__global__ void __launch_bounds__(256) attn_kernel(const TA* qkv, float* y32, TA* z, int T, int D, int H, int B,
                                                   const float* g, const float* b, const float* ss, int ss_ld) {
#ifdef DSHEG_EMU
  float* sm = reinterpret_cast<float*>(emu::self().cta->smem);
#else
  extern __shared__ float sm[];
#endif
  constexpr int LD = HD + 1;
  constexpr int NPART = 256 / HD;
  float* Qs = sm;
  float* Ks = Qs + T * LD;
  float* Vs = Ks + T * LD;
  float* As = Vs + T * LD;          // [HD][LD]
  float* red = As + HD * LD;        // [NPART][HD]
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const int smp = blockIdx.x;
  const size_t row0 = (size_t)smp * T;
  const int col = tid % HD, part = tid / HD;

  for (int h = 0; h < H; ++h) {
    // ---- stage the head's Q, K, V tiles (fp32 in smem)
    for (int e = tid; e < T * HD; e += 256) {
      const int t = e / HD, d = e % HD;
      const TA* p = qkv + (row0 + t) * (size_t)(3 * D) + h * HD + d;
      Qs[t * LD + d] = AT<TA>::ld(p);
      Ks[t * LD + d] = AT<TA>::ld(p + D);
      Vs[t * LD + d] = AT<TA>::ld(p + 2 * D);
    }
    __syncthreads();
    // ---- softmax over time for every K column (tr:123, dim=1)
    float m = -INFINITY;
    for (int t = part; t < T; t += NPART) m = fmaxf(m, Ks[t * LD + col]);
    red[part * HD + col] = m;
    __syncthreads();
    m = red[col];
#pragma unroll
    for (int p2 = 1; p2 < NPART; ++p2) m = fmaxf(m, red[p2 * HD + col]);
    __syncthreads();
    float s = 0.f;
    for (int t = part; t < T; t += NPART) {
      const float e = __expf(Ks[t * LD + col] - m);
      Ks[t * LD + col] = e;
      s += e;
    }
    red[part * HD + col] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < NPART; ++p2) s += red[p2 * HD + col];
    const float inv = 1.f / s;
    for (int t = part; t < T; t += NPART) Ks[t * LD + col] *= inv;
    // ---- softmax over the head dim for every Q row (tr:122, dim=-1): one warp per row
    for (int t = warp; t < T; t += 8) {
      float mx = -INFINITY;
      for (int d = lane; d < HD; d += 32) mx = fmaxf(mx, Qs[t * LD + d]);
      mx = warp_max(mx);
      float sq = 0.f;
      for (int d = lane; d < HD; d += 32) {
        const float e = __expf(Qs[t * LD + d] - mx);
        Qs[t * LD + d] = e;
        sq += e;
      }
      sq = 1.f / warp_sum(sq);
      for (int d = lane; d < HD; d += 32) Qs[t * LD + d] *= sq;
    }
    __syncthreads();
    // ---- A[d][l] = sum_t K'[t][d] V[t][l]
    for (int d = part; d < HD; d += NPART) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc = fmaf(Ks[t * LD + d], Vs[t * LD + col], acc);
      As[d * LD + col] = acc;
    }
    __syncthreads();
    // ---- Y[t][l] = sum_d Q'[t][d] A[d][l]
    for (int t = part; t < T; t += NPART) {
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) acc = fmaf(Qs[t * LD + d], As[d * LD + col], acc);
      y32[(row0 + t) * (size_t)D + h * HD + col] = acc;
    }
    __syncthreads();
  }
  // ---- StylizationBlock prologue over the full rows (all heads done)
  const int sidx = smp % B;
  const float* sc = ss + (size_t)sidx * ss_ld;
  for (int t = warp; t < T; t += 8)
    ln_mod_silu_row<float, TA>(y32 + (row0 + t) * (size_t)D, z + (row0 + t) * (size_t)D, D, g, b, sc, sc + D, lane);
}

template <int HD>
inline size_t attn_smem_bytes(int T) { return (size_t)(3 * T * (HD + 1) + HD * (HD + 1) + 256) * sizeof(float); }

// ---------------------------------------------------------------------------------------------
// Glue
// ---------------------------------------------------------------------------------------------
// xin[r, 0..ld) = cast(x[r*Dtot + off + j]) for j < feats, 0 for the K padding.
template <typename TA>
__global__ void cast_pad_kernel(const float* x, int Dtot, int off, int feats, TA* xin, int ld, int n_rows) {
  DSHEG_PDL_ENTER();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_rows * ld) return;
  const int r = (int)(i / ld), j = (int)(i % ld);
  AT<TA>::st(xin + i, j < feats ? x[(size_t)r * Dtot + off + j] : 0.f);
}

// mel staging per window: aud256[:, 0:A] = mel ; a0 = 2*mel (the cond_residual doubling of the
// xf=None audio layer, transformer.py:302-303,337-338).
template <typename TA>
__global__ void mel_stage_kernel(const float* mel, int A, TA* aud256, int ld256, TA* a0, int n_rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_rows * A) return;
  const int r = (int)(i / A), j = (int)(i % A);
  const float v = mel[i];
  AT<TA>::st(aud256 + (size_t)r * ld256 + j, v);
  AT<TA>::st(a0 + i, 2.f * v);
}

// out layer epilogue (tr:585-586) + x0 prediction for the gesture net's conditioning (tr:717-725,749):
//   eps = G==2 ? o_u + s*(o_c - o_u) : o      written to eps_out[r*Dtot + off + j]
//   expr[r, j] = a*x - b*eps  (only when expr != nullptr; padding columns zeroed)
// Step scalars live in a 4-float device buffer {t_orig, a, b, cond_scale} written by step_params_kernel, so that a
// captured CUDA graph of the denoiser is valid for every diffusion step (nothing step-dependent is baked in).
__global__ void step_params_kernel(float* prm, float t, float a, float b, float s) {
  DSHEG_PDL_ENTER();
  if (threadIdx.x == 0) { prm[0] = t; prm[1] = a; prm[2] = b; prm[3] = s; }
}

template <typename TA>
__global__ void cfg_mix_kernel(const float* o, int ldo, int n_rows, int feats, int two, const float* prm, float* eps_out,
                               const float* x, int Dtot, int off, TA* expr, int ld_expr) {
  DSHEG_PDL_ENTER();
  const float a = prm[1], b = prm[2], s = prm[3];
  const int wcols = expr ? ld_expr : feats;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_rows * wcols) return;
  const int r = (int)(i / wcols), j = (int)(i % wcols);
  if (j >= feats) { AT<TA>::st(expr + (size_t)r * ld_expr + j, 0.f); return; }
  float e;
  if (two) {
    const float u = o[(size_t)r * ldo + j], c = o[(size_t)(r + n_rows) * ldo + j];
    e = u + s * (c - u);
  } else {
    e = o[(size_t)r * ldo + j];
  }
  const size_t xi = (size_t)r * Dtot + off + j;
  eps_out[xi] = e;
  if (expr) AT<TA>::st(expr + (size_t)r * ld_expr + j, a * x[xi] - b * e);
}

// sinusoidal timestep embedding (tr:42-59): [cos(t*f) | sin(t*f)], f uploaded from the host so the
// arguments are bit-identical to torch's.
__global__ void sinus_kernel(const float* prm, const float* freqs, int half, float* out) {
  DSHEG_PDL_ENTER();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const float arg = prm[0] * freqs[i];
  out[i] = cosf(arg);
  out[half + i] = sinf(arg);
}

// Batched GEMV: out[p][n] = act( sum_k in[p][k] * W[p][n][k] + b[p][n] ), one warp per (p, n).
// in_silu: apply SiLU to the input on the fly (StylizationBlock.emb_layers, tr:75-78).
struct GemvProb { const float* in; const float* w; const float* b; float* out; };
struct GemvBatch { GemvProb p[3]; };
__global__ void gemv_kernel(GemvBatch gb, int N, int K, int act, int in_silu) {
  DSHEG_PDL_ENTER();
  const GemvProb pr = gb.p[blockIdx.y];
  const int n = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (n >= N) return;
  const float* w = pr.w + (size_t)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float v = pr.in[k];
    if (in_silu) v = silu_f(v);
    acc = fmaf(v, w[k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) pr.out[n] = apply_act(acc + pr.b[n], act);
}

// embs[b, :] = SiLU(temb + pide[b, :])   (tr:559 then the SiLU of every emb_layers, tr:75-78)
template <typename TA>
__global__ void embs_kernel(const float* temb, const float* pide, TA* embs, int B, int E) {
  DSHEG_PDL_ENTER();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * E) return;
  AT<TA>::st(embs + i, silu_f(temb[i % E] + pide[i]));
}

// Conv1d(k=3, pad=1, no bias) over time on [B,T,Cin] -> [B,T,128] with folded BN bias + optional GELU
// (hubert_encoder, tr:436-442).  w layout [3][Cin][128] (co contiguous).  8 frames per CTA.
constexpr int HC_TR = 8, HC_CO = 128;
template <typename TOUT>
__global__ void __launch_bounds__(HC_CO) hubconv_kernel(const float* in, int Cin, const float* w, const float* bias,
                                                        int act, TOUT* out, int ldo, int T, int tiles_per_sample) {
#ifdef DSHEG_EMU
  float* xs = reinterpret_cast<float*>(emu::self().cta->smem);
#else
  extern __shared__ float xs[];  // [(HC_TR+2)][Cin]
#endif
  const int smp = blockIdx.x / tiles_per_sample, t0 = (blockIdx.x % tiles_per_sample) * HC_TR;
  const int co = threadIdx.x;
  for (int e = threadIdx.x; e < (HC_TR + 2) * Cin; e += HC_CO) {
    const int rr = e / Cin, c = e % Cin;
    const int t = t0 + rr - 1;
    xs[e] = (t >= 0 && t < T) ? in[((size_t)smp * T + t) * Cin + c] : 0.f;
  }
  __syncthreads();
  float acc[HC_TR];
#pragma unroll
  for (int i = 0; i < HC_TR; ++i) acc[i] = 0.f;
  for (int dk = 0; dk < 3; ++dk) {
    for (int c = 0; c < Cin; ++c) {
      const float wv = w[((size_t)dk * Cin + c) * HC_CO + co];
#pragma unroll
      for (int i = 0; i < HC_TR; ++i) acc[i] = fmaf(xs[(i + dk) * Cin + c], wv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < HC_TR; ++i) {
    const int t = t0 + i;
    if (t < T) {
      float v = acc[i] + (bias ? bias[co] : 0.f);
      v = apply_act(v, act);
      AT<TOUT>::st(out + ((size_t)smp * T + t) * ldo + co, v);
    }
  }
}

}  // namespace dsheg
