// fp32-accumulate SIMT GEMM with the shared GemmDesc epilogue.
//
// Role: (1) the arithmetic engine of the strict-fp32 parity mode (the reference runs torch
// fp32 with TF32 off), (2) the once-per-window fp32 ops in bf16 mode (person-id
// MLP), and (3) a bisecting aid: DSHEG_GEMM_ENGINE=simt runs the bf16-mode graph on this
// kernel instead of the tcgen05 one.  It is NOT the performance path (see gemm_tc.cuh).
#pragma once
#include "common.cuh"

namespace dsheg {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename TA, typename TW>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc d) {
#ifdef DSHEG_EMU   // tests/emu: the two static tiles live in the emulated CTA's shared-memory window (launch_gemm_simt passes their size)
  float (*As)[SG_BM + 4] = reinterpret_cast<float (*)[SG_BM + 4]>(emu::self().cta->smem);
  float (*Ws)[SG_BN + 4] = reinterpret_cast<float (*)[SG_BN + 4]>(emu::self().cta->smem + sizeof(float) * SG_BK * (SG_BM + 4));
#else
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
#endif
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const TW* W = reinterpret_cast<const TW*>(d.w);
  int koff = 0;  // padded K offset of the current segment inside W
  for (int s = 0; s < d.nseg; ++s) {
    const TA* A = reinterpret_cast<const TA*>(d.a[s].ptr);
    const int lda = d.a[s].ld, ks = d.a[s].k;
    for (int k0 = 0; k0 < ks; k0 += SG_BK) {
#pragma unroll
      for (int e = tid; e < SG_BM * SG_BK; e += 256) {
        const int r = e / SG_BK, kk = e % SG_BK;
        const int m = m0 + r, k = k0 + kk;
        As[kk][r] = (m < d.M && k < ks) ? AT<TA>::ld(A + (size_t)m * lda + k) : 0.f;
        const int n = n0 + r;
        Ws[kk][r] = (n < d.N && k < ks) ? AT<TW>::ld(W + (size_t)n * d.Kp + koff + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SG_BK; ++kk) {
        float a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
    koff += (ks + 63) / 64 * 64;
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
    const float mu = d.csum ? d.mu[m] : 0.f, rstd = d.csum ? d.rstd[m] : 1.f;
    const int mr = d.res_mod > 0 ? m % d.res_mod : m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.N) continue;
      float v = acc[i][j];
      if (d.csum) v = rstd * (v - mu * d.csum[n]);
      if (d.bias) v += d.bias[n];
      v = apply_act(v, d.act);
      if (d.res) {
        v += d.res_f32 ? reinterpret_cast<const float*>(d.res)[(size_t)mr * d.ldr + n]
                       : AT<TA>::ld(reinterpret_cast<const TA*>(d.res) + (size_t)mr * d.ldr + n);
      }
      const size_t o = (size_t)m * d.ldo + n;
      if (d.out_f32) {
        reinterpret_cast<float*>(d.out)[o] = v;
        if (d.out2) reinterpret_cast<float*>(d.out2)[o] = v;
      } else {
        AT<TA>::st(reinterpret_cast<TA*>(d.out) + o, v);
        if (d.out2) AT<TA>::st(reinterpret_cast<TA*>(d.out2) + o, v);
      }
    }
  }
}

template <typename TA, typename TW>
inline cudaError_t launch_gemm_simt(const GemmDesc& d, cudaStream_t st) {
  dim3 grid((d.N + SG_BN - 1) / SG_BN, (d.M + SG_BM - 1) / SG_BM);
  const auto kern = gemm_simt_kernel<TA, TW>;
#ifdef DSHEG_EMU
  DSHEG_LAUNCH_PLAIN(kern, grid, 256, sizeof(float) * SG_BK * (SG_BM + 4 + SG_BN + 4), st, d);
#else
  DSHEG_LAUNCH_PLAIN(kern, grid, 256, 0, st, d);
#endif
  return cudaGetLastError();
}

}  // namespace dsheg
