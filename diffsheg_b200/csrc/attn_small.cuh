// Linear-attention core + StylizationBlock prologue for the AUDIO encoder layer (encoder_aud: D = 128, 8 heads of 16, T <= 96),
// bf16 in / out.  Default for the audio layer since its first hardware run (round 2: -7 ms per B = 950 step; DSHEG_ATTN_AUD=0 opts out); also runs on the CPU emulator (tests/test_emu_kernels.py).
// Same mathematics as every attention kernel here (reference transformer.py:112-130 + :86-97):
//   K' = softmax_t(K)   Q' = softmax_d(Q)   A_h = K'_h^T V_h  [16 x 16]   Y_h = Q'_h A_h   z = SiLU(LN_128(Y) * (1 + scale) + shift)
//
// Why: the generic SIMT kernel (kernels.cuh attn_kernel<bf16, 16>) walks the 8 heads one after the other with 256 threads, six CTA
// barriers per head, scalar loads and an fp32 round trip of Y through global memory: 336 us per call at the headline batch
// (profiles/r01/final_launches_summary.txt) for 85 MB of algorithmic traffic = 0.25 TB/s, 1.5 % of a denoiser call.  This kernel
// handles ALL heads at once: one 512-thread CTA per sample, Q / K / V staged once as fp32 in shared memory with 16-byte loads,
// column softmax over all 128 columns in parallel (its 1/sum folded into A), A and Y as register-blocked 1 x 4 strips, five CTA
// barriers in total, LayerNorm from shared memory, 8-byte coalesced bf16 stores.  The FLOPs are negligible (0.7 MFLOP per sample):
// no tensor cores, the kernel only has to stream 90 KB per sample.
#pragma once
#include "common.cuh"

namespace dsheg {
namespace asmall {

constexpr int D = 128, NH = 8, HD = 16, TP = 96;
constexpr int NTHREADS = 512;
constexpr int LD = D + 4;                 // fp32 row stride of the staged tiles (16-byte aligned rows, rows land on different banks)
constexpr int ALD = 20;                   // row stride of A[h][d][.] (float4-aligned)
constexpr int AHS = HD * ALD + 16;        // head stride of A: consecutive heads sit 16 banks apart (conflict-free float4 reads in step 4)
constexpr int NPART = NTHREADS / D;       // 4 row partitions for the column-parallel passes
inline int smem_bytes(int T) { return (3 * T * LD + NH * AHS + 2 * NPART * D + D) * 4; }

__device__ __forceinline__ float bflo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfhi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_d128_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ z, int T, int B, const float* __restrict__ ln_g,
                 const float* __restrict__ ln_b, const float* __restrict__ ss, int ss_ld) {
  DSHEG_PDL_ENTER();
#ifdef DSHEG_EMU
  float* sm = reinterpret_cast<float*>(emu::self().cta->smem);
#else
  extern __shared__ __align__(16) float sm[];
#endif
  float* Qs = sm;                          // [T][LD]  q, then exp(q - rowmax) / rowsum
  float* Ks = Qs + T * LD;                 // [T][LD]  k, then exp(k - colmax) (unnormalised), then Y
  float* Vs = Ks + T * LD;                 // [T][LD]
  float* As = Vs + T * LD;                 // [NH][AHS] = [NH][HD][ALD] + padding
  float* red = As + NH * AHS;         // [2][NPART][D] column partials (max, then sum)
  float* inv = red + 2 * NPART * D;        // [D] 1 / column sum
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int smp = blockIdx.x;
  const size_t row0 = (size_t)smp * T;

  // ---- 1. stage Q | K | V (bf16 -> fp32): 48 chunks of 8 columns per row, 16-byte global loads
  for (int e = tid; e < T * 48; e += NTHREADS) {
    const int t = e / 48, c = e % 48;
    const uint4 u = *reinterpret_cast<const uint4*>(qkv + (row0 + t) * (size_t)(3 * D) + c * 8);
    float* dst = (c < 16 ? Qs : (c < 32 ? Ks : Vs)) + t * LD + (c & 15) * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(bflo(u.x), bfhi(u.x), bflo(u.y), bfhi(u.y));
    *reinterpret_cast<float4*>(dst + 4) = make_float4(bflo(u.z), bfhi(u.z), bflo(u.w), bfhi(u.w));
  }
  __syncthreads();

  // ---- 2a. softmax over time for every K column (tr:123), all 128 columns at once; thread = (column, row partition)
  const int col = tid & (D - 1), part = tid >> 7;
  {
    float m = -INFINITY;
    for (int t = part; t < T; t += NPART) m = fmaxf(m, Ks[t * LD + col]);
    red[part * D + col] = m;
  }
  // ---- 2b. softmax over the 16 channels of a head for every Q row (tr:122): thread = (row, head), normalised in place
  for (int p = tid; p < T * NH; p += NTHREADS) {
    float* qr = Qs + (p >> 3) * LD + (p & 7) * HD;
    float v[HD];
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 f = *reinterpret_cast<const float4*>(qr + d);
      v[d] = f.x; v[d + 1] = f.y; v[d + 2] = f.z; v[d + 3] = f.w;
    }
    float mx = v[0];
#pragma unroll
    for (int d = 1; d < HD; ++d) mx = fmaxf(mx, v[d]);
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) { v[d] = __expf(v[d] - mx); s += v[d]; }
    const float r = 1.f / s;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(qr + d) = make_float4(v[d] * r, v[d + 1] * r, v[d + 2] * r, v[d + 3] * r);
  }
  __syncthreads();
  {
    float m = red[col];
#pragma unroll
    for (int p2 = 1; p2 < NPART; ++p2) m = fmaxf(m, red[p2 * D + col]);
    float s = 0.f;
    for (int t = part; t < T; t += NPART) {
      const float e = __expf(Ks[t * LD + col] - m);
      Ks[t * LD + col] = e;
      s += e;
    }
    red[(NPART + part) * D + col] = s;
  }
  __syncthreads();
  if (tid < D) {
    float s = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < NPART; ++p2) s += red[(NPART + p2) * D + tid];
    inv[tid] = 1.f / s;
  }
  // ---- 3. A_h[d][l] = (1 / colsum[d]) sum_t e[t][d] V[t][l]: one 1 x 4 strip per thread (8 heads x 16 d x 4 strips = 512)
  {
    const int h = tid >> 6, d = (tid >> 2) & 15, l4 = (tid & 3) * 4;
    const float* kp = Ks + h * HD + d;
    const float* vp = Vs + h * HD + l4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float k = kp[t * LD];
      const float4 v = *reinterpret_cast<const float4*>(vp + t * LD);
      a0 = fmaf(k, v.x, a0); a1 = fmaf(k, v.y, a1); a2 = fmaf(k, v.z, a2); a3 = fmaf(k, v.w, a3);
    }
    __syncthreads();   // inv[] complete (and every thread is done reading the column partials)
    const float r = inv[h * HD + d];
    *reinterpret_cast<float4*>(As + h * AHS + d * ALD + l4) = make_float4(a0 * r, a1 * r, a2 * r, a3 * r);
  }
  __syncthreads();     // A complete; nobody reads K' any more: its tile receives Y

  // ---- 4. Y[t][h 16 + l] = sum_d Q'[t][h 16 + d] A_h[d][l]: 1 x 4 strips, 32 strips per row
  for (int s = tid; s < T * 32; s += NTHREADS) {
    const int t = s >> 5, c4 = (s & 31) * 4, h = c4 >> 4, l4 = c4 & 15;
    const float* qp = Qs + t * LD + h * HD;
    const float* ap = As + h * AHS + l4;
    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(qp + d);   // 16-byte loads: a quarter-warp touches two heads only
      const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 a = *reinterpret_cast<const float4*>(ap + (d + e) * ALD);
        y0 = fmaf(qv[e], a.x, y0); y1 = fmaf(qv[e], a.y, y1); y2 = fmaf(qv[e], a.z, y2); y3 = fmaf(qv[e], a.w, y3);
      }
    }
    *reinterpret_cast<float4*>(Ks + t * LD + c4) = make_float4(y0, y1, y2, y3);
  }
  __syncthreads();

  // ---- 5. StylizationBlock prologue: LN(128) * (1 + scale) + shift, SiLU; one warp per row, 4 columns per lane
  {
    const float* sc = ss + (size_t)(smp % B) * ss_ld;
    const int c = lane * 4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(ln_g + c)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + c));
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c)), h4 = __ldg(reinterpret_cast<const float4*>(sc + D + c));
    for (int t = warp; t < T; t += NTHREADS / 32) {
      const float4 y = *reinterpret_cast<const float4*>(Ks + t * LD + c);
      float s = (y.x + y.y) + (y.z + y.w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.f / D);
      const float d0 = y.x - mean, d1 = y.y - mean, d2 = y.z - mean, d3 = y.w - mean;
      float var = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = rsqrtf(var * (1.f / D) + 1e-5f);
      float o0 = fmaf(d0 * rstd, g4.x, b4.x), o1 = fmaf(d1 * rstd, g4.y, b4.y), o2 = fmaf(d2 * rstd, g4.z, b4.z), o3 = fmaf(d3 * rstd, g4.w, b4.w);
      o0 = fmaf(o0, 1.f + s4.x, h4.x); o1 = fmaf(o1, 1.f + s4.y, h4.y); o2 = fmaf(o2, 1.f + s4.z, h4.z); o3 = fmaf(o3, 1.f + s4.w, h4.w);
      o0 = o0 / (1.f + __expf(-o0)); o1 = o1 / (1.f + __expf(-o1)); o2 = o2 / (1.f + __expf(-o2)); o3 = o3 / (1.f + __expf(-o3));
      *reinterpret_cast<uint2*>(z + (row0 + t) * (size_t)D + c) = make_uint2(pack_bf2(o0, o1), pack_bf2(o2, o3));
    }
  }
}

}  // namespace asmall
}  // namespace dsheg
