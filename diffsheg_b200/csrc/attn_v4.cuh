// Linear-attention core + StylizationBlock prologue, bf16, as a CLUSTER OF TWO half-sample CTAs (opt-in: DSHEG_ATTN=v4).
//
// Same arithmetic and per-head dataflow as attn_v3.cuh (reference transformer.py:112-130 + :86-97); what changes is the
// decomposition.  v3 runs one 512-thread / 205 KB CTA per sample, so exactly one CTA fits an SM and its phases (cp.async
// fill -> column softmax -> two MMA passes -> LayerNorm/SiLU store) are serialised on that SM: the memory pipe idles during
// the compute phases and vice versa (ncu, profiles/r01: 43 % issue slots, 30 % of HBM).  Here a sample is split over the two
// CTAs of a cluster -- CTA r owns heads 4r..4r+3, 256 threads, 104 KB -- so TWO CTAs (of different samples, in different
// phases) share an SM.  The only cross-head quantity, the LayerNorm mean / variance over all 512 columns of a row, is
// exchanged through distributed shared memory: each CTA reduces (sum, sum of squares) over its 256 columns per row, stores
// the pair into its peer's smem (st.shared::cluster), one cluster barrier, then normalises its own 256 columns.
// HBM traffic is unchanged: read q,k,v + write z = 4 * T * 512 * 2 bytes per sample.
//
// STATUS: written after round 1's GPU budget was spent -- compiles for sm_100a, never executed.  Not selected by default.
#pragma once
#include "attn_v3.cuh"

namespace dsheg {
namespace av4 {

using av3::TP; using av3::HD; using av3::D; using av3::TILE_BYTES;
using av3::smem_addr; using av3::cp_async16; using av3::cp_async_wait_all; using av3::ldsm_x4; using av3::ldsm_x4_trans;
using av3::mma_bf16; using av3::pair_sync; using av3::pack2; using av3::unpack2; using av3::ex2f; using av3::swz; using av3::load_q_tile;

constexpr int NH_CTA = 4;                 // heads per CTA
constexpr int NTHREADS = 256;
constexpr int RED_FLOATS = NH_CTA * 2 * HD;
constexpr int STAT_FLOATS = 2 * TP;       // (sum, sumsq) per row
constexpr int SMEM_BYTES = NH_CTA * 2 * TILE_BYTES + 2 * RED_FLOATS * 4 + 2 * STAT_FLOATS * 4;   // 98304 + 4096 + 1536

using prims::cluster_rank; using prims::cluster_barrier; using prims::st_peer_f32x2; using av3::tanh_approx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 2)
attn_v4_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ z, int T, int B, const float* __restrict__ ln_g,
               const float* __restrict__ ln_b, const float* __restrict__ ss, int ss_ld) {
  DSHEG_DYN_SMEM(sm, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hl = warp >> 1, half = warp & 1;           // local head, warp of the pair
  const int g = lane >> 2, q = lane & 3;
  const uint32_t rank = cluster_rank();
  const int smp = blockIdx.x >> 1;
  const int head = (int)rank * NH_CTA + hl;            // global head
  const size_t row0 = (size_t)smp * T;
  uint8_t* Ks = sm + hl * 2 * TILE_BYTES;
  uint8_t* Vs = Ks + TILE_BYTES;
  float* red = reinterpret_cast<float*>(sm + NH_CTA * 2 * TILE_BYTES);  // [NH_CTA][2][HD] column max partials
  float* red2 = red + RED_FLOATS;                                        // [NH_CTA][2][HD] column sum partials
  float* st_mine = red2 + RED_FLOATS;                                    // [TP][2] this CTA's row partials
  float* st_peer = st_mine + STAT_FLOATS;                                // [TP][2] written by the peer CTA
  const uint32_t ks_addr = smem_addr(Ks), vs_addr = smem_addr(Vs);
  const bf16* qhead = qkv + row0 * (3 * D) + head * HD;
  const int n_mt = (T + 15) >> 4;
  const int Tpad = n_mt * 16;

  // ---- 1. K and V head tiles -> smem (cp.async); the pair splits the rows by parity
  for (int i = lane; i < ((T - half + 1) >> 1) * 8; i += 32) {
    const int r = 2 * (i >> 3) + half, c = i & 7;
    const bf16* src = qhead + (size_t)r * (3 * D) + c * 8;
    cp_async16(ks_addr + swz(r, c), src + D);
    cp_async16(vs_addr + swz(r, c), src + 2 * D);
  }
  if (half == 0) {
    for (int i = lane; i < (Tpad - T) * 8; i += 32) {
      const int r = T + (i >> 3), c = i & 7;
      *reinterpret_cast<uint4*>(Vs + swz(r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  // ---- 2. first Q m-tile (global -> registers) overlaps the cp.async latency
  uint32_t qa[4][4];
  if (half < n_mt) load_q_tile(qhead, half * 16, T, g, q, qa);
  cp_async_wait_all();
  pair_sync(hl);

  // ---- 3. softmax over time per K column (identical to v3 step 3)
  {
    const int c = lane >> 2, w = lane & 3;
    const int rsplit = (T + 1) >> 1;
    const int r_lo = half ? rsplit : 0, r_hi = half ? T : rsplit;
    __nv_bfloat162 mx2 = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll 8
    for (int r = r_lo; r < r_hi; ++r)
      mx2 = __hmax2(mx2, *reinterpret_cast<const __nv_bfloat162*>(Ks + swz(r, c) + w * 4));
    float m0 = __bfloat162float(mx2.x), m1 = __bfloat162float(mx2.y);
    float* myred = red + (hl * 2 + half) * HD;
    const float* otred = red + (hl * 2 + (half ^ 1)) * HD;
    myred[2 * lane] = m0;
    myred[2 * lane + 1] = m1;
    pair_sync(hl);
    m0 = fmaxf(m0, otred[2 * lane]);
    m1 = fmaxf(m1, otred[2 * lane + 1]);
    const float L2E = 1.4426950408889634f;
    const float n0 = -m0 * L2E, n1 = -m1 * L2E;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int r = r_lo; r < r_hi; ++r) {
      uint32_t* p = reinterpret_cast<uint32_t*>(Ks + swz(r, c) + w * 4);
      const float2 v = unpack2(*p);
      const float e0 = ex2f(fmaf(v.x, L2E, n0)), e1 = ex2f(fmaf(v.y, L2E, n1));
      s0 += e0;
      s1 += e1;
      *p = pack2(e0, e1);
    }
    float* myred2 = red2 + (hl * 2 + half) * HD;
    myred2[2 * lane] = s0;
    myred2[2 * lane + 1] = s1;
    if (half == 1) {
      for (int i = lane; i < (Tpad - T) * 8; i += 32) {
        const int r = T + (i >> 3), cc = i & 7;
        *reinterpret_cast<uint4*>(Ks + swz(r, cc)) = make_uint4(0, 0, 0, 0);
      }
    }
    pair_sync(hl);
  }

  // ---- 4. A^T[l][d] = sum_t V[t][l] K'[t][d] for this warp's l-half (identical to v3 step 4)
  float acc[2][8][4];
  {
    const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { acc[mi][nt][0] = acc[mi][nt][1] = acc[mi][nt][2] = acc[mi][nt][3] = 0.f; }
    for (int kt = 0; kt < n_mt; ++kt) {
      uint32_t a0[4], a1[4];
      {
        const int r = kt * 16 + rr + ((mat >> 1) << 3);
        ldsm_x4_trans(vs_addr + swz(r, 4 * half + (mat & 1)), a0[0], a0[1], a0[2], a0[3]);
        ldsm_x4_trans(vs_addr + swz(r, 4 * half + 2 + (mat & 1)), a1[0], a1[1], a1[2], a1[3]);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int r = kt * 16 + rr + ((mat & 1) << 3), c = 2 * np + (mat >> 1);
        ldsm_x4_trans(ks_addr + swz(r, c), b0, b1, b2, b3);
        mma_bf16(acc[0][2 * np], a0, b0, b1);
        mma_bf16(acc[0][2 * np + 1], a0, b2, b3);
        mma_bf16(acc[1][2 * np], a1, b0, b1);
        mma_bf16(acc[1][2 * np + 1], a1, b2, b3);
      }
    }
  }
  pair_sync(hl);  // both warps are done reading K' and V: K's tile now receives A^T, V's tile Y
  {
    const float* sa = red2 + (hl * 2) * HD;
    const float* sb = sa + HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d0 = 8 * nt + 2 * q;
      const float i0 = 1.f / (sa[d0] + sb[d0]), i1 = 1.f / (sa[d0 + 1] + sb[d0 + 1]);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int l = 32 * half + 16 * mi + g;
        *reinterpret_cast<uint32_t*>(Ks + swz(l, nt) + q * 4) = pack2(acc[mi][nt][0] * i0, acc[mi][nt][1] * i1);
        *reinterpret_cast<uint32_t*>(Ks + swz(l + 8, nt) + q * 4) = pack2(acc[mi][nt][2] * i0, acc[mi][nt][3] * i1);
      }
    }
  }
  pair_sync(hl);  // A^T[l][d] (bf16, 64 x 64) complete

  // ---- 5. Y[t][l] = softmax_d(Q)[t][:] . A for this warp's m-tiles; bf16 Y -> V tile (identical to v3 step 5)
  {
    const int mat = lane >> 3, rr = lane & 7;
    for (int mt = half; mt < n_mt; mt += 2) {
      __nv_bfloat162 ma = __floats2bfloat162_rn(-INFINITY, -INFINITY), mb = ma;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ma = __hmax2(ma, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&qa[ks][0]), *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][2])));
        mb = __hmax2(mb, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&qa[ks][1]), *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][3])));
      }
      float mx0 = fmaxf(__bfloat162float(ma.x), __bfloat162float(ma.y)), mx1 = fmaxf(__bfloat162float(mb.x), __bfloat162float(mb.y));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float L2E = 1.4426950408889634f;
      const float n0 = -mx0 * L2E, n1 = -mx1 * L2E;
      float sm0 = 0.f, sm1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 a0 = unpack2(qa[ks][0]), a1 = unpack2(qa[ks][1]), a2 = unpack2(qa[ks][2]), a3 = unpack2(qa[ks][3]);
        const float e00 = ex2f(fmaf(a0.x, L2E, n0)), e01 = ex2f(fmaf(a0.y, L2E, n0));
        const float e20 = ex2f(fmaf(a2.x, L2E, n0)), e21 = ex2f(fmaf(a2.y, L2E, n0));
        const float e10 = ex2f(fmaf(a1.x, L2E, n1)), e11 = ex2f(fmaf(a1.y, L2E, n1));
        const float e30 = ex2f(fmaf(a3.x, L2E, n1)), e31 = ex2f(fmaf(a3.y, L2E, n1));
        sm0 += (e00 + e01) + (e20 + e21);
        sm1 += (e10 + e11) + (e30 + e31);
        pa[ks][0] = pack2(e00, e01);
        pa[ks][1] = pack2(e10, e11);
        pa[ks][2] = pack2(e20, e21);
        pa[ks][3] = pack2(e30, e31);
      }
      sm0 += __shfl_xor_sync(0xffffffffu, sm0, 1); sm0 += __shfl_xor_sync(0xffffffffu, sm0, 2);
      sm1 += __shfl_xor_sync(0xffffffffu, sm1, 1); sm1 += __shfl_xor_sync(0xffffffffu, sm1, 2);
      if (mt + 2 < n_mt) load_q_tile(qhead, (mt + 2) * 16, T, g, q, qa);
      float y[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { y[nt][0] = y[nt][1] = y[nt][2] = y[nt][3] = 0.f; }
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const int r = 16 * np + rr + ((mat >> 1) << 3), c = 2 * kd + (mat & 1);
          ldsm_x4(ks_addr + swz(r, c), b0, b1, b2, b3);
          mma_bf16(y[2 * np], pa[kd], b0, b1);
          mma_bf16(y[2 * np + 1], pa[kd], b2, b3);
        }
      }
      const float r0 = 1.f / sm0, r1 = 1.f / sm1;
      const int ra = mt * 16 + g, rb = ra + 8;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(Vs + swz(ra, nt) + q * 4) = pack2(y[nt][0] * r0, y[nt][1] * r0);
        *reinterpret_cast<uint32_t*>(Vs + swz(rb, nt) + q * 4) = pack2(y[nt][2] * r1, y[nt][3] * r1);
      }
    }
  }
  __syncthreads();  // this CTA's 4 heads of Y are in smem

  // ---- 6. LayerNorm(512) * (1 + scale) + shift, SiLU over this CTA's 256 columns; row statistics via the peer CTA
  {
    const int hh = lane >> 3, c = lane & 7;   // lane covers local head hh, 16-B chunk c (8 columns)
    const uint8_t* Yh = sm + hh * 2 * TILE_BYTES + TILE_BYTES;
    // 6a. partial (sum, sum of squares) over 256 columns per row -> own table and the peer's
    for (int t = warp; t < T; t += NTHREADS / 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(Yh + swz(t, c));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      float s = 0.f, sq = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 p2 = unpack2(w[e]);
        s += p2.x + p2.y;
        sq = fmaf(p2.x, p2.x, fmaf(p2.y, p2.y, sq));
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o2); sq += __shfl_xor_sync(0xffffffffu, sq, o2); }
      if (lane == 0) {
        st_mine[2 * t] = s;
        st_mine[2 * t + 1] = sq;
        st_peer_f32x2(smem_addr(st_peer + 2 * t), rank ^ 1u, s, sq);
      }
    }
    // per-column constants of this lane's 8 columns (folded like v3): t = (v-mean)*rstd*2G + 2Bc, h = t/2
    const int col0 = ((int)rank * NH_CTA + hh) * HD + c * 8;
    const float* sc = ss + (size_t)(smp % B) * ss_ld;
    float G[8], Bc[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ln_g + col0 + e)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + col0 + e));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc + col0 + e)), d4 = __ldg(reinterpret_cast<const float4*>(sc + D + col0 + e));
      G[e] = 0.5f * a.x * (1.f + c4.x); G[e + 1] = 0.5f * a.y * (1.f + c4.y); G[e + 2] = 0.5f * a.z * (1.f + c4.z); G[e + 3] = 0.5f * a.w * (1.f + c4.w);
      Bc[e] = 0.5f * fmaf(b4.x, 1.f + c4.x, d4.x); Bc[e + 1] = 0.5f * fmaf(b4.y, 1.f + c4.y, d4.y);
      Bc[e + 2] = 0.5f * fmaf(b4.z, 1.f + c4.z, d4.z); Bc[e + 3] = 0.5f * fmaf(b4.w, 1.f + c4.w, d4.w);
    }
    cluster_barrier();  // both tables complete and visible (release / acquire at cluster scope); no remote access after this
    // 6b. normalise, modulate, SiLU, store: a warp writes 512 contiguous bytes per row
    for (int t = warp; t < T; t += NTHREADS / 32) {
      const float s = st_mine[2 * t] + st_peer[2 * t], sq = st_mine[2 * t + 1] + st_peer[2 * t + 1];
      const float mean = s * (1.f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.f / D) - mean * mean, 0.f) + 1e-5f);
      const float nmr = -mean * rstd;
      const uint4 u = *reinterpret_cast<const uint4*>(Yh + swz(t, c));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 p2 = unpack2(w[e]);
        const float h0 = fmaf(fmaf(p2.x, rstd, nmr), G[2 * e], Bc[2 * e]);
        const float h1 = fmaf(fmaf(p2.y, rstd, nmr), G[2 * e + 1], Bc[2 * e + 1]);
        const float t0 = tanh_approx(h0), t1 = tanh_approx(h1);
        o[e] = pack2(fmaf(h0, t0, h0), fmaf(h1, t1, h1));
      }
      *reinterpret_cast<uint4*>(z + (row0 + t) * (size_t)D + col0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace av4
}  // namespace dsheg
