// Linear-attention core + StylizationBlock prologue, bf16 (D = 512, 8 heads of 64, T <= 96) -- sixth generation: the HIGH-OCCUPANCY
// kernel for static-shift numerators.  Opt-in (DSHEG_ATTN=v6 together with DSHEG_EXPO=1) until its first hardware run; validated
// thread for thread on the CPU emulator (tests/test_emu_kernels.py).
//
// Input contract = attn_v5<CL, PRE = 2>: the Q and K columns of `qkv` hold exp(value - static shift) (ACT_EXPO epilogue of the QKV
// GEMM, gemm_tc.cuh; shifts proven safe by pack.py:expo_shift), V is plain.  Same mathematics (transformer.py:112-130 + :86-97):
//   A = K'^T V / colsum(K')  [64 x 64]     Y = Q' A / rowsum(Q')     z = SiLU(LN_512(Y) * (1 + scale) + shift)
//
// Why another kernel: v3 / v5 hold 16 warps per SM (512 threads x 128 registers = the whole register file) and issue on 43 % of
// the slots.  With the softmaxes gone from the kernel, the register hogs are the A^T accumulators (64 per lane) and the packed
// rows the LayerNorm pass keeps across the cluster barrier.  v6 gives a head FOUR warps instead of two:
//   * A^T: one 32 x 32 quadrant (l-half x d-half) per warp (32 accumulator registers instead of 64); the K' column sums ride on
//     the same fragments (ones . K' on the tensor core), 16 columns per warp;
//   * Y = Q' A: a warp owns one l-half and every second m-tile (16 accumulator registers); Q fragments are loaded just in time;
//   * LayerNorm pass: 8 columns per lane (16 registers of folded constants), rows re-read from shared memory after the
//     cluster barrier instead of being carried across it;
// so the kernel fits 64 registers: __launch_bounds__(256, 4) as clusters of 4 CTAs (2 heads each, 53 KB) or (512, 2) as clusters of 2
// (4 heads each, 101 KB) = 32 warps per SM either way -- twice the latency-hiding of v5 at the same shared-memory footprint.  The price is redundancy: a quadrant warp
// re-reads its half of V and of K' (+14 % ldmatrix) and the two l-halves both load Q and sum its rows (+5 % mma).
// HBM traffic is unchanged: read q', k', v + write z = 4 * T * 512 * 2 bytes per sample.
#pragma once
#include "attn_v5.cuh"

namespace dsheg {
namespace av6 {

using av3::TP; using av3::HD; using av3::D; using av3::TILE_BYTES;
using av3::pack2; using av3::swz;
using av5::bf_pair; using av5::BF2_ONES; using av5::load_q;
using prims::smem_addr; using prims::cp_async16; using prims::cp_async_commit; using prims::cp_async_wait_group;
using prims::ldsm_x4; using prims::ldsm_x4_trans; using prims::mma_bf16; using prims::tanh_approx; using prims::rcp_approx;
using prims::ffma2; using prims::fadd2; using prims::fmul2;

constexpr int WPH = 4;                        // warps per head
template <int CL> struct Cfg {                // CL = CTAs per sample: a cluster of 4 (2 heads, 256 threads) or 2 (4 heads, 512), or ONE CTA of 1024
  static_assert(CL == 1 || CL == 2 || CL == 4, "1, 2 or 4 CTAs per sample");
  static constexpr int NH_CTA = 8 / CL;              // heads per CTA
  static constexpr int NWARPS = NH_CTA * WPH;        // 8 / 16
  static constexpr int NTHREADS = 32 * NWARPS;       // 256 / 512
  static constexpr int COLS = D / CL;                // LayerNorm columns owned by this CTA (128 / 256)
  static constexpr int LPR = CL == 1 ? 32 : COLS / 8;   // lanes per row in the LayerNorm pass (8 columns per lane; CL = 1: two-phase pass)
  static constexpr int RPI = 32 / LPR;               // rows per warp iteration (2 / 1)
  static constexpr int LN_ITERS = CL == 1 ? 2 * TP / NWARPS : TP / (NWARPS * RPI);   // 6
  static constexpr int SUM_BYTES = NH_CTA * HD * 4;      // [head][64] column sums of K'
  static constexpr int STAT_BYTES = CL * TP * 8;         // [source rank][row] (sum, sum of squares); CL = 1: [row] (mean, rstd)
  static constexpr int SMEM_BYTES = NH_CTA * 2 * TILE_BYTES + SUM_BYTES + STAT_BYTES;   // 52 736 / 101 376
  static constexpr int CTAS_PER_SM = 1024 / NTHREADS;    // 32 warps per SM either way
  static_assert((CL == 1 || TP % (NWARPS * RPI) == 0) && TP % 16 == 0 && (2 * TP) % NWARPS == 0, "row schedule");
};

__device__ __forceinline__ void head_sync(int hl) { prims::named_bar_sync<32 * WPH>(hl + 1); }

template <int CL>
__global__ void __launch_bounds__(Cfg<CL>::NTHREADS, Cfg<CL>::CTAS_PER_SM)
attn_v6_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ z, int T, int B, const float* __restrict__ ln_g,
               const float* __restrict__ ln_b, const float* __restrict__ ss, int ss_ld) {
  DSHEG_PDL_ENTER();
  using C = Cfg<CL>;
  constexpr int NH_CTA = C::NH_CTA, NWARPS = C::NWARPS, COLS = C::COLS, LPR = C::LPR, RPI = C::RPI, LN_ITERS = C::LN_ITERS, SUM_BYTES = C::SUM_BYTES;
  DSHEG_DYN_SMEM(sm, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hl = warp >> 2, wq = warp & 3;             // local head, warp of the head's quartet
  const int g = lane >> 2, q = lane & 3;
  const int mat = lane >> 3, rr = lane & 7;            // ldmatrix: matrix index / row inside the matrix
  const uint32_t rank = CL > 1 ? prims::cluster_rank() : 0u;
  if (CL > 1) prims::cluster_arrive_relaxed();         // "this CTA runs": matched by the cluster_wait before the first remote store
  const int smp = blockIdx.x / CL;
  const int head = (int)rank * NH_CTA + hl;            // global head
  const size_t row0 = (size_t)smp * T;
  uint8_t* Ks = sm + hl * 2 * TILE_BYTES;
  uint8_t* Vs = Ks + TILE_BYTES;
  float* colsum = reinterpret_cast<float*>(sm + NH_CTA * 2 * TILE_BYTES);                    // [NH_CTA][64]
  float2* stat = reinterpret_cast<float2*>(sm + NH_CTA * 2 * TILE_BYTES + SUM_BYTES);       // [CL][TP]
  const uint32_t ks_addr = smem_addr(Ks), vs_addr = smem_addr(Vs);
  const bf16* qhead = qkv + row0 * (3 * D) + head * HD;
  const int n_mt = (T + 15) >> 4;  // 16-row tiles that contain valid frames
  const int Tpad = n_mt * 16;

  // ---- 1. K' and V head tiles -> smem.  The 128 threads of a head: row (th >> 3) + 16 j, 16-byte chunk th & 7, so (row & 7) --
  //         the swizzle -- is a per-thread constant and both addresses advance by compile-time immediates.
  {
    const int th = tid & (32 * WPH - 1);
    const int rl = th >> 3, c = th & 7;
    const uint32_t so = swz(rl, c);
    const bf16* gp = qhead + (size_t)rl * (3 * D) + c * 8;
#pragma unroll
    for (int j = 0; j < TP / 16; ++j)
      if (rl + 16 * j < T) cp_async16(ks_addr + so + j * 2048, gp + (size_t)j * (16 * 3 * D) + D);
#pragma unroll
    for (int j = 0; j < TP / 16; ++j)
      if (rl + 16 * j < T) cp_async16(vs_addr + so + j * 2048, gp + (size_t)j * (16 * 3 * D) + 2 * D);
    cp_async_commit();
  }
  if (wq < 2) {   // rows T .. Tpad-1 of both tiles contribute zeros (disjoint from the rows the cp.async fills)
    uint8_t* tile = wq == 0 ? Ks : Vs;
    for (int i = lane; i < (Tpad - T) * 8; i += 32) {
      const int r = T + (i >> 3), c = i & 7;
      *reinterpret_cast<uint4*>(tile + swz(r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_group<0>();   // this thread's chunks
  head_sync(hl);              // ... and those of the head's other threads: K' and V complete in smem

  // ---- 2. A^T[l][d] = sum_t V[t][l] K'[t][d]: the head's four warps take the four 32 x 32 quadrants (l-half lq, d-half dq), so a
  //         warp reads half of V and half of K' per k-step (4 ldmatrix.x4 for 8 mma).  On the same K' fragments: the column sums of
  //         K' for d = 32 dq + 16 lq .. + 15 (ones[16 x 16] . K'[16 x 8] on the tensor core)
  {
    const int lq = wq & 1, dq = wq >> 1;
    float acc[2][4][4], cs[2][4];
    const uint32_t ones[4] = {BF2_ONES, BF2_ONES, BF2_ONES, BF2_ONES};
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[mi][nt][0] = acc[mi][nt][1] = acc[mi][nt][2] = acc[mi][nt][3] = 0.f; }
    cs[0][0] = cs[0][1] = cs[0][2] = cs[0][3] = cs[1][0] = cs[1][1] = cs[1][2] = cs[1][3] = 0.f;
#pragma unroll 1
    for (int kt = 0; kt < n_mt; ++kt) {  // 16 frames per k-step
      uint32_t a0[4], a1[4];
      {
        const int r = kt * 16 + rr + ((mat >> 1) << 3);
        ldsm_x4_trans(vs_addr + swz(r, 4 * lq + (mat & 1)), a0[0], a0[1], a0[2], a0[3]);
        ldsm_x4_trans(vs_addr + swz(r, 4 * lq + 2 + (mat & 1)), a1[0], a1[1], a1[2], a1[3]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // two d n-tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(ks_addr + swz(kt * 16 + rr + ((mat & 1) << 3), 4 * dq + 2 * np + (mat >> 1)), b0, b1, b2, b3);
        mma_bf16(acc[0][2 * np], a0, b0, b1);
        mma_bf16(acc[0][2 * np + 1], a0, b2, b3);
        mma_bf16(acc[1][2 * np], a1, b0, b1);
        mma_bf16(acc[1][2 * np + 1], a1, b2, b3);
        if (np == lq) {   // warp-uniform: the two warps of a d-half share its column sums
          mma_bf16(cs[0], ones, b0, b1);
          mma_bf16(cs[1], ones, b2, b3);
        }
      }
    }
    if (g == 0) {   // every accumulator row holds the same sums; row 0 publishes them
      *reinterpret_cast<float2*>(colsum + hl * HD + 32 * dq + 16 * lq + 2 * q) = make_float2(cs[0][0], cs[0][1]);
      *reinterpret_cast<float2*>(colsum + hl * HD + 32 * dq + 16 * lq + 8 + 2 * q) = make_float2(cs[1][0], cs[1][1]);
    }
    head_sync(hl);  // all four warps are done reading K' and V and have published their column sums: K's tile receives A^T
    const float* csum = colsum + hl * HD + 32 * dq;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 s2 = *reinterpret_cast<const float2*>(csum + 8 * nt + 2 * q);
      const float2 inv = make_float2(rcp_approx(s2.x), rcp_approx(s2.y));
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int l = 32 * lq + 16 * mi + g;
        const float2 lo = fmul2(make_float2(acc[mi][nt][0], acc[mi][nt][1]), inv), hi = fmul2(make_float2(acc[mi][nt][2], acc[mi][nt][3]), inv);
        *reinterpret_cast<uint32_t*>(Ks + swz(l, 4 * dq + nt) + q * 4) = pack2(lo.x, lo.y);
        *reinterpret_cast<uint32_t*>(Ks + swz(l + 8, 4 * dq + nt) + q * 4) = pack2(hi.x, hi.y);
      }
    }
  }
  head_sync(hl);  // A^T[l][d] (bf16, 64 x 64) complete; V's tile receives Y

  // ---- 3. Y[t][l] = Q'[t][:] . A / rowsum(Q') for l-half lh = wq & 1 and m-tiles mt = (wq >> 1), +2, +4; bf16 Y -> V tile.
  //         The row sums come out of the tensor core too (Q' . ones): exactly the bf16 weights the product uses.
  {
    const int lh = wq & 1;
    for (int mt = wq >> 1; mt < n_mt; mt += 2) {
      uint32_t pa[4][4];
      load_q(qhead, mt * 16, T, g, q, pa);
      float y[4][4], rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { y[nt][0] = y[nt][1] = y[nt][2] = y[nt][3] = 0.f; }
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
        mma_bf16(rs, pa[kd], BF2_ONES, BF2_ONES);
#pragma unroll
        for (int np = 0; np < 2; ++np) {  // B fragments of two l n-tiles per ldmatrix.x4 from A^T[l][d]
          uint32_t b0, b1, b2, b3;
          ldsm_x4(ks_addr + swz(32 * lh + 16 * np + rr + ((mat >> 1) << 3), 2 * kd + (mat & 1)), b0, b1, b2, b3);
          mma_bf16(y[2 * np], pa[kd], b0, b1);
          mma_bf16(y[2 * np + 1], pa[kd], b2, b3);
        }
      }
      const int ra = mt * 16 + g, rb = ra + 8;
      // zero-filled Q rows beyond T have zero sums: keep their Y rows at 0
      const float r0 = ra >= T ? 0.f : rcp_approx(rs[0]), r1 = rb >= T ? 0.f : rcp_approx(rs[2]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 ya = fmul2(make_float2(y[nt][0], y[nt][1]), make_float2(r0, r0)), yb = fmul2(make_float2(y[nt][2], y[nt][3]), make_float2(r1, r1));
        *reinterpret_cast<uint32_t*>(Vs + swz(ra, 4 * lh + nt) + q * 4) = pack2(ya.x, ya.y);
        *reinterpret_cast<uint32_t*>(Vs + swz(rb, 4 * lh + nt) + q * 4) = pack2(yb.x, yb.y);
      }
    }
  }
  __syncthreads();  // this CTA's heads of Y are in smem

  // ---- 4. StylizationBlock prologue: LN(512) * (1 + scale) + shift, SiLU over this CTA's 128 columns.  A lane covers 8 columns
  //         of one row; 16 lanes make a row, a warp handles 2 rows per iteration.  The LayerNorm (sum, sum of squares) of a row
  //         is the only cross-CTA quantity: exchanged through distributed shared memory, summed in rank order.
  if constexpr (CL == 1) {
    // One CTA holds whole rows, no cluster: (a) statistics -- a warp per row, 16 columns per lane, no constants needed -- into
    // stat[row] = (mean, rstd); CTA barrier; (b) normalise half-rows: warp parity selects the column half (so the folded
    // constants are loaded once per lane), 8 columns per lane, 32 lanes x 8 columns = 256 columns per warp pass.
    for (int t = warp; t < T; t += NWARPS) {
      const int hh = lane >> 2, c0 = (lane & 3) * 2;
      const uint8_t* Yh = sm + hh * 2 * TILE_BYTES + TILE_BYTES;
      const uint4 u0 = *reinterpret_cast<const uint4*>(Yh + swz(t, c0));
      const uint4 u1 = *reinterpret_cast<const uint4*>(Yh + swz(t, c0 + 1));
      const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float2 v = bf_pair(w[e]); s2 = fadd2(s2, v); q2 = ffma2(v, v, q2); }
      float s = s2.x + s2.y, sq = q2.x + q2.y;
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o2); sq += __shfl_xor_sync(0xffffffffu, sq, o2); }
      if (lane == 0) {
        const float mean = s * (1.f / D);
        // y is O(1) (a convex combination of V rows), so E[x^2] - mean^2 is safe in fp32
        stat[t] = make_float2(mean, rsqrtf(fmaxf(sq * (1.f / D) - mean * mean, 0.f) + 1e-5f));
      }
    }
    __syncthreads();
    const int hsel = warp & 1;                              // column half of this warp
    const int hh = 4 * hsel + (lane >> 3), c0 = lane & 7;   // head, 16-byte chunk of that head's row
    const uint8_t* Yh = sm + hh * 2 * TILE_BYTES + TILE_BYTES;
    const int col0 = 256 * hsel + lane * 8;                 // global column
    const float* sc = ss + (size_t)(smp % B) * ss_ld;
    float2 G[4], Bc[4];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ln_g + col0 + e)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + col0 + e));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc + col0 + e)), d4 = __ldg(reinterpret_cast<const float4*>(sc + D + col0 + e));
      G[e / 2] = make_float2(0.5f * a.x * (1.f + c4.x), 0.5f * a.y * (1.f + c4.y));
      G[e / 2 + 1] = make_float2(0.5f * a.z * (1.f + c4.z), 0.5f * a.w * (1.f + c4.w));
      Bc[e / 2] = make_float2(0.5f * fmaf(b4.x, 1.f + c4.x, d4.x), 0.5f * fmaf(b4.y, 1.f + c4.y, d4.y));
      Bc[e / 2 + 1] = make_float2(0.5f * fmaf(b4.z, 1.f + c4.z, d4.z), 0.5f * fmaf(b4.w, 1.f + c4.w, d4.w));
    }
#pragma unroll 1
    for (int t = warp >> 1; t < T; t += NWARPS / 2) {
      const float2 st2 = stat[t];
      const float2 rs2 = make_float2(st2.y, st2.y), nm2 = make_float2(-st2.x * st2.y, -st2.x * st2.y);
      const uint4 u = *reinterpret_cast<const uint4*>(Yh + swz(t, c0));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // SiLU(x) = h + h*tanh(h), h = x/2 (exact identity; MUFU.TANH)
        const float2 h = ffma2(ffma2(bf_pair(w[e]), rs2, nm2), G[e], Bc[e]);
        const float2 r = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
        o[e] = pack2(r.x, r.y);
      }
      // the 32 lanes of a warp write 512 contiguous bytes of the row
      *reinterpret_cast<uint4*>(z + (row0 + t) * (size_t)D + col0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else {
    const int sub = lane & (LPR - 1), rsel = lane / LPR;
    const int hh = sub >> 3, c0 = sub & 7;                  // local head, 16-byte chunk of that head's row
    const uint8_t* Yh = sm + hh * 2 * TILE_BYTES + TILE_BYTES;
    const int col0 = (int)rank * COLS + sub * 8;           // global column
    const float* sc = ss + (size_t)(smp % B) * ss_ld;
    // per-column constants folded once:  h = t/2,  t = ((v-mean)*rstd*g + b)*(1+scale) + shift = (v-mean)*rstd*2G + 2Bc
    float2 G[4], Bc[4];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ln_g + col0 + e)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + col0 + e));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc + col0 + e)), d4 = __ldg(reinterpret_cast<const float4*>(sc + D + col0 + e));
      G[e / 2] = make_float2(0.5f * a.x * (1.f + c4.x), 0.5f * a.y * (1.f + c4.y));
      G[e / 2 + 1] = make_float2(0.5f * a.z * (1.f + c4.z), 0.5f * a.w * (1.f + c4.w));
      Bc[e / 2] = make_float2(0.5f * fmaf(b4.x, 1.f + c4.x, d4.x), 0.5f * fmaf(b4.y, 1.f + c4.y, d4.y));
      Bc[e / 2 + 1] = make_float2(0.5f * fmaf(b4.z, 1.f + c4.z, d4.z), 0.5f * fmaf(b4.w, 1.f + c4.w, d4.w));
    }
    // 4a. partial statistics of this CTA's 128 columns, reduced over the row's 16 lanes
    float ps[LN_ITERS], pq[LN_ITERS];
#pragma unroll
    for (int i = 0; i < LN_ITERS; ++i) {
      const int t = (i * NWARPS + warp) * RPI + rsel;
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      if (t < T) {
        const uint4 u = *reinterpret_cast<const uint4*>(Yh + swz(t, c0));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 v = bf_pair(w[e]); s2 = fadd2(s2, v); q2 = ffma2(v, v, q2); }
      }
      float s = s2.x + s2.y, sq = q2.x + q2.y;
#pragma unroll
      for (int o2 = LPR / 2; o2 > 0; o2 >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o2); sq += __shfl_xor_sync(0xffffffffu, sq, o2); }
      ps[i] = s; pq[i] = sq;
    }
    // 4b. publish into every CTA's table (slot = source rank), one cluster barrier
    prims::cluster_wait();   // every CTA of the cluster has started: its shared memory may be written
#pragma unroll
    for (int i = 0; i < LN_ITERS; ++i) {
      const int t = (i * NWARPS + warp) * RPI + rsel;
      if (t < T && sub < CL) {   // lane `sub` of the row serves CTA (rank + sub) % CL; sub == 0 is the local table
        const uint32_t dst = (rank + (uint32_t)sub) % CL;
        prims::st_peer_f32x2(smem_addr(stat + rank * TP + t), dst, ps[i], pq[i]);
      }
    }
    prims::cluster_barrier();   // all tables complete and visible (release / acquire at cluster scope); no remote access after this
    // 4c. total statistics in rank order (bit-identical in every CTA of the sample); rows re-read from smem, normalised, stored
#pragma unroll 1
    for (int i = 0; i < LN_ITERS; ++i) {
      const int t = (i * NWARPS + warp) * RPI + rsel;
      if (t < T) {
        float s = 0.f, sq = 0.f;
#pragma unroll
        for (int r = 0; r < CL; ++r) { const float2 st2 = stat[r * TP + t]; s += st2.x; sq += st2.y; }
        const float mean = s * (1.f / D);
        // y is O(1) (a convex combination of V rows), so E[x^2] - mean^2 is safe in fp32
        const float rstd = rsqrtf(fmaxf(sq * (1.f / D) - mean * mean, 0.f) + 1e-5f);
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
        const uint4 u = *reinterpret_cast<const uint4*>(Yh + swz(t, c0));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // SiLU(x) = h + h*tanh(h), h = x/2 (exact identity; MUFU.TANH)
          const float2 h = ffma2(ffma2(bf_pair(w[e]), rs2, nm2), G[e], Bc[e]);
          const float2 r = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
          o[e] = pack2(r.x, r.y);
        }
        // the 16 lanes of a row write 256 contiguous bytes
        *reinterpret_cast<uint4*>(z + (row0 + t) * (size_t)D + col0) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

#ifndef DSHEG_EMU
// host launcher: thread-block clusters of CL CTAs (one cluster per sample)
template <int CL>
inline cudaError_t launch_attn_v6(const bf16* qkv, bf16* z, int n_samples, int T, int ssB, const float* ln_g, const float* ln_b,
                                  const float* ss, int ss_ld, cudaStream_t st) {
  using C = Cfg<CL>;
  auto kern = attn_v6_kernel<CL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_samples * CL); cfg.blockDim = dim3(C::NTHREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
#if DSHEG_PDL_ATTRS
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
#endif
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, qkv, z, T, ssB, ln_g, ln_b, ss, ss_ld);
}
#endif  // DSHEG_EMU

}  // namespace av6
}  // namespace dsheg
