// Linear-attention core + StylizationBlock prologue, bf16 (D = 512, 8 heads of 64, T <= 96) -- PERSISTENT, TMA-STAGED, WARP-SPECIALISED.
//
// Mathematics (reference transformer.py:112-130 + :86-97), input contract = the ACT_EXPO epilogue of the QKV GEMM (gemm_tc.cuh):
// the Q and K columns hold exp(value - static shift) (softmax is shift-invariant; pack.py:expo_shift proves the range), V is plain:
//   A = K'^T V / colsum(K')  [64 x 64 per head]     Y = Q' A / rowsum(Q')     z = SiLU(LN_512(Y) * (1 + scale) + shift)
//
// Why this shape.  Round 2's first TMA kernel (attn_tma, deleted; profiles/r02/ci1, call9) had all 16 compute warps of the one CTA an SM
// can hold walk through the same phases together -- A^T, Y, a CTA-wide barrier, the LayerNorm pass -- so the tensor pipe idled during
// LayerNorm, the FP32 / MUFU pipes idled during the products, and every dependency stall was exposed at 4 warps per scheduler (ncu:
// issue slots 34 % busy, 0.42 of the copy peak in the sampling loop).  Two samples cannot be resident (Y of one sample is 96 KB of
// shared memory).  Here the phases belong to DIFFERENT warps working on DIFFERENT samples, and Y never exists in shared memory:
//   * 8 "A warps" turn the K' / V tiles of one head after the other (TMA ring, two heads deep; the refill duty rotates over the A
//     warps; the same head of the next sample is prefetched into L2) into the normalised bf16 A^T of that head -- a 32 x 16 tile per
//     warp; the column sums of K' come out of the same fragments (ones . K') in exactly the accumulator layout, so nothing is
//     exchanged between the warps.  The tile goes straight into the head's shared-memory slot (8 x 8 KB) when both readers of the
//     previous sample have released it (mbarrier pair per head; the test is a warp vote: both branches hold warp collectives), and
//     is otherwise PARKED in tensor memory (8 words per thread) and copied in a second phase: nothing in phase 1 waits for the Y
//     warps, so the A warps can run a whole sample ahead.
//     (History, profiles/r02/call10 .. call15: v1 accumulated at most ONE head ahead of the release and was serial with the Y warps,
//     2.36 TB/s; v2 had 4 A warps with 32 x 32 quadrants -- one warp per SM sub-partition issuing 620 instructions per head at 0.17
//     IPC was the critical path; half tiles refilled by the last-arriving warp through a shared-memory atomic lost to this version.)
//   * 16 "Y warps" = 8 heads x 2 row halves.  Warp (h, half) owns head h of up to three 16-frame tiles: its Q' box (48 frames x 64
//     columns, 6 KB) arrives by a TMA load the warp issues ITSELF for the next sample the moment its last product has consumed the
//     current one, so the reload has the whole LayerNorm part to land.  Per tile: Y = Q' A on mma.sync from ldmatrix fragments (A^T
//     rows are stored permuted so that a thread's 16 output columns are two contiguous runs of 8), row sums on the tensor core, and
//     the per-row LayerNorm partials (sum, sum of squares over the head's 64 columns) straight from the fp32 accumulators into a
//     [row][head] table.  Finished tiles are PARKED IN TENSOR MEMORY (tcgen05.st, thread-private columns: TMEM as a 192 KB register
//     spill space for mma.sync warps; micro-benchmark scripts/ubench), which frees the A^T slot after the last product instead of
//     after the LayerNorm part and keeps every warp within 80 registers (24 warps per SM).
//   * one named barrier per row half and sample publishes the partials; then every warp normalises, modulates and applies SiLU to
//     its own tiles (fp32 Y back from tensor memory, 16 columns at a time, never rounded to bf16 before the LayerNorm) and writes z
//     with 16-byte stores: a store instruction covers 64 contiguous bytes of 8 rows.  No CTA-wide barrier, no LayerNorm pass over
//     shared memory, no shuffles beyond one quad reduction per tile.
// Measured (profiles/r02/ci3): 3.23 TB/s in the sampling loop at 1 492 MHz = 0.49 of the copy peak, 168 - 174 us = 0.60 - 0.62 alone
// under ncu (attn_tma: 2.7 / 207 us); 35 k warp-instructions per sample, issue slots 40 % busy; the A warps (223 instructions and
// 2.9 k cycles per head, 19 % of it waiting for K' / V tiles) are the critical path.  `rev` walks the CTA's samples last-to-first
// (engine.cu: alternating row walk).  HBM traffic: read q', k', v + write z = 4 * T * 512 * 2 bytes per sample.
#pragma once
#include <type_traits>

#include "attn_v3.cuh"
#include "gemm_tc.cuh"   // tc:: mbarrier / TMA primitives, tensor-map encoder entry point

namespace dsheg {
namespace aws {

using av3::TP; using av3::HD; using av3::D; using av3::TILE_BYTES;
using av3::pack2; using av3::swz;
using prims::ldsm_x4; using prims::ldsm_x4_trans; using prims::mma_bf16; using prims::tanh_approx; using prims::rcp_approx;
using prims::ffma2; using prims::fadd2; using prims::fmul2;
constexpr uint32_t BF2_ONES = 0x3F803F80u;      // bf16x2 (1.0, 1.0)

constexpr int NH = 8;                          // heads
constexpr int NYW = 2 * NH;                    // Y warps: head = warp & 7, row half = warp >> 3
constexpr int NAW = 8;                         // A warps (the LAST warps of the CTA): l-half = wq & 1, 16-column d-slice = wq >> 1
constexpr int NTHREADS = 32 * (NYW + NAW);     // 6 warps per SM sub-partition: 80 registers each
constexpr int MH = TP / 32;                    // 16-frame tiles per row half (3)
constexpr int QBOX_BYTES = MH * 16 * 128;      // one Y warp's Q' box: 48 frames x 64 columns
constexpr int A_BYTES = HD * 128;              // one head's A^T (64 x 64 bf16)
constexpr int NST = 4;                         // K' / V ring slots = two heads
constexpr int Q_OFF = 0;                                      // [NYW] Q' boxes
constexpr int A_OFF = Q_OFF + NYW * QBOX_BYTES;               // [NH] A^T slots
constexpr int RING_OFF = A_OFF + NH * A_BYTES;                // [NST] K' / V tiles
constexpr int STAT_OFF = RING_OFF + NST * TILE_BYTES;         // [sample parity][half][tile][row 16][head 8] float2 (sum, sumsq)
constexpr int STAT_BYTES = 2 * 2 * MH * 16 * NH * 8;
constexpr int BAR_OFF = STAT_OFF + STAT_BYTES;                // kv_full[NST] kv_empty[NST / 2] a_full[NH] a_empty[NH] q_full[NYW] (mbarriers)
constexpr int NBAR = NST + NST / 2 + 2 * NH + NYW;
constexpr int TMEM_SLOT_OFF = BAR_OFF + NBAR * 8;
constexpr int SMEM_BYTES = ((TMEM_SLOT_OFF + 4 + 127) / 128) * 128;
constexpr int TMEM_COLS = 512;                 // per lane quadrant: 4 Y warps x MH parked tiles x 32 columns, then 2 A warps x 8 heads x 8 columns
constexpr int YPARK_COLS = MH * 32;
constexpr int APARK_COL = (NYW / 4) * YPARK_COLS;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
static_assert(Q_OFF % 1024 == 0 && QBOX_BYTES % 1024 == 0 && A_OFF % 1024 == 0 && A_BYTES % 1024 == 0 && RING_OFF % 1024 == 0 && TILE_BYTES % 1024 == 0,
              "SWIZZLE_128B tiles start on 1024-byte boundaries");
static_assert(APARK_COL + (NAW / 4) * NH * 8 <= TMEM_COLS, "parking space");

// mbarrier wait that SLEEPS (suspend-time hint) instead of spinning: a waiting warp must not take issue slots from the working ones.
// Watchdog like tc::mbar_wait: a lost arrive becomes a launch error after 4 s, not a hung GPU.
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  if (tc::mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  while (!tc::mbar_try_wait_hint(bar, parity, 20000u)) {
    const uint64_t now = tc::globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) tc::trap();
  }
}

// Who arrives on the A^T hand-over barriers: EVERY lane (default).  One arrive by lane 0 after __syncwarp is sufficient under the PTX
// memory model (release is cumulative over the stores the warp barrier ordered before it) and was the shipped form until the last
// hardware session, but compute-sanitizer racecheck tracks happens-before per thread and reported the A^T hand-over as a RAW hazard
// (profiles/r02/ci3); with every storing / loading lane arriving itself the same run reports 0 hazards (profiles/r02/diag).
// -DDSHEG_AWS_ALL_LANES_ARRIVE=0 restores the single arrive.
#if !defined(DSHEG_AWS_ALL_LANES_ARRIVE) || DSHEG_AWS_ALL_LANES_ARRIVE
constexpr int ARRIVALS_PER_WARP = 32;
__device__ __forceinline__ void warp_arrive(uint32_t bar, int) { tc::mbar_arrive(bar); }
#else
constexpr int ARRIVALS_PER_WARP = 1;
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) { if (lane == 0) tc::mbar_arrive(bar); }
#endif

__device__ __forceinline__ void half_sync(int half) { prims::named_bar_sync<32 * NH>(1 + half); }     // ids 1, 2: the 8 Y warps of a row half

// shared-memory row of A^T that holds output column l of the head: n-tile nt = 4 (l >> 5) + ((l >> 1) & 3), row 2 ((l >> 3) & 3) + (l & 1)
// of the tile -- thread q of an mma quad then owns l = 32 a + 8 q + {0 .. 7}, a = 0, 1: two 16-byte runs of the output row.
__device__ __forceinline__ int a_row(int l) { return 8 * (4 * (l >> 5) + ((l >> 1) & 3)) + 2 * ((l >> 3) & 3) + (l & 1); }

// Self-attention: tmQ and tmKV both view the fused qkv tensor [rows, 1536] (boxes of 16 mh and 16 n_kt frames), kcol = 512, vcol = 1024.
// Cross-attention (transformer.py:133-166): tmQ views q' [n_samples * T, 512], tmKV views [n_samples * Tkv, 1024] (k' | v), kcol = 0, vcol = 512.
__global__ void __launch_bounds__(NTHREADS, 1)
attn_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int kcol, int vcol,
               bf16* __restrict__ z, int n_samples, int T, int Tkv, int B,
               const float* __restrict__ ln_g, const float* __restrict__ ln_b, const float* __restrict__ ss, int ss_ld, int rev) {
  DSHEG_PDL_ENTER();
  DSHEG_TC_DYN_SMEM(sm);
  const uint32_t sbase = tc::smem_u32(sm);
  if (sbase & 1023u) tc::trap();   // SWIZZLE_128B tiles: the dynamic smem base must be 1024-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mt = (T + 15) >> 4;            // 16-frame tiles of Q' / Y that contain valid frames
  const int n_kt = (Tkv + 15) >> 4;          // ... of K' / V
  const int mh = (n_mt + 1) >> 1;            // tiles of the first row half = height of a Q' box in tiles
  const uint32_t q_tx = (uint32_t)mh * 16u * 128u;      // bytes one TMA box delivers (frames beyond T arrive as zeros)
  const uint32_t kv_tx = (uint32_t)n_kt * 16u * 128u;
  auto kv_full = [&](int s) { return sbase + BAR_OFF + 8u * s; };
  auto kv_empty = [&](int p) { return sbase + BAR_OFF + 8u * (NST + p); };
  auto a_full = [&](int h) { return sbase + BAR_OFF + 8u * (NST + NST / 2 + h); };
  auto a_empty = [&](int h) { return sbase + BAR_OFF + 8u * (NST + NST / 2 + NH + h); };
  auto q_full = [&](int w) { return sbase + BAR_OFF + 8u * (NST + NST / 2 + 2 * NH + w); };
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) tc::mbar_init(kv_full(s), 1);
    for (int p2 = 0; p2 < NST / 2; ++p2) tc::mbar_init(kv_empty(p2), NAW);
    for (int h = 0; h < NH; ++h) { tc::mbar_init(a_full(h), NAW * ARRIVALS_PER_WARP); tc::mbar_init(a_empty(h), 2 * ARRIVALS_PER_WARP); }
    for (int w = 0; w < NYW; ++w) tc::mbar_init(q_full(w), 1);
    tc::fence_mbarrier_init();
  }
  if (warp == NYW) tc::tmem_alloc<1, TMEM_COLS>(sbase + TMEM_SLOT_OFF);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const volatile uint32_t*>(sm + TMEM_SLOT_OFF);
  const int n_iter = ((int)blockIdx.x < n_samples) ? (n_samples - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // samples of this CTA
  // j-th sample of this CTA; rev: last samples first (start where the producer of q' / k' / v finished: those rows are still in L2)
  auto smp_at = [&](int j) { const int s = (int)blockIdx.x + j * (int)gridDim.x; return rev ? n_samples - 1 - s : s; };
  const int g = lane >> 2, q = lane & 3;           // mma fragment coordinates
  const int mat = lane >> 3, rr = lane & 7;        // ldmatrix: matrix index / row inside the matrix

  if (warp < NYW) {
    // ========================================================== Y warps ==========================================================
    const int h = warp & 7, half = warp >> 3;
    const int mt0 = half * mh;                                   // first tile of this warp
    const int nu = n_mt - mt0 < mh ? (n_mt - mt0 > 0 ? n_mt - mt0 : 0) : mh;   // tiles of this warp (0 .. 3; the same for the 8 warps of a half)
    const uint32_t qb_addr = sbase + Q_OFF + warp * QBOX_BYTES, as_addr = sbase + A_OFF + h * A_BYTES;
    const uint32_t park = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * YPARK_COLS);
    if (nu > 0 && lane == 0 && n_iter > 0) {
      tc::prefetch_tensormap(&tmQ);
      tc::mbar_arrive_expect_tx(q_full(warp), q_tx);
      tc::tma_load_3d_hint(&tmQ, q_full(warp), qb_addr, h * HD, mt0 * 16, smp_at(0), tc::L2_EVICT_FIRST);
    }
    const int colA = h * HD + 8 * q;     // this thread's output columns: [colA, colA + 8) and [colA + 32, colA + 40)
#pragma unroll 1
    for (int i = 0; i < n_iter; ++i) {
      const int smp = smp_at(i);
      const uint32_t par = (uint32_t)(i & 1);
      float2* const stat = reinterpret_cast<float2*>(sm + STAT_OFF) + (size_t)((par * 2 + half) * MH) * 16 * NH;
      const float* const sc = ss + (size_t)(smp % B) * ss_ld + colA;   // the sample's modulation row (scale | shift) at this thread's columns
      if (nu > 0) {
        prims::prefetch_l1(sc); prims::prefetch_l1(sc + 32); prims::prefetch_l1(sc + D); prims::prefetch_l1(sc + D + 32);   // read after the products
        wait_bar(q_full(warp), par);
      }
      wait_bar(a_full(h), par);
#pragma unroll 1
      for (int u = 0; u < nu; ++u) {
        // ---- Y[t][l] = Q'[t][:] . A: the 16 x 64 tile of (tile u, head h); row sums of Q' on the tensor core (Q' . ones)
        uint32_t pa[4][4];
#pragma unroll
        for (int kd = 0; kd < 4; ++kd)   // A fragments: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15) of the 16 x 16 block
          ldsm_x4(qb_addr + swz(u * 16 + rr + ((mat & 1) << 3), 2 * kd + (mat >> 1)), pa[kd][0], pa[kd][1], pa[kd][2], pa[kd][3]);
        float y[8][4], rs[4] = {0.f, 0.f, 0.f, 0.f};   // rs: rows g, g + 8 in [0], [2]
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { y[nt][0] = y[nt][1] = y[nt][2] = y[nt][3] = 0.f; }
#pragma unroll
        for (int kd = 0; kd < 4; ++kd) {
          mma_bf16(rs, pa[kd], BF2_ONES, BF2_ONES);
#pragma unroll
          for (int np = 0; np < 4; ++np) {   // B fragments of two n-tiles per ldmatrix.x4 from the (row-permuted) A^T[l][d]
            uint32_t b0, b1, b2, b3;
            ldsm_x4(as_addr + swz(16 * np + rr + ((mat >> 1) << 3), 2 * kd + (mat & 1)), b0, b1, b2, b3);
            mma_bf16(y[2 * np], pa[kd], b0, b1);
            mma_bf16(y[2 * np + 1], pa[kd], b2, b3);
          }
        }
        if (u == nu - 1) {
          // the last product of this sample has consumed every Q' and A^T fragment (the mma results depend on them): hand the A^T
          // slot back to the A warps and fetch the next sample's Q' box into this warp's own (now dead, only ever read) box
          __syncwarp();
          warp_arrive(a_empty(h), lane);
          if (lane == 0) {
            if (i + 1 < n_iter) {
              tc::mbar_arrive_expect_tx(q_full(warp), q_tx);
              tc::tma_load_3d_hint(&tmQ, q_full(warp), qb_addr, h * HD, mt0 * 16, smp_at(i + 1), tc::L2_EVICT_FIRST);
            }
          }
          __syncwarp();
        }
        const int ra = (mt0 + u) * 16 + g, rb = ra + 8;
        // zero-filled Q' rows beyond T have zero sums: keep their Y rows at 0
        const float r0 = ra >= T ? 0.f : rcp_approx(rs[0]), r1 = rb >= T ? 0.f : rcp_approx(rs[2]);
        const float2 r02 = make_float2(r0, r0), r12 = make_float2(r1, r1);
        float2 sa = make_float2(0.f, 0.f), sb = sa, qa = sa, qb = sa;
        uint32_t pk[32];   // the tile as it is parked: column 4 nt + j of the warp's parking row = y[nt][j]
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float2 ya = fmul2(make_float2(y[nt][0], y[nt][1]), r02), yb = fmul2(make_float2(y[nt][2], y[nt][3]), r12);
          pk[4 * nt] = __float_as_uint(ya.x); pk[4 * nt + 1] = __float_as_uint(ya.y);
          pk[4 * nt + 2] = __float_as_uint(yb.x); pk[4 * nt + 3] = __float_as_uint(yb.y);
          sa = fadd2(sa, ya); qa = ffma2(ya, ya, qa);
          sb = fadd2(sb, yb); qb = ffma2(yb, yb, qb);
        }
        tc::tmem_st32(park + (uint32_t)(u * 32), pk);   // parked in tensor memory until the LayerNorm part
        float s0 = sa.x + sa.y, q0 = qa.x + qa.y, s1 = sb.x + sb.y, q1 = qb.x + qb.y;
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {   // the 4 lanes of a quad hold the 64 columns of rows g / g + 8
          s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        if (q == 0) {
          stat[(u * 16 + g) * NH + h] = make_float2(s0, q0);
          stat[(u * 16 + g + 8) * NH + h] = make_float2(s1, q1);
        }
      }
      if (nu == 0) {   // a row half without frames (T <= 16 ...): keep the A^T hand-shake in step
        __syncwarp();
        warp_arrive(a_empty(h), lane);
        __syncwarp();
        continue;
      }
      float2 G[4], Bc[4];
      float4 raw[4];   // the sample's (scale | shift) values of one run: in flight well before they are folded
      auto load_raw = [&](int a) {
        raw[0] = __ldg(reinterpret_cast<const float4*>(sc + 32 * a)); raw[1] = __ldg(reinterpret_cast<const float4*>(sc + 32 * a + 4));
        raw[2] = __ldg(reinterpret_cast<const float4*>(sc + D + 32 * a)); raw[3] = __ldg(reinterpret_cast<const float4*>(sc + D + 32 * a + 4));
      };
      auto load_consts = [&](int a) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 s4 = raw[e], t4 = raw[2 + e];
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(ln_g + colA + 32 * a + 4 * e)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + colA + 32 * a + 4 * e));
          const float sx = 1.f + s4.x, sy = 1.f + s4.y, sz = 1.f + s4.z, sw = 1.f + s4.w;
          G[2 * e] = make_float2(0.5f * g4.x * sx, 0.5f * g4.y * sy);
          G[2 * e + 1] = make_float2(0.5f * g4.z * sz, 0.5f * g4.w * sw);
          Bc[2 * e] = make_float2(0.5f * fmaf(b4.x, sx, t4.x), 0.5f * fmaf(b4.y, sy, t4.y));
          Bc[2 * e + 1] = make_float2(0.5f * fmaf(b4.z, sz, t4.z), 0.5f * fmaf(b4.w, sw, t4.w));
        }
      };
      load_raw(0);       // L2 round trip under the barrier wait
      half_sync(half);   // the partials of all 8 heads of this half's rows are in the table
      // ---- LayerNorm + modulation + SiLU:  t = ((v - mean) rstd g + b)(1 + scale) + shift = 2 ((v - mean) rstd G + Bc),  SiLU(t) = h + h tanh(h), h = t / 2
      float2 rn[MH][2];   // per tile and row (g, g + 8): (rstd, -mean rstd)
#pragma unroll
      for (int u = 0; u < MH; ++u)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          rn[u][r] = make_float2(0.f, 0.f);
          if (u < nu) {
            const float4* sp = reinterpret_cast<const float4*>(stat + (u * 16 + g + 8 * r) * NH);
            const float4 p0 = sp[0], p1 = sp[1], p2 = sp[2], p3 = sp[3];   // heads (0,1) (2,3) (4,5) (6,7): (sum, sumsq) pairs
            const float s = ((p0.x + p0.z) + (p1.x + p1.z)) + ((p2.x + p2.z) + (p3.x + p3.z));
            const float sq = ((p0.y + p0.w) + (p1.y + p1.w)) + ((p2.y + p2.w) + (p3.y + p3.w));
            const float mean = s * (1.f / D);
            // y is O(1) (a convex combination of V rows), so E[x^2] - mean^2 is safe in fp32
            const float rstd = rsqrtf(fmaxf(sq * (1.f / D) - mean * mean, 0.f) + 1e-5f);
            rn[u][r] = make_float2(rstd, -mean * rstd);
          }
        }
      const size_t row0 = (size_t)smp * T;
      // steps st = 0 .. 2 nu - 1: (run a, tile u) = (st / nu, st % nu), the thread's run a = columns [colA + 32 a, + 8) = n-tiles 4 a .. 4 a + 3 of
      // tile u = 16 parked columns; a step's load is in flight while the previous step is processed.
      // steps st = 0 .. 2 NU - 1, fully unrolled for the warp's tile count NU (so every register index is a compile-time constant)
      auto ln_part = [&](auto nu_c) {
        constexpr int NU = decltype(nu_c)::value;
        uint32_t w[2][16];
        load_consts(0);
        tc::tmem_ld16_issue(park, w[0]);
#pragma unroll
        for (int st = 0; st < 2 * NU; ++st) {
          const int a = st / NU, u = st % NU;
          tc::tmem_wait_ld();                                    // w[st & 1] = step st
          if (st + 1 < 2 * NU) tc::tmem_ld16_issue(park + (uint32_t)(((st + 1) % NU) * 32 + ((st + 1) / NU) * 16), w[(st + 1) & 1]);
          if (st == NU) { load_raw(1); load_consts(1); }   // (loading run 1's values a step early costs 360 bytes of spills at 80 registers)
          const int ra = (mt0 + u) * 16 + g;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float2 rs2 = make_float2(rn[u][r].x, rn[u][r].x), nm2 = make_float2(rn[u][r].y, rn[u][r].y);
            uint32_t o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float2 v = make_float2(__uint_as_float(w[st & 1][4 * c + 2 * r]), __uint_as_float(w[st & 1][4 * c + 2 * r + 1]));
              const float2 hv = ffma2(ffma2(v, rs2, nm2), G[c], Bc[c]);
              const float2 sv = ffma2(hv, make_float2(tanh_approx(hv.x), tanh_approx(hv.y)), hv);
              o[c] = pack2(sv.x, sv.y);
            }
            const int t = ra + 8 * r;
            if (t < T) *reinterpret_cast<uint4*>(z + (row0 + t) * (size_t)D + colA + 32 * a) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      };
      static_assert(MH == 3, "the dispatch below spells out 1 .. 3 tiles");
      if (nu == 3) ln_part(std::integral_constant<int, 3>{});
      else if (nu == 2) ln_part(std::integral_constant<int, 2>{});
      else ln_part(std::integral_constant<int, 1>{});
    }
  } else {
    // ========================================================== A warps ==========================================================
    const int wq = warp - NYW;
    const int lq = wq & 1, d8 = wq >> 1;      // this warp's 32 x 16 tile of A^T[l][d]: l-half lq, d = 16 d8 .. + 15
    const uint32_t apark = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(APARK_COL + (wq >> 2) * NH * 8);
    const uint32_t ones[4] = {BF2_ONES, BF2_ONES, BF2_ONES, BF2_ONES};
    // fragment addresses of k-step 0 (k-step kt adds kt * 2048 bytes: 16 rows of 128 bytes, the XOR pattern repeats every 8 rows)
    const int vr = rr + ((mat >> 1) << 3), kr = rr + ((mat & 1) << 3);
    const uint32_t v_off0 = swz(vr, 4 * lq + (mat & 1)), v_off1 = swz(vr, 4 * lq + 2 + (mat & 1));
    const uint32_t k_off = swz(kr, 2 * d8 + (mat >> 1));
    // byte offsets of this thread's 8 words inside a head's A^T slot: [nt][mi][rows g | g + 8] (row-permuted, swizzled)
    uint32_t a_off[8];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int l = 32 * lq + 16 * mi + g;
        a_off[4 * nt + 2 * mi] = swz(a_row(l), 2 * d8 + nt) + q * 4;
        a_off[4 * nt + 2 * mi + 1] = swz(a_row(l + 8), 2 * d8 + nt) + q * 4;
      }
    const int n_heads = n_iter * NH;          // (sample, head) units of this CTA, in order
    auto issue = [&](int hc) {                // one lane: K' and V tiles of unit hc -> ring slots 2 (hc & 1), + 1
      const int smp = smp_at(hc >> 3), hh = hc & 7, s0 = 2 * (hc & 1);
      tc::mbar_arrive_expect_tx(kv_full(s0), kv_tx);
      tc::tma_load_3d_hint(&tmKV, kv_full(s0), sbase + RING_OFF + s0 * TILE_BYTES, kcol + hh * HD, 0, smp, tc::L2_EVICT_FIRST);
      tc::mbar_arrive_expect_tx(kv_full(s0 + 1), kv_tx);
      tc::tma_load_3d_hint(&tmKV, kv_full(s0 + 1), sbase + RING_OFF + (s0 + 1) * TILE_BYTES, vcol + hh * HD, 0, smp, tc::L2_EVICT_FIRST);
      // pull the same head of the NEXT sample from HBM into L2 (the ring is only two units deep: it then covers L2 latency, not HBM latency)
      if ((hc >> 3) + 1 < n_iter) {
        tc::tma_prefetch_l2_3d(&tmKV, kcol + hh * HD, 0, smp_at((hc >> 3) + 1));
        tc::tma_prefetch_l2_3d(&tmKV, vcol + hh * HD, 0, smp_at((hc >> 3) + 1));
      }
    };
    if (wq == 0 && lane == 0) {
      tc::prefetch_tensormap(&tmKV);
      for (int hc = 0; hc < 2 && hc < n_heads; ++hc) issue(hc);
    }
#pragma unroll 1
    for (int i = 0; i < n_iter; ++i) {
      // ---- phase 1: A^T of the sample's 8 heads, normalised and packed to bf16 (independent of the Y warps)
      uint32_t parked = 0;   // heads of this sample that wait in tensor memory for their slot
#pragma unroll 1
      for (int hh = 0; hh < NH; ++hh) {
        const int hc = i * NH + hh, s0 = 2 * (hc & 1);
        const uint32_t par = (uint32_t)((hc >> 1) & 1);
        const uint32_t ks_addr = sbase + RING_OFF + s0 * TILE_BYTES, vs_addr = ks_addr + TILE_BYTES;
        wait_bar(kv_full(s0), par);
        wait_bar(kv_full(s0 + 1), par);
        // A^T[l][d] = sum_t V[t][l] K'[t][d]; cs = ones . K' = the column sums of K' in the layout of the accumulator columns
        float acc[2][2][4], cs[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          cs[nt][0] = cs[nt][1] = cs[nt][2] = cs[nt][3] = 0.f;
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) { acc[mi][nt][0] = acc[mi][nt][1] = acc[mi][nt][2] = acc[mi][nt][3] = 0.f; }
        }
        // two fragment sets in ping-pong (no register copies): a k-step's fragments are loaded under the products of the step before
        uint32_t fa0[2][4], fb0[4], fa1[2][4], fb1[4];
        auto load_frags = [&](uint32_t (&fa)[2][4], uint32_t (&fb)[4], int kt) {
          const uint32_t o = (uint32_t)kt * 2048u;
          ldsm_x4_trans(vs_addr + v_off0 + o, fa[0][0], fa[0][1], fa[0][2], fa[0][3]);
          ldsm_x4_trans(vs_addr + v_off1 + o, fa[1][0], fa[1][1], fa[1][2], fa[1][3]);
          ldsm_x4_trans(ks_addr + k_off + o, fb[0], fb[1], fb[2], fb[3]);
        };
        auto products = [&](const uint32_t (&fa)[2][4], const uint32_t (&fb)[4]) {
          mma_bf16(acc[0][0], fa[0], fb[0], fb[1]);
          mma_bf16(acc[0][1], fa[0], fb[2], fb[3]);
          mma_bf16(acc[1][0], fa[1], fb[0], fb[1]);
          mma_bf16(acc[1][1], fa[1], fb[2], fb[3]);
          mma_bf16(cs[0], ones, fb[0], fb[1]);
          mma_bf16(cs[1], ones, fb[2], fb[3]);
        };
        load_frags(fa0, fb0, 0);
        if (n_kt == TP / 16) {   // the full-length window (T = 81 .. 96): six k-steps, fully unrolled
#pragma unroll
          for (int kt = 0; kt < TP / 16; kt += 2) {
            load_frags(fa1, fb1, kt + 1);
            products(fa0, fb0);
            if (kt + 2 < TP / 16) load_frags(fa0, fb0, kt + 2);
            products(fa1, fb1);
          }
        } else {
#pragma unroll 1
          for (int kt = 0; kt < n_kt; kt += 2) {   // 16 frames per k-step, two k-steps per iteration
            const bool two = kt + 1 < n_kt;        // warp-uniform
            if (two) load_frags(fa1, fb1, kt + 1);
            products(fa0, fb0);
            if (two) {
              if (kt + 2 < n_kt) load_frags(fa0, fb0, kt + 2);
              products(fa1, fb1);
            }
          }
        }
        // every fragment of this unit's K' / V tiles has been consumed by a product: the two ring slots go back to the refill duty
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(kv_empty(hc & 1));
        uint32_t pk[8];   // [nt][mi][rows g | g + 8]: bf16 pairs of columns d = 16 d8 + 8 nt + 2 q, + 1
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float2 inv = make_float2(rcp_approx(cs[nt][0]), rcp_approx(cs[nt][1]));
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const float2 lo = fmul2(make_float2(acc[mi][nt][0], acc[mi][nt][1]), inv), hi = fmul2(make_float2(acc[mi][nt][2], acc[mi][nt][3]), inv);
            pk[4 * nt + 2 * mi] = pack2(lo.x, lo.y);
            pk[4 * nt + 2 * mi + 1] = pack2(hi.x, hi.y);
          }
        }
        // the head's slot is normally free by now (both readers of the previous sample are done with it): write it directly.  Only a
        // head that finishes EARLY (the A warps running ahead of the Y warps) is parked in tensor memory and copied in phase 2.
        uint8_t* const as_ptr = sm + A_OFF + hh * A_BYTES;
        if (__all_sync(0xffffffffu, (int)tc::mbar_try_wait(a_empty(hh), (uint32_t)((i & 1) ^ 1)))) {   // warp-uniform: both branches hold warp collectives
#pragma unroll
          for (int e = 0; e < 8; ++e) *reinterpret_cast<uint32_t*>(as_ptr + a_off[e]) = pk[e];
          __syncwarp();
          warp_arrive(a_full(hh), lane);   // release: this warp's tile of A^T is written
        } else {
          tc::tmem_st8(apark + (uint32_t)(hh * 8), pk);
          parked |= 1u << hh;
        }
        // refill duty rotates over the A warps: by now the others have normally arrived, so the wait is short
        if (wq == (hc & 7) && lane == 0 && hc + 2 < n_heads) {
          wait_bar(kv_empty(hc & 1), par);
          issue(hc + 2);
        }
        __syncwarp();
      }
      // ---- phase 2: parked heads -> their A^T slots, each as soon as both readers of the previous sample have released it
#pragma unroll 1
      for (int hh = 0; hh < NH; ++hh) {
        if (!((parked >> hh) & 1u)) continue;   // warp-uniform
        uint32_t pk[8];
        tc::tmem_ld8(apark + (uint32_t)(hh * 8), pk);
        wait_bar(a_empty(hh), (uint32_t)((i & 1) ^ 1));
        uint8_t* const as_ptr = sm + A_OFF + hh * A_BYTES;
#pragma unroll
        for (int e = 0; e < 8; ++e) *reinterpret_cast<uint32_t*>(as_ptr + a_off[e]) = pk[e];
        __syncwarp();
        warp_arrive(a_full(hh), lane);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == NYW) tc::tmem_dealloc<1, TMEM_COLS>(tmem_base);
}

#ifndef DSHEG_EMU
// (column, frame, sample) view of a bf16 tensor [n_samples * T, cols]; box = 64 columns x Tpad frames of one sample.
// box_frames: height of the box in frames (default: all Tpad frames of the sample; Q' is loaded by row halves).
inline bool make_frames_tmap(CUtensorMap* map, const void* base, int cols, int n_samples, int T, std::string* err, int box_frames = -1) {
  struct Key { const void* p; int c, n, t, b; bool operator==(const Key& o) const { return p == o.p && c == o.c && n == o.n && t == o.t && b == o.b; } };
  struct Hash { size_t operator()(const Key& k) const { return reinterpret_cast<size_t>(k.p) ^ ((size_t)k.n * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.t << 48) ^ ((size_t)k.c << 32) ^ ((size_t)k.b << 56); } };
  static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;
  const Key k{base, cols, n_samples, T, box_frames};
  auto it = cache.find(k);
  if (it != cache.end()) { *map = it->second; return true; }
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (cols % 8)) { *err = "attention operand not 16-byte aligned"; return false; }
  const int Tpad = box_frames > 0 ? box_frames : (T + 15) / 16 * 16;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)n_samples};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)T * cols * 2};
  cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)Tpad, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled (3-D frames view) failed, CUresult " + std::to_string((int)r); return false; }
  if (cache.size() > 1024) cache.clear();
  cache.emplace(k, *map);
  return true;
}

inline cudaError_t launch_attn_ws_qkv(const CUtensorMap& mq, const CUtensorMap& mkv, int kcol, int vcol, bf16* z, int n_samples, int T, int Tkv,
                                      int ssB, const float* ln_g, const float* ln_b, const float* ss, int ss_ld, int num_sms, cudaStream_t st, int rev = 0) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = n_samples < num_sms ? n_samples : num_sms;
  DSHEG_LAUNCH(attn_ws_kernel, grid, NTHREADS, SMEM_BYTES, st, mq, mkv, kcol, vcol, z, n_samples, T, Tkv, ssB, ln_g, ln_b, ss, ss_ld, rev);
  return cudaGetLastError();
}

inline int q_box_frames(int T) { return 16 * ((((T + 15) >> 4) + 1) >> 1); }

// self-attention on the fused projection qkv [n_samples * T, 1536] (q' | k' | v)
inline cudaError_t launch_attn_ws(const bf16* qkv, bf16* z, int n_samples, int T, int ssB, const float* ln_g, const float* ln_b,
                                  const float* ss, int ss_ld, int num_sms, cudaStream_t st, std::string* err, int rev = 0) {
  CUtensorMap mq, mkv;
  if (!make_frames_tmap(&mq, qkv, 3 * D, n_samples, T, err, q_box_frames(T)) || !make_frames_tmap(&mkv, qkv, 3 * D, n_samples, T, err))
    return cudaErrorInvalidValue;
  return launch_attn_ws_qkv(mq, mkv, D, 2 * D, z, n_samples, T, T, ssB, ln_g, ln_b, ss, ss_ld, num_sms, st, rev);
}

// cross-attention (transformer.py:133-166): q' [n_samples * T, 512] from the motion stream, kv [n_samples * Tkv, 1024] (k' | v) from the conditioning
inline cudaError_t launch_cross_attn_ws(const bf16* q, const bf16* kv, bf16* z, int n_samples, int T, int Tkv, int ssB, const float* ln_g,
                                        const float* ln_b, const float* ss, int ss_ld, int num_sms, cudaStream_t st, std::string* err) {
  CUtensorMap mq, mkv;
  if (!make_frames_tmap(&mq, q, D, n_samples, T, err, q_box_frames(T)) || !make_frames_tmap(&mkv, kv, 2 * D, n_samples, Tkv, err))
    return cudaErrorInvalidValue;
  return launch_attn_ws_qkv(mq, mkv, 0, D, z, n_samples, T, Tkv, ssB, ln_g, ln_b, ss, ss_ld, num_sms, st);
}
#elif defined(DSHEG_EMU_RUNTIME)
// tests/emu whole-engine build: the same self-attention launch on the emulated 3-D tensor maps (emu_tc_prims.h)
inline cudaError_t launch_attn_ws(const bf16* qkv, bf16* z, int n_samples, int T, int ssB, const float* ln_g, const float* ln_b,
                                  const float* ss, int ss_ld, int num_sms, cudaStream_t st, std::string* err, int rev = 0) {
  if ((reinterpret_cast<uintptr_t>(qkv) & 15)) { *err = "attention operand not 16-byte aligned"; return cudaErrorInvalidValue; }
  auto frames_map = [&](int box_frames) {
    CUtensorMap m;
    m.base = qkv; m.cols = (uint64_t)(3 * D); m.rows = (uint64_t)T; m.ld_bytes = (uint64_t)(3 * D) * 2;
    m.box_cols = HD; m.box_rows = (uint32_t)box_frames; m.swizzle_bytes = 128;
    m.n2 = (uint64_t)n_samples; m.ld2_bytes = (uint64_t)T * (3 * D) * 2;
    return m;
  };
  const int n_mt = (T + 15) >> 4, mh = (n_mt + 1) >> 1;
  const CUtensorMap mq = frames_map(16 * mh), mkv = frames_map(16 * n_mt);
  const int grid = n_samples < num_sms ? n_samples : num_sms;
  DSHEG_LAUNCH(attn_ws_kernel, grid, NTHREADS, SMEM_BYTES, st, mq, mkv, D, 2 * D, z, n_samples, T, T, ssB, ln_g, ln_b, ss, ss_ld, rev);
  return cudaGetLastError();
}
#endif  // DSHEG_EMU

}  // namespace aws
}  // namespace dsheg
