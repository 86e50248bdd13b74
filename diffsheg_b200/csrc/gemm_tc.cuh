// tcgen05 (UMMA) bf16 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] . W[N,K]^T ), fp32 accumulate.
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a multi-stage smem ring
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16 per instruction), accumulators live in
//     TMEM (2 x BN columns, double-buffered: the epilogue of tile i overlaps the mainloop of tile i+1)
//   * BN/16 epilogue warps (4 TMEM lane quadrants x BN/64 column groups) read TMEM with tcgen05.ld (one
//     accumulator row per thread, 64 columns per warp) and apply the fused epilogue -- LayerNorm fold, bias,
//     SiLU / GELU, residual or positional add, optional duplicate store for the two CFG halves.  The epilogue
//     is specialised at compile time (no per-element branches); bias / column sums are staged in smem once
//     per tile; every warp owns a 4 KB swizzled staging box: the bf16 residual box arrives by TMA LOAD (issued
//     before the accumulator is ready), results leave by TMA STORE (cp.async.bulk.tensor, bulk groups), so all
//     global traffic of the epilogue is full-line and asynchronous and warps never synchronise with each other.
//     (History in profiles/r01: v0 = 4 generic epilogue warps, IPC 0.12, 5-11 % tensor pipe; v2 = 16
//     specialised warps with direct 16-byte global accesses, lg_throttle / long_scoreboard bound.)
//   * persistent: one CTA per SM walks tiles n-fastest so concurrently running CTAs share the A row-panel
//     through L2; W panels (<= 2 MB) stay L2-resident
//   * A is a VIRTUAL CONCAT of up to 4 row-major segments (one tensor map each): the feat_proj input
//     cat(h, audio, hubert, expr) (transformer.py:304-310) is never materialised.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2-3 idle (2 = W producer in the split-ring build), 4.. = epilogue.
#pragma once
#include <unordered_map>

#include "tc_prims.cuh"

namespace dsheg {
namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_ACC = 2;
constexpr int UMMA_K = 16;
constexpr int CPW = 64;            // accumulator columns per epilogue warp (= one 128-byte bf16 row segment)
constexpr int STG_BYTES_WIDE = 32 * CPW * 2;   // per-warp staging box: 32 rows x 64 bf16 (128-byte rows, SWIZZLE_128B)
constexpr int STG_BYTES_NARROW = 32 * 32 * 2;  // 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B), used twice per tile
constexpr int BAR_BYTES = 512;

// CG = 1: one CTA per tile (128 x BN).  CG = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile:
// each CTA stages its 128 A rows and HALF of the W tile (128 of the 256 N rows); the pair's tensor cores read both
// halves, so the smem fill + operand-read traffic per MMA drops by a third.  ncu on the CG = 1 kernel showed it
// bound by shared-memory bandwidth (TMA fill 96 B/clk + UMMA operand reads 96 B/clk against ~128 B/clk).
// NARROW (pair kernels without a bf16 residual): the epilogue box shrinks to 2 KB per warp and is used once per
// 32-column chunk, which frees 32 KB for a FIFTH pipeline stage -- the pair mainloop is pipeline-depth bound
// (2 / 3 / 4 stages: 793 / 1085 / 1234 TF/s on the qkv shape, profiles/r01/final_gemm_sweep_pair_stages*.txt).
// LONGK (K >= 768, pair kernels): the mainloop dominates, so the kernel runs at the smem limit -- no alignment slack (the
// dynamic smem base must be 1024-aligned: checked, traps otherwise), single-buffered bias/csum vectors (one extra epilogue
// barrier per tile) -> 6 stages with narrow boxes, 5 with wide (residual) boxes.  K = 512 tiles are epilogue-sensitive and
// keep 4 stages, wide boxes and double-buffered vectors (profiles/r01/v8_gemm_sweep.txt vs final_gemm_sweep.txt).
// ROWLN (ACT_LNMS): full-row LayerNorm epilogue over both accumulator stages; 5 stages, narrow boxes, and a vector block that
// holds bias | gamma/2 | beta/2 for all 512 columns plus double-buffered per-(column group, row) statistics partials.
template <int BN, int CG = 1, bool NARROW = false, bool LONGK = false, bool ROWLN = false> struct Cfg {
  static_assert(!ROWLN || (BN == 256 && NARROW && (CG == 1 || LONGK)), "the full-row LayerNorm epilogue: 256-wide tiles, narrow boxes; pairs need K >= 768");
  static_assert(CG == 1 || BN == 256, "the CTA-pair kernel uses 256-wide tiles");
  static constexpr int NE = BN / 16;  // epilogue warps
  static constexpr int NUM_THREADS = 128 + NE * 32;
  // pair kernels: 4 ring stages for K = 512 (epilogue-sensitive: wide boxes, double-buffered vectors), 5 / 6 for K >= 768.
  // Measured and rejected on B200 (profiles/r02/call1): a fifth stage for K = 512 (-2 % in the loop), separate A / W rings
  // with two producer threads (7 + 3 and 6 + 4: -2 ... -5 %), L2 prefetch of the next A panel (-4 %).
  static constexpr int PAIR_STAGES = 4;
  static constexpr int STG_BYTES = NARROW ? STG_BYTES_NARROW : STG_BYTES_WIDE;
  static constexpr bool AT_LIMIT = CG == 2 && LONGK;
  static constexpr int STAGES = CG == 2 ? (LONGK ? ((NARROW && !ROWLN) ? PAIR_STAGES + 2 : PAIR_STAGES + 1) : PAIR_STAGES)
                                        : (BN == 128 ? 5 : 3);
  // pair kernels run at the smem limit: the dynamic smem base is required to be 1024-aligned (checked, traps otherwise)
  // and the per-tile bias/csum vectors are single-buffered (one extra epilogue barrier per tile)
  static constexpr int ALIGN_SLACK = AT_LIMIT ? 0 : 1024;
  static constexpr int NVEC = AT_LIMIT ? 1 : NUM_ACC;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;   // W rows staged by ONE CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = NUM_ACC * BN;
  static constexpr int VEC_BYTES = ROWLN ? 3 * (2 * BN) * 4 + 2 * (BN / CPW) * BM * 8 : NVEC * 2 * BN * 4;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int NBAR_RING = 2 * STAGES;   // ring barriers (full + empty)
  static_assert((NBAR_RING + 2 * NUM_ACC + NE + 1) * 8 <= BAR_BYTES, "barrier block too small");
  static constexpr int SMEM_BYTES = PIPE_BYTES + NE * STG_BYTES + ALIGN_SLACK + BAR_BYTES + VEC_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
};

enum { RES_NONE = 0, RES_BF16 = 1, RES_F32_MOD = 2 };

struct Params {
  int M, N, num_kb, nseg;
  int seg_kb_start[5];
  int tiles_m, tiles_n, rev;
  const float* bias; const float* csum; const float* mu; const float* rstd;
  const void* res; int ldr, res_mod;
  void* out; int ldo; void* out2;
  float2* ps_out; const float* nullc; int n_uncond;
  const float2* ps_in; const float2* cs_in; int ps_slots; float ps_invP;
  const float* eshift; int expo_cols;   // ACT_EXPO (appended likewise)
  const float* lnms_g; const float* lnms_b; const float* lnms_ss; int lnms_ld, lnms_B, lnms_T;   // ACT_LNMS (appended likewise)
};

// (L2 eviction hints on the operand loads -- activations / residuals evict-first, weights evict-last -- LOST on hardware: the A panel
//  is re-read by the n-tiles of its row panel and evict-first throws it out in between; 830 -> 787 TF/s, profiles/r02/call17.)

// ---- spin on an mbarrier phase (PTX primitives: tc_prims.cuh) ---------------------------------------------------------------
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  uint64_t t0 = 0;
  while (true) {
    done = mbar_try_wait(bar, parity);   // (try_wait with a suspend-time hint -- sleeping instead of spinning -- measured no difference here: call13)
    if (done) break;
    // watchdog: a lost arrive / wrong descriptor becomes a launch error after 4 s, not a hung GPU
    if ((++spins & 0x3FFu) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) trap();
    }
  }
}

// K-major, 128B-swizzled operand tile [rows][64 bf16] (8-row groups of 1024 B): SBO = 1024 B,
// LBO unused, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // version = 1
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Epilogue activations for bf16 outputs (MUFU.TANH, |abs err| <= 2^-10.99 on tanh, i.e. below the bf16
// output resolution):  silu(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2 (exact identity);
// gelu(x) ~= h + h*tanh(0.79788456 x + 0.03567741 x^3)  (tanh form, |gelu_tanh - gelu_erf| < 5e-4).
template <int ACT> __device__ __forceinline__ float act_fast(float x) {
  if (ACT == ACT_SILU) { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }
  if (ACT == ACT_GELU) {
    const float h = 0.5f * x;
    const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
    return fmaf(h, tanh_fast(u), h);
  }
  return x;
}

// The LayerNorm fold, bias add and activations of the generic epilogue run on packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2, column
// pairs): the K = 512 tiles are epilogue-sensitive and the GELU epilogue would spend 7 scalar fp32 instructions per element
// (3.5 packed + MUFU.TANH).  Same arithmetic, same rounding as the scalar form; measured on B200: ffn1 1165 -> 1224 TF/s
// isolated, +1.6 % in the sampling loop (profiles/r02/call1).
template <int ACT> __device__ __forceinline__ float2 act_fast2(float2 x) {
  if (ACT == ACT_SILU) {
    const float2 h = fmul2(x, make_float2(0.5f, 0.5f));
    return ffma2(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
  }
  if (ACT == ACT_GELU) {
    const float2 h = fmul2(x, make_float2(0.5f, 0.5f));
    const float2 u = fmul2(x, ffma2(make_float2(0.0356774081f, 0.0356774081f), fmul2(x, x), make_float2(0.7978845608f, 0.7978845608f)));
    return ffma2(h, make_float2(tanh_fast(u.x), tanh_fast(u.y)), h);
  }
  return x;
}

template <int BN, bool LN, int ACT, int RES, bool OUTF32, int CG, bool LONGK_>
__global__ void __launch_bounds__(Cfg<BN, CG>::NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmOut2, const __grid_constant__ CUtensorMap tmRes, const Params p) {
  DSHEG_PDL_TRIGGER();   // PDL build: the next kernel may begin its own prologue now (it still waits for this grid's completion)
  constexpr bool LONGK = LONGK_ && CG == 2 && !OUTF32;
  constexpr bool ROWLN = ACT == ACT_LNMS;   // both n-tiles of a row panel per CTA (pair), LayerNorm over the full 2 * BN-column row
  constexpr bool NARROW = (LONGK && RES != RES_BF16) || ROWLN;
  using C = Cfg<BN, CG, NARROW, LONGK, ROWLN>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, NE = C::NE, STG_BYTES = C::STG_BYTES;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // rank 0 of a pair = leader (issues the MMAs)
  const int cta_stride = gridDim.x / CG, cta_first = blockIdx.x / CG;   // tiles are walked per CTA (pair)
  DSHEG_TC_DYN_SMEM(smem_raw);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
  if (C::ALIGN_SLACK == 0 && smem_base != smem_u32(smem_raw)) trap();   // pair kernels have no room for slack
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + C::PIPE_BYTES;
  const uint32_t bar_base = stg_base + NE * STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (C::NBAR_RING + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (C::NBAR_RING + NUM_ACC + a); };
  auto res_bar = [&](int e) { return bar_base + 8u * (C::NBAR_RING + 2 * NUM_ACC + e); };
  const uint32_t tmem_slot = bar_base + 8u * (C::NBAR_RING + 2 * NUM_ACC + NE);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - smem_base));
  float* vecs = reinterpret_cast<float*>(gen_base + (bar_base - smem_base) + BAR_BYTES);  // [NUM_ACC][2][BN]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_tiles = p.tiles_m * p.tiles_n;
  // persistent tile walk (tiles are numbered n-fastest).  Default: every cta_stride-th tile.  ROWLN (tiles_n == 2): a CTA pair
  // takes BOTH n-tiles of every cta_stride-th row panel back to back, so accumulator stage s always holds columns [s BN, (s+1) BN)
  const int tile_first = ROWLN ? cta_first * 2 : cta_first;
  auto tile_next = [&](int tile) { return ROWLN ? ((tile & 1) ? tile + 2 * cta_stride - 1 : tile + 1) : tile + cta_stride; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NUM_ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NE * CG); }
    for (int e = 0; e < NE; ++e) mbar_init(res_bar(e), 1);
    fence_mbarrier_init();
    prefetch_tensormap(&tmA0);
    prefetch_tensormap(&tmW);
    if (!OUTF32) prefetch_tensormap(&tmOut);
  }
  if (warp == 1) {  // TMEM allocation: one full warp (the same warp id in both CTAs of a pair), which also owns the dealloc
    tmem_alloc<CG, C::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // PDL build: barriers, TMEM and tensor-map prefetches above touch nothing a previous kernel wrote; operands, residuals,
  // statistics and outputs below do
  DSHEG_PDL_WAIT();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t leader_full0 = CG == 2 ? mapa_rank(full_bar(0), 0) : 0u;
      for (int tile = tile_first; tile < num_tiles; tile = tile_next(tile)) {
        const int m_pan = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        const int m_blk = (p.rev ? p.tiles_m - 1 - m_pan : m_pan) * CG + (int)rank;
        int seg = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          while (kb >= p.seg_kb_start[seg + 1]) ++seg;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const CUtensorMap* ma = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : (seg == 2 ? &tmA2 : &tmA3));
          if (CG == 2) {
            // both CTAs fill their own smem; all bytes are credited to the leader's full barrier
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            const uint32_t lb = leader_full0 + 8u * stage;
            tma_load_2d_pair(ma, lb, sa, (kb - p.seg_kb_start[seg]) * BK, m_blk * BM);
            tma_load_2d_pair(&tmW, lb, sb, kb * BK, n_blk * BN + (int)rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(ma, full_bar(stage), sa, (kb - p.seg_kb_start[seg]) * BK, m_blk * BM);
            tma_load_2d(&tmW, full_bar(stage), sb, kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile = tile_next(tile)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);        // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 B (= UMMA_K bf16) inside the 128B swizzle atom: +2 in 16-byte units
            if (CG == 2) tc_mma_bf16_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::IDESC, (kb | k) != 0);
            else tc_mma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::IDESC, (kb | k) != 0);
          }
          if (CG == 2) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));  // frees the smem slot(s) when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) tc_commit_pair(tfull_bar(acc)); else tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue(s)
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================= epilogue (NE warps, each 32 rows x 64 columns) =================
    const int e = warp - 4;
    const int q = warp & 3;   // TMEM lane quadrant this warp may touch (warp id % 4)
    const int cg = e >> 2;    // column group
    const int etid = threadIdx.x - 128;
    const uint32_t stg = stg_base + e * STG_BYTES;
    uint8_t* stg_gen = gen_base + (stg - smem_base);
    int acc = 0; uint32_t acc_phase = 0, res_phase = 0;
    const uint32_t leader_tempty0 = CG == 2 ? mapa_rank(tempty_bar(0), 0) : 0u;
    if constexpr (ROWLN) {
      // ===== ACT_LNMS: z = SiLU(LN(acc + bias) * (1 + scale) + shift) over the FULL row (tr:92-96 fused into ffn.linear2) =====
      // A thread owns one row and the 64-column group `cg` of BOTH accumulator stages (columns s*BN + cg*64 ..): pass 1 reads
      // them for (sum, sum of squares), the four column-group warps of a lane quadrant exchange partials through smem, pass 2
      // re-reads TMEM (cheap), normalises, modulates with the row's sample, applies SiLU and stores 32-column boxes by TMA.
      constexpr int NC = 2 * BN;                       // row width (= p.N)
      float* bvec = vecs;                              // [NC] bias
      float* ghalf = vecs + NC;                        // [NC] gamma / 2   (SiLU(x) = h + h tanh(h) with h = x / 2)
      float* bhalf = vecs + 2 * NC;                    // [NC] beta / 2
      float2* part = reinterpret_cast<float2*>(vecs + 3 * NC);   // [2][BN / CPW][BM]
      for (int i = etid; i < NC; i += NE * 32) {
        bvec[i] = p.bias ? __ldg(p.bias + i) : 0.f;
        ghalf[i] = 0.5f * __ldg(p.lnms_g + i);
        bhalf[i] = 0.5f * __ldg(p.lnms_b + i);
      }
      epi_bar_sync<NE * 32>();
      int pbuf = 0;
      for (int unit = cta_first; unit < p.tiles_m; unit += cta_stride) {
        const int m0 = ((p.rev ? p.tiles_m - 1 - unit : unit) * CG + (int)rank) * BM + q * 32;
        const int m = m0 + lane;
        const bool row_ok = m < p.M;
        // ---- pass 1: statistics of this thread's 2 x 64 columns; stage 0 is read while the MMAs still fill stage 1
        float sx = 0.f, sq = 0.f;
#pragma unroll
        for (int stn = 0; stn < 2; ++stn) {
          mbar_wait(tfull_bar(stn), acc_phase);
          tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < CPW / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stn * BN + cg * CPW + ch * 32), r);
            const float* bb = bvec + stn * BN + cg * CPW + ch * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bb + j);
              const float v0 = __uint_as_float(r[j]) + b4.x, v1 = __uint_as_float(r[j + 1]) + b4.y;
              const float v2 = __uint_as_float(r[j + 2]) + b4.z, v3 = __uint_as_float(r[j + 3]) + b4.w;
              sx += (v0 + v1) + (v2 + v3);
              sq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, sq))));
            }
          }
        }
        float2* mypart = part + (size_t)pbuf * (BN / CPW) * BM;
        mypart[cg * BM + q * 32 + lane] = make_float2(sx, sq);
        // one barrier per panel: the partials are double-buffered, so nobody can overwrite buffer b before every warp has
        // read it (a warp writes b again only after the NEXT panel's barrier, which needs all reads of this one)
        epi_bar_sync<NE * 32>();
        sx = 0.f; sq = 0.f;
#pragma unroll
        for (int c = 0; c < BN / CPW; ++c) { const float2 t2 = mypart[c * BM + q * 32 + lane]; sx += t2.x; sq += t2.y; }   // fixed order
        pbuf ^= 1;
        const float mean = sx * (1.f / NC);
        const float rstd = rsqrtf(fmaxf(sq * (1.f / NC) - mean * mean, 0.f) + 1e-5f);
        const float nmr = -mean * rstd;
        const float* sc = p.lnms_ss + (size_t)((row_ok ? m / p.lnms_T : 0) % p.lnms_B) * p.lnms_ld;
        // ---- pass 2: normalise, modulate, SiLU, store
#pragma unroll 1
        for (int stn = 0; stn < 2; ++stn) {
#pragma unroll
          for (int ch = 0; ch < CPW / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stn * BN + cg * CPW + ch * 32), r);
            if (ch == CPW / 32 - 1) {   // last TMEM read of this stage by this warp: hand it back to the MMA issuer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(leader_tempty0 + 8u * stn); else mbar_arrive(tempty_bar(stn));
              }
            }
            const int n0 = stn * BN + cg * CPW + ch * 32;
            // 2 KB box (32 rows x 64 B, SWIZZLE_64B: chunk c of row r at c ^ ((r >> 1) & 3)), reused for every 32-column chunk
            if (lane == 0) bulk_wait_read0();   // the previous chunk's TMA store has finished reading the box
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 4; ++u) {       // 8 columns at a time: compute, pack, store (keeps the live registers low)
              float o[8];
#pragma unroll
              for (int k4 = 0; k4 < 8; k4 += 4) {
                const int c = n0 + u * 8 + k4;
                const float4 b4 = *reinterpret_cast<const float4*>(bvec + c);
                const float4 g4 = *reinterpret_cast<const float4*>(ghalf + c), e4 = *reinterpret_cast<const float4*>(bhalf + c);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c)), h4 = __ldg(reinterpret_cast<const float4*>(sc + NC + c));
                const float bi[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {e4.x, e4.y, e4.z, e4.w};
                const float s1[4] = {s4.x, s4.y, s4.z, s4.w}, s2[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float one_s = 1.f + s1[e2];
                  // h = x / 2,  x = ((v - mean) rstd gamma + beta) (1 + scale) + shift
                  const float hh = fmaf(fmaf(__uint_as_float(r[u * 8 + k4 + e2]) + bi[e2], rstd, nmr), gg[e2] * one_s, fmaf(be[e2], one_s, 0.5f * s2[e2]));
                  o[k4 + e2] = fmaf(hh, tanh_fast(hh), hh);
                }
              }
              uint4 w;
              w.x = pack_bf16x2(o[0], o[1]);
              w.y = pack_bf16x2(o[2], o[3]);
              w.z = pack_bf16x2(o[4], o[5]);
              w.w = pack_bf16x2(o[6], o[7]);
              *reinterpret_cast<uint4*>(stg_gen + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = w;
            }
            fence_async_smem();                 // generic-proxy smem writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmOut, stg, n0, m0);
              bulk_commit();
            }
          }
        }
        acc_phase ^= 1;
      }
    } else
    for (int tile = tile_first; tile < num_tiles; tile = tile_next(tile)) {
      const int m_pan = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      const int m_blk = (p.rev ? p.tiles_m - 1 - m_pan : m_pan) * CG + (int)rank;
      const int n_tile0 = n_blk * BN;
      // ---- stage per-column vectors for this tile (double-buffered with the accumulator stage)
      float* vb = vecs + (C::NVEC == 1 ? 0 : acc) * 2 * BN;
      if (C::NVEC == 1) epi_bar_sync<NE * 32>();   // single buffer: everyone is done reading the previous tile's vectors
      for (int i = etid; i < BN; i += NE * 32) {
        const int n = n_tile0 + i;
        const float bv = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
        vb[i] = bv;
        if (LN) vb[BN + i] = n < p.N ? __ldg(p.csum + n) : 0.f;
        else if (RES != RES_NONE && p.nullc) vb[BN + i] = bv + (n < p.N ? __ldg(p.nullc + n) : 0.f);
        if (ACT == ACT_EXPO && LN && n < p.expo_cols) {   // exponent columns: everything pre-scaled by log2(e), shift folded into the bias
          vb[i] = (bv - __ldg(p.eshift + n)) * 1.4426950408889634f;
          vb[BN + i] *= 1.4426950408889634f;
        }
      }
      epi_bar_sync<NE * 32>();
      const int m0 = m_blk * BM + q * 32;     // first row of this warp's box
      const int nc0 = n_tile0 + cg * CPW;     // first column of this warp's box
      if (!OUTF32 && !NARROW) {
        if (lane == 0) bulk_wait_read0();     // the previous tile's TMA store has finished reading the box
        __syncwarp();
        if (RES == RES_BF16 && lane == 0) {   // residual box by TMA, in flight while the mainloop runs
          mbar_arrive_expect_tx(res_bar(e), STG_BYTES);
          tma_load_2d(&tmRes, res_bar(e), stg, nc0, m0);
        }
      }
      // ---- per-row operands, fetched before the accumulator is ready
      const int m = m0 + lane;
      const bool row_ok = m < p.M;
      float rs = 1.f, rm = 0.f;  // v = rs * acc + rm * csum[n] + bias[n]
      if (LN && row_ok) {
        if (p.ps_in) {  // statistics from the producer epilogue's per-64-column partials (fixed summation order)
          float sx = 0.f, sq = 0.f;
          const float4* pp = reinterpret_cast<const float4*>(p.ps_in + (size_t)m * p.ps_slots);
          for (int u = 0; u < p.ps_slots / 2; ++u) { const float4 f = __ldg(pp + u); sx += f.x + f.z; sq += f.y + f.w; }
          if (p.cs_in) { const float2 c = __ldg(p.cs_in + m); sx += c.x; sq += c.y; }
          const float mean = sx * p.ps_invP;
          rs = rsqrtf(fmaxf(sq * p.ps_invP - mean * mean, 0.f) + 1e-5f);
          rm = -rs * mean;
        } else {
          rs = __ldg(p.rstd + m); rm = -rs * __ldg(p.mu + m);
        }
      }
      // ACT_EXPO: this warp's 64 columns are exponent columns (Q or K) or plain ones (V) as a whole (expo_cols % 64 == 0);
      // with the staged vectors pre-scaled, scaling rs alone yields log2(e) * (v - eshift) from the same two FFMAs
      const bool expo = ACT == ACT_EXPO && LN && nc0 < p.expo_cols;
      if (ACT == ACT_EXPO && expo) rs *= 1.4426950408889634f;
      // rows of the CFG-null half take bias + nullc (slot 1) instead of bias (slot 0)
      const int bsel = (!LN && RES != RES_NONE && p.nullc && m < p.n_uncond) ? BN : 0;
      float psum = 0.f, psq = 0.f;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (RES == RES_BF16) { mbar_wait(res_bar(e), res_phase); res_phase ^= 1; }
#pragma unroll
      for (int ch = 0; ch < CPW / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cg * CPW + ch * 32), r);
        if (ch == CPW / 32 - 1) {   // last TMEM read of this warp is complete: hand the accumulator stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(leader_tempty0 + 8u * acc); else mbar_arrive(tempty_bar(acc));
          }
        }
        const int n0 = nc0 + ch * 32;
        const float* bvec = vb + cg * CPW + ch * 32 + (LN ? 0 : bsel);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bvec + j);
          {
            float2 p0 = make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), p1 = make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (LN) {
              const float4 c4 = *reinterpret_cast<const float4*>(bvec + BN + j);
              const float2 rs2 = make_float2(rs, rs), rm2 = make_float2(rm, rm);
              p0 = ffma2(rs2, p0, ffma2(rm2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
              p1 = ffma2(rs2, p1, ffma2(rm2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
            } else {
              p0 = fadd2(p0, make_float2(b4.x, b4.y));
              p1 = fadd2(p1, make_float2(b4.z, b4.w));
            }
            if (ACT == ACT_EXPO && expo) { p0 = make_float2(ex2_fast(p0.x), ex2_fast(p0.y)); p1 = make_float2(ex2_fast(p1.x), ex2_fast(p1.y)); }
            p0 = act_fast2<ACT>(p0); p1 = act_fast2<ACT>(p1);
            v[j] = p0.x; v[j + 1] = p0.y; v[j + 2] = p1.x; v[j + 3] = p1.y;
          }
        }
        if (RES == RES_F32_MOD) {
          if (row_ok) {
            const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + (size_t)(m % p.res_mod) * p.ldr + n0);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 f = __ldg(rp + u);
              v[u * 4] += f.x; v[u * 4 + 1] += f.y; v[u * 4 + 2] += f.z; v[u * 4 + 3] += f.w;
            }
          }
        }
        if (OUTF32) {
          if (row_ok) {
            float* op = reinterpret_cast<float*>(p.out) + (size_t)m * p.ldo + n0;
            if (n0 + 32 <= p.N && (p.ldo & 3) == 0) {
#pragma unroll
              for (int u = 0; u < 8; ++u) reinterpret_cast<float4*>(op)[u] = make_float4(v[u * 4], v[u * 4 + 1], v[u * 4 + 2], v[u * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + j < p.N) op[j] = v[j];
            }
          }
        } else if (NARROW) {
          // 2 KB box (32 rows x 64 B, SWIZZLE_64B: chunk c of row r at c ^ ((r >> 1) & 3)), reused for every 32-column chunk
          if (RES != RES_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { psum += v[j]; psq = fmaf(v[j], v[j], psq); }
          }
          if (lane == 0) bulk_wait_read0();   // the previous chunk's / tile's TMA store has finished reading the box
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]);
            w.y = pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]);
            w.z = pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]);
            w.w = pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]);
            *reinterpret_cast<uint4*>(stg_gen + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = w;
          }
          fence_async_smem();                 // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, stg, n0, m0);
            if (p.out2) tma_store_2d(&tmOut2, stg, n0, m0);
            bulk_commit();
          }
        } else {
          // this thread's row of the box: 16-byte chunk c of row `lane` lives at chunk position c ^ (lane & 7)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4* sp = reinterpret_cast<uint4*>(stg_gen + lane * 128 + (((ch * 4 + u) ^ (lane & 7)) << 4));
            if (RES == RES_BF16) {
              const uint4 w = *sp;
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[k2]);
                v[u * 8 + k2 * 2] += __bfloat162float(h2.x);
                v[u * 8 + k2 * 2 + 1] += __bfloat162float(h2.y);
              }
            }
            if (RES != RES_NONE) {
#pragma unroll
              for (int k2 = 0; k2 < 8; ++k2) { psum += v[u * 8 + k2]; psq = fmaf(v[u * 8 + k2], v[u * 8 + k2], psq); }
            }
            uint4 w;
            w.x = pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]);
            w.y = pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]);
            w.z = pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]);
            w.w = pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]);
            *sp = w;
          }
        }
      }
      if (RES != RES_NONE && !OUTF32 && p.ps_out && row_ok) p.ps_out[(size_t)m * (p.N / CPW) + (nc0 / CPW)] = make_float2(psum, psq);
      if (!OUTF32 && !NARROW) {
        fence_async_smem();                           // generic-proxy smem writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, stg, nc0, m0);         // rows >= M / columns >= N are clipped by the tensor map
          if (p.out2) tma_store_2d(&tmOut2, stg, nc0, m0);
          bulk_commit();
        }
      }
      if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
    }
    if (!OUTF32 && lane == 0) bulk_wait0();           // all stores complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG, C::TMEM_COLS>(tmem_base);
  }
}

// ---- host side ---------------------------------------------------------------------------------
#ifndef DSHEG_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}
#endif  // DSHEG_EMU

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 64], 128B swizzle,
// out-of-bounds elements read as zero (ragged M / N / K tails need no padding in memory).
inline bool make_tmap_uncached(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols, std::string* err);

// Encoding a tensor map costs ~1 us of host time and a GEMM needs up to 8: cache them (workspace pointers are
// stable for the life of a handle), which matters for the launch-bound small-batch configurations.
inline bool make_tmap(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, std::string* err, int box_cols = BK) {
  struct Key {
    const void* p; int r, c, l, b;
    bool operator==(const Key& o) const { return p == o.p && r == o.r && c == o.c && l == o.l && b == o.b; }
  };
  struct Hash {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.p);
      h ^= (size_t)k.r * 0x9E3779B97F4A7C15ull; h ^= (size_t)k.c * 0xC2B2AE3D27D4EB4Full;
      h ^= ((size_t)k.l << 20) ^ ((size_t)k.b << 7);
      return h;
    }
  };
  static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;
  const Key k{ptr, rows, cols, ld, box_rows * 1024 + box_cols};
  auto it = cache.find(k);
  if (it != cache.end()) { *map = it->second; return true; }
  if (!make_tmap_uncached(map, ptr, rows, cols, ld, box_rows, box_cols, err)) return false;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(k, *map);
  return true;
}

inline bool make_tmap_uncached(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols, std::string* err) {
  // box = [box_rows x 64 columns] (128-byte rows, SWIZZLE_128B): operand tiles and wide epilogue boxes;
  // box = [box_rows x 32 columns] (64-byte rows, SWIZZLE_64B): narrow epilogue boxes
#ifdef DSHEG_EMU   // tests/emu: the emulated tensor map records the same geometry (emu_tc_prims.h)
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) { *err = "TMA operand not 16-byte aligned"; return false; }
  map->base = ptr; map->cols = (uint64_t)cols; map->rows = (uint64_t)rows; map->ld_bytes = (uint64_t)ld * 2;
  map->box_cols = (uint32_t)box_cols; map->box_rows = (uint32_t)box_rows; map->swizzle_bytes = box_cols == 32 ? 64 : 128;
  return true;
#else
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) { *err = "TMA operand not 16-byte aligned"; return false; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r); return false; }
  return true;
#endif
}

#ifdef DSHEG_EMU
inline std::string& g_emu_error() { static std::string e; return e; }
#endif

template <int BN, bool LN, int ACT, int RES, bool OUTF32, int CG, bool LONGK_ = false>
inline cudaError_t launch_variant(const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, LN, ACT, RES, OUTF32, CG, LONGK_>;
  using C = Cfg<BN, CG, ((LONGK_ && CG == 2 && !OUTF32 && RES != RES_BF16) || ACT == ACT_LNMS), (LONGK_ && CG == 2 && !OUTF32), ACT == ACT_LNMS>;
#ifdef DSHEG_EMU   // tests/emu: run the grid on the thread-level emulator (clusters of CG CTAs)
  (void)st;
  const CUtensorMap m0 = maps[0], m1 = maps[1], m2 = maps[2], m3 = maps[3], m4 = maps[4], m5 = maps[5], m6 = maps[6], m7 = maps[7];
  static std::string emu_err;
  const bool ok = emu::run_grid(grid, C::NUM_THREADS, CG, C::SMEM_BYTES, [=] { kern(m0, m1, m2, m3, m4, m5, m6, m7, p); }, &emu_err);
  if (!ok) { g_emu_error() = emu_err; tma_stores_drained(); return cudaErrorLaunchFailure; }
  if (!tma_stores_drained()) { g_emu_error() = "a TMA store was still pending (no cp.async.bulk.wait_group) when its thread exited"; return cudaErrorLaunchFailure; }
  return cudaSuccess;
#else
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (CG == 1) {
    DSHEG_LAUNCH(kern, grid, C::NUM_THREADS, C::SMEM_BYTES, st, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C::NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#if DSHEG_PDL_ATTRS   // experiment build: programmatic dependent launch (the kernel executes DSHEG_PDL_WAIT after its prologue)
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 2;
#endif
  return cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
#endif
}

template <int BN, int CG>
inline cudaError_t dispatch(const GemmDesc& d, const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st,
                            std::string* err, bool longk = false) {
  const bool ln = d.csum != nullptr;
  const int res = !d.res ? RES_NONE : (d.res_f32 ? RES_F32_MOD : RES_BF16);
  if (d.out_f32) {
    if (!ln && d.act == ACT_NONE && res == RES_NONE && !d.out2) return launch_variant<BN, false, ACT_NONE, RES_NONE, true, CG>(maps, p, grid, st);
  } else if (CG == 2 && longk && !ln && d.act == ACT_NONE && res == RES_BF16) {   // long K, residual: wide boxes, 5 stages
    return launch_variant<BN, false, ACT_NONE, RES_BF16, false, CG, true>(maps, p, grid, st);
  } else if (CG == 2 && longk && res == RES_NONE) {   // long K loops: 2 KB epilogue boxes, 6 stages
    if constexpr (BN == 256 && CG == 2) {
      if (!ln && d.act == ACT_LNMS) return launch_variant<BN, false, ACT_LNMS, RES_NONE, false, CG, true>(maps, p, grid, st);
    }
    if (ln && d.act == ACT_NONE) return launch_variant<BN, true, ACT_NONE, RES_NONE, false, CG, true>(maps, p, grid, st);
    if (ln && d.act == ACT_SILU) return launch_variant<BN, true, ACT_SILU, RES_NONE, false, CG, true>(maps, p, grid, st);
    if (!ln && d.act == ACT_NONE) return launch_variant<BN, false, ACT_NONE, RES_NONE, false, CG, true>(maps, p, grid, st);
    if (!ln && d.act == ACT_GELU) return launch_variant<BN, false, ACT_GELU, RES_NONE, false, CG, true>(maps, p, grid, st);
    if (!ln && d.act == ACT_SILU) return launch_variant<BN, false, ACT_SILU, RES_NONE, false, CG, true>(maps, p, grid, st);
  } else if (ln) {
    if (d.act == ACT_EXPO && res == RES_NONE) return launch_variant<BN, true, ACT_EXPO, RES_NONE, false, CG>(maps, p, grid, st);
    if (d.act == ACT_NONE && res == RES_NONE) return launch_variant<BN, true, ACT_NONE, RES_NONE, false, CG>(maps, p, grid, st);
    if (d.act == ACT_SILU && res == RES_NONE) return launch_variant<BN, true, ACT_SILU, RES_NONE, false, CG>(maps, p, grid, st);
  } else {
    if (d.act == ACT_NONE && res == RES_NONE) return launch_variant<BN, false, ACT_NONE, RES_NONE, false, CG>(maps, p, grid, st);
    if (d.act == ACT_NONE && res == RES_BF16) return launch_variant<BN, false, ACT_NONE, RES_BF16, false, CG>(maps, p, grid, st);
    if (d.act == ACT_NONE && res == RES_F32_MOD) return launch_variant<BN, false, ACT_NONE, RES_F32_MOD, false, CG>(maps, p, grid, st);
    if (d.act == ACT_GELU && res == RES_NONE) return launch_variant<BN, false, ACT_GELU, RES_NONE, false, CG>(maps, p, grid, st);
    if (d.act == ACT_SILU && res == RES_NONE) return launch_variant<BN, false, ACT_SILU, RES_NONE, false, CG>(maps, p, grid, st);
    if constexpr (BN == 256 && CG == 1) {   // small batches: the full-row LayerNorm epilogue on single CTAs (128 rows x 512 columns of TMEM)
      if (d.act == ACT_LNMS && res == RES_NONE) return launch_variant<BN, false, ACT_LNMS, RES_NONE, false, CG>(maps, p, grid, st);
    }
  }
  *err = "no tcgen05 GEMM variant for this epilogue combination";
  return cudaErrorInvalidValue;
}

inline int g_cg_override() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DSHEG_TC_CG"); v = e ? atoi(e) : 0; }
  return v;
}

inline int g_bn_override() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DSHEG_TC_BN"); v = e ? atoi(e) : 0; }
  return v;
}

inline cudaError_t launch_gemm_tc(const GemmDesc& d, int num_sms, cudaStream_t st, std::string* err, int bn_force = 0, int cg_force = 0) {
  // vector paths need 16-byte aligned rows; every engine buffer satisfies this
  if (!d.out_f32 && ((d.ldo % 8) || (d.N % 64))) { *err = "bf16-output GEMM needs N % 64 == 0 and ldo % 8 == 0"; return cudaErrorInvalidValue; }
  if (d.res && !d.res_f32 && (d.ldr % 8)) { *err = "bf16 residual needs ldr % 8 == 0"; return cudaErrorInvalidValue; }
  if (d.res && d.res_f32 && ((d.ldr % 4) || d.res_mod <= 0)) { *err = "fp32 residual needs ldr % 4 == 0 and res_mod > 0"; return cudaErrorInvalidValue; }
  int bn = bn_force ? bn_force : g_bn_override();
  if (bn != 128 && bn != 256) {
    bn = (d.N % 256 == 0) ? 256 : 128;
    // single-clip regime (a handful of row tiles): 128-wide tiles double the CTAs that stream W and deepen the ring to 5 stages;
    // measured on B200 at B = 1: 787 -> 941 frames/s (BEAT), 1780 -> 2115 (SHOW) -- profiles/r02/call2/r2_configs1_bn128.jsonl
    const int tiles256 = ((d.M + BM - 1) / BM) * ((d.N + 255) / 256);
    if (bn == 256 && d.act != ACT_LNMS && tiles256 * 4 < num_sms) bn = 128;
  }
  // CTA pairs (cta_group::2) for the big row counts; tiny problems keep the single-CTA kernel (more tiles in flight)
  int cg = cg_force ? cg_force : g_cg_override();
  if (cg != 1 && cg != 2) cg = (bn == 256 && d.M >= 4096 && d.Kp >= 512) ? 2 : 1;   // short K loops: single CTAs win (profiles/r01 sweep)
  if (bn != 256) cg = 1;
  if (d.act == ACT_LNMS && d.Kp < 768) cg = 1;   // the pair form of the full-row LayerNorm epilogue is built on the long-K configuration
  // K >= 768: the mainloop dominates -> deeper ring (and narrow epilogue boxes when there is no bf16 residual box to load)
  const bool longk = cg == 2 && !d.out_f32 && d.Kp >= 768 && !(d.res && d.res_f32);
  const bool narrow = (longk && !d.res) || d.act == ACT_LNMS;
  Params p{};
  CUtensorMap maps[8];
  p.M = d.M; p.N = d.N; p.nseg = d.nseg;
  int kb = 0;
  for (int s = 0; s < d.nseg; ++s) {
    p.seg_kb_start[s] = kb;
    if (!make_tmap(&maps[s], d.a[s].ptr, d.M, d.a[s].k, d.a[s].ld, BM, err)) return cudaErrorInvalidValue;
    kb += (d.a[s].k + BK - 1) / BK;
  }
  for (int s = d.nseg; s < 5; ++s) p.seg_kb_start[s] = kb;
  for (int s = d.nseg; s < 4; ++s) maps[s] = maps[0];
  p.num_kb = kb;
  if (kb * BK != d.Kp) { *err = "GEMM weight K padding does not match the A segments"; return cudaErrorInvalidValue; }
  if (!make_tmap(&maps[4], d.w, d.N, d.Kp, d.Kp, bn / cg, err)) return cudaErrorInvalidValue;
  maps[5] = maps[6] = maps[7] = maps[4];
  if (!d.out_f32) {  // epilogue boxes: 32 rows x 64 columns of the bf16 output / residual (32 columns for NARROW kernels)
    const int bc = narrow ? 32 : BK;
    if (!make_tmap(&maps[5], d.out, d.M, d.N, d.ldo, 32, err, bc)) return cudaErrorInvalidValue;
    if (d.out2 && !make_tmap(&maps[6], d.out2, d.M, d.N, d.ldo, 32, err, bc)) return cudaErrorInvalidValue;
    if (d.res && !d.res_f32 && !make_tmap(&maps[7], d.res, d.M, d.N, d.ldr, 32, err)) return cudaErrorInvalidValue;
  }
  p.tiles_m = (d.M + BM * cg - 1) / (BM * cg); p.tiles_n = (d.N + bn - 1) / bn;   // tiles per CTA (pair)
  p.rev = d.rev;
  p.bias = d.bias; p.csum = d.csum; p.mu = d.mu; p.rstd = d.rstd;
  p.res = d.res; p.ldr = d.ldr; p.res_mod = d.res_mod;
  p.out = d.out; p.ldo = d.ldo; p.out2 = d.out2;
  p.ps_out = d.ps_out; p.nullc = d.nullc; p.n_uncond = d.n_uncond;
  p.ps_in = d.ps_in; p.cs_in = d.cs_in; p.ps_slots = d.ps_slots; p.ps_invP = d.ps_P > 0 ? 1.0f / (float)d.ps_P : 0.f;
  p.lnms_g = d.lnms_g; p.lnms_b = d.lnms_b; p.lnms_ss = d.lnms_ss; p.lnms_ld = d.lnms_ld; p.lnms_B = d.lnms_B; p.lnms_T = d.lnms_T;
  if (d.act == ACT_LNMS && (bn != 256 || (cg == 2 && !longk) || d.N != 2 * bn || d.csum || d.res || d.out_f32 || d.out2 || !d.lnms_g || !d.lnms_b ||
                            !d.lnms_ss || d.lnms_B <= 0 || d.lnms_T <= 0 || (d.lnms_ld % 4))) {
    *err = "ACT_LNMS needs 256-wide tiles with N == 512 (CTA pairs: K >= 768), bf16 output, no LN fold / residual / duplicate store, and the LayerNorm / modulation operands";
    return cudaErrorInvalidValue;
  }
  p.eshift = d.eshift; p.expo_cols = d.expo_cols;
  if (d.act == ACT_EXPO && (!d.csum || !d.eshift || d.expo_cols <= 0 || (d.expo_cols % CPW) || d.expo_cols > d.N || d.res || d.out_f32 || d.Kp >= 768)) {
    *err = "ACT_EXPO needs an LN-fold bf16-output GEMM with K < 768, no residual, eshift and expo_cols % 64 == 0";
    return cudaErrorInvalidValue;
  }
  if ((d.ps_out || d.nullc) && (d.out_f32 || !d.res)) { *err = "fused LN statistics need a bf16-output residual GEMM"; return cudaErrorInvalidValue; }
  if (d.ps_in && (!d.csum || (d.ps_slots & 1))) { *err = "ps_in needs an LN-fold GEMM and an even slot count"; return cudaErrorInvalidValue; }
  const int tiles = d.act == ACT_LNMS ? p.tiles_m : p.tiles_m * p.tiles_n;   // ACT_LNMS: a pair owns whole row panels (both n-tiles)
  const int units = num_sms / cg;   // CTAs (or CTA pairs) that can be resident
  const int grid = (tiles < units ? tiles : units) * cg;
  if (cg == 2) return dispatch<256, 2>(d, maps, p, grid, st, err, longk);
  return bn == 256 ? dispatch<256, 1>(d, maps, p, grid, st, err) : dispatch<128, 1>(d, maps, p, grid, st, err);
}

}  // namespace tc
}  // namespace dsheg
