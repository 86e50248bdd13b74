// TF32 tensor-core GEMM (tcgen05.mma kind::tf32 + TMA + TMEM) with the shared GemmDesc epilogue, fp32 in / fp32 out.
//
// Role: the arithmetic engine of the "tf32" precision mode -- the middle mode between the strict-fp32 parity engine
// (gemm_simt.cuh; the reference itself runs torch fp32 with TF32 off) and the all-bf16 performance mode
// (gemm_tc.cuh): activations, residual stream, LayerNorm statistics and every epilogue stay fp32; only the two MMA operands
// are rounded to TF32 (10-bit mantissa, round-to-nearest by the TMA load: CU_TENSOR_MAP_DATA_TYPE_TFLOAT32), accumulation is
// fp32 in TMEM.  out[M, N] = epilogue(A[M, K] . W[N, K]^T), A = virtual concat of up to 4 fp32 segments.
//
// Shape: PERSISTENT like the bf16 engine (one CTA or CTA pair per SM walks the output tiles, n fastest), K in blocks of 32 fp32
// (128-byte rows, SWIZZLE_128B -- byte-for-byte the operand geometry of the bf16 kernel: 8 TF32 elements = 32 bytes per MMA k-step),
// 4-stage TMA ring, TWO accumulator stages in TMEM (2 x 256 columns): the fp32 epilogue of tile i -- exact erf GELU, fp32 residual /
// output rows of 1 - 6 KB, several times the mainloop when it ran alone -- runs on 16 warps under the mainloop of tile i + 1.
// Warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 4..19 = epilogue (TMEM lane quadrant = warp id % 4, column group =
// (warp - 4) / 4; a warp hands its accumulator stage back right after its last tcgen05.ld).  The epilogue does its per-row /
// per-column math in the TMEM layout (thread = row), then turns each 32-column chunk through a private 4.5 KB staging buffer so that
// the residual loads and the stores are 128-byte coalesced: 8 lanes x 16 bytes per row, 4 rows per instruction.
// History (profiles/r02): 4 epilogue warps, one tile per CTA, one CTA per SM: tensor pipe 5 - 12 % busy (call9) -> 8 epilogue warps,
// two CTAs per SM (call11: 2 x) -> residual values of a chunk loaded up front (t32: + 6 %) -> this persistent form.
#pragma once
#include "gemm_tc.cuh"

namespace dsheg {
namespace t32 {

using namespace tc;

constexpr int TBM = 128, TBK = 32;                     // rows per CTA, fp32 elements per k-block (128-byte rows)
constexpr int T_NE = 16;                                 // epilogue warps (warps 4 .. 19)
constexpr int T_NUM_THREADS = 128 + 32 * T_NE;
constexpr int T_NUM_ACC = 2;                             // accumulator stages in TMEM
constexpr int T_TURN_LD = 36;                            // floats per staged row (144 B: 16-byte aligned, conflict-free)
constexpr int T_TURN_BYTES = 32 * T_TURN_LD * 4;         // one warp's staging rows
constexpr int T_BAR_BYTES = 128;

// CG = 1: a CTA per 128 x 128 tile (small problems).  CG = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile, like
//         the bf16 engine's pair kernels: each CTA stages its 128 A rows and HALF of the W tile (128 of the 256 N rows), the leader
//         issues 256 x 256 x 8 MMAs that read both CTAs' shared memory, each CTA's TMEM holds its 128 rows x 2 x 256 columns.
template <int CG> struct TCfg {
  static constexpr int BN = CG == 2 ? 256 : 128;         // tile width
  static constexpr int CPW = BN / (T_NE / 4);            // columns per epilogue warp (64 / 32)
  static constexpr int W_ROWS = BN / CG;                 // W rows staged by ONE CTA
  static constexpr int STAGES = 4;
  static constexpr int A_BYTES = TBM * TBK * 4, B_BYTES = W_ROWS * TBK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int VEC_BYTES = T_NUM_ACC * 2 * BN * 4;   // bias | csum of a tile's columns, double-buffered with the accumulator stage
  static constexpr int SMEM_BYTES = PIPE_BYTES + T_NE * T_TURN_BYTES + T_BAR_BYTES + VEC_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = T_NUM_ACC * BN;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static_assert((2 * STAGES + 2 * T_NUM_ACC + 1) * 8 <= T_BAR_BYTES, "barrier area");
  // kind::tf32 instruction descriptor: D = f32 (bit 4), A = B = TF32 (format 2 at [7,10) and [10,13)), both K-major,
  // N >> 3 at [17,23), M >> 4 at [24,29)
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TBM * CG) >> 4) << 24);
};

#ifndef DSHEG_EMU
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
#else   // tests/emu: host model of kind::tf32 (emu_tc_prims.h)
inline void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { tc_mma_tf32_emu(1, tmem_d, da, db, idesc, acc); }
inline void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { tc_mma_tf32_emu(2, tmem_d, da, db, idesc, acc); }
#endif

struct T32Params {
  int M, N, num_kb, nseg;
  int seg_kb_start[5];
  int tiles_m, tiles_n;
  const float* bias; const float* csum; const float* mu; const float* rstd;
  int act;
  const float* res; int ldr, res_mod;
  float* out; int ldo; float* out2;
};

template <int CG>
__global__ void __launch_bounds__(T_NUM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                 const __grid_constant__ CUtensorMap tmW, const T32Params p) {
  using C = TCfg<CG>;
  constexpr int BN = C::BN, STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, A_BYTES = C::A_BYTES, CPW = C::CPW;
  DSHEG_TC_DYN_SMEM(smem_raw);
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // rank 0 of a pair = leader (issues the MMAs)
  const int cta_stride = gridDim.x / CG, cta_first = blockIdx.x / CG;   // tiles are walked per CTA (pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-B alignment
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::PIPE_BYTES + T_NE * T_TURN_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + T_NUM_ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * T_NUM_ACC);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - smem_base));
  float* vecs = reinterpret_cast<float*>(gen_base + (bar_base - smem_base) + T_BAR_BYTES);   // [T_NUM_ACC][2][BN]: bias | csum
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_tiles = p.tiles_m * p.tiles_n;   // tiles per CTA (pair), numbered n fastest: concurrent CTAs share the A panel in L2

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < T_NUM_ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), T_NE * CG); }
    fence_mbarrier_init();
    prefetch_tensormap(&tmA0);
    prefetch_tensormap(&tmW);
  }
  if (warp == 1) tmem_alloc<CG, C::TMEM_COLS>(tmem_slot);   // the same warp id in both CTAs of a pair
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer (each CTA fills its own smem; a pair credits all bytes to the leader's barrier) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t leader_full0 = CG == 2 ? mapa_rank(full_bar(0), 0) : 0u;
      for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
        const int m_blk = (tile / p.tiles_n) * CG + (int)rank, n_blk = tile % p.tiles_n;
        int seg = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          while (kb >= p.seg_kb_start[seg + 1]) ++seg;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const CUtensorMap* ma = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : (seg == 2 ? &tmA2 : &tmA3));
          // W's K axis lays every segment out padded to 64 = two k-blocks, so k-block kb of the walk IS W's k-block kb
          if (CG == 2) {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            const uint32_t lb = leader_full0 + 8u * stage;
            tma_load_2d_pair(ma, lb, sa, (kb - p.seg_kb_start[seg]) * TBK, m_blk * TBM);
            tma_load_2d_pair(&tmW, lb, sb, kb * TBK, n_blk * BN + (int)rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(ma, full_bar(stage), sa, (kb - p.seg_kb_start[seg]) * TBK, m_blk * TBM);
            tma_load_2d(&tmW, full_bar(stage), sb, kb * TBK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // the epilogue warps (of both CTAs) have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {   // 8 TF32 = 32 bytes per k-step: +2 in 16-byte descriptor units
            if (CG == 2) tc_mma_tf32_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::IDESC, (kb | k) != 0);
            else tc_mma_tf32(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::IDESC, (kb | k) != 0);
          }
          if (CG == 2) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));   // frees the smem slot(s) when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) tc_commit_pair(tfull_bar(acc)); else tc_commit(tfull_bar(acc));   // accumulator complete -> epilogue(s)
        if (++acc == T_NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================= epilogue: warps 4..19, TMEM lane quadrant = warp id % 4, column group = (warp - 4) / 4 =================
    const int e = warp - 4, qd = warp & 3, cgp = e >> 2;
    const int etid = threadIdx.x - 128;
    const bool ln = p.csum != nullptr;
    float* turn = reinterpret_cast<float*>(gen_base + C::PIPE_BYTES) + e * 32 * T_TURN_LD;   // this warp's 32 x 36 staging rows
    const int tr = lane >> 3, tcg = lane & 7;              // coalesced layout: row 4 it + tr of the chunk, columns 4 tcg .. + 3
    const uint32_t leader_tempty0 = CG == 2 ? mapa_rank(tempty_bar(0), 0) : 0u;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
      const int m_blk = (tile / p.tiles_n) * CG + (int)rank, n_blk = tile % p.tiles_n;
      // ---- per-column vectors of this tile, double-buffered with the accumulator stage (a warp can be at most one tile ahead of another)
      float* vbias = vecs + acc * 2 * BN;
      float* vcsum = vbias + BN;
      for (int i = etid; i < BN; i += 32 * T_NE) {
        const int n = n_blk * BN + i;
        vbias[i] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
        vcsum[i] = (ln && n < p.N) ? __ldg(p.csum + n) : 0.f;
      }
      epi_bar_sync<32 * T_NE>();
      const int m_row = m_blk * TBM + qd * 32 + lane;        // the row this thread holds in the TMEM layout
      const float mu = (ln && m_row < p.M) ? __ldg(p.mu + m_row) : 0.f;
      const float rstd = (ln && m_row < p.M) ? __ldg(p.rstd + m_row) : 1.f;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < CPW / 32; ++ch) {
        const int c0 = cgp * CPW + ch * 32;                  // first column of the chunk inside the tile
        const int n0 = n_blk * BN + c0;
        const bool chunk_ok = n0 < p.N;                      // warp-uniform
        const int n = n0 + 4 * tcg;
        const bool vec_ok = n + 3 < p.N && (p.ldo & 3) == 0 && (!p.res || (p.ldr & 3) == 0);
        // the chunk's residual values first: 8 independent 16-byte loads per thread in flight under the TMEM load, the epilogue math
        // and the turn through shared memory (inside the store loop they would be 8 SERIAL round trips: the compiler cannot hoist a
        // load above the previous iteration's store to a possibly aliasing pointer)
        float4 rv[8];
        if (p.res && vec_ok && chunk_ok) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int m = m_blk * TBM + qd * 32 + 4 * it + tr;
            const int mr = p.res_mod > 0 ? m % p.res_mod : m;
            rv[it] = (m < p.M) ? __ldg(reinterpret_cast<const float4*>(p.res + (size_t)mr * p.ldr + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        if (ch == CPW / 32 - 1) {   // this warp's last TMEM read of the tile is complete: hand the accumulator stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(leader_tempty0 + 8u * acc); else mbar_arrive(tempty_bar(acc));
          }
        }
        if (!chunk_ok) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float v[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int nn = n0 + j + q4;
            float x = __uint_as_float(r[j + q4]);
            if (nn < p.N) {
              if (ln) x = rstd * (x - mu * vcsum[c0 + j + q4]);
              x = apply_act(x + vbias[c0 + j + q4], p.act);
            }
            v[q4] = x;
          }
          *reinterpret_cast<float4*>(turn + lane * T_TURN_LD + j) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = 4 * it + tr;
          const int m = m_blk * TBM + qd * 32 + rl;
          if (m >= p.M || n >= p.N) continue;
          float4 x = *reinterpret_cast<const float4*>(turn + rl * T_TURN_LD + 4 * tcg);
          const size_t o = (size_t)m * p.ldo + n;
          const int mr = p.res_mod > 0 ? m % p.res_mod : m;
          if (vec_ok) {
            if (p.res) { x.x += rv[it].x; x.y += rv[it].y; x.z += rv[it].z; x.w += rv[it].w; }
            *reinterpret_cast<float4*>(p.out + o) = x;
            if (p.out2) *reinterpret_cast<float4*>(p.out2 + o) = x;
          } else {   // ragged N / unaligned rows: element by element
            if (p.res) {
              x.x += p.res[(size_t)mr * p.ldr + n];
              if (n + 1 < p.N) x.y += p.res[(size_t)mr * p.ldr + n + 1];
              if (n + 2 < p.N) x.z += p.res[(size_t)mr * p.ldr + n + 2];
              if (n + 3 < p.N) x.w += p.res[(size_t)mr * p.ldr + n + 3];
            }
            p.out[o] = x.x;
            if (n + 1 < p.N) p.out[o + 1] = x.y;
            if (n + 2 < p.N) p.out[o + 2] = x.z;
            if (n + 3 < p.N) p.out[o + 3] = x.w;
            if (p.out2) {
              p.out2[o] = x.x;
              if (n + 1 < p.N) p.out2[o + 1] = x.y;
              if (n + 2 < p.N) p.out2[o + 2] = x.z;
              if (n + 3 < p.N) p.out2[o + 3] = x.w;
            }
          }
        }
        __syncwarp();   // the staging rows are rewritten by the next chunk
      }
      if (++acc == T_NUM_ACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG, C::TMEM_COLS>(tmem_base);
  }
}

// fp32 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 32 columns] (128-byte rows,
// SWIZZLE_128B); elements are rounded to TF32 by the load; out-of-bounds elements read as zero.
inline bool make_tmap_f32(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, std::string* err) {
#ifdef DSHEG_EMU   // tests/emu: the emulated tensor map records the same geometry (emu_tc_prims.h)
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 4)) { *err = "TMA operand not 16-byte aligned"; return false; }
  map->base = ptr; map->cols = (uint64_t)cols; map->rows = (uint64_t)rows; map->ld_bytes = (uint64_t)ld * 4;
  map->box_cols = (uint32_t)TBK; map->box_rows = (uint32_t)box_rows; map->swizzle_bytes = 128; map->elem_bytes = 4;
  return true;
#else
  struct Key { const void* p; int r, c, l, b; bool operator==(const Key& o) const { return p == o.p && r == o.r && c == o.c && l == o.l && b == o.b; } };
  struct Hash { size_t operator()(const Key& k) const { return reinterpret_cast<size_t>(k.p) ^ ((size_t)k.r * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.c * 0xC2B2AE3D27D4EB4Full) ^ ((size_t)k.l << 20) ^ ((size_t)k.b << 7); } };
  static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;
  const Key k{ptr, rows, cols, ld, box_rows};
  auto it = cache.find(k);
  if (it != cache.end()) { *map = it->second; return true; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled (tf32) failed, CUresult " + std::to_string((int)r); return false; }
  if (cache.size() > 8192) cache.clear();
  cache.emplace(k, *map);
  return true;
#endif
}

// true when the TMA path can take this GEMM (16-byte aligned bases and row strides, fp32 in / fp32 out); the caller falls
// back to the SIMT fp32 kernel otherwise (small odd-shaped projections: more precise anyway)
inline bool tf32_eligible(const GemmDesc& d) {
  if (d.nseg < 1 || d.nseg > 4 || d.M < 1 || d.N < 1) return false;
  if ((reinterpret_cast<uintptr_t>(d.w) & 15) || (d.Kp & 3)) return false;
  for (int s = 0; s < d.nseg; ++s)
    if ((reinterpret_cast<uintptr_t>(d.a[s].ptr) & 15) || (d.a[s].ld & 3) || d.a[s].k < 1) return false;
  if (d.ps_out || d.ps_in || d.act > ACT_GELU) return false;   // fused statistics / softmax / LayerNorm epilogues: bf16 engine only
  return true;
}

// A, W, residual, out: fp32 (the "tf32" mode keeps every activation in fp32); residual / out may be fp32 by type or by flag
#ifdef DSHEG_EMU
inline std::string& g_emu_error_tf32() { static std::string e; return e; }
#endif

template <int CG>
inline cudaError_t launch_tf32_variant(const CUtensorMap* maps, const T32Params& p, int tiles, int num_sms, cudaStream_t st) {
  auto kern = gemm_tf32_kernel<CG>;
  const int slots = num_sms / CG > 0 ? num_sms / CG : 1;          // persistent: one CTA (pair) per SM (pair) walks the tiles
  const int units = tiles < slots ? tiles : slots;
#ifdef DSHEG_EMU   // tests/emu: run the grid on the thread-level emulator (clusters of CG CTAs)
  (void)st;
  const CUtensorMap m0 = maps[0], m1 = maps[1], m2 = maps[2], m3 = maps[3], m4 = maps[4];
  static std::string emu_err;
  if (!emu::run_grid(units * CG, T_NUM_THREADS, CG, TCfg<CG>::SMEM_BYTES, [=] { kern(m0, m1, m2, m3, m4, p); }, &emu_err)) {
    g_emu_error_tf32() = emu_err;
    return cudaErrorLaunchFailure;
  }
  return cudaSuccess;
#else
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg<CG>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CG); cfg.blockDim = dim3(T_NUM_THREADS); cfg.dynamicSmemBytes = TCfg<CG>::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (CG == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], p);
#endif
}

// A, W, residual, out: fp32 (the "tf32" mode keeps every activation in fp32); residual / out may be fp32 by type or by flag.
// cg_force: 0 = automatic (CTA pairs for M >= 2048 and N >= 256), 1 / 2 = force
inline cudaError_t launch_gemm_tf32(const GemmDesc& d, int num_sms, cudaStream_t st, std::string* err, int cg_force = 0) {
  const int cg = cg_force ? cg_force : ((d.M >= 2048 && d.N >= 256) ? 2 : 1);
  const int bn = cg == 2 ? 256 : 128;
  T32Params p{};
  p.M = d.M; p.N = d.N; p.nseg = d.nseg;
  CUtensorMap maps[5];
  int koff = 0;
  for (int s = 0; s < d.nseg; ++s) {
    // segment s covers W columns [koff, koff + k) with koff a multiple of 64 = 2 k-blocks: k-block bookkeeping in units of 32
    p.seg_kb_start[s] = koff / TBK;
    if (!make_tmap_f32(&maps[s], d.a[s].ptr, d.M, d.a[s].k, d.a[s].ld, TBM, err)) return cudaErrorInvalidValue;
    koff += (d.a[s].k + 63) / 64 * 64;
  }
  // the producer walks k-blocks 0 .. num_kb-1 and maps them to (segment, block inside the segment) through seg_kb_start;
  // blocks that lie entirely in a segment's zero padding multiply zeros (at most one per segment)
  p.seg_kb_start[d.nseg] = koff / TBK;
  for (int s = d.nseg + 1; s < 5; ++s) p.seg_kb_start[s] = 1 << 30;
  p.num_kb = koff / TBK;
  for (int s = d.nseg; s < 4; ++s) maps[s] = maps[0];
  if (!make_tmap_f32(&maps[4], d.w, d.N, d.Kp, d.Kp, bn / cg, err)) return cudaErrorInvalidValue;
  p.tiles_n = (d.N + bn - 1) / bn;
  const int tiles_m = (d.M + TBM * cg - 1) / (TBM * cg);
  p.tiles_m = tiles_m;
  p.bias = d.bias; p.csum = d.csum; p.mu = d.mu; p.rstd = d.rstd; p.act = d.act;
  p.res = reinterpret_cast<const float*>(d.res); p.ldr = d.ldr; p.res_mod = d.res_mod;
  p.out = reinterpret_cast<float*>(d.out); p.ldo = d.ldo; p.out2 = reinterpret_cast<float*>(d.out2);
  return cg == 2 ? launch_tf32_variant<2>(maps, p, tiles_m * p.tiles_n, num_sms, st) : launch_tf32_variant<1>(maps, p, tiles_m * p.tiles_n, num_sms, st);
}

}  // namespace t32
}  // namespace dsheg
