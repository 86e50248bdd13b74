// TEST INFRASTRUCTURE ONLY -- diffsheg_b200/csrc/attn_ws.cuh (the warp-specialised TMA attention kernel: mbarriers, 3-D TMA boxes,
// mma.sync / ldmatrix, tensor-memory parking) compiled for the thread-level emulator behind a C ABI for tests/test_emu_kernels.py.
#include "attn_ws.cuh"

using namespace dsheg;

static std::string g_err;
extern "C" const char* emu_attn_ws_last_error() { return g_err.c_str(); }

static CUtensorMap frames_map(const uint16_t* base, int cols, int n_samples, int T, int box_frames) {
  CUtensorMap m;
  m.base = base; m.cols = (uint64_t)cols; m.rows = (uint64_t)T; m.ld_bytes = (uint64_t)cols * 2;
  m.box_cols = av3::HD; m.box_rows = (uint32_t)box_frames; m.swizzle_bytes = 128;
  m.n2 = (uint64_t)n_samples; m.ld2_bytes = (uint64_t)T * cols * 2;
  return m;
}

// q [n_samples * T, q_cols] (Q' at column 0), kv [n_samples * Tkv, kv_cols] (K' at kcol, V at vcol); self-attention passes the fused
// qkv tensor twice (q_cols = kv_cols = 1536, kcol = 512, vcol = 1024).  `grid` = number of persistent CTAs.
extern "C" int emu_attention_ws(const uint16_t* q, int q_cols, const uint16_t* kv, int kv_cols, int kcol, int vcol, uint16_t* z, int n_samples,
                                int T, int Tkv, int ssB, const float* ln_g, const float* ln_b, const float* ss, int ss_ld, int grid, int rev) {
  g_err.clear();
  const int n_mt = (T + 15) >> 4, mh = (n_mt + 1) >> 1, n_kt = (Tkv + 15) >> 4;
  const CUtensorMap mq = frames_map(q, q_cols, n_samples, T, 16 * mh), mkv = frames_map(kv, kv_cols, n_samples, Tkv, 16 * n_kt);
  bf16* zo = reinterpret_cast<bf16*>(z);
  if (grid > n_samples) grid = n_samples;
  const bool ok = emu::run_grid(grid, aws::NTHREADS, 1, aws::SMEM_BYTES,
                                [=] { aws::attn_ws_kernel(mq, mkv, kcol, vcol, zo, n_samples, T, Tkv, ssB, ln_g, ln_b, ss, ss_ld, rev); }, &g_err);
  return ok ? 0 : 1;
}
