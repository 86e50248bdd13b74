// TEST INFRASTRUCTURE ONLY -- the attention kernel SOURCES of diffsheg_b200/csrc compiled for the host emulator
// (g++ -DDSHEG_EMU) behind a small C ABI for tests/test_emu_kernels.py.  Pointers are host pointers; bf16 travels as uint16.
#include <algorithm>

#include "attn_v3.cuh"
#include "attn_small.cuh"
#include "frontend.cuh"
#include "postprocess.cuh"
#include "sampler.cuh"

using namespace dsheg;

static std::string g_err;

extern "C" const char* emu_last_error() { return g_err.c_str(); }

// dynamic per-lane operation counts since the last reset: {cp.async 16 B, ldmatrix, mma.sync, ex2, tanh, rcp, packed fp32}
extern "C" void emu_op_counters(unsigned long long* out, int reset) {
  prims::OpCounters& c = prims::op_counters();
  const unsigned long long v[7] = {c.cp_async16, c.ldsm, c.mma, c.ex2, c.tanh, c.rcp, c.packed_fp32};
  for (int i = 0; i < 7; ++i) out[i] = v[i];
  if (reset) c = prims::OpCounters{};
}

// variant: 3 = attn_v3 (one CTA per sample; the per-layer fallback of the engine).  The default kernel (attn_tma.cuh: TMA, mbarriers)
// is validated on hardware: op-level parity, memcheck and racecheck under compute-sanitizer (profiles/r02).
extern "C" int emu_attention(int variant, const uint16_t* qkv, uint16_t* z, int n_samples, int T, int ssB, const float* ln_g,
                             const float* ln_b, const float* ss, int ss_ld, const float* qsum) {
  (void)qsum;
  g_err.clear();
  prims::async_copies().clear();
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  bf16* zo = reinterpret_cast<bf16*>(z);
  bool ok = false;
  if (variant == 3) {
    ok = emu::run_grid(n_samples, av3::NTHREADS, 1, av3::SMEM_BYTES, [=] { av3::attn_v3_kernel(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else {
    g_err = "unknown attention variant";
  }
  return ok ? 0 : 1;
}

// audio-layer attention (D = 128, 8 heads of 16): qkv [n_samples, T, 384], z [n_samples, T, 128]
extern "C" int emu_attention_d128(const uint16_t* qkv, uint16_t* z, int n_samples, int T, int ssB, const float* ln_g, const float* ln_b,
                                  const float* ss, int ss_ld) {
  g_err.clear();
  prims::async_copies().clear();
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  bf16* zo = reinterpret_cast<bf16*>(z);
  return emu::run_grid(n_samples, asmall::NTHREADS, 1, asmall::smem_bytes(T), [=] { asmall::attn_d128_kernel(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err) ? 0 : 1;
}

// ---- elementwise kernels: a small grid of 256-thread CTAs, grid-stride like on the device --------------------------------------
static int ew(long long n, std::function<void()> body) {
  g_err.clear();
  const int grid = (int)std::max<long long>(1, std::min<long long>(4, (n + 255) / 256));
  return emu::run_grid(grid, 256, 1, 0, body, &g_err) ? 0 : 1;
}

extern "C" int emu_inv_standardize(const float* x, int ldx, const float* mean, const float* stdv, float* out, int ldo, long long rows, int D) {
  return ew(rows * D, [=] { inv_standardize_kernel(x, ldx, mean, stdv, out, ldo, rows, D); });
}
extern "C" int emu_beat_axis_angle(const float* x, int ldx, const float* mean_aa, const float* std_aa, const float* mean_pose,
                                   const float* std_pose, float* euler_deg, float* out_norm, long long rows, int joints) {
  return ew(rows * joints, [=] { beat_axis_angle_kernel(x, ldx, mean_aa, std_aa, mean_pose, std_pose, euler_deg, out_norm, rows, joints); });
}
extern "C" int emu_ddim_step(const float* x, const float* eps, float* x_out, float* pred_out, long long n, int T, int D, float a, float b,
                             float sqrt_acp, float sqrt_1m_acp, const float* gt, const unsigned char* mask, const float* noise2,
                             int blend, int overlap_len) {
  DdimArgs p{x, eps, x_out, pred_out, n, T, D, a, b, sqrt_acp, sqrt_1m_acp, gt, mask, noise2, blend, overlap_len};
  return ew(n, [=] { ddim_step_kernel(p); });
}
extern "C" int emu_undo_step(const float* x, const float* noise, float* out, long long n, float c1, float c2) {
  return ew(n, [=] { undo_step_kernel(x, noise, out, n, c1, c2); });
}
extern "C" int emu_ddpm_step(const float* x, const float* eps, const float* noise, float* out, float* pred_out, long long n, float a,
                             float b, float c1, float c2, float sigma) {
  return ew(n, [=] { ddpm_step_kernel(x, eps, noise, out, pred_out, n, a, b, c1, c2, sigma); });
}
extern "C" int emu_repaint_merge(const float* x, const float* gt, const unsigned char* mask, const float* noise, float* out, long long n,
                                 float c1, float c2) {
  return ew(n, [=] { repaint_merge_kernel(x, gt, mask, noise, out, n, c1, c2); });
}

// mel power spectrogram (csrc/frontend.cuh): one CTA per frame
extern "C" int emu_mel_spectrogram(const float* audio, long long n_samples, int hop, int pad_mode, const float* window, const float* basis,
                                   const int* range, int n_mels, float* out, int n_frames) {
  g_err.clear();
  return emu::run_grid(n_frames, fe::NTHREADS, 1, fe::SMEM_BYTES,
                       [=] { fe::mel_power_kernel(audio, n_samples, hop, pad_mode, window, basis, range, n_mels, out); }, &g_err) ? 0 : 1;
}
