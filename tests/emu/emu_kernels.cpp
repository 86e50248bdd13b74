// TEST INFRASTRUCTURE ONLY -- the attention kernel SOURCES of diffsheg_b200/csrc compiled for the host emulator
// (g++ -DDSHEG_EMU) behind a small C ABI for tests/test_emu_kernels.py.  Pointers are host pointers; bf16 travels as uint16.
#include "attn_v3.cuh"
#include "attn_v4.cuh"
#include "attn_v5.cuh"

using namespace dsheg;

static std::string g_err;

extern "C" const char* emu_last_error() { return g_err.c_str(); }

// variant: 3 = attn_v3 (one CTA per sample), 4 = attn_v4 (cluster of two half-sample CTAs), 51 / 52 / 54 = attn_v5<CL = 1 / 2 / 4>
extern "C" int emu_attention(int variant, const uint16_t* qkv, uint16_t* z, int n_samples, int T, int ssB, const float* ln_g,
                             const float* ln_b, const float* ss, int ss_ld) {
  g_err.clear();
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  bf16* zo = reinterpret_cast<bf16*>(z);
  bool ok = false;
  if (variant == 3) {
    ok = emu::run_grid(n_samples, av3::NTHREADS, 1, av3::SMEM_BYTES, [=] { av3::attn_v3_kernel(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else if (variant == 4) {
    ok = emu::run_grid(2 * n_samples, av4::NTHREADS, 2, av4::SMEM_BYTES, [=] { av4::attn_v4_kernel(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else if (variant == 51) {
    ok = emu::run_grid(n_samples, av5::Cfg<1>::NTHREADS, 1, av5::Cfg<1>::SMEM_BYTES, [=] { av5::attn_v5_kernel<1>(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else if (variant == 52) {
    ok = emu::run_grid(2 * n_samples, av5::Cfg<2>::NTHREADS, 2, av5::Cfg<2>::SMEM_BYTES, [=] { av5::attn_v5_kernel<2>(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else if (variant == 54) {
    ok = emu::run_grid(4 * n_samples, av5::Cfg<4>::NTHREADS, 4, av5::Cfg<4>::SMEM_BYTES, [=] { av5::attn_v5_kernel<4>(q, zo, T, ssB, ln_g, ln_b, ss, ss_ld); }, &g_err);
  } else {
    g_err = "unknown attention variant";
  }
  return ok ? 0 : 1;
}
