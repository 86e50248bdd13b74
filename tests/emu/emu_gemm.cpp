// TEST INFRASTRUCTURE ONLY -- diffsheg_b200/csrc/gemm_tc.cuh (kernel AND its host-side launch code: tensor maps, tile walk,
// variant dispatch) compiled for the thread-level emulator behind a C ABI for tests/test_emu_gemm.py.
#include "gemm_tc.cuh"

using namespace dsheg;

struct EmuGemmArgs {
  int32_t M, N, nseg;
  int32_t seg_k[4], seg_ld[4];
  const void* seg_ptr[4];
  const void* w;
  int32_t Kp;
  const float *bias, *csum, *mu, *rstd;
  int32_t act;
  const void* res;
  int32_t ldr, res_mod, res_f32;
  void* out;
  int32_t ldo, out_f32;
  void* out2;
  float* ps_out;
  const float* nullc;
  int32_t n_uncond;
  const float* ps_in;
  const float* cs_in;
  int32_t ps_slots, ps_P;
  int32_t num_sms, bn_force, cg_force;
  const float* eshift;
  int32_t expo_cols;
  const float *lnms_g, *lnms_b, *lnms_ss;
  int32_t lnms_ld, lnms_B, lnms_T;
  int32_t rev;
};

static std::string g_err;
extern "C" const char* emu_gemm_last_error() { return g_err.c_str(); }

extern "C" int emu_gemm_tc(const EmuGemmArgs* a) {
  GemmDesc d;
  d.nseg = a->nseg;
  for (int s = 0; s < a->nseg; ++s) { d.a[s].ptr = a->seg_ptr[s]; d.a[s].ld = a->seg_ld[s]; d.a[s].k = a->seg_k[s]; }
  d.M = a->M; d.N = a->N; d.w = a->w; d.Kp = a->Kp;
  d.bias = a->bias; d.csum = a->csum; d.mu = a->mu; d.rstd = a->rstd; d.act = a->act;
  d.res = a->res; d.ldr = a->ldr; d.res_mod = a->res_mod; d.res_f32 = a->res_f32;
  d.out = a->out; d.ldo = a->ldo; d.out_f32 = a->out_f32; d.out2 = a->out2;
  d.ps_out = reinterpret_cast<float2*>(a->ps_out); d.nullc = a->nullc; d.n_uncond = a->n_uncond;
  d.ps_in = reinterpret_cast<const float2*>(a->ps_in); d.cs_in = reinterpret_cast<const float2*>(a->cs_in);
  d.ps_slots = a->ps_slots; d.ps_P = a->ps_P;
  d.eshift = a->eshift; d.expo_cols = a->expo_cols;
  d.lnms_g = a->lnms_g; d.lnms_b = a->lnms_b; d.lnms_ss = a->lnms_ss; d.lnms_ld = a->lnms_ld; d.lnms_B = a->lnms_B; d.lnms_T = a->lnms_T;
  d.rev = a->rev;
  std::string terr;
  const cudaError_t e = tc::launch_gemm_tc(d, a->num_sms, nullptr, &terr, a->bn_force, a->cg_force);
  if (e != cudaSuccess) { g_err = terr + " " + tc::g_emu_error(); return 1; }
  g_err.clear();
  return 0;
}
