// TEST INFRASTRUCTURE ONLY -- diffsheg_b200/csrc/gemm_tf32.cuh (kernel AND its host-side launch code: TFLOAT32 tensor maps, segment
// bookkeeping, CTA-pair dispatch) compiled for the thread-level emulator behind a C ABI for tests/test_emu_gemm.py.
#include "gemm_tf32.cuh"

using namespace dsheg;

struct EmuGemmTf32Args {
  int32_t M, N, nseg;
  int32_t seg_k[4], seg_ld[4];
  const float* seg_ptr[4];
  const float* w;
  int32_t Kp;
  const float *bias, *csum, *mu, *rstd;
  int32_t act;
  const float* res;
  int32_t ldr, res_mod;
  float* out;
  int32_t ldo;
  float* out2;
  int32_t cg_force, num_sms;
};

static std::string g_err;
extern "C" const char* emu_gemm_tf32_last_error() { return g_err.c_str(); }

extern "C" int emu_gemm_tf32(const EmuGemmTf32Args* a) {
  GemmDesc d;
  d.nseg = a->nseg;
  for (int s = 0; s < a->nseg; ++s) { d.a[s].ptr = a->seg_ptr[s]; d.a[s].ld = a->seg_ld[s]; d.a[s].k = a->seg_k[s]; }
  d.M = a->M; d.N = a->N; d.w = a->w; d.Kp = a->Kp;
  d.bias = a->bias; d.csum = a->csum; d.mu = a->mu; d.rstd = a->rstd; d.act = a->act;
  d.res = a->res; d.ldr = a->ldr; d.res_mod = a->res_mod; d.res_f32 = 1;
  d.out = a->out; d.ldo = a->ldo; d.out_f32 = 1; d.out2 = a->out2;
  if (!t32::tf32_eligible(d)) { g_err = "not eligible for the tf32 TMA path"; return 2; }
  std::string terr;
  const cudaError_t e = t32::launch_gemm_tf32(d, a->num_sms, nullptr, &terr, a->cg_force);
  if (e != cudaSuccess) { g_err = terr + " " + t32::g_emu_error_tf32(); return 1; }
  g_err.clear();
  return 0;
}
