// Output post-processing on the device: what the trainers do to the sampled motion in numpy / torch-CPU
// after a per-window D2H copy, done here on the resident [rows, D] fp32 sample so that ONE D2H of the final arrays remains.
//   * inv_standardize          datasets/show.py:157-162 as used by trainers/ddpm_show_trainer.py:913-918
//   * BEAT axis-angle branch   trainers/ddpm_beat_trainer.py:1056-1062 with datasets/rotation_converter.py:204-233
//                              (axis-angle -> quaternion), :251-280 (-> matrix), :342-385 / :299-329 (-> XYZ Euler)
// Elementwise and HBM-bound (8-12 B per element); the arithmetic keeps the reference's operation order with explicit
// round-to-nearest mul/add (no FMA contraction) and the precise sinf/cosf/atan2f/asinf of libdevice.
#pragma once
#include "common.cuh"

namespace dsheg {

// out[r, c] = x[r, c] * std[c] + mean[c]   (show.py:159); x may be a column window of a wider tensor (ldx)
__global__ void inv_standardize_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                       const float* __restrict__ stdv, float* __restrict__ out, int ldo, long long rows, int D) {
  const long long n = rows * D, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = i / D;
    const int c = (int)(i - r * D);
    out[r * ldo + c] = __fadd_rn(__fmul_rn(x[r * ldx + c], __ldg(stdv + c)), __ldg(mean + c));
  }
}

// one thread per (row, joint): 3 normalised axis-angle channels in, 3 Euler degrees + 3 re-normalised values out
__global__ void beat_axis_angle_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mean_aa,
                                       const float* __restrict__ std_aa, const float* __restrict__ mean_pose,
                                       const float* __restrict__ std_pose, float* __restrict__ euler_deg,
                                       float* __restrict__ out_norm, long long rows, int joints) {
  const long long n = rows * joints, stride = (long long)gridDim.x * blockDim.x;
  const int C = joints * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = i / joints;
    const int c0 = (int)(i - r * joints) * 3;
    const float* xp = x + r * ldx + c0;
    // denorm_out = out_motions * std_pose_axis_angle + mean_pose_axis_angle                       (beat:1057)
    const float vx = __fadd_rn(__fmul_rn(xp[0], __ldg(std_aa + c0)), __ldg(mean_aa + c0));
    const float vy = __fadd_rn(__fmul_rn(xp[1], __ldg(std_aa + c0 + 1)), __ldg(mean_aa + c0 + 1));
    const float vz = __fadd_rn(__fmul_rn(xp[2], __ldg(std_aa + c0 + 2)), __ldg(mean_aa + c0 + 2));
    // axis-angle -> quaternion (rc:217-232)
    const float angle = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    const float half = __fmul_rn(angle, 0.5f);
    const float s = angle < 1e-6f ? __fsub_rn(0.5f, __fdiv_rn(__fmul_rn(angle, angle), 48.f)) : __fdiv_rn(sinf(half), angle);
    const float qr = cosf(half), qi = __fmul_rn(vx, s), qj = __fmul_rn(vy, s), qk = __fmul_rn(vz, s);
    // quaternion -> the five matrix entries XYZ Euler extraction reads (rc:262-277)
    const float two_s = __fdiv_rn(2.0f, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qi, qi)), __fmul_rn(qj, qj)), __fmul_rn(qk, qk)));
    const float m00 = __fsub_rn(1.f, __fmul_rn(two_s, __fadd_rn(__fmul_rn(qj, qj), __fmul_rn(qk, qk))));
    const float m01 = __fmul_rn(two_s, __fsub_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
    const float m02 = __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
    const float m12 = __fmul_rn(two_s, __fsub_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
    const float m22 = __fsub_rn(1.f, __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qj, qj))));
    // matrix -> Euler 'XYZ' (rc:364-384): atan2(-M12, M22), asin(M02), atan2(-M01, M00); NaN for |M02| > 1 like torch.asin
    float e[3] = {atan2f(-m12, m22), asinf(m02), atan2f(-m01, m00)};
    const float RAD2DEG = 57.29577951308232f;                                                       // beat:1060
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float deg = __fmul_rn(e[k], RAD2DEG);
      if (euler_deg) euler_deg[r * C + c0 + k] = deg;
      if (out_norm) out_norm[r * C + c0 + k] = __fdiv_rn(__fsub_rn(deg, __ldg(mean_pose + c0 + k)), __ldg(std_pose + c0 + k));   // beat:1061
    }
  }
}

// ---- SURVEY 8 row f1 (the part that needs no network weights): HuBERT features resampled to the motion frame rate ----------------
// F.interpolate(x.swapaxes(-1,-2), size=n_out, mode='linear', align_corners=True).swapaxes(-1,-2)  (show:1082, datasets/show.py:98,
// datasets/beat.py:445): out[b, i, c] = (1 - l) in[b, i0, c] + l in[b, i0 + 1, c], src = i (n_in - 1) / (n_out - 1), i0 = floor(src),
// l = src - i0 -- the operation order of ATen's upsample_linear1d (fp32 index arithmetic).  One thread per 4 channels.
__global__ void resample_linear_kernel(const float* __restrict__ in, float* __restrict__ out, int n_in, int n_out, int C, long long total4) {
  const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  const int C4 = C >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int c4 = (int)(i % C4);
    const long long r = i / C4;
    const int io = (int)(r % n_out);
    const long long b = r / n_out;
    const float src = __fmul_rn(scale, (float)io);
    const int i0 = (int)src;
    const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    const float l1 = __fsub_rn(src, (float)i0), l0 = __fsub_rn(1.0f, l1);
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + (b * n_in + i0) * (long long)C) + c4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(in + (b * n_in + i1) * (long long)C) + c4);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(l0, a.x), __fmul_rn(l1, d.x)); o.y = __fadd_rn(__fmul_rn(l0, a.y), __fmul_rn(l1, d.y));
    o.z = __fadd_rn(__fmul_rn(l0, a.z), __fmul_rn(l1, d.z)); o.w = __fadd_rn(__fmul_rn(l0, a.w), __fmul_rn(l1, d.w));
    reinterpret_cast<float4*>(out + (b * n_out + io) * (long long)C)[c4] = o;
  }
}

}  // namespace dsheg
