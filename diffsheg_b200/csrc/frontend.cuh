// Audio front-end, the part without network weights: the mel power spectrogram the trainers compute with
//   librosa.feature.melspectrogram(y=aud, sr=18000, hop_length=1200, n_mels=128)
// (trainers/ddpm_show_trainer.py:1063, trainers/ddpm_beat_trainer.py:1244, datasets/beat.py:371; librosa 0.9.2 per
// assets/environment.yml:54) -- inside the reference's published FPS timing (show:1062-1065).  librosa's defaults make this
//   frames of n_fft = 2048 samples every `hop`, centred (the signal padded by 1024 on both sides), periodic Hann window,
//   |rfft|^2 (power = 2.0), then the [n_mels, 1025] Slaney filterbank:  mel[m, f] = sum_k basis[m, k] |X_f[k]|^2.
// One CTA per frame: the windowed frame goes to shared memory in bit-reversed order, 11 radix-2 stages (1024 butterflies each,
// twiddles from a shared-memory table filled with sincospif: exact argument reduction), the power spectrum overwrites the real
// parts, and each mel band sums its own bin range [lo, hi) (the host passes the ranges: a Slaney band is a short triangle).
// 24 KB of shared memory, 256 threads; a 60 s clip is 900 CTAs of ~0.2 MFLOP -- launch-latency sized, which is the point: the
// reference spends ~10 ms of host time per clip here.  fp32 throughout (the reference's float64 FFT is rounded to complex64 before
// the power is taken; the fp32 butterfly network stays within 1e-5 of it relative to the largest band, tests/test_wave_frontend.py).
#pragma once
#include "common.cuh"

namespace dsheg {
namespace fe {

constexpr int NFFT = 2048, LOG2N = 11, NBINS = NFFT / 2 + 1, NTHREADS = 256;
constexpr int SMEM_BYTES = (2 * NFFT + NFFT) * (int)sizeof(float);   // re | im | twiddle (cos, sin) x 1024
enum { PAD_CONSTANT = 0, PAD_REFLECT = 1 };                           // np.pad modes of librosa.stft's `pad_mode`

__device__ __forceinline__ int bit_reverse11(int j) {
  int r = 0;
#pragma unroll
  for (int b = 0; b < LOG2N; ++b) r |= ((j >> b) & 1) << (LOG2N - 1 - b);
  return r;
}

__device__ __forceinline__ void twiddle(int t, float* c, float* s) {   // exp(-2 pi i t / NFFT)
#ifdef DSHEG_EMU
  const double a = -2.0 * 3.14159265358979323846 * (double)t / (double)NFFT;
  *c = (float)cos(a); *s = (float)sin(a);
#else
  sincospif(-(float)t * (2.0f / (float)NFFT), s, c);
#endif
}

// audio [n_samples]; window [NFFT]; basis [n_mels, NBINS]; range [n_mels][2] = first / one-past-last non-zero bin of each band;
// out [n_frames, n_mels] (frame-major: the layout the trainers reach with np.swapaxes, show:1066)
__global__ void __launch_bounds__(NTHREADS) mel_power_kernel(const float* __restrict__ audio, long long n_samples, int hop, int pad_mode,
                                                             const float* __restrict__ window, const float* __restrict__ basis,
                                                             const int* __restrict__ range, int n_mels, float* __restrict__ out) {
#ifdef DSHEG_EMU
  float* sm = reinterpret_cast<float*>(emu::self().cta->smem);
#else
  extern __shared__ float sm[];
#endif
  float* re = sm;
  float* im = sm + NFFT;
  float* twc = sm + 2 * NFFT;          // cos(2 pi t / NFFT), t < 1024
  float* tws = twc + NFFT / 2;         // -sin(2 pi t / NFFT)
  const int tid = threadIdx.x;
  const long long f = blockIdx.x;
  const long long start = f * hop - NFFT / 2;
  for (int j = tid; j < NFFT; j += NTHREADS) {
    long long s = start + j;
    float v = 0.f;
    if (pad_mode == PAD_REFLECT) {     // np.pad(mode='reflect'): the edge sample is not repeated
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
      if (s >= 0 && s < n_samples) v = __ldg(audio + s);
    } else if (s >= 0 && s < n_samples) {
      v = __ldg(audio + s);
    }
    const int r = bit_reverse11(j);
    re[r] = __fmul_rn(v, __ldg(window + j));
    im[r] = 0.f;
  }
  for (int t = tid; t < NFFT / 2; t += NTHREADS) twiddle(t, twc + t, tws + t);
  __syncthreads();
  // decimation in time: stage s joins blocks of half = 2^s points; butterfly (a, b) -> (a + w b, a - w b), w = exp(-2 pi i k / (2 half))
#pragma unroll 1
  for (int s = 0; s < LOG2N; ++s) {
    const int half = 1 << s, tstep = (NFFT / 2) >> s;
    for (int bfl = tid; bfl < NFFT / 2; bfl += NTHREADS) {
      const int k = bfl & (half - 1);
      const int i0 = ((bfl >> s) << (s + 1)) + k, i1 = i0 + half;
      const float wc = twc[k * tstep], ws = tws[k * tstep];
      const float br = re[i1], bi = im[i1];
      const float tr = __fsub_rn(__fmul_rn(wc, br), __fmul_rn(ws, bi));
      const float ti = __fadd_rn(__fmul_rn(wc, bi), __fmul_rn(ws, br));
      const float ar = re[i0], ai = im[i0];
      re[i0] = __fadd_rn(ar, tr); im[i0] = __fadd_rn(ai, ti);
      re[i1] = __fsub_rn(ar, tr); im[i1] = __fsub_rn(ai, ti);
    }
    __syncthreads();
  }
  for (int k = tid; k < NBINS; k += NTHREADS) {
    const float a = re[k], b = im[k];
    re[k] = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));      // |X[k]|^2; bins 0 .. 1024 only (the input is real)
  }
  __syncthreads();
  for (int m = tid; m < n_mels; m += NTHREADS) {
    const int lo = __ldg(range + 2 * m), hi = __ldg(range + 2 * m + 1);
    const float* bm = basis + (size_t)m * NBINS;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(bm + k), re[k], acc);
    out[f * n_mels + m] = acc;
  }
}

}  // namespace fe
}  // namespace dsheg
