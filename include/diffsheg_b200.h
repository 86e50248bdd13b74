/*
 * diffsheg_b200 -- C ABI of the B200-native DiffSHEG sampling hot path.
 *
 * One shared library (libdiffsheg_b200.so, sm_100a only).  Plain pointers and sizes: no
 * torch / C++ types cross this boundary.  All data pointers are DEVICE pointers unless the
 * argument name starts with `host_`; `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success and a non-zero status otherwise; the message is
 * available from dsheg_last_error().  Nothing here throws, aborts or synchronises the
 * device (except dsheg_load_tensor / dsheg_finalize_weights, which are set-up calls).
 *
 * The reference (JeremyCJM/DiffSHEG, pure Python) has no FFI; each entry point replaces
 * the Python code cited beside it (paths relative to the reference repository).
 */
#ifndef DIFFSHEG_B200_H_
#define DIFFSHEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Version 2 appends cond_projection / no_cond_residual to dsheg_config; a version-1 caller (abi_version = 1, the 15-field struct)
 * is still accepted and gets the shipped defaults. */
#define DSHEG_ABI_VERSION 2

/* FP32: strict parity mode (SIMT fp32 GEMMs).  BF16: performance mode (tcgen05 bf16 GEMMs, bf16 activations).
 * TF32: fp32 activations / residual stream / statistics / epilogues with tcgen05 kind::tf32 GEMMs (operands rounded to
 *       TF32 by the TMA load, fp32 accumulation) -- the precise-and-fast middle mode. */
enum { DSHEG_PREC_FP32 = 0, DSHEG_PREC_BF16 = 1, DSHEG_PREC_TF32 = 2 };
enum { DSHEG_DTYPE_F32 = 0, DSHEG_DTYPE_BF16 = 1 };
/* opt.cond_projection (options/base_options.py:21; models/transformer.py:262-263,281-289,304-324): what feat_proj of every
 * MotionTransformer layer consumes and computes.  *_INCLUDEX: cat(x, audio, hubert[, expression]); *_EXCLUDEX: the conditioning
 * only (the input is then always added back, transformer.py:302,337).  MLP: LayerNorm -> Linear -> SiLU -> Linear; LINEAR: one
 * Linear.  'none' is not offered: the reference's own layers raise NotImplementedError for it (transformer.py:323-324). */
enum { DSHEG_COND_MLP_INCLUDEX = 0, DSHEG_COND_LINEAR_INCLUDEX = 1, DSHEG_COND_MLP_EXCLUDEX = 2, DSHEG_COND_LINEAR_EXCLUDEX = 3 };

/* Frozen subset of the reference `opt` Namespace + build_models() arguments
 * (runner.py:32-45, runner.py:124-222, options/base_options.py:16-128). */
typedef struct dsheg_config {
  int32_t abi_version;     /* DSHEG_ABI_VERSION */
  int32_t dim_pose;        /* gesture channels: 129 (SHOW) / 141 (BEAT); opt.split_pos */
  int32_t expression_dim;  /* 103 / 51 */
  int32_t audio_dim;       /* mel channels, 128 */
  int32_t hubert_dim;      /* 1024 */
  int32_t aud_latent_dim;  /* 256 */
  int32_t latent_dim;      /* 512 */
  int32_t num_layers;      /* 8 */
  int32_t num_heads;       /* 8 */
  int32_t ff_size;         /* 1024 */
  int32_t style_dim;       /* 4 / 30 */
  int32_t classifier_free; /* opt.classifier_free */
  int32_t precision;       /* DSHEG_PREC_* : arithmetic of the per-step GEMMs/activations */
  int32_t max_batch;       /* workspace is sized for max_batch x max_frames (x2 under CFG) */
  int32_t max_frames;
  /* ---- ABI version 2 (zero = the shipped configuration) ---- */
  int32_t cond_projection;  /* DSHEG_COND_*; 0 = mlp_includeX (the default of options/base_options.py:21) */
  int32_t no_cond_residual; /* 1 = opt.cond_residual False: feat_proj's result replaces x instead of being added to it
                               (transformer.py:302,337; ignored by the *_EXCLUDEX projections, which always add) */
} dsheg_config;

typedef struct dsheg_handle dsheg_handle;

/* Last error message of `h` (or of the last failed dsheg_create when h == NULL). */
const char* dsheg_last_error(const dsheg_handle* h);

/* Replaces UniDiffuser.__init__ (models/transformer.py:590-699) + .to(device).  */
int dsheg_create(const dsheg_config* cfg, int device, dsheg_handle** out);
void dsheg_destroy(dsheg_handle* h);

/* Weight upload: packed tensors by name (the packed contract is documented in
 * diffsheg_b200/pack.py, built from checkpoint['encoder'], trainers/ddpm_show_trainer.py:
 * 278-292).  host_data is HOST memory, copied synchronously. */
int dsheg_load_tensor(dsheg_handle* h, const char* key, const void* host_data, int32_t dtype,
                      const int64_t* shape, int32_t ndim);
/* Resolve every tensor the configured network needs; fails listing the first missing key. */
int dsheg_finalize_weights(dsheg_handle* h);

/* Per-window, step-invariant work: person-id MLP (transformer.py:453-457,559),
 * hubert_encoder conv stack for both nets (transformer.py:436-442,512-515), mel staging.
 * mel [B,T,audio_dim], hubert [B,T,hubert_dim], person_id [B,style_dim], fp32 contiguous. */
int dsheg_prepare_window(dsheg_handle* h, const float* mel, const float* hubert, const float* person_id,
                         int32_t B, int32_t T, void* stream);

/* One denoiser call: UniDiffuser.forward (transformer.py:728-770) for a batch-uniform
 * ORIGINAL timestep t_orig (gaussian_diffusion.py:1196 builds t = [i]*B; respace.py:119-124
 * maps it), a = sqrt_recip_alphas_cumprod[t], b = sqrt_recipm1_alphas_cumprod[t]
 * (gaussian_diffusion.py:527-532).  x, eps_out: [B,T,dim_pose+expression_dim] fp32. */
int dsheg_denoise(dsheg_handle* h, const float* x, int32_t t_orig, float a, float b, float cond_scale,
                  float* eps_out, void* stream);

/* Kernel launches issued by this handle since creation (bench.py reports the delta). */
int64_t dsheg_launch_count(const dsheg_handle* h);

/* Per-kernel-class device timing for bench.py's roofline pass: between begin and end every GEMM,
 * attention and row-wise (LayerNorm/statistics) launch is bracketed by CUDA events on the launching
 * stream.  end() synchronises and returns, for class c in {0 GEMM, 1 attention, 2 row-wise}:
 * ms[c] summed kernel time, work[c] algorithmic FLOPs (c = 0) or bytes (c = 1, 2), count[c] launches. */
int dsheg_profile_begin(dsheg_handle* h);
int dsheg_profile_end(dsheg_handle* h, double* ms, double* work, int64_t* count);

/* ---- stateless sampler-step kernels (fp32, elementwise, HBM-bound) ---------------------- */

/* ddim_sample, eta = 0 (gaussian_diffusion.py:976-1066): x_out = DDIM update of (x, eps),
 * then, when gt/mask are given, the RePaint merge :1034-1056 with noise2 and the linear
 * overlap blend (blend != 0, first overlap_len frames).  n = B*T*D elements, frame stride
 * D, T frames per sample.  gt/mask/noise2 may be NULL (no repaint).  mask: 1 byte/element.
 * gt and mask given but noise2 NULL: gt already holds the NOISY known frames (--same_overlap_noisy, :1040-1042: the tail the
 * previous window saved at this very step) and is merged as is. */
int dsheg_ddim_step(const float* x, const float* eps, float* x_out, int64_t n, int32_t T, int32_t D,
                    float sqrt_recip_ac, float sqrt_recipm1_ac, float sqrt_ac_prev,
                    float sqrt_one_minus_ac_prev, const float* gt, const uint8_t* mask,
                    const float* noise2, int32_t blend, int32_t overlap_len, float* pred_xstart_out,
                    void* stream);

/* _undo (gaussian_diffusion.py:467-473): x_out = sqrt(1-beta) x + sqrt(beta) noise. */
int dsheg_undo_step(const float* x, const float* noise, float* x_out, int64_t n, float sqrt_one_minus_beta,
                    float sqrt_beta, void* stream);

/* p_sample (gaussian_diffusion.py:747-774 with q_posterior :475-497): posterior mean from
 * (x, eps) + nonzero * exp(0.5 logvar) * noise. */
int dsheg_ddpm_step(const float* x, const float* eps, const float* noise, float* x_out, int64_t n,
                    float sqrt_recip_ac, float sqrt_recipm1_ac, float coef1, float coef2,
                    float sigma /* exp(0.5*logvar) or 0 at t == 0 */, float* pred_xstart_out, void* stream);

/* p_sample's harmonize pre-merge (gaussian_diffusion.py:727-745):
 * x_out = mask ? sqrt_ac*gt + sqrt_one_minus_ac*noise : x. */
int dsheg_repaint_merge(const float* x, const float* gt, const uint8_t* mask, const float* noise,
                        float* x_out, int64_t n, float sqrt_ac, float sqrt_one_minus_ac, void* stream);

/* ---- output post-processing on the resident sample -------------------------- */

/* datasets/show.py:157-162 inv_standardize as trainers/ddpm_show_trainer.py:913-918 applies it to the sampled motion:
 * out[r,c] = x[r,c] * std[c] + mean[c].  x / out are fp32 [rows, D] with row strides ldx / ldo (so the gesture /
 * expression split of show:920-921 is a column window: pass x + split_pos, D = expression_dim, ldx = net_dim_pose). */
int dsheg_inv_standardize(const float* x, int32_t ldx, const float* mean, const float* std, float* out, int32_t ldo,
                          int64_t rows, int32_t D, void* stream);

/* trainers/ddpm_beat_trainer.py:1056-1062 (--axis_angle): de-normalise the axis-angle gesture, convert every joint with
 * datasets/rotation_converter.py:282-296 (axis-angle -> quaternion -> matrix -> XYZ Euler), radians -> degrees, and
 * re-normalise with the Euler statistics.  x fp32 [rows, >= C] (row stride ldx), C = 3 * joints.
 * euler_deg [rows, C] (what result2target_vis writes to BVH, beat:1076) and out_norm [rows, C] (the saved .npy,
 * beat:1061) are dense; either may be NULL. */
int dsheg_beat_axis_angle_to_euler(const float* x, int32_t ldx, const float* mean_aa, const float* std_aa,
                                   const float* mean_pose, const float* std_pose, float* euler_deg, float* out_norm,
                                   int64_t rows, int32_t C, void* stream);

/* Audio front-end, the part without network weights: HuBERT features [B, n_in, C] resampled to the motion
 * frame rate [B, n_out, C] -- F.interpolate(mode='linear', align_corners=True) along the frame axis
 * (trainers/ddpm_show_trainer.py:1082, datasets/show.py:98, datasets/beat.py:445).  fp32, C % 4 == 0. */
int dsheg_resample_linear(const float* in, float* out, int32_t B, int32_t n_in, int32_t n_out, int32_t C, void* stream);

/* Audio front-end, mel half: librosa.feature.melspectrogram(y=aud, sr=18000, hop_length=1200, n_mels=128) as the
 * trainers call it (trainers/ddpm_show_trainer.py:1063, trainers/ddpm_beat_trainer.py:1244, datasets/beat.py:371): centred frames of
 * n_fft = 2048 samples every `hop`, times `window` [n_fft], |rfft|^2, times the filterbank mel_basis [n_mels, n_fft/2 + 1] whose band m
 * is non-zero on bins mel_range[2m] .. mel_range[2m+1]-1.  pad_mode: 0 = zeros ('constant'), 1 = 'reflect' (np.pad modes of
 * librosa.stft).  audio fp32 [n_samples]; out fp32 [n_frames, n_mels], frame-major (the layout after np.swapaxes, show:1066),
 * n_frames <= 1 + n_samples / hop.  Window and filterbank are the caller's (diffsheg_b200/frontend.py builds librosa's). */
int dsheg_mel_spectrogram(const float* audio, int64_t n_samples, int32_t n_fft, int32_t hop, int32_t pad_mode, const float* window,
                          const float* mel_basis, const int32_t* mel_range, int32_t n_mels, float* out, int32_t n_frames, void* stream);

/* ---- op-level entry points used by the parity tests -------------------------------------- */

/* out[M,N] = act(A[M,K] W[N,K]^T + bias) (+ residual); precision selects the SIMT fp32, the
 * tcgen05 bf16 or the tcgen05 tf32 engine (fp32 in / out like the SIMT one).  A/W/out/residual are fp32 device arrays; the bf16 engine converts
 * operands, residual and (when N % 32 == 0) the output to bf16 internally, exactly the layout the
 * denoiser uses (test path only).  act: 0 none, 1 SiLU, 2 GELU. */
int dsheg_op_linear(int32_t precision, const float* A, const float* W, const float* bias,
                    const float* residual, float* out, int32_t M, int32_t N, int32_t K, int32_t act,
                    void* stream);

/* The fused epilogue modes of the tcgen05 engine on one GEMM (test path only; bf16 operands / output, N % 64 == 0, K % 64 == 0):
 *   mode 4: out = rstd (A W^T - mu csum) + bias with its leading i0 columns written as exp(. - eshift[n]) -- the softmax
 *           numerators of transformer.py:122-123 with static shifts; aux0 = mu [M], aux1 = rstd [M], aux2 = eshift [i0],
 *           aux3 = csum [N] (row sums of the bf16-rounded W);
 *   mode 5: out = SiLU(LayerNorm_N(A W^T + bias) (1 + scale) + shift) -- ffn.linear2 + StylizationBlock prologue
 *           (transformer.py:178-181, 92-96), N == 512 (CTA pairs for M >= 4096 and K >= 768, single CTAs otherwise); aux0 = gamma [N], aux1 = beta [N],
 *           aux2 = [i1][i0] table (scale at [0,N), shift at [N,2N), row = (m / i2) mod i1), i0 = row stride, i2 = frames per sample. */
int dsheg_op_linear_fused(int32_t mode, const float* A, const float* W, const float* bias, const float* aux0,
                          const float* aux1, const float* aux2, const float* aux3, float* out, int32_t M, int32_t N,
                          int32_t K, int32_t i0, int32_t i1, int32_t i2, void* stream);

/* Device timing of one tcgen05 GEMM shape (tuning / roofline aid): mode 0 bias, 1 LN-fold+bias,
 * 2 LN-fold+bias+SiLU, 3 bias+bf16 residual, 4 bias+GELU, 5 LN-fold+bias with exponential leading columns (ACT_EXPO),
 * 6 bias + full-row LayerNorm / modulate / SiLU (ACT_LNMS, N == 512); bn 0 = auto, 128 or 256. */
int dsheg_bench_gemm(int32_t M, int32_t N, int32_t K, int32_t mode, int32_t bn, int32_t iters, float* ms_out);

/* Linear self-attention core + Stylization prologue on one [Bn,T,3D] qkv tensor
 * (transformer.py:122-128 then :92-96 up to the SiLU): z = SiLU(LN(y)*(1+scale)+shift). */
int dsheg_op_attention(const float* qkv, const float* ln_g, const float* ln_b, const float* scale_shift,
                       float* z, int32_t Bn, int32_t T, int32_t D, int32_t H, void* stream);

/* Same op as the tf32 mode runs it (attn_tf32.cuh: fp32 arrays, heads of 64; the two products on mma.sync TF32, softmaxes, sums and
 * LayerNorm exact fp32). */
int dsheg_op_attention_tf32(const float* qkv, const float* ln_g, const float* ln_b, const float* scale_shift,
                            float* z, int32_t Bn, int32_t T, int32_t D, int32_t H, void* stream);

/* Same op on the bf16 fast path (D = 512, 8 heads, T <= 96): qkv and z are bf16 arrays.
 * numerators = 1: the engine's default kernel (attn_ws.cuh: persistent, TMA-staged, warp-specialised); the Q and K columns of qkv already hold the
 *                 softmax NUMERATORS exp(value - shift) that the ACT_EXPO epilogue of the QKV GEMM writes (any per-(row, head)
 *                 shift for Q, any per-(sample, column) shift for K: softmax is shift-invariant);
 * numerators = 0: plain q, k, v; both softmaxes run inside the kernel (attn_v3.cuh, the per-layer fallback). */
int dsheg_op_attention_bf16(const void* qkv, const float* ln_g, const float* ln_b, const float* scale_shift,
                            void* z, int32_t Bn, int32_t T, int32_t numerators, void* stream);

/* LinearTemporalCrossAttention core + Stylization prologue (transformer.py:133-166, then :92-96 up to the SiLU), bf16:
 * q [Bn, T, 512] holds the Q numerators exp(query - shift), kv [Bn, N, 1024] = (K numerators exp(key - shift) | value) of the
 * conditioning sequence (N frames, any N <= 96), z [Bn, T, 512].  Same kernel as numerators = 1 above with separate sources. */
int dsheg_op_cross_attention_bf16(const void* q, const void* kv, const float* ln_g, const float* ln_b,
                                  const float* scale_shift, void* z, int32_t Bn, int32_t T, int32_t N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSHEG_B200_H_ */
