/*
 * b2v.h -- C ABI of libb2v.so, the B200 (sm_100a) implementation of the latent-diffusion sampling hot path of
 * Kkuntal990/video-to-video-diffusion.  This is the drop-in boundary: the reference has no FFI layer of its own
 * (pure PyTorch), so each entry point below states the reference Python interface it stands behind
 * (paths are relative to the reference repository root).  The thin Python mirror of the reference classes in
 * video-to-video-diffusion_b200/ binds exactly these symbols through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer to contiguous fp32 in the reference's own layout (NCDHW)
 *     unless stated otherwise; timesteps are int64 like the reference's `t`;
 *   - the caller owns all I/O buffers; an object owns its repacked fp16 weights, workspaces and CUDA graphs;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), ordered by it; an object is bound to
 *     the device current at creation and is not re-entrant;
 *   - return value 0 = success, <0 = error; b2v_last_error() returns the message of the calling thread's last error;
 *   - there is no CPU fallback: every entry point fails if no sm_100 device is present.
 */
#ifndef B2V_H
#define B2V_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2V_ABI_VERSION 2

const char* b2v_last_error(void);
int b2v_abi_version(void);
/* number of kernels launched by this library in the calling process so far (bench.py's gpu_launches) */
long long b2v_launch_count(void);

/* ---------------------------------------------------------------- U-Net denoiser ---------------------------
 * models/unet3d.py:227-413  UNet3D(latent_dim, model_channels, num_res_blocks, attention_levels, channel_mult,
 *                                   num_heads, time_embed_dim).forward(x, t, c)                                  */
typedef struct b2v_unet b2v_unet;
typedef struct {
  int latent_dim;
  int model_channels;
  int num_res_blocks;
  int num_levels;
  int channel_mult[8];
  int attention_mask; /* bit l set <=> level l has TemporalAttention (attention_levels) */
  int num_heads;      /* accepted for parity; the reference's attention output does not depend on it */
  int time_embed_dim;
} b2v_unet_desc;

int b2v_unet_create(b2v_unet** out, const b2v_unet_desc* desc);
void b2v_unet_destroy(b2v_unet* u);
/* state_dict entry by its reference key (e.g. "down_blocks.0.0.0.conv1.conv.weight"); data = HOST fp32, copied */
int b2v_unet_load_weight(b2v_unet* u, const char* key, const float* data, const int64_t* shape, int ndim);
/* repack all weights (fp16 K-major tiles, folded attention); fails listing the first missing key */
int b2v_unet_finalize(b2v_unet* u);
/* UNet3D.forward (models/unet3d.py:357): x, c: (B,L,T,h,w); t: (B,) int64 device; eps_out: (B,L,T,h,w) */
int b2v_unet_forward(b2v_unet* u, const float* x, const int64_t* t, const float* c, float* eps_out, int B, int T,
                     int h, int w, void* stream);

/* ---------------------------------------------------------------- samplers ---------------------------------
 * inference/sampler.py:241-336  DDIMSampler.sample : the whole loop, one CUDA-graph replay per step, no host sync.
 *   z_init      : the reference's initial torch.randn(shape) draw (made by the Python layer so seeds agree)
 *   timesteps   : HOST int64[n], the reversed subset from DDIMSampler._get_timesteps (:221-239)
 *   alphas_cumprod : HOST fp32[n_train], GaussianDiffusion.alphas_cumprod (models/diffusion.py:49)
 *   noise       : NULL for eta == 0, else DEVICE fp32 [n][numel(z)] = the per-step torch.randn_like draws
 *   the five NaN/Inf guards of the reference are applied on device; *nan_flag (DEVICE int, may be NULL) is set
 *   if any of them fired.                                                                                     */
int b2v_ddim_sample(b2v_unet* u, const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w,
                    const int64_t* timesteps, int n, const float* alphas_cumprod, int n_train, float eta,
                    const float* noise, int* nan_flag, void* stream);
/* models/diffusion.py:270-367 / inference/sampler.py:35-61  DDPM ancestral sampling, step-wise so that the caller
 * can supply the reference's per-step torch.randn_like draw:
 *   begin(z_init, cond) ; for t = T-1..0: step(t, coef[8], noise) ; end(z_out)
 *   coef = {sqrt_one_minus_alphas_cumprod[t], sqrt_alphas_cumprod[t], posterior_mean_coef1[t],
 *           posterior_mean_coef2[t], (t != 0), exp(0.5*posterior_log_variance_clipped[t]), 0, 0}  (HOST fp32)   */
int b2v_sampler_begin(b2v_unet* u, const float* z_init, const float* cond, int B, int T, int h, int w, void* stream);
int b2v_ddpm_step(b2v_unet* u, int64_t t, const float* coef, const float* noise, void* stream);
int b2v_sampler_end(b2v_unet* u, float* z_out, void* stream);
/* models/diffusion.py:340-367  GaussianDiffusion.p_sample_loop / inference/sampler.py:35-61 DDPMSampler.sample: the
 * WHOLE ancestral loop in one call -- one CUDA-graph replay per step (U-Net + fused posterior update), the timestep,
 * coefficient row and time-embedding row selected by a device-side step counter, no host synchronisation.
 *   coef  : HOST fp32 [n][8], row t = the b2v_ddpm_step coefficients of TIMESTEP t; loop step s runs t = n-1-s
 *   noise : DEVICE fp32 [n][numel(z)], row s = the torch.randn_like draw of loop step s (what the Python layer
 *           passes so that a seed reproduces the reference's noise stream), or NULL: N(0,1) draws are generated on
 *           the device from `seed` (Philox4x32-10 keyed by seed, counter = (element / 4, step), Box-Muller)
 * b2v_ddpm_run is the same loop in chunks, between b2v_sampler_begin / b2v_sampler_end, for callers whose per-step
 * noise does not fit in memory at once: steps [first, first + count), noise = [count][numel]; chunks must be
 * contiguous and start at first = 0 (which uploads coef and computes the time embeddings of all n steps).      */
int b2v_ddpm_sample(b2v_unet* u, const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w,
                    const float* coef, int n, const float* noise, uint64_t seed, void* stream);
int b2v_ddpm_run(b2v_unet* u, const float* coef, int n, int first, int count, const float* noise, uint64_t seed,
                 void* stream);
/* inference/sampler.py:221-239  DDIMSampler._get_timesteps: every (n_train // steps)-th timestep plus n_train-1 if
 * missing, descending; writes at most cap entries to HOST out and returns the count (steps or steps + 1), <0 on error */
int b2v_ddim_timesteps(int n_train, int num_inference_steps, int64_t* out, int cap);
/* out[i] = the i-th N(0,1) draw of loop step `step` of the device generator b2v_ddpm_sample uses (DEVICE fp32 [n])  */
int b2v_philox_normal(float* out, uint64_t seed, int step, long long n, void* stream);

/* ---------------------------------------------------------------- VAE ---------------------------------------
 * models/vae.py:207-306  SliceInterpolationVAE(in_channels, latent_dim, base_channels, scaling_factor)        */
typedef struct b2v_vae b2v_vae;
typedef struct {
  int in_channels;
  int latent_dim;
  int base_channels;
  float scaling_factor;
} b2v_vae_desc;

int b2v_vae_create(b2v_vae** out, const b2v_vae_desc* desc);
void b2v_vae_destroy(b2v_vae* v);
int b2v_vae_load_weight(b2v_vae* v, const char* key, const float* data, const int64_t* shape, int ndim);
int b2v_vae_finalize(b2v_vae* v);
/* encode (models/vae.py:235-247): x (B,Cin,T,H,W) -> z (B,L,T,H/4,W/4), already multiplied by scaling_factor */
int b2v_vae_encode(b2v_vae* v, const float* x, float* z, int B, int T, int H, int W, void* stream);
/* decode (models/vae.py:249-260): z (B,L,T,h,w) -> x (B,Cin,T,4h,4w), tanh-bounded */
int b2v_vae_decode(b2v_vae* v, const float* z, float* x, int B, int T, int h, int w, void* stream);

/* ---------------------------------------------------------------- end to end -------------------------------
 * models/model.py:230-343  VideoToVideoDiffusion.generate(v_in, sampler, num_inference_steps, target_depth):
 *   NaN check of v_in -> vae.encode -> guard -> trilinear depth upsample T_in -> T_out (skipped when equal) -> guard
 *   -> sampler over the U-Net -> guard -> vae.decode -> guard, all on `stream`, no host synchronisation; the
 *   reference's five NaN/Inf checkpoints run on the device and set *nan_flag (DEVICE int, may be NULL).
 *   v_in (B,Cin,T_in,H,W) -> v_out (B,Cin,T_out,H,W); z_init (B,L,T_out,H/4,W/4) = the sampler's initial randn draw.
 * The schedule is passed as host tables so that the caller's GaussianDiffusion buffers are the single source.      */
typedef struct {
  int sampler;                 /* 0 = DDIM (b2v_ddim_sample), 1 = DDPM (b2v_ddpm_sample) */
  int n;                       /* DDIM: entries of `timesteps`; DDPM: number of training timesteps */
  const int64_t* timesteps;    /* DDIM: HOST int64[n] */
  const float* alphas_cumprod; /* DDIM: HOST fp32[n_train] */
  int n_train;
  float eta;                   /* DDIM */
  const float* ddpm_coef;      /* DDPM: HOST fp32 [n][8] */
  const float* noise;          /* DEVICE per-step draws [n][numel(z)] or NULL (see b2v_ddim_sample / b2v_ddpm_sample) */
  uint64_t seed;               /* DDPM with noise == NULL */
} b2v_sampler_cfg;
int b2v_generate(b2v_unet* u, b2v_vae* v, const b2v_sampler_cfg* cfg, const float* v_in, const float* z_init,
                 float* v_out, int B, int T_in, int T_out, int H, int W, int* nan_flag, void* stream);

/* models/diffusion.py:81-190  training forward (no backward): q_sample and the per-sample terms of training_loss.
 *   b2v_q_sample : z_t = sqrt_alphas_cumprod[t_b] * z_0 + sqrt_one_minus_alphas_cumprod[t_b] * noise
 *                  (t: DEVICE int64 [B]; the two tables: DEVICE fp32 [n_train]; per_sample = numel / B)
 *   b2v_eps_mse  : out[b] = (sum mask*(eps_pred-noise)^2, sum mask) per sample (DEVICE fp32 [B][2]); mask: DEVICE fp32
 *                  (B, C, T) broadcast over the HW innermost elements, or NULL (all ones).  Deterministic reduction.
 * The Min-SNR-5 weighting and the batch mean are B-element arithmetic done by the caller (models/diffusion.py:148-190). */
int b2v_q_sample(const float* z0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1m_ac,
                 float* zt, int B, long long per_sample, void* stream);
size_t b2v_eps_mse_ws_bytes(int B); /* bytes of DEVICE workspace b2v_eps_mse needs for batch B (8-byte aligned) */
int b2v_eps_mse(const float* eps_pred, const float* noise, const float* mask, float* out, void* ws, size_t ws_bytes,
                int B, long long per_sample, long long HW, void* stream);

/* ---------------------------------------------------------------- glue ops ----------------------------------
 * models/model.py:284-289  F.interpolate(z, (Dout, h, w), 'trilinear', align_corners=False) with h, w unchanged */
int b2v_upsample_depth(const float* in, float* out, int BC, int Din, int Dout, int HW, void* stream);

/* inference/sampler.py:379-451  sliding-window stitching: acc[bc, d0+d, h0+h, w0+w] += patch * gd[d]*gh[h]*gw[w],
 * wsum likewise (patch: (BC,pd,ph,pw); acc, wsum: (BC,D,H,W); gd/gh/gw: the separable Gaussian window, device fp32),
 * then acc /= (wsum + 1e-8).  Windows accumulated by concurrent launches must not overlap.                     */
int b2v_stitch_accumulate(const float* patch, float* acc, float* wsum, const float* gd, const float* gh,
                          const float* gw, int BC, int pd, int ph, int pw, int D, int H, int W, int d0, int h0, int w0,
                          void* stream);
int b2v_stitch_normalize(float* acc, const float* wsum, long long n, void* stream);

/* utils/metrics.py:125-193  calculate_video_metrics: a, b (BC, T, H, W) fp32 in [0, max_val]; out: DEVICE fp32 [T][2]
 * = per depth slice (sum of squared error, sum of the 11x11 box-filter SSIM map) over all (bc, h, w); per-tile partials
 * are folded in a fixed order (deterministic); the scratch for them is a stream-ordered allocation on `stream`     */
int b2v_video_metrics(const float* a, const float* b, float* out, int BC, int T, int H, int W, float max_val,
                      void* stream);

/* data/patch_slice_interpolation_dataset.py:163-181 + data/slice_interpolation_dataset.py:575-592, on the device:
 * crop [z0,z1) x [y0,y0+ph) x [x0,x0+pw) of vol (D,H,W), apply f(v) = a*clip(v,lo,hi)+b (CT windowing / range map),
 * resample the depth axis to pd slices like F.interpolate(trilinear, align_corners=False); out: (pd,ph,pw)      */
int b2v_extract_patch(const float* vol, float* out, int D, int H, int W, int z0, int z1, int y0, int x0, int pd, int ph,
                      int pw, float lo, float hi, float a, float b, void* stream);

/* per-op timing of the last planned program of an object, written as JSON text into buf:
 *   [{"name": "...", "ms": .., "flops": .., "bytes": ..}, ...]  (averaged over iters CUDA-event-timed runs)   */
int b2v_unet_profile(b2v_unet* u, int iters, char* buf, size_t cap, void* stream);
int b2v_vae_profile(b2v_vae* v, int which /*0 encode, 1 decode*/, int iters, char* buf, size_t cap, void* stream);

/* debugging / bisecting: (re)run the first n_ops ops of the last planned program eagerly, then copy the primary
 * output of op `index` (device -> dst, a DEVICE buffer of cap bytes); returns the byte count, <0 on error.
 * prog: 0 = U-Net forward | 1 = VAE encode | 2 = VAE decode.  b2v_debug_op_name returns "" past the end.    */
long long b2v_debug_op_output(void* obj, int prog, int n_ops, int index, void* dst, size_t cap, void* stream);
const char* b2v_debug_op_name(void* obj, int prog, int index);

/* ---------------------------------------------------------------- op level (parity tests, building blocks) --
 * activations here are NDHWC fp16 ("cl16") device buffers                                                     */
typedef struct b2v_conv b2v_conv;
/* kind: 0 Conv3d k3 s1 p1 | 1 Conv3d k1 | 2 Conv3d k(3,4,4) s(1,2,2) p1 | 3 ConvTranspose3d k(3,4,4) s(1,2,2) p1
 * weight/bias: HOST fp32 in the torch layout of that module; cin1 > 0 = second input (channel concat)        */
int b2v_conv_create(b2v_conv** out, int kind, const float* weight, const float* bias, int cin0, int cin1, int cout);
void b2v_conv_destroy(b2v_conv* c);
/* GroupNorm statistics cross the op-level API as int64 [N][groups][2] = (sum, sum of squares) * 2^20 (Q43.20 fixed
 * point): integer atomics make the cross-CTA accumulation order-independent, so results are bitwise reproducible.
 * in0/in1: cl16 [N][D][H][W][cin]; out: cl16 (out_fp32 == 0) or NCDHW fp32 (out_fp32 == 1);
 * stats: NULL or int64 [N][groups][2], ACCUMULATED (zero it first) with the output's per-(sample, group) sums   */
int b2v_conv_forward(b2v_conv* c, const void* in0, const void* in1, void* out, int out_fp32, int64_t* stats,
                     int groups, int act_tanh, int N, int D, int H, int W, void* stream);
int b2v_nc32_to_cl16(const float* in, void* out, int B, int C, int Cpad, long long S, void* stream);
int b2v_cl16_to_nc32(const void* in, float* out, int B, int C, int Cpad, long long S, void* stream);
/* out = silu(gn(y)) + temb (mode 0) or silu(gn(y) + res) (mode 1); stats_in as produced by b2v_conv_forward */
int b2v_gn_apply(const void* y, void* out, const int64_t* stats_in, const float* gamma, const float* beta,
                 const float* temb, const void* res, int B, long long S, int C, int G, int mode, int64_t* stats_out,
                 int G_out, void* stream);
int b2v_gn_stats(const void* x, int B, long long S, int C, int G, int64_t* stats, void* stream);
/* Fused tail of an attention-followed ResBlock3D + the TemporalAttention block (models/unet3d.py:126-133,163-194):
 *   y <- silu(GN_G2(y) + res);  y <- y + Wpv * sum_t GN_Ga(y) + bias   (in place, cl16 [B][T][P][C])
 * wt: device fp16 [C][C], wt[c][co] = (Wproj*Wv)[co][c];  bias: device fp32 [C] = T*Wproj*bv + bproj
 * stats_mid: zeroed int64 [B][Ga][2] (receives the statistics of the ResBlock output)
 * tsum_ws: fp32 workspace of tsum_cap >= B*5*P*C elements                                                      */
int b2v_res_attn_tail(void* y, const void* res, const int64_t* stats_in, const float* gamma2, const float* beta2, int G2,
                      const float* gamma_a, const float* beta_a, int Ga, const void* wt, const float* bias,
                      int64_t* stats_mid, float* tsum_ws, long long tsum_cap, int B, int T, int P, int C, void* stream);
/* DDIM update of one step, coef = device fp32[8] {c1,c2,c3,c4,sigma,..} (inference/sampler.py:299-329) */
int b2v_ddim_update(float* z, const float* eps, const float* noise, const float* coef, long long n, int* nan_flag,
                    void* stream);
/* DDPM ancestral update of one step (models/diffusion.py:287-338), in place on z; coef = HOST fp32[8] as in
 * b2v_ddpm_step; noise = DEVICE fp32 [n] (this step's torch.randn_like draw)                                   */
int b2v_ddpm_update(float* z, const float* eps, const float* noise, const float* coef, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2V_H */
