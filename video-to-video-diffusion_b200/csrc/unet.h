// U-Net denoiser object behind b2v_unet (see include/b2v.h)
#pragma once
#include "../../include/b2v.h"
#include "engine.h"

namespace b2v {

struct ResW {  // ResBlock3D (models/unet3d.py:77-133)
  ConvLayer conv1, conv2, res;
  GNW n1, n2;
  bool has_res = false;
  int cin0 = 0, cin1 = 0, cout = 0;
  int temb_off = 0;  // row offset of this block's time_mlp projection in the concatenated table
};
struct AttnW {  // TemporalAttention (models/unet3d.py:136-194), folded
  GNW norm;
  ConvLayer pv;
  __half* Wt = nullptr;  // [c][co] fp16 transpose of the folded Wp*Wv (fused attn_proj_add path; owned by UNet::ds)
  std::vector<float> u, bp;
  int C = 0;
};
struct ULevel {
  std::vector<ResW> res;
  std::vector<AttnW> attn;
  ConvLayer resample;
  bool has_resample = false;
};

struct UProgram {
  int B = 0, T = 0, h = 0, w = 0;
  long long numel = 0;
  DeviceStore ds;
  Pool pool;
  std::vector<Op> core;
  Program fwd, ddim, ddpm;  // one U-Net evaluation | + DDIM update + advance | + DDPM update + advance
  float *x_in = nullptr, *c_in = nullptr, *eps = nullptr;
  long long *t_dev = nullptr, *t_table = nullptr;
  SamplerCtl* ctl = nullptr;  // device loop state of the sampler graphs (step counter, NaN flag, noise source)
  int *step_dev = nullptr, *nan_dev = nullptr;  // = &ctl->step, &ctl->nan_flag
  int loop_pos = 0, loop_n = 0;                  // host mirror of ctl->step / loop length of the running sampler
  float *coef_table = nullptr, *silu_temb = nullptr, *proj = nullptr;
  stat_t* stats = nullptr;
  size_t stats_cap = 0;
  TembSource temb;            // set before every run: per-sample table (forward) or all-steps table (DDIM graph)
  float *silu_all = nullptr, *proj_all = nullptr;  // time embedding of every DDIM step, computed once per sample() call
  int all_cap = 0;
  long long stamp = 0;  // last use (least-recently-used program is evicted when the cache is full)
  int desc_rows = 0;  // rows of one time-embedding projection table (= UNet::proj_rows)
};

struct UNet {
  b2v_unet_desc desc;
  WeightMap wm;
  bool finalized = false;
  DeviceStore ds;
  float *freqs = nullptr, *W1 = nullptr, *B1 = nullptr, *W2 = nullptr, *B2 = nullptr, *Wproj = nullptr,
        *Bproj = nullptr;
  int proj_rows = 0;
  ConvLayer conv_in, conv_out;
  GNW out_norm;
  std::vector<ULevel> enc, dec;
  ResW mid1, mid2;
  AttnW mid_attn;
  std::map<std::string, std::unique_ptr<UProgram>> progs;
  UProgram* last = nullptr;
  UProgram* active = nullptr;
  long long clock = 0;

  int finalize();
  UProgram* program(int B, int T, int h, int w);
  int forward(const float* x, const long long* t, const float* c, float* eps_out, int B, int T, int h, int w,
              cudaStream_t st);
  int sampler_begin(const float* z_init, const float* cond, int B, int T, int h, int w, cudaStream_t st);
  int ddim_sample(const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w,
                  const long long* timesteps, int n, const float* ac, int n_train, float eta, const float* noise,
                  int* nan_flag, cudaStream_t st);
  int ddpm_step(long long t, const float* coef, const float* noise, cudaStream_t st);
  // loop steps [first, first + count) of an n-step ancestral loop (step s handles timestep n-1-s)
  int ddpm_run(const float* coef, int n, int first, int count, const float* noise, unsigned long long seed,
               cudaStream_t st);
  int ddpm_sample(const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w, const float* coef,
                  int n, const float* noise, unsigned long long seed, cudaStream_t st);
  int temb_tables(UProgram* up, int n, cudaStream_t st);  // time embedding of every loop step from up->t_table
  int sampler_end(float* z_out, cudaStream_t st);
  ~UNet();
};

}  // namespace b2v
