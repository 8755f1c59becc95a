// VAE encoder / decoder object behind b2v_vae (see include/b2v.h)
#pragma once
#include "../../include/b2v.h"
#include "engine.h"

namespace b2v {

struct VResW {  // models/vae.py ResBlock3D :38-56
  ConvLayer conv1, conv2;
  GNW n1, n2;
};
struct VBlockW {  // Conv3DBlock / DownsampleBlock / UpsampleBlock: conv -> GN(8) -> SiLU
  ConvLayer conv;
  GNW norm;
};

struct VProgram {
  int B = 0, T = 0, H = 0, W = 0;  // encode: pixel dims; decode: latent dims
  long long in_numel = 0, out_numel = 0;
  DeviceStore ds;
  Pool pool;
  Program prog;
  float *in = nullptr, *out = nullptr;
  stat_t* stats = nullptr;
  size_t stats_cap = 0;
};

struct VAE {
  b2v_vae_desc desc;
  WeightMap wm;
  bool finalized = false;
  DeviceStore ds;
  // encoder (models/vae.py:100-147)
  VBlockW e_in, e_down1, e_down2;
  VResW e_res1[2], e_res2[2], e_mid[2];
  ConvLayer e_out;  // conv_out folded with quant_conv and scaling_factor
  // decoder (models/vae.py:150-204)
  float *pq_w = nullptr, *pq_b = nullptr;
  VBlockW d_in, d_up2, d_up3;
  VResW d_mid[2], d_res2[2], d_res3[2];
  ConvLayer d_out;
  std::map<std::string, std::unique_ptr<VProgram>> progs;
  VProgram* last[2] = {nullptr, nullptr};

  int finalize();
  VProgram* program(int which, int B, int T, int H, int W);
  int encode(const float* x, float* z, int B, int T, int H, int W, cudaStream_t st);
  int decode(const float* z, float* x, int B, int T, int h, int w, cudaStream_t st);
  ~VAE();
};

}  // namespace b2v
