// SliceInterpolationVAE encoder / decoder (reference models/vae.py:100-260) on the conv + GN-apply kernels.
#include "vae.h"

namespace b2v {

static int load_block(VBlockW& b, int kind, const WeightMap& wm, const std::string& pre, int cin, int cout,
                      DeviceStore& ds) {
  if (load_conv(b.conv, kind, wm, pre + ".conv", cin, 0, cout)) return -1;
  return load_gn(b.norm, wm, pre + ".norm", cout, 8, ds);
}
static int load_vres(VResW& r, const WeightMap& wm, const std::string& pre, int C, DeviceStore& ds) {
  if (load_conv(r.conv1, CONV_K3, wm, pre + ".conv1.conv", C, 0, C)) return -1;
  if (load_gn(r.n1, wm, pre + ".conv1.norm", C, 8, ds)) return -1;
  if (load_conv(r.conv2, CONV_K3, wm, pre + ".conv2.0", C, 0, C)) return -1;
  return load_gn(r.n2, wm, pre + ".conv2.1", C, 8, ds);
}

int VAE::finalize() {
  if (finalized) return 0;
  const int bc = desc.base_channels, L = desc.latent_dim, Cin = desc.in_channels;
  if (bc % 64) return fail("vae base_channels must be a multiple of 64");
  if (L > 16 || Cin > 16) return fail("vae latent_dim / in_channels must be <= 16");
  std::string err;
  if (conv_setup_kernels(err)) return fail(err);
  // ---- encoder
  if (load_block(e_in, CONV_K3_PACKALL, wm, "encoder.conv_in", Cin, bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(e_res1[i], wm, "encoder.down1." + std::to_string(i), bc, ds)) return -1;
  if (load_block(e_down1, CONV_DOWN, wm, "encoder.down1.2", bc, 2 * bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(e_res2[i], wm, "encoder.down2." + std::to_string(i), 2 * bc, ds)) return -1;
  if (load_block(e_down2, CONV_DOWN, wm, "encoder.down2.2", 2 * bc, 4 * bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(e_mid[i], wm, "encoder.mid." + std::to_string(i), 4 * bc, ds)) return -1;
  {
    // z = scaling * quant_conv(conv_out(h))  (models/vae.py:145-146,244-246): both linear, folded exactly
    const int C4 = 4 * bc;
    const HostTensor *cw = need(wm, "encoder.conv_out.weight", 8LL * C4 * 27), *cb = need(wm, "encoder.conv_out.bias", 8),
                     *qw = need(wm, "encoder.quant_conv.weight", (long long)L * 8),
                     *qb = need(wm, "encoder.quant_conv.bias", L);
    if (!cw || !cb || !qw || !qb) return -1;
    std::vector<float> W((size_t)L * C4 * 27), Bv(L);
    const double sc = desc.scaling_factor;
    for (int l = 0; l < L; ++l) {
      double bb = qb->data[l];
      for (int j = 0; j < 8; ++j) bb += (double)qw->data[l * 8 + j] * cb->data[j];
      Bv[l] = (float)(sc * bb);
      for (long long k = 0; k < (long long)C4 * 27; ++k) {
        double s = 0;
        for (int j = 0; j < 8; ++j) s += (double)qw->data[l * 8 + j] * cw->data[(size_t)j * C4 * 27 + k];
        W[(size_t)l * C4 * 27 + k] = (float)(sc * s);
      }
    }
    if (conv_layer_init(e_out, CONV_K3, W.data(), Bv.data(), C4, 0, L, err)) return fail("encoder.conv_out: " + err);
  }
  // ---- decoder
  {
    const HostTensor *pw = need(wm, "decoder.post_quant_conv.weight", 8LL * L),
                     *pb = need(wm, "decoder.post_quant_conv.bias", 8);
    if (!pw || !pb) return -1;
    pq_w = ds.upload(pw->data);
    pq_b = ds.upload(pb->data);
  }
  if (load_block(d_in, CONV_K3_PACKW, wm, "decoder.conv_in", 8, 4 * bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(d_mid[i], wm, "decoder.mid." + std::to_string(i), 4 * bc, ds)) return -1;
  if (load_block(d_up2, CONV_UPT, wm, "decoder.up2_upsample", 4 * bc, 2 * bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(d_res2[i], wm, "decoder.up2_res." + std::to_string(i), 2 * bc, ds)) return -1;
  if (load_block(d_up3, CONV_UPT, wm, "decoder.up3_upsample", 2 * bc, bc, ds)) return -1;
  for (int i = 0; i < 2; ++i)
    if (load_vres(d_res3[i], wm, "decoder.up3_res." + std::to_string(i), bc, ds)) return -1;
  if (load_conv(d_out, CONV_K3, wm, "decoder.conv_out", bc, 0, Cin)) return -1;
  wm.clear();
  finalized = true;
  return 0;
}

VAE::~VAE() {
  progs.clear();
  auto fb = [](VBlockW& b) { conv_layer_free(b.conv); };
  auto fr = [](VResW& r) {
    conv_layer_free(r.conv1);
    conv_layer_free(r.conv2);
  };
  fb(e_in), fb(e_down1), fb(e_down2), fb(d_in), fb(d_up2), fb(d_up3);
  for (int i = 0; i < 2; ++i) fr(e_res1[i]), fr(e_res2[i]), fr(e_mid[i]), fr(d_mid[i]), fr(d_res2[i]), fr(d_res3[i]);
  conv_layer_free(e_out);
  conv_layer_free(d_out);
}

struct VBuild {
  Builder b;
  VBuild(VProgram& p) : b(p.prog.ops, p.pool, p.B, p.stats, p.stats_cap) { b.ds = &p.ds; }
  // conv -> GN(8) -> SiLU ; consumes x
  Act block(const std::string& name, const VBlockW& w, Act& x) {
    stat_t* s = b.new_stats(8);
    Act y = b.conv(name + ".conv", w.conv, x, nullptr, s, 8);
    b.free(x);
    b.gn_apply(name + ".gn_silu", y, s, w.norm, -1, nullptr, 0, nullptr, 0);
    return y;
  }
  // models/vae.py:50-56 ; consumes x
  Act res(const std::string& name, const VResW& r, Act& x) {
    stat_t* s1 = b.new_stats(8);
    Act y1 = b.conv(name + ".conv1", r.conv1, x, nullptr, s1, 8);
    b.gn_apply(name + ".gn1_silu", y1, s1, r.n1, -1, nullptr, 0, nullptr, 0);
    stat_t* s2 = b.new_stats(8);
    Act y2 = b.conv(name + ".conv2", r.conv2, y1, nullptr, s2, 8);
    b.free(y1);
    b.gn_apply(name + ".gn2_res_silu", y2, s2, r.n2, -1, &x, 1, nullptr, 0);
    b.free(x);
    return y2;
  }
};

static int build_vae_program(VAE& v, VProgram& vp, int which) {
  const int B = vp.B, T = vp.T, H = vp.H, W = vp.W, L = v.desc.latent_dim, Cin = v.desc.in_channels;
  vp.stats_cap = (size_t)B * 16 * 64;
  vp.stats = (stat_t*)vp.ds.alloc(vp.stats_cap * sizeof(stat_t));
  if (which == 0) {
    if (H % 4 || W % 4) return fail("vae.encode: H and W must be multiples of 4");
    vp.in_numel = (long long)B * Cin * T * H * W;
    vp.out_numel = (long long)B * L * T * (H / 4) * (W / 4);
  } else {
    vp.in_numel = (long long)B * L * T * H * W;
    vp.out_numel = (long long)B * Cin * T * (4 * H) * (4 * W);
  }
  vp.in = (float*)vp.ds.alloc(vp.in_numel * 4);
  vp.out = (float*)vp.ds.alloc(vp.out_numel * 4);
  if (!vp.in || !vp.out || !vp.stats) return fail("out of device memory (VAE buffers)");
  VBuild vb(vp);
  Builder& b = vb.b;
  {
    float* stats = reinterpret_cast<float*>(vp.stats);
    const size_t bytes = vp.stats_cap * sizeof(stat_t);
    Op op;
    op.name = "zero_stats";
    op.run = [=](cudaStream_t st) { launch_zero(stats, (long long)(bytes / sizeof(float)), st); };
    b.ops.push_back(std::move(op));
  }
  if (which == 0) {
    const int Cp = v.e_in.conv.cin0_pad;
    Act packed = b.alloc(Cp, T, H, W);
    if (!b.ok) return -1;
    {
      const float* in = vp.in;
      __half* o = packed.p;
      Op op;
      op.name = "enc.pack_in";
      op.bytes = (double)vp.in_numel * 4 + (double)B * T * H * W * Cp * 2;
      op.run = [=](cudaStream_t st) { launch_pack_vae_enc_in(in, o, B, Cin, Cp, T, H, W, st); };
      b.ops.push_back(std::move(op));
    }
    Act x = vb.block("enc.conv_in", v.e_in, packed);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("enc.down1." + std::to_string(i), v.e_res1[i], x);
    x = vb.block("enc.down1.2", v.e_down1, x);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("enc.down2." + std::to_string(i), v.e_res2[i], x);
    x = vb.block("enc.down2.2", v.e_down2, x);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("enc.mid." + std::to_string(i), v.e_mid[i], x);
    if (!b.ok) return -1;
    b.conv("enc.conv_out_quant", v.e_out, x, nullptr, nullptr, 0, vp.out);
    b.free(x);
  } else {
    Act packed = b.alloc(64, T, H, W);
    if (!b.ok) return -1;
    {
      const float *in = vp.in, *pw = v.pq_w, *pb = v.pq_b;
      const float sc = v.desc.scaling_factor;
      __half* o = packed.p;
      Op op;
      op.name = "dec.post_quant_pack";
      op.bytes = (double)vp.in_numel * 4 + (double)B * T * H * W * 128;
      op.run = [=](cudaStream_t st) { launch_pack_vae_dec_in(in, pw, pb, sc, o, B, L, T, H, W, st); };
      b.ops.push_back(std::move(op));
    }
    Act x = vb.block("dec.conv_in", v.d_in, packed);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("dec.mid." + std::to_string(i), v.d_mid[i], x);
    x = vb.block("dec.up2_upsample", v.d_up2, x);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("dec.up2_res." + std::to_string(i), v.d_res2[i], x);
    x = vb.block("dec.up3_upsample", v.d_up3, x);
    for (int i = 0; i < 2 && b.ok; ++i) x = vb.res("dec.up3_res." + std::to_string(i), v.d_res3[i], x);
    if (!b.ok) return -1;
    b.conv("dec.conv_out_tanh", v.d_out, x, nullptr, nullptr, 0, vp.out, ACT_TANH);
    b.free(x);
  }
  return b.ok ? 0 : -1;
}

VProgram* VAE::program(int which, int B, int T, int H, int W) {
  if (!finalized && finalize()) return nullptr;
  const std::string key = std::to_string(which) + ":" + std::to_string(B) + "x" + std::to_string(T) + "x" +
                          std::to_string(H) + "x" + std::to_string(W);
  auto it = progs.find(key);
  if (it != progs.end()) return last[which] = it->second.get();
  if (progs.size() >= 4) {
    progs.clear();
    last[0] = last[1] = nullptr;
  }
  std::unique_ptr<VProgram> vp(new VProgram());
  vp->B = B;
  vp->T = T;
  vp->H = H;
  vp->W = W;
  if (build_vae_program(*this, *vp, which)) return nullptr;
  last[which] = vp.get();
  progs[key] = std::move(vp);
  return last[which];
}

int VAE::encode(const float* x, float* z, int B, int T, int H, int W, cudaStream_t st) {
  VProgram* vp = program(0, B, T, H, W);
  if (!vp) return -1;
  B2V_CUDA(cudaMemcpyAsync(vp->in, x, vp->in_numel * 4, cudaMemcpyDeviceToDevice, st));
  if (vp->prog.run(st)) return -1;
  B2V_CUDA(cudaMemcpyAsync(z, vp->out, vp->out_numel * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int VAE::decode(const float* z, float* x, int B, int T, int h, int w, cudaStream_t st) {
  VProgram* vp = program(1, B, T, h, w);
  if (!vp) return -1;
  B2V_CUDA(cudaMemcpyAsync(vp->in, z, vp->in_numel * 4, cudaMemcpyDeviceToDevice, st));
  if (vp->prog.run(st)) return -1;
  B2V_CUDA(cudaMemcpyAsync(x, vp->out, vp->out_numel * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace b2v
