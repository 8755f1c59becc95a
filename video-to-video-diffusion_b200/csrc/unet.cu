// U-Net denoiser + samplers.  Structure follows the reference module tree (models/unet3d.py:240-332) so that a
// reference state_dict loads key by key; execution is a flat op list per input shape, replayed as a CUDA graph.
#include "unet.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>

namespace b2v {

// ------------------------------------------------------------------ weights
static int load_res(ResW& r, const WeightMap& wm, const std::string& pre, int cin0, int cin1, int cout, int td,
                    DeviceStore& ds, std::vector<float>& proj_w, std::vector<float>& proj_b) {
  r.cin0 = cin0;
  r.cin1 = cin1;
  r.cout = cout;
  if (load_conv(r.conv1, CONV_K3, wm, pre + ".conv1.conv", cin0, cin1, cout)) return -1;
  const int g1 = (cout % 8 == 0) ? 8 : groups32(cout);  // reference Conv3DBlock (models/unet3d.py:58)
  if (load_gn(r.n1, wm, pre + ".conv1.norm", cout, g1, ds)) return -1;
  if (load_conv(r.conv2, CONV_K3, wm, pre + ".conv2.0", cout, 0, cout)) return -1;
  if (load_gn(r.n2, wm, pre + ".conv2.1", cout, groups32(cout), ds)) return -1;
  r.has_res = (cin0 + cin1 != cout);
  if (r.has_res && load_conv(r.res, CONV_K1, wm, pre + ".residual_conv", cin0, cin1, cout)) return -1;
  const HostTensor* tw = need(wm, pre + ".time_mlp.1.weight", (long long)cout * td);
  const HostTensor* tb = need(wm, pre + ".time_mlp.1.bias", cout);
  if (!tw || !tb) return -1;
  r.temb_off = (int)proj_b.size();
  proj_w.insert(proj_w.end(), tw->data.begin(), tw->data.end());
  proj_b.insert(proj_b.end(), tb->data.begin(), tb->data.end());
  return 0;
}

// TemporalAttention fold (see ew_kernels.cu: attn_tsum): Wpv = Wp*Wv, u = Wp*bv (bias = T*u + bp at plan time)
static int load_attn(AttnW& a, const WeightMap& wm, const std::string& pre, int C, DeviceStore& ds) {
  a.C = C;
  if (load_gn(a.norm, wm, pre + ".norm", C, groups32(C), ds)) return -1;
  const HostTensor* qkv = need(wm, pre + ".qkv.weight", 3LL * C * C);
  const HostTensor* qb = need(wm, pre + ".qkv.bias", 3LL * C);
  const HostTensor* pw = need(wm, pre + ".proj_out.weight", (long long)C * C);
  const HostTensor* pb = need(wm, pre + ".proj_out.bias", C);
  if (!qkv || !qb || !pw || !pb) return -1;
  const float* Wv = qkv->data.data() + 2LL * C * C;
  const float* bv = qb->data.data() + 2LL * C;
  const float* Wp = pw->data.data();
  std::vector<float> Wpv((size_t)C * C, 0.f);
  std::vector<double> row(C);
  for (int o = 0; o < C; ++o) {
    for (int i = 0; i < C; ++i) row[i] = 0.0;
    for (int k = 0; k < C; ++k) {
      const double p = Wp[(size_t)o * C + k];
      const float* wv = Wv + (size_t)k * C;
      for (int i = 0; i < C; ++i) row[i] += p * wv[i];
    }
    for (int i = 0; i < C; ++i) Wpv[(size_t)o * C + i] = (float)row[i];
  }
  a.u.assign(C, 0.f);
  a.bp.assign(pb->data.begin(), pb->data.end());
  for (int o = 0; o < C; ++o) {
    double s = 0;
    for (int k = 0; k < C; ++k) s += (double)Wp[(size_t)o * C + k] * bv[k];
    a.u[o] = (float)s;
  }
  std::string err;
  if (conv_layer_init(a.pv, CONV_K1, Wpv.data(), nullptr, C, 0, C, err)) return fail(pre + ": " + err);
  // transposed fp16 copy for the fused attn_proj_add kernel: Wt[c][co] = Wpv[co][c]
  std::vector<__half> wt((size_t)C * C);
  for (int o = 0; o < C; ++o)
    for (int i = 0; i < C; ++i) wt[(size_t)i * C + o] = conv_operand(Wpv[(size_t)o * C + i]);
  a.Wt = (__half*)ds.alloc(wt.size() * sizeof(__half));
  if (!a.Wt || cudaMemcpy(a.Wt, wt.data(), wt.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(pre + ": device alloc failed for the folded attention weights");
  return 0;
}

int UNet::finalize() {
  if (finalized) return 0;
  const b2v_unet_desc& d = desc;
  const int mc = d.model_channels, td = d.time_embed_dim, L = d.latent_dim, NL = d.num_levels;
  if (mc % 64) return fail("model_channels must be a multiple of 64");
  if (2 * L * 3 > 64) return fail("latent_dim too large for the packed conv_in (2*latent_dim*3 must be <= 64)");
  if (td > 4096 || mc > 4096) return fail("time_embed_dim / model_channels too large");
  std::string err;
  if (conv_setup_kernels(err)) return fail(err);

  // time embedding (models/unet3d.py:25-48)
  {
    const int half = mc / 2;
    std::vector<float> fr(half);
    const float c = (float)(-(log(10000.0) / (half - 1)));
    for (int i = 0; i < half; ++i) fr[i] = expf((float)i * c);
    freqs = ds.upload(fr);
    const HostTensor *w1 = need(wm, "time_embed.time_mlp.1.weight", (long long)td * mc),
                     *b1 = need(wm, "time_embed.time_mlp.1.bias", td),
                     *w2 = need(wm, "time_embed.time_mlp.3.weight", (long long)td * td),
                     *b2 = need(wm, "time_embed.time_mlp.3.bias", td);
    if (!w1 || !b1 || !w2 || !b2) return -1;
    W1 = ds.upload(w1->data);
    B1 = ds.upload(b1->data);
    W2 = ds.upload(w2->data);
    B2 = ds.upload(b2->data);
  }
  std::vector<float> pw, pb;
  if (load_conv(conv_in, CONV_K3_PACKW, wm, "conv_in", 2 * L, 0, mc)) return -1;

  int ch = mc;
  enc.resize(NL);
  std::vector<int> level_ch(NL);
  for (int l = 0; l < NL; ++l) {
    const int out_ch = mc * d.channel_mult[l];
    const bool at = (d.attention_mask >> l) & 1;
    enc[l].res.resize(d.num_res_blocks);
    if (at) enc[l].attn.resize(d.num_res_blocks);
    for (int i = 0; i < d.num_res_blocks; ++i) {
      const std::string pre = "down_blocks." + std::to_string(l) + "." + std::to_string(i);
      if (load_res(enc[l].res[i], wm, pre + ".0", ch, 0, out_ch, td, ds, pw, pb)) return -1;
      if (at && load_attn(enc[l].attn[i], wm, pre + ".1", out_ch, ds)) return -1;
      ch = out_ch;
    }
    level_ch[l] = ch;
    enc[l].has_resample = (l < NL - 1);
    if (enc[l].has_resample &&
        load_conv(enc[l].resample, CONV_DOWN, wm, "down_samples." + std::to_string(l) + ".conv", ch, 0, ch))
      return -1;
  }
  if (load_res(mid1, wm, "mid_block1", ch, 0, ch, td, ds, pw, pb)) return -1;
  if (load_attn(mid_attn, wm, "mid_attn", ch, ds)) return -1;
  if (load_res(mid2, wm, "mid_block2", ch, 0, ch, td, ds, pw, pb)) return -1;
  dec.resize(NL);
  for (int j = 0; j < NL; ++j) {
    const int lvl = NL - 1 - j;
    const int out_ch = mc * d.channel_mult[lvl];
    const bool at = (d.attention_mask >> lvl) & 1;
    dec[j].res.resize(d.num_res_blocks + 1);
    if (at) dec[j].attn.resize(d.num_res_blocks + 1);
    for (int i = 0; i < d.num_res_blocks + 1; ++i) {
      const std::string pre = "up_blocks." + std::to_string(j) + "." + std::to_string(i);
      const int skip_ch = (i == 0) ? mc * d.channel_mult[lvl] : 0;  // models/unet3d.py:303-306
      if (load_res(dec[j].res[i], wm, pre + ".0", ch, skip_ch, out_ch, td, ds, pw, pb)) return -1;
      if (at && load_attn(dec[j].attn[i], wm, pre + ".1", out_ch, ds)) return -1;
      ch = out_ch;
    }
    dec[j].has_resample = (j < NL - 1);
    if (dec[j].has_resample &&
        load_conv(dec[j].resample, CONV_UPT, wm, "up_samples." + std::to_string(j) + ".conv", ch, 0, ch))
      return -1;
  }
  if (load_gn(out_norm, wm, "conv_out.0", ch, groups32(ch), ds)) return -1;
  if (load_conv(conv_out, CONV_K3, wm, "conv_out.2", ch, 0, L)) return -1;
  proj_rows = (int)pb.size();
  Wproj = ds.upload(pw);
  Bproj = ds.upload(pb);
  if (!Wproj || !Bproj || !W1 || !W2) return fail("device alloc failed for time-embedding weights");
  wm.clear();
  finalized = true;
  return 0;
}

UNet::~UNet() {
  progs.clear();
  auto fr = [](ResW& r) {
    conv_layer_free(r.conv1);
    conv_layer_free(r.conv2);
    conv_layer_free(r.res);
  };
  for (auto& lv : enc) {
    for (auto& r : lv.res) fr(r);
    for (auto& a : lv.attn) conv_layer_free(a.pv);
    conv_layer_free(lv.resample);
  }
  for (auto& lv : dec) {
    for (auto& r : lv.res) fr(r);
    for (auto& a : lv.attn) conv_layer_free(a.pv);
    conv_layer_free(lv.resample);
  }
  fr(mid1);
  fr(mid2);
  conv_layer_free(mid_attn.pv);
  conv_layer_free(conv_in);
  conv_layer_free(conv_out);
}

// ------------------------------------------------------------------ program
static bool attn_fused(int C) {
  static const bool off = getenv("B2V_ATTN_UNFUSED") != nullptr;  // A/B switch: the four-launch form
  return !off && attn_fused_supported(C);
}

struct UBuild {
  UNet& u;
  UProgram& up;
  Builder b;
  UBuild(UNet& u_, UProgram& p) : u(u_), up(p), b(p.core, p.pool, p.B, p.stats, p.stats_cap) {
    b.temb_src = &p.temb;
    b.ds = &p.ds;
  }

  // ResBlock3D.forward (models/unet3d.py:116-133)
  // tsum_out != nullptr: the block feeds a TemporalAttention -- its tail also emits the depth sums (fused path)
  Act res(const std::string& name, const ResW& r, const Act& x, const Act* skip, stat_t** out_stats, int G_out,
          float** tsum_out = nullptr, int* ts_out = nullptr) {
    stat_t* s1 = b.new_stats(r.n1.G);
    Act y1 = b.conv(name + ".conv1", r.conv1, x, skip, s1, r.n1.G);
    Act rr;
    if (r.has_res) rr = b.conv(name + ".residual_conv", r.res, x, skip, nullptr, 0);
    b.gn_apply(name + ".gn1_silu_temb", y1, s1, r.n1, r.temb_off, nullptr, 0, nullptr, 0);
    stat_t* s2 = b.new_stats(r.n2.G);
    Act y2 = b.conv(name + ".conv2", r.conv2, y1, nullptr, s2, r.n2.G);
    b.free(y1);
    stat_t* so = out_stats ? b.new_stats(G_out) : nullptr;
    if (out_stats) *out_stats = so;
    if (tsum_out && so && b.ok) {
      const int B = up.B, T = y2.D, P = y2.H * y2.W, C = y2.C, G = r.n2.G;
      const int TS = attn_tsum_splits(B, T, P, C);
      // depth sums [B][TS][P][C] fp32, followed by the attention GEMM's fp16 output [B][P][C]
      float* tsum = (float*)b.pool.get((size_t)B * (TS + 1) * P * C * sizeof(float));
      if (!tsum) {
        b.ok = false;
        return y2;
      }
      *tsum_out = tsum;
      *ts_out = TS;
      __half* yp = y2.p;
      const __half* rp = r.has_res ? rr.p : x.p;
      const float *ga = r.n2.gamma, *be = r.n2.beta;
      Op op;
      op.name = name + ".gn2_res_silu_tsum";
      op.bytes = (double)B * T * P * C * 2.0 * 3.0;
      op.out = yp;
      op.out_bytes = (size_t)B * T * P * C * 2;
      op.run = [=](cudaStream_t st) {
        launch_gn_res_tsum(yp, rp, s2, ga, be, B, T, P, C, G, 1e-5f, so, G_out, tsum, TS, st);
      };
      b.ops.push_back(std::move(op));
    } else {
      b.gn_apply(name + ".gn2_res_silu", y2, s2, r.n2, -1, r.has_res ? &rr : &x, 1, so, G_out);
    }
    if (r.has_res) b.free(rr);
    return y2;
  }

  // TemporalAttention.forward (models/unet3d.py:163-194), folded: x += Wpv * sum_t GN(x) + (T*u + bp)
  void attn(const std::string& name, const AttnW& a, Act& x, const stat_t* stats_x, float* tsum = nullptr, int TS = 0) {
    const int B = up.B, T = x.D, P = x.H * x.W, C = x.C;
    if (tsum) {
      std::vector<float> bias(C, 0.f);
      for (int i = 0; i < C; ++i) bias[i] = (float)T * a.u[i] + a.bp[i];
      const float* bias_d = up.ds.upload(bias);
      if (!bias_d) {
        b.ok = false;
        return;
      }
      __half* xp = x.p;
      const __half* wt = a.Wt;
      const float *ga = a.norm.gamma, *be = a.norm.beta;
      const int G = a.norm.G;
      Op op;
      op.name = name + ".proj_add";
      op.bytes = (double)B * T * P * C * 2.0 * 2.0;
      // algorithmic FLOPs of the reference block: qkv (C->3C) and proj (C->C) 1x1 convs over all T positions
      op.flops = 2.0 * B * T * P * (double)C * (4.0 * C);
      op.out = xp;
      op.out_bytes = (size_t)B * T * P * C * 2;
      op.launches = 2;
      __half* g_ws = (__half*)(tsum + (size_t)B * TS * P * C);
      op.run = [=](cudaStream_t st) {
        launch_attn_proj_add(xp, tsum, TS, stats_x, ga, be, wt, bias_d, g_ws, B, T, P, C, G, 1e-5f, st);
      };
      b.ops.push_back(std::move(op));
      b.pool.put(tsum);
      return;
    }
    Act s = b.alloc(C, 1, x.H, x.W);
    if (!b.ok) return;
    {
      const __half* xp = x.p;
      __half* sp = s.p;
      const float *ga = a.norm.gamma, *be = a.norm.beta;
      const int G = a.norm.G;
      Op op;
      op.name = name + ".gn_tsum";
      op.bytes = (double)B * T * P * C * 2.0;
      op.run = [=](cudaStream_t st) { launch_attn_tsum(xp, stats_x, ga, be, sp, B, T, P, C, G, 1e-5f, st); };
      b.ops.push_back(std::move(op));
    }
    std::vector<float> bias(a.pv.cout_pad, 0.f);
    for (int i = 0; i < C; ++i) bias[i] = (float)T * a.u[i] + a.bp[i];
    const float* bias_d = up.ds.upload(bias);
    Act y = b.conv(name + ".pv_gemm", a.pv, s, nullptr, nullptr, 0, nullptr, ACT_NONE, bias_d);
    // algorithmic FLOPs of the reference block: qkv (C->3C) and proj (C->C) 1x1 convs over all T positions
    b.ops.back().flops = 2.0 * B * T * P * (double)C * (4.0 * C);
    b.free(s);
    if (!b.ok) return;
    {
      __half* xp = x.p;
      const __half* yp = y.p;
      Op op;
      op.name = name + ".add_bcast";
      op.bytes = (double)B * T * P * C * 2.0 * 2.0;
      op.run = [=](cudaStream_t st) { launch_add_bcast_t(xp, yp, B, T, P, C, st); };
      b.ops.push_back(std::move(op));
    }
    b.free(y);
  }
};

static int build_unet_program(UNet& u, UProgram& up) {
  const b2v_unet_desc& d = u.desc;
  const int B = up.B, T = up.T, h = up.h, w = up.w, L = d.latent_dim, NL = d.num_levels;
  for (int l = 0; l < NL - 1; ++l)
    if ((h >> l) & 1 || (w >> l) & 1) return fail("latent H, W must be divisible by 2^(levels-1)");
  const long long numel = (long long)B * L * T * h * w;
  up.numel = numel;
  up.desc_rows = u.proj_rows;
  up.x_in = (float*)up.ds.alloc(numel * 4);
  up.c_in = (float*)up.ds.alloc(numel * 4);
  up.eps = (float*)up.ds.alloc(numel * 4);
  up.t_dev = (long long*)up.ds.alloc(sizeof(long long) * (B < 64 ? 64 : B));
  up.ctl = (SamplerCtl*)up.ds.alloc(sizeof(SamplerCtl));
  up.step_dev = up.ctl ? &up.ctl->step : nullptr;
  up.nan_dev = up.ctl ? &up.ctl->nan_flag : nullptr;
  up.t_table = (long long*)up.ds.alloc(sizeof(long long) * 4096);
  up.coef_table = (float*)up.ds.alloc(sizeof(float) * 8 * 4096);
  up.stats_cap = (size_t)B * 64 * 256;
  up.stats = (stat_t*)up.ds.alloc(up.stats_cap * sizeof(stat_t));
  up.silu_temb = (float*)up.ds.alloc((size_t)B * d.time_embed_dim * 4);
  up.proj = (float*)up.ds.alloc((size_t)B * u.proj_rows * 4);
  if (!up.x_in || !up.c_in || !up.eps || !up.stats || !up.proj || !up.coef_table || !up.ctl || !up.t_dev ||
      !up.t_table || !up.silu_temb)
    return fail("out of device memory (U-Net buffers)");

  UBuild ub(u, up);
  Builder& b = ub.b;
  {
    float* stats = reinterpret_cast<float*>(up.stats);
    const size_t bytes = up.stats_cap * sizeof(stat_t);
    Op op;
    op.name = "zero_stats";
    op.bytes = (double)bytes;
    op.run = [=](cudaStream_t st) { launch_zero(stats, (long long)(bytes / sizeof(float)), st); };
    b.ops.push_back(std::move(op));
  }
  {
    const long long* tp = up.t_dev;
    const float *fr = u.freqs, *W1 = u.W1, *B1 = u.B1, *W2 = u.W2, *B2 = u.B2, *Wp = u.Wproj, *Bp = u.Bproj;
    float *st_ = up.silu_temb, *pj = up.proj;
    const int rows = u.proj_rows, dim = d.model_channels, td = d.time_embed_dim;
    Op op;
    op.name = "time_embed";
    op.launches = 2;
    op.bytes = ((double)rows * td + (double)td * td + (double)td * dim) * 4.0;
    op.run = [=](cudaStream_t st) {
      launch_temb(tp, nullptr, nullptr, fr, W1, B1, W2, B2, st_, Wp, Bp, pj, rows, dim, td, B, st);
    };
    b.ops.push_back(std::move(op));
  }
  Act packed = b.alloc(64, T, h, w);
  if (!b.ok) return -1;
  {
    const float *x = up.x_in, *c = up.c_in;
    __half* o = packed.p;
    Op op;
    op.name = "pack_conv_in";
    op.bytes = (double)numel * 8.0 + (double)B * T * h * w * 128.0;
    op.run = [=](cudaStream_t st) { launch_pack_unet_in(x, c, o, B, L, T, h, w, st); };
    b.ops.push_back(std::move(op));
  }
  Act cur = b.conv("conv_in", u.conv_in, packed, nullptr, nullptr, 0);
  b.free(packed);

  std::vector<Act> skips;
  stat_t* cur_stats = nullptr;
  bool cur_is_skip = false;
  for (int l = 0; l < NL && b.ok; ++l) {
    auto& lv = u.enc[l];
    for (size_t i = 0; i < lv.res.size() && b.ok; ++i) {
      const std::string nm = "down" + std::to_string(l) + "." + std::to_string(i);
      const bool at = !lv.attn.empty();
      float* tsum = nullptr;
      int TS = 0;
      const bool fz = at && attn_fused(lv.attn[i].C);
      Act nxt = ub.res(nm, lv.res[i], cur, nullptr, at ? &cur_stats : nullptr, at ? lv.attn[i].norm.G : 0,
                       fz ? &tsum : nullptr, &TS);
      b.free(cur);
      cur = nxt;
      if (at) ub.attn(nm + ".attn", lv.attn[i], cur, cur_stats, tsum, TS);
    }
    skips.push_back(cur);
    cur_is_skip = true;
    if (lv.has_resample) {
      cur = b.conv("downsample" + std::to_string(l), lv.resample, cur, nullptr, nullptr, 0);
      cur_is_skip = false;
    }
  }
  if (!b.ok) return -1;
  {
    float* tsum = nullptr;
    int TS = 0;
    Act nxt = ub.res("mid1", u.mid1, cur, nullptr, &cur_stats, u.mid_attn.norm.G,
                     attn_fused(u.mid_attn.C) ? &tsum : nullptr, &TS);
    if (!cur_is_skip) b.free(cur);
    cur = nxt;
    ub.attn("mid.attn", u.mid_attn, cur, cur_stats, tsum, TS);
    nxt = ub.res("mid2", u.mid2, cur, nullptr, nullptr, 0);
    b.free(cur);
    cur = nxt;
  }
  for (int j = 0; j < NL && b.ok; ++j) {
    auto& lv = u.dec[j];
    for (size_t i = 0; i < lv.res.size() && b.ok; ++i) {
      const std::string nm = "up" + std::to_string(j) + "." + std::to_string(i);
      const bool at = !lv.attn.empty();
      const bool last = (j == NL - 1 && i + 1 == lv.res.size());
      stat_t** so = (at || last) ? &cur_stats : nullptr;
      const int Go = at ? lv.attn[i].norm.G : (last ? u.out_norm.G : 0);
      Act nxt;
      float* tsum = nullptr;
      int TS = 0;
      float** tso = (at && attn_fused(lv.attn[i].C)) ? &tsum : nullptr;
      if (i == 0) {
        Act skip = skips.back();
        skips.pop_back();
        nxt = ub.res(nm, lv.res[i], cur, &skip, so, Go, tso, &TS);
        b.free(skip);
      } else {
        nxt = ub.res(nm, lv.res[i], cur, nullptr, so, Go, tso, &TS);
      }
      b.free(cur);
      cur = nxt;
      if (at) {
        ub.attn(nm + ".attn", lv.attn[i], cur, cur_stats, tsum, TS);
        if (last) {  // conv_out's GroupNorm needs statistics of the post-attention tensor
          cur_stats = b.new_stats(u.out_norm.G);
          const __half* xp = cur.p;
          stat_t* so2 = cur_stats;
          const long long S = cur.S();
          const int C = cur.C, G = u.out_norm.G;
          Op op;
          op.name = "conv_out.stats";
          op.bytes = (double)B * S * C * 2.0;
          op.run = [=](cudaStream_t st) { launch_gn_stats(xp, B, S, C, G, so2, st); };
          b.ops.push_back(std::move(op));
        }
      }
    }
    if (lv.has_resample) {
      Act nxt = b.conv("upsample" + std::to_string(j), lv.resample, cur, nullptr, nullptr, 0);
      b.free(cur);
      cur = nxt;
    }
  }
  if (!b.ok) return -1;
  // conv_out: GroupNorm -> SiLU -> Conv3d (models/unet3d.py:328-332)
  b.gn_apply("conv_out.gn_silu", cur, cur_stats, u.out_norm, -1, nullptr, 0, nullptr, 0);
  b.conv("conv_out.conv", u.conv_out, cur, nullptr, nullptr, 0, up.eps);
  b.free(cur);
  if (!b.ok) return -1;

  up.fwd.ops = up.core;
  // sampler steps: core (the time embedding of every step is computed once per sample() call: see temb_tables), the
  // scheduler update reading its coefficient row / noise through the device control block, advance
  for (auto& o : up.core)
    if (o.name != "time_embed") {
      up.ddim.ops.push_back(o);
      up.ddpm.ops.push_back(o);
    }
  {
    float* z = up.x_in;
    const float* e = up.eps;
    const float* ct = up.coef_table;
    SamplerCtl* ctl = up.ctl;
    int* nf = up.nan_dev;
    Op op;
    op.name = "ddim_update";
    op.launches = 2;
    op.bytes = (double)numel * 12.0;
    op.run = [=](cudaStream_t st) {
      launch_ddim_update(z, e, nullptr, ct, ctl, 0, numel, nf, st);
      launch_advance_step(ctl, st);
    };
    up.ddim.ops.push_back(op);
    op.name = "ddpm_update";
    op.bytes = (double)numel * 16.0;
    op.run = [=](cudaStream_t st) {
      launch_ddpm_update(z, e, nullptr, ct, ctl, Coef8{}, numel, st);
      launch_advance_step(ctl, st);
    };
    up.ddpm.ops.push_back(std::move(op));
  }
  return 0;
}

UProgram* UNet::program(int B, int T, int h, int w) {
  if (!finalized && finalize()) return nullptr;
  const std::string key = std::to_string(B) + "x" + std::to_string(T) + "x" + std::to_string(h) + "x" + std::to_string(w);
  auto it = progs.find(key);
  if (it != progs.end()) {
    last = it->second.get();
    last->stamp = ++clock;
    return last;
  }
  if (progs.size() >= 6) {  // bound memory: evict the least recently used shape only (a ragged-batch sweep alternates
    auto victim = progs.begin();  // between two or three shapes and must not thrash graph capture)
    for (auto i = progs.begin(); i != progs.end(); ++i)
      if (i->second->stamp < victim->second->stamp) victim = i;
    if (victim->second.get() == active) active = nullptr;
    if (victim->second.get() == last) last = nullptr;
    cudaDeviceSynchronize();  // its buffers may still be in use by queued work
    progs.erase(victim);
  }
  std::unique_ptr<UProgram> up(new UProgram());
  up->B = B;
  up->T = T;
  up->h = h;
  up->w = w;
  if (build_unet_program(*this, *up)) return nullptr;
  up->stamp = ++clock;
  last = up.get();
  progs[key] = std::move(up);
  return last;
}

int UNet::forward(const float* x, const long long* t, const float* c, float* eps_out, int B, int T, int h, int w,
                  cudaStream_t st) {
  UProgram* up = program(B, T, h, w);
  if (!up) return -1;
  const size_t bytes = up->numel * 4;
  B2V_CUDA(cudaMemcpyAsync(up->x_in, x, bytes, cudaMemcpyDeviceToDevice, st));
  B2V_CUDA(cudaMemcpyAsync(up->c_in, c, bytes, cudaMemcpyDeviceToDevice, st));
  B2V_CUDA(cudaMemcpyAsync(up->t_dev, t, sizeof(long long) * B, cudaMemcpyDeviceToDevice, st));
  up->temb = TembSource{up->proj, proj_rows, nullptr, 0};
  if (up->fwd.run(st)) return -1;
  B2V_CUDA(cudaMemcpyAsync(eps_out, up->eps, bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int UNet::sampler_begin(const float* z_init, const float* cond, int B, int T, int h, int w, cudaStream_t st) {
  UProgram* up = program(B, T, h, w);
  if (!up) return -1;
  const size_t bytes = up->numel * 4;
  B2V_CUDA(cudaMemcpyAsync(up->x_in, z_init, bytes, cudaMemcpyDeviceToDevice, st));
  B2V_CUDA(cudaMemcpyAsync(up->c_in, cond, bytes, cudaMemcpyDeviceToDevice, st));
  B2V_CUDA(cudaMemsetAsync(up->ctl, 0, sizeof(SamplerCtl), st));
  up->loop_pos = up->loop_n = 0;
  active = up;
  return 0;
}

// time embedding + all ResBlock projections for every loop step at once ([n][rows] table, one launch pair)
int UNet::temb_tables(UProgram* up, int n, cudaStream_t st) {
  if (up->all_cap < n) {
    B2V_CUDA(cudaStreamSynchronize(st));  // the old tables may be in use by queued work
    if (up->silu_all) up->ds.release(up->silu_all);
    if (up->proj_all) up->ds.release(up->proj_all);
    up->all_cap = 0;
    up->silu_all = (float*)up->ds.alloc((size_t)n * desc.time_embed_dim * sizeof(float));
    up->proj_all = (float*)up->ds.alloc((size_t)n * proj_rows * sizeof(float));
    if (!up->silu_all || !up->proj_all) return fail("out of device memory (time-embedding table)");
    up->all_cap = n;
    for (Program* pr : {&up->ddim, &up->ddpm})
      if (pr->exec) {  // the captured graphs hold the old table pointer
        cudaGraphExecDestroy(pr->exec);
        pr->exec = nullptr;
      }
  }
  launch_temb(up->t_table, nullptr, nullptr, freqs, W1, B1, W2, B2, up->silu_all, Wproj, Bproj, up->proj_all, proj_rows,
              desc.model_channels, desc.time_embed_dim, n, st);
  g_launches += 2;
  up->temb = TembSource{up->proj_all, 0, up->step_dev, proj_rows};
  return 0;
}

// coefficient table in the reference's fp32 operation order (inference/sampler.py:295-325)
int UNet::ddim_sample(const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w,
                      const long long* timesteps, int n, const float* ac, int n_train, float eta, const float* noise,
                      int* nan_flag, cudaStream_t st) {
  if (n < 1 || n > 4096) return fail("ddim_sample: 1 <= n <= 4096 timesteps");
  if (eta > 0.f && !noise) return fail("ddim_sample: eta > 0 needs the per-step noise draws");
  if (!timesteps || !ac) return fail("ddim_sample: null schedule");
  if (sampler_begin(z_init, cond, B, T, h, w, st)) return -1;
  UProgram* up = active;
  std::vector<float> coef((size_t)n * 8, 0.f);
  std::vector<long long> ts(timesteps, timesteps + n);
  for (int i = 0; i < n; ++i) {
    if (ts[i] < 0 || ts[i] >= n_train) return fail("ddim_sample: timestep out of range");
    const float a_t = ac[ts[i]];
    const float a_prev = (i < n - 1) ? ac[ts[i + 1]] : 1.0f;
    float* c = &coef[(size_t)i * 8];
    c[0] = sqrtf(1.0f - a_t + 1e-8f);
    c[1] = sqrtf(a_t + 1e-8f) + 1e-8f;
    c[2] = sqrtf(a_prev + 1e-8f);
    c[3] = sqrtf(1.0f - a_prev + 1e-8f);
    if (eta > 0.f)
      c[4] = eta * sqrtf((1.0f - a_prev + 1e-8f) / (1.0f - a_t + 1e-8f) * (1.0f - a_t / (a_prev + 1e-8f)));
  }
  // pageable-host copies are staged by the runtime before returning, so the vectors may go out of scope
  B2V_CUDA(cudaMemcpyAsync(up->t_table, ts.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
  B2V_CUDA(cudaMemcpyAsync(up->coef_table, coef.data(), sizeof(float) * 8 * n, cudaMemcpyHostToDevice, st));
  if (eta > 0.f) {  // the stochastic variant reads step i's draw from row i of the caller's buffer
    SamplerCtl c{};
    c.noise = noise;
    B2V_CUDA(cudaMemcpyAsync(up->ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
  }
  if (temb_tables(up, n, st)) return -1;
  for (int i = 0; i < n; ++i)
    if (up->ddim.run(st)) return -1;
  B2V_CUDA(cudaMemcpyAsync(z_out, up->x_in, up->numel * 4, cudaMemcpyDeviceToDevice, st));
  if (nan_flag) B2V_CUDA(cudaMemcpyAsync(nan_flag, up->nan_dev, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return check_launches("ddim_sample");
}

// GaussianDiffusion.p_sample_loop (models/diffusion.py:340-367) as one graph replay per step.  coef: HOST rows indexed
// by TIMESTEP (ddpm coefficient rows, see b2v.h); loop step s handles timestep n-1-s.
int UNet::ddpm_run(const float* coef, int n, int first, int count, const float* noise, unsigned long long seed,
                   cudaStream_t st) {
  UProgram* up = active;
  if (!up) return fail("ddpm_run: call b2v_sampler_begin first");
  if (n < 1 || n > 4096) return fail("ddpm_run: 1 <= n <= 4096 timesteps");
  if (first < 0 || count < 0 || first + count > n) return fail("ddpm_run: step range outside the loop");
  if (first != up->loop_pos) return fail("ddpm_run: chunks must be contiguous (expected first = " + std::to_string(up->loop_pos) + ")");
  if (first == 0) {
    if (!coef) return fail("ddpm_run: null coefficient table");
    std::vector<float> rows((size_t)n * 8);
    std::vector<long long> ts(n);
    for (int s = 0; s < n; ++s) {
      ts[s] = n - 1 - s;
      for (int j = 0; j < 8; ++j) rows[(size_t)s * 8 + j] = coef[(size_t)(n - 1 - s) * 8 + j];
    }
    B2V_CUDA(cudaMemcpyAsync(up->t_table, ts.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
    B2V_CUDA(cudaMemcpyAsync(up->coef_table, rows.data(), sizeof(float) * 8 * n, cudaMemcpyHostToDevice, st));
    if (temb_tables(up, n, st)) return -1;
    up->loop_n = n;
  } else if (n != up->loop_n) {
    return fail("ddpm_run: loop length changed between chunks");
  }
  if (count == 0) return 0;
  {  // this chunk's noise source; step / nan_flag are owned by the device (offsetof: the first two ints stay untouched)
    SamplerCtl c{};
    c.noise_first = first;
    c.noise = noise;
    c.seed = seed;
    const size_t off = offsetof(SamplerCtl, noise_first);
    B2V_CUDA(cudaMemcpyAsync((char*)up->ctl + off, (const char*)&c + off, sizeof(SamplerCtl) - off,
                             cudaMemcpyHostToDevice, st));
  }
  up->temb = TembSource{up->proj_all, 0, up->step_dev, proj_rows};
  for (int i = 0; i < count; ++i)
    if (up->ddpm.run(st)) return -1;
  up->loop_pos = first + count;
  return check_launches("ddpm_run");
}

int UNet::ddpm_sample(const float* z_init, const float* cond, float* z_out, int B, int T, int h, int w,
                      const float* coef, int n, const float* noise, unsigned long long seed, cudaStream_t st) {
  if (sampler_begin(z_init, cond, B, T, h, w, st)) return -1;
  if (ddpm_run(coef, n, 0, n, noise, seed, st)) return -1;
  return sampler_end(z_out, st);
}

int UNet::ddpm_step(long long t, const float* coef, const float* noise, cudaStream_t st) {
  UProgram* up = active;
  if (!up) return fail("ddpm_step: call b2v_sampler_begin first");
  if (!noise || !coef) return fail("ddpm_step: null noise / coefficients");
  launch_set_t(up->t_dev, nullptr, nullptr, t, up->B, st);
  up->temb = TembSource{up->proj, proj_rows, nullptr, 0};
  if (up->fwd.run(st)) return -1;
  Coef8 c8;
  for (int i = 0; i < 8; ++i) c8.v[i] = coef[i];
  launch_ddpm_update(up->x_in, up->eps, noise, nullptr, nullptr, c8, up->numel, st);
  g_launches += 2;
  return check_launches("ddpm_step");
}

int UNet::sampler_end(float* z_out, cudaStream_t st) {
  UProgram* up = active;
  if (!up) return fail("sampler_end: no active sampler");
  B2V_CUDA(cudaMemcpyAsync(z_out, up->x_in, up->numel * 4, cudaMemcpyDeviceToDevice, st));
  active = nullptr;
  return 0;
}

}  // namespace b2v
