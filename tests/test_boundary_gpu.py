"""The drop-in boundary on the B200: the single-call C entry points (b2v_generate, b2v_ddpm_sample / b2v_ddpm_run),
driven from raw ctypes calls only, against the mirror classes and the oracle; the DDPM update op; the device noise
generator; generate_batch; the training forward (q_sample + Min-SNR-5 loss)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import TINY_UNET, golden, rel_l2, tiny_unet
from oracle import ref_port as R

pytestmark = pytest.mark.gpu


def _sd(m, dev):
    return {k: v.to(dev) for k, v in m.state_dict().items()}


def _tiny_model(dev):
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    torch.manual_seed(g["seed"])
    return VideoToVideoDiffusion(g["config"]).eval().to(dev), g


def _c_objects(m):
    """build b2v_unet / b2v_vae from a state_dict with nothing but C-ABI calls (what INTEGRATION.md level 2 shows)"""
    from v2v_b200 import _lib
    L = _lib.lib()

    def load(kind, desc, sd):
        h = ctypes.c_void_p()
        _lib.check(getattr(L, f"b2v_{kind}_create")(ctypes.byref(h), ctypes.byref(desc)), "create")
        for key, val in sd.items():
            w = val.detach().to("cpu", torch.float32).contiguous()
            shape = (ctypes.c_int64 * max(1, w.dim()))(*w.shape)
            _lib.check(getattr(L, f"b2v_{kind}_load_weight")(h, key.encode(), ctypes.c_void_p(w.data_ptr()), shape,
                                                             w.dim()), key)
        _lib.check(getattr(L, f"b2v_{kind}_finalize")(h), "finalize")
        return h

    u = load("unet", m.unet._desc(), m.unet.state_dict())
    v = load("vae", _lib.VAEDesc(m.vae.in_channels, m.vae.latent_dim, m.vae.base_channels, float(m.vae.scaling_factor)),
             m.vae.state_dict())
    return L, u, v


def test_generate_from_c_abi_calls_only(cuda_dev):
    """the whole path -- encode, depth upsample, DDIM loop, decode, NaN guards -- as ONE b2v_generate call on objects
    built through the C ABI, no mirror class on the execution path; equals VideoToVideoDiffusion.generate bit for bit"""
    from v2v_b200 import _lib
    m, g = _tiny_model(cuda_dev)
    L, u, v = _c_objects(m)
    try:
        v_in = g["v_in"].to(cuda_dev).contiguous()
        B, C, T_in, H, W = v_in.shape
        T_out, steps = g["target_depth"], g["steps"]
        lat = (B, m.vae.latent_dim, T_out, H // 4, W // 4)
        ts = (ctypes.c_int64 * 64)()
        n = L.b2v_ddim_timesteps(1000, steps, ts, 64)
        assert n == steps + 1
        acp = m.diffusion.alphas_cumprod.detach().float().cpu().contiguous()
        cfg = _lib.SamplerCfg()
        cfg.sampler, cfg.n, cfg.n_train, cfg.eta = 0, n, acp.numel(), 0.0
        cfg.timesteps = ctypes.cast(ts, ctypes.POINTER(ctypes.c_int64))
        cfg.alphas_cumprod = ctypes.cast(acp.data_ptr(), ctypes.POINTER(ctypes.c_float))
        torch.manual_seed(g["sample_seed"])
        torch.randn(lat, device=cuda_dev)  # the reference's discarded draw
        z_init = torch.randn(lat, device=cuda_dev)
        out = torch.empty((B, C, T_out, H, W), device=cuda_dev)
        flag = torch.ones(1, dtype=torch.int32, device=cuda_dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.b2v_generate(u, v, ctypes.byref(cfg), ctypes.c_void_p(v_in.data_ptr()), ctypes.c_void_p(z_init.data_ptr()),
                            ctypes.c_void_p(out.data_ptr()), B, T_in, T_out, H, W, ctypes.c_void_p(flag.data_ptr()), st)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        assert flag.item() == 0 and torch.isfinite(out).all()
        torch.manual_seed(g["sample_seed"])
        via_mirror = m.generate(v_in, "ddim", steps, target_depth=T_out)
        assert torch.equal(out, via_mirror)
        torch.manual_seed(g["sample_seed"])
        with torch.no_grad():
            ref = R.generate(_sd(m, cuda_dev), g["config"], v_in, "ddim", steps, target_depth=T_out)
        assert rel_l2(out, ref) < 0.2
        # NaN in the input: zeroed like the reference (models/model.py:261-263) and flagged
        bad = v_in.clone()
        bad[0, 0, 0, 0, :3] = float("nan")
        rc = L.b2v_generate(u, v, ctypes.byref(cfg), ctypes.c_void_p(bad.data_ptr()), ctypes.c_void_p(z_init.data_ptr()),
                            ctypes.c_void_p(out.data_ptr()), B, T_in, T_out, H, W, ctypes.c_void_p(flag.data_ptr()), st)
        assert rc == 0 and flag.item() == 1 and torch.isfinite(out).all()
        # error behaviour: bad sampler id, mismatching shape
        cfg.sampler = 7
        assert L.b2v_generate(u, v, ctypes.byref(cfg), ctypes.c_void_p(v_in.data_ptr()), ctypes.c_void_p(z_init.data_ptr()),
                              ctypes.c_void_p(out.data_ptr()), B, T_in, T_out, H, W, None, st) != 0
        assert "unknown sampler" in _lib.last_error()
    finally:
        L.b2v_unet_destroy(u)
        L.b2v_vae_destroy(v)


def test_generate_same_depth_ragged_batch_and_shape_errors(cuda_dev):
    """generate() without a target depth (no upsample stage), batch 3, non-square 24 x 40 slices (partial TMA boxes in
    every layer); shape contract errors surface as ValueError / RuntimeError before anything is launched"""
    m, g = _tiny_model(cuda_dev)
    gen = torch.Generator().manual_seed(71)
    v = (torch.rand((3, 1, 5, 24, 40), generator=gen) * 2 - 1).to(cuda_dev)
    torch.manual_seed(2)
    got = m.generate(v, "ddim", 3)
    assert got.shape == (3, 1, 5, 24, 40) and torch.isfinite(got).all() and m.last_nan_flag.item() == 0
    torch.manual_seed(2)
    with torch.no_grad():
        ref = R.generate(_sd(m, cuda_dev), g["config"], v, "ddim", 3)
    n = lambda a: (a.clamp(-1, 1) + 1) / 2  # noqa: E731
    assert rel_l2(got, ref) < 0.2 and R.psnr(n(got), n(ref)) > 25.0, (rel_l2(got, ref), R.psnr(n(got), n(ref)))
    torch.manual_seed(2)
    assert torch.equal(m.generate(v, "ddim", 3), got)
    with pytest.raises(ValueError):
        m.generate(v[:, :, :, :18], "ddim", 3)  # H % 4 != 0
    with pytest.raises(ValueError):
        m.generate(v.expand(3, 2, 5, 24, 40).contiguous(), "ddim", 3)  # channel count
    with pytest.raises(RuntimeError):
        m.generate(v.cpu(), "ddim", 3)
    with pytest.raises(ValueError):
        m.generate(v, "euler")


def test_program_cache_evicts_least_recently_used_shape_only(cuda_dev):
    """more input shapes than cached programs (6): the least recently used plan is dropped, the others keep their
    captured graphs, and a re-planned shape reproduces its first result bit for bit (clear-all thrashing)"""
    m = tiny_unet(0).to(cuda_dev)
    gen = torch.Generator().manual_seed(3)
    shapes = [(1, 4, 2, 4, 4), (1, 4, 3, 4, 4), (2, 4, 2, 4, 6), (1, 4, 4, 8, 8), (1, 4, 2, 8, 4), (3, 4, 2, 4, 4),
              (1, 4, 5, 6, 6), (2, 4, 3, 6, 4)]
    data, first = [], []
    for shp in shapes:
        x, c = torch.randn(shp, generator=gen).to(cuda_dev), torch.randn(shp, generator=gen).to(cuda_dev)
        t = torch.randint(0, 1000, (shp[0],), generator=gen).to(cuda_dev)
        data.append((x, t, c))
        first.append(m(x, t, c))
    for i in (0, 7, 3, 1, 0):  # evicted and still-cached shapes alike
        x, t, c = data[i]
        assert torch.equal(m(x, t, c), first[i]), shapes[i]
    # a sampler keeps working across evictions of other shapes
    from v2v_b200.inference import DDIMSampler
    from v2v_b200.models import GaussianDiffusion
    diff = GaussianDiffusion("cosine", 1000).to(cuda_dev)
    torch.manual_seed(1)
    a = DDIMSampler(diff, m).sample(shapes[3], data[3][2], 3, cuda_dev, progress=False)
    for i in (4, 5, 6, 7, 0, 1, 2):
        m(*data[i])
    torch.manual_seed(1)
    assert torch.equal(DDIMSampler(diff, m).sample(shapes[3], data[3][2], 3, cuda_dev, progress=False), a)


def test_ddpm_update_bit_exact_and_teacher_forced_step(cuda_dev):
    """the ancestral update reproduces the reference's eager fp32 chain bit for bit (models/diffusion.py:287-338), and
    one full DDPM step (U-Net + update) teacher-forced on the reference's z_t stays within the eps tolerance"""
    from v2v_b200 import _lib, ops
    from v2v_b200.models import GaussianDiffusion
    diff = GaussianDiffusion("cosine", 1000).to(cuda_dev)
    rows = diff.ddpm_coefficients()
    buf = {k: v.to(cuda_dev) for k, v in R.diffusion_buffers("cosine", 1000).items()}
    g = torch.Generator().manual_seed(4)
    z = torch.randn((2, 4, 4, 8, 8), generator=g).to(cuda_dev)
    eps = torch.randn((2, 4, 4, 8, 8), generator=g).to(cuda_dev)
    noise = torch.randn((2, 4, 4, 8, 8), generator=g).to(cuda_dev)
    for t_idx in (999, 998, 500, 20, 1, 0):
        ref = R.ddpm_step(z, eps, t_idx, buf, noise)
        got = ops.ddpm_update(z.clone(), eps, noise, rows[t_idx].tolist())
        assert torch.equal(got, ref), (t_idx, (got - ref).abs().max().item())
        # the reference module's own p_sample formula (mirror's generic path) gives the same bits
        mean, _, logvar = diff.p_mean_variance(lambda *_: eps, z, torch.full((2,), t_idx, device=cuda_dev), None)
        assert torch.equal(mean + (0.0 if t_idx == 0 else 1.0) * torch.exp(0.5 * logvar) * noise, ref)
    # teacher-forced full step through the sampler entry points
    m = tiny_unet(0).to(cuda_dev)
    sd = _sd(m, cuda_dev)
    cond = golden("ddpm_tiny.pt")["cond"].to(cuda_dev)
    zt = torch.randn((1, 4, 4, 8, 8), generator=g).to(cuda_dev)
    nz = torch.randn((1, 4, 4, 8, 8), generator=g).to(cuda_dev)
    L = _lib.lib()
    for t_idx in (999, 400, 0):
        with torch.no_grad():
            eps_ref = R.unet_forward(sd, TINY_UNET, zt, torch.tensor([t_idx], device=cuda_dev), cond)
        ref = R.ddpm_step(zt, eps_ref, t_idx, buf, nz)
        out = torch.empty_like(zt)
        h = m.native(cuda_dev)
        _lib.check(L.b2v_sampler_begin(h, _lib.dptr(zt), _lib.dptr(cond), 1, 4, 8, 8, _lib.stream()), "begin")
        coef = (ctypes.c_float * 8)(*rows[t_idx].tolist())
        _lib.check(L.b2v_ddpm_step(h, t_idx, coef, _lib.dptr(nz), _lib.stream()), "step")
        _lib.check(L.b2v_sampler_end(h, _lib.dptr(out), _lib.stream()), "end")
        assert rel_l2(out, ref) < 1e-2, (t_idx, rel_l2(out, ref))


def test_ddpm_whole_loop_equals_stepwise_and_device_noise(cuda_dev):
    """b2v_ddpm_sample (one call, graph per step, device step counter) == the step-wise API on the same noise, bit for
    bit; chunked b2v_ddpm_run likewise; with noise == NULL the kernel's own Philox draws equal the ones
    b2v_philox_normal reports, which equal the numpy restatement"""
    from oracle import philox as P
    from v2v_b200 import _lib, ops
    from v2v_b200.inference import DDPMSampler
    from v2v_b200.models import GaussianDiffusion
    m = tiny_unet(0).to(cuda_dev)
    n = 12
    diff = GaussianDiffusion("cosine", n).to(cuda_dev)
    rows = diff.ddpm_coefficients()
    coef = ctypes.cast(rows.data_ptr(), ctypes.POINTER(ctypes.c_float))
    cond = golden("ddpm_tiny.pt")["cond"].to(cuda_dev)
    shape = (1, 4, 4, 8, 8)
    g = torch.Generator().manual_seed(11)
    z_init = torch.randn(shape, generator=g).to(cuda_dev)
    noise = torch.randn((n,) + shape, generator=g).to(cuda_dev)
    L, h = _lib.lib(), m.native(cuda_dev)
    whole = torch.empty_like(z_init)
    _lib.check(L.b2v_ddpm_sample(h, _lib.dptr(z_init), _lib.dptr(cond), _lib.dptr(whole), 1, 4, 8, 8, coef, n,
                                 _lib.dptr(noise), 0, _lib.stream()), "ddpm_sample")
    step = torch.empty_like(z_init)
    _lib.check(L.b2v_sampler_begin(h, _lib.dptr(z_init), _lib.dptr(cond), 1, 4, 8, 8, _lib.stream()), "begin")
    for s in range(n):
        t_idx = n - 1 - s
        c8 = (ctypes.c_float * 8)(*rows[t_idx].tolist())
        _lib.check(L.b2v_ddpm_step(h, t_idx, c8, _lib.dptr(noise[s]), _lib.stream()), "step")
    _lib.check(L.b2v_sampler_end(h, _lib.dptr(step), _lib.stream()), "end")
    assert torch.equal(whole, step)
    chunked = torch.empty_like(z_init)
    _lib.check(L.b2v_sampler_begin(h, _lib.dptr(z_init), _lib.dptr(cond), 1, 4, 8, 8, _lib.stream()), "begin")
    for first, count in ((0, 5), (5, 4), (9, 3)):
        _lib.check(L.b2v_ddpm_run(h, coef, n, first, count, _lib.dptr(noise[first:first + count].contiguous()), 0,
                                  _lib.stream()), "run")
    assert L.b2v_ddpm_run(h, coef, n, 3, 2, _lib.dptr(noise[:2].contiguous()), 0, _lib.stream()) != 0  # not contiguous
    assert "contiguous" in _lib.last_error()
    _lib.check(L.b2v_sampler_end(h, _lib.dptr(chunked), _lib.stream()), "end")
    assert torch.equal(chunked, whole)
    # oracle loop on the same noise (free-running, 12 steps)
    sd = _sd(m, cuda_dev)
    buf = {k: v.to(cuda_dev) for k, v in R.diffusion_buffers("cosine", n).items()}
    z = z_init.clone()
    with torch.no_grad():
        for s in range(n):
            t_idx = n - 1 - s
            z = R.ddpm_step(z, R.unet_forward(sd, TINY_UNET, z, torch.tensor([t_idx], device=cuda_dev), cond), t_idx, buf,
                            noise[s])
    assert rel_l2(whole, z) < 3e-2, rel_l2(whole, z)
    # the mirror's p_sample_loop draws the same stream with torch and reaches the same bits, also when chunked
    import os
    torch.manual_seed(5)
    a = DDPMSampler(diff, m).sample(shape, cond, cuda_dev, progress=False)
    os.environ["B2V_DDPM_NOISE_BUDGET_GB"] = str(3.5 * z_init.numel() * 4 / (1 << 30))  # 3 steps per chunk
    try:
        torch.manual_seed(5)
        b = DDPMSampler(diff, m).sample(shape, cond, cuda_dev, progress=False)
    finally:
        del os.environ["B2V_DDPM_NOISE_BUDGET_GB"]
    assert torch.equal(a, b)
    # device generator
    seed = 0x1234ABCD5678
    nz = torch.stack([ops.philox_normal(z_init.numel(), seed, s, cuda_dev).view(shape) for s in range(n)])
    ref_nz = np.stack([P.normal(z_init.numel(), seed, s) for s in range(n)]).reshape(nz.shape)
    assert np.abs(nz.cpu().numpy() - ref_nz).max() < 2e-5
    dev_noise = torch.empty_like(z_init)
    _lib.check(L.b2v_ddpm_sample(h, _lib.dptr(z_init), _lib.dptr(cond), _lib.dptr(dev_noise), 1, 4, 8, 8, coef, n, None,
                                 seed, _lib.stream()), "ddpm_sample")
    from_buf = torch.empty_like(z_init)
    _lib.check(L.b2v_ddpm_sample(h, _lib.dptr(z_init), _lib.dptr(cond), _lib.dptr(from_buf), 1, 4, 8, 8, coef, n,
                                 _lib.dptr(nz.contiguous()), 0, _lib.stream()), "ddpm_sample")
    assert torch.equal(dev_noise, from_buf)
    torch.manual_seed(5)
    c = diff.p_sample_loop(m, shape, cond, cuda_dev, device_rng_seed=seed)
    assert torch.isfinite(c).all() and not torch.equal(c, a)


def test_generate_ddpm_single_call_equals_staged(cuda_dev):
    """generate(sampler='ddpm') as one b2v_generate call == the staged path (chunked noise), and tracks the oracle"""
    import os
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    cfg = dict(g["config"], diffusion_timesteps=10)
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(cfg).eval().to(cuda_dev)
    v_in = g["v_in"].to(cuda_dev)
    torch.manual_seed(3)
    one = m.generate(v_in, "ddpm", target_depth=g["target_depth"])
    os.environ["B2V_DDPM_NOISE_BUDGET_GB"] = "1e-9"
    try:
        torch.manual_seed(3)
        staged = m.generate(v_in, "ddpm", target_depth=g["target_depth"])
    finally:
        del os.environ["B2V_DDPM_NOISE_BUDGET_GB"]
    assert torch.equal(one, staged)
    torch.manual_seed(3)
    with torch.no_grad():
        ref = R.generate(_sd(m, cuda_dev), cfg, v_in, "ddpm", target_depth=g["target_depth"])
    n = lambda v: (v.clamp(-1, 1) + 1) / 2  # noqa: E731
    assert rel_l2(one, ref) < 0.2 and R.psnr(n(one), n(ref)) > 25.0


@pytest.mark.parametrize("sampler", ["ddim", "ddpm"])
def test_generate_batch_on_gpu(cuda_dev, sampler):
    """inference/generate.py:98-155 generate_batch: same-depth encode -> sample -> decode (no depth upsample)"""
    from v2v_b200.inference import DDIMSampler, DDPMSampler
    from v2v_b200.inference.generate import generate_batch
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    cfg = dict(g["config"], diffusion_timesteps=1000 if sampler == "ddim" else 8)
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(cfg).eval()
    gen = torch.Generator().manual_seed(17)
    vids = torch.rand((2, 1, 3, 16, 24), generator=gen) * 2 - 1  # CPU input: generate_batch moves model and data
    torch.manual_seed(8)
    got = generate_batch(m, vids, sampler_type=sampler, num_inference_steps=4, device=cuda_dev)
    assert got.shape == (2, 1, 3, 16, 24) and got.is_cuda and torch.isfinite(got).all()
    # the same composition by hand through the mirror: identical bits
    torch.manual_seed(8)
    z_in = m.vae.encode(vids.to(cuda_dev))
    if sampler == "ddim":
        z0 = DDIMSampler(m.diffusion, m.unet).sample(z_in.shape, z_in, 4, cuda_dev, progress=False)
    else:
        z0 = DDPMSampler(m.diffusion, m.unet).sample(z_in.shape, z_in, cuda_dev, progress=False)
    assert torch.equal(got, m.vae.decode(z0))
    # and against the oracle's composition of the reference ops with the same seed on the same device
    sd = _sd(m, cuda_dev)
    vae_cfg, unet_cfg, diff_cfg = R.resolve_config(cfg)
    model = lambda z, t, c: R.unet_forward(sd, unet_cfg, z, t, c, "unet.")  # noqa: E731
    buffers = R.diffusion_buffers(**diff_cfg)
    torch.manual_seed(8)
    with torch.no_grad():
        zr = R.vae_encode(sd, vids.to(cuda_dev), vae_cfg["scaling_factor"], "vae.")
        assert rel_l2(z_in, zr) < 1e-2
        if sampler == "ddim":
            z0r = R.ddim_sample(model, buffers, tuple(zr.shape), zr, 4, cuda_dev)
        else:
            z0r = R.ddpm_sample(model, buffers, tuple(zr.shape), zr, cuda_dev)
        ref = R.vae_decode(sd, z0r, vae_cfg["scaling_factor"], "vae.")
    n = lambda v: (v.clamp(-1, 1) + 1) / 2  # noqa: E731
    assert rel_l2(got, ref) < 0.2 and R.psnr(n(got), n(ref)) > 25.0, (rel_l2(got, ref), R.psnr(n(got), n(ref)))
    with pytest.raises(ValueError):
        generate_batch(m, vids, sampler_type="euler", device=cuda_dev)


def test_training_forward_loss_matches_reference_formula(cuda_dev):
    """q_sample + U-Net + Min-SNR-5 weighted eps-MSE (models/diffusion.py:81-190), forward only.
    The draws (t, noise) come from torch in the reference's order, so the same seed gives the oracle the same batch."""
    import torch.nn.functional as F
    from v2v_b200.models import GaussianDiffusion
    m = tiny_unet(0).to(cuda_dev)
    sd = _sd(m, cuda_dev)
    diff = GaussianDiffusion("cosine", 1000).to(cuda_dev)
    buf = {k: v.to(cuda_dev) for k, v in R.diffusion_buffers("cosine", 1000).items()}
    g = torch.Generator().manual_seed(23)
    B = 3
    z0 = torch.randn((B, 4, 4, 8, 8), generator=g).to(cuda_dev)
    c = torch.randn((B, 4, 4, 8, 8), generator=g).to(cuda_dev)

    def reference_loss(model, mask):
        t = torch.randint(0, 1000, (B,), device=cuda_dev, dtype=torch.long)
        noise = torch.randn_like(z0)
        e = lambda a: a[t].view(B, 1, 1, 1, 1)  # noqa: E731
        z_t = e(buf["sqrt_alphas_cumprod"]) * z0 + e(buf["sqrt_one_minus_alphas_cumprod"]) * noise
        pred = model(z_t, t, c)
        snr = buf["alphas_cumprod"][t] / (1 - buf["alphas_cumprod"][t] + 1e-8)
        w = torch.clamp(snr, max=5.0) / (snr + 1e-8)
        if mask is None:
            return (F.mse_loss(pred, noise, reduction="none").reshape(B, -1).mean(dim=1) * w).mean(), z_t, t
        me = mask[..., None, None].expand_as(pred)
        mm = (pred - noise) ** 2 * me
        nv = me.reshape(B, -1).sum(dim=1)
        if (nv == nv[0]).all():
            return ((mm.sum() / me.sum()) * w).mean(), z_t, t
        return torch.stack([(mm[i].sum() / nv[i]) * w[i] if nv[i] > 0 else torch.tensor(0.0, device=cuda_dev)
                            for i in range(B)]).mean(), z_t, t

    full = torch.ones((B, 4, 4), device=cuda_dev)
    same = full.clone()
    same[:, :, 3:] = 0
    ragged = full.clone()
    ragged[0, :, 2:] = 0
    ragged[2] = 0
    oracle_model = lambda z, t, cc: R.unet_forward(sd, TINY_UNET, z, t, cc)  # noqa: E731
    for mask in (None, same, ragged):
        torch.manual_seed(77)
        with torch.no_grad():
            ref, z_t_ref, t_ref = reference_loss(oracle_model, mask)
        torch.manual_seed(77)
        loss, info = diff.training_loss(m, z0, c, mask=mask)
        assert abs(loss.item() - ref.item()) <= 2e-2 * abs(ref.item()), (loss.item(), ref.item())
        assert info["mse"] == info["total"] == loss.item()
        # teacher-forced on the same U-Net output the reduction itself is exact to fp32 rounding
        torch.manual_seed(77)
        with torch.no_grad():
            ref_same_model, _, _ = reference_loss(m, mask)
        assert abs(loss.item() - ref_same_model.item()) <= 2e-6 * max(1.0, abs(ref_same_model.item()))
    # q_sample alone: bit-exact against the reference's eager expression (two multiplies, one add, per-sample t)
    t = torch.tensor([0, 500, 999], device=cuda_dev)
    nz = torch.randn((B, 4, 4, 8, 8), generator=g).to(cuda_dev)
    zt, nz_out = diff.q_sample(z0, t, nz)
    e = lambda a: a[t].view(B, 1, 1, 1, 1)  # noqa: E731
    assert nz_out is nz and torch.equal(zt, e(buf["sqrt_alphas_cumprod"]) * z0 + e(buf["sqrt_one_minus_alphas_cumprod"]) * nz)
    torch.manual_seed(1)
    a = diff.training_loss(m, z0, c)[0]
    torch.manual_seed(1)
    assert torch.equal(a, diff.training_loss(m, z0, c)[0])  # deterministic reduction
