"""Model-level parity on the B200, through the mirror classes and the C ABI:
  * against the golden vectors of the unmodified reference (tests/golden, tiny configs),
  * against the oracle (oracle/ref_port.py) run in true fp32 on the same GPU at BASELINE.json's full shapes,
  * teacher-forced per-step eps rel-L2 <= 1e-2 and final-volume PSNR within 0.05 dB (north_star tolerances).
"""
import pytest
import torch

from helpers import TINY_UNET, golden, rel_l2, tiny_unet, tiny_vae
from oracle import ref_port as R

pytestmark = pytest.mark.gpu
EPS_TOL = 1e-2  # north_star: per-step noise-prediction relative L2 (fp16 operands, fp32 accumulate)


def _sd(m, dev):
    return {k: v.to(dev) for k, v in m.state_dict().items()}


def test_unet_forward_golden_and_oracle(cuda_dev):
    g = golden("unet_tiny.pt")
    m = tiny_unet(g["seed"]).to(cuda_dev)
    eps = m(g["x"].to(cuda_dev), g["t"].to(cuda_dev), g["c"].to(cuda_dev))
    assert eps.shape == g["eps"].shape and eps.dtype == torch.float32
    assert rel_l2(eps.cpu(), g["eps"]) < EPS_TOL, rel_l2(eps.cpu(), g["eps"])
    with torch.no_grad():
        ref = R.unet_forward(_sd(m, cuda_dev), TINY_UNET, g["x"].to(cuda_dev), g["t"].to(cuda_dev), g["c"].to(cuda_dev))
    assert rel_l2(eps, ref) < EPS_TOL
    # replaying the captured graph gives the SAME BITS: GroupNorm statistics are accumulated as fixed-point integers
    # (order-independent) and every in-CTA reduction runs in a fixed order -- see DESIGN.md "Reproducibility"
    eps2 = m(g["x"].to(cuda_dev), g["t"].to(cuda_dev), g["c"].to(cuda_dev))
    assert torch.equal(eps2, eps)
    # samples of a batch do not interact (another batch size takes another tile schedule, hence not bitwise)
    one = m(g["x"][:1].to(cuda_dev), g["t"][:1].to(cuda_dev), g["c"][:1].to(cuda_dev))
    assert rel_l2(one, eps[:1]) < 5e-3


def test_unet_rejects_cpu_input():
    m = tiny_unet(0)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 4, 8, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 4, 4, 8, 8))


def test_vae_golden(cuda_dev):
    g = golden("vae_tiny.pt")
    vae = tiny_vae(g["seed"]).to(cuda_dev)
    z = vae.encode(g["v"].to(cuda_dev))
    rec = vae.decode(g["z"].to(cuda_dev))
    assert z.shape == g["z"].shape and rec.shape == g["recon"].shape
    assert rel_l2(z.cpu(), g["z"]) < EPS_TOL, rel_l2(z.cpu(), g["z"])
    assert rel_l2(rec.cpu(), g["recon"]) < EPS_TOL, rel_l2(rec.cpu(), g["recon"])
    rec2, z2 = vae(g["v"].to(cuda_dev))
    assert torch.equal(z2, z) and rec2.shape == rec.shape
    assert torch.equal(vae.decode(g["z"].to(cuda_dev)), rec)


def test_ddim_teacher_forced_and_final(cuda_dev):
    from v2v_b200.inference import DDIMSampler
    from v2v_b200.models import GaussianDiffusion
    g = golden("ddim_tiny.pt")
    m = tiny_unet(0).to(cuda_dev)
    cond = g["cond"].to(cuda_dev)
    worst = 0.0
    for s in g["steps"]:  # teacher-forced on the reference's own trajectory
        e = m(s["z"].to(cuda_dev), torch.tensor([s["t"]], device=cuda_dev), cond)
        worst = max(worst, rel_l2(e.cpu(), s["eps"]))
    assert worst < EPS_TOL, worst
    diff = GaussianDiffusion("cosine", 1000).to(cuda_dev)
    smp = DDIMSampler(diff, m)
    assert smp._get_timesteps(5).tolist() == [999, 800, 600, 400, 200, 0]
    # free-running loop from the reference's initial noise (CPU and CUDA randn streams differ, so inject it)
    z0 = g["steps"][0]["z"].to(cuda_dev)
    orig = torch.randn
    try:
        torch.randn = lambda *a, **k: z0.clone()
        z = smp.sample((1, 4, 4, 8, 8), cond, 5, cuda_dev, progress=False)
    finally:
        torch.randn = orig
    assert rel_l2(z.cpu(), g["z_final"]) < 5e-2, rel_l2(z.cpu(), g["z_final"])
    assert smp.last_nan_flag.item() == 0


def test_ddim_loop_equals_oracle_loop_on_gpu(cuda_dev):
    """same seed on the same device => same noise stream as the oracle's (reference-ordered) draws, eta 0 and > 0"""
    from v2v_b200.inference import DDIMSampler
    from v2v_b200.models import GaussianDiffusion
    m = tiny_unet(0).to(cuda_dev)
    sd = _sd(m, cuda_dev)
    cond = golden("ddim_tiny.pt")["cond"].to(cuda_dev)
    model = lambda z, t, c: R.unet_forward(sd, TINY_UNET, z, t, c)  # noqa: E731
    buf = R.diffusion_buffers("cosine", 1000)
    diff = GaussianDiffusion("cosine", 1000).to(cuda_dev)
    for eta in (0.0, 0.5):
        torch.manual_seed(7)
        with torch.no_grad():
            ref = R.ddim_sample(model, buf, (1, 4, 4, 8, 8), cond, 5, cuda_dev, eta=eta)
        torch.manual_seed(7)
        got = DDIMSampler(diff, m).sample((1, 4, 4, 8, 8), cond, 5, cuda_dev, eta=eta, progress=False)
        assert rel_l2(got, ref) < 5e-2, (eta, rel_l2(got, ref))


def test_ddim_generic_model_path(cuda_dev):
    """DDIMSampler accepts any callable model (as the reference does); the update still runs on our kernel"""
    from v2v_b200.inference import DDIMSampler
    from v2v_b200.models import GaussianDiffusion
    diff = GaussianDiffusion("cosine", 1000)
    cond = torch.zeros((1, 4, 2, 4, 4), device=cuda_dev)
    fake = lambda z, t, c: 0.5 * z  # noqa: E731
    torch.manual_seed(3)
    got = DDIMSampler(diff, fake).sample((1, 4, 2, 4, 4), cond, 4, cuda_dev, progress=False)
    torch.manual_seed(3)
    ref = R.ddim_sample(fake, R.diffusion_buffers("cosine", 1000), (1, 4, 2, 4, 4), cond, 4, cuda_dev)
    assert torch.equal(got, ref)  # fp32 update is bit-exact against the reference formula


def test_ddpm_short_schedule(cuda_dev):
    from v2v_b200.inference import DDPMSampler
    from v2v_b200.models import GaussianDiffusion
    m = tiny_unet(0).to(cuda_dev)
    sd = _sd(m, cuda_dev)
    cond = golden("ddpm_tiny.pt")["cond"].to(cuda_dev)
    model = lambda z, t, c: R.unet_forward(sd, TINY_UNET, z, t, c)  # noqa: E731
    torch.manual_seed(9)
    with torch.no_grad():
        ref = R.ddpm_sample(model, R.diffusion_buffers("cosine", 12), (1, 4, 4, 8, 8), cond, cuda_dev)
    torch.manual_seed(9)
    got = DDPMSampler(GaussianDiffusion("cosine", 12).to(cuda_dev), m).sample((1, 4, 4, 8, 8), cond, cuda_dev,
                                                                              progress=False)
    assert rel_l2(got, ref) < 3e-2, rel_l2(got, ref)


def test_generate_tiny_against_reference_golden(cuda_dev):
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(g["config"]).eval().to(cuda_dev)
    # oracle on the GPU with the same seed = the reference's RNG order on this device
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        ref = R.generate(_sd(m, cuda_dev), g["config"], g["v_in"].to(cuda_dev), "ddim", g["steps"],
                         target_depth=g["target_depth"])
    torch.manual_seed(g["sample_seed"])
    got = m.generate(g["v_in"].to(cuda_dev), "ddim", g["steps"], target_depth=g["target_depth"])
    assert got.shape == g["v_out"].shape
    # free-running through the reference's x0 clamp (+-10 after a x9880 division at t=999) is chaotic, so this is
    # a loose end-to-end sanity bound; the strict gates are the teacher-forced steps and the PSNR test below
    assert rel_l2(got, ref) < 0.2, rel_l2(got, ref)
    n = lambda v: (v.clamp(-1, 1) + 1) / 2  # noqa: E731
    assert R.psnr(n(got), n(ref)) > 25.0
    with pytest.raises(ValueError):
        m.generate(g["v_in"].to(cuda_dev), "euler")


# ------------------------------------------------------------------------------------- BASELINE.json shapes
@pytest.fixture(scope="module")
def bench_model(cuda_dev):
    import os
    import yaml
    from v2v_b200.models import VideoToVideoDiffusion
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "golden", "slice_interpolation_full_medium.yaml")))
    torch.manual_seed(0)
    m = VideoToVideoDiffusion(cfg).eval().to(cuda_dev)
    return m, cfg


@pytest.mark.timeout(600)
def test_full_unet_step_vs_fp32_oracle(cuda_dev, bench_model):
    """config 1 shape: latent (1,8,48,48,48); per-step eps rel-L2 <= 1e-2 at early / middle / late timesteps"""
    m, cfg = bench_model
    _, unet_cfg, _ = R.resolve_config(cfg)
    sd = {k[len("unet."):]: v for k, v in m.state_dict().items() if k.startswith("unet.")}
    g = torch.Generator().manual_seed(5)
    x = torch.randn((1, 8, 48, 48, 48), generator=g).to(cuda_dev)
    c = torch.randn((1, 8, 48, 48, 48), generator=g).to(cuda_dev)
    for tv in (999, 500, 0):
        t = torch.tensor([tv], device=cuda_dev)
        got = m.unet(x, t, c)
        with torch.no_grad():
            ref = R.unet_forward(sd, unet_cfg, x, t, c)
        err = rel_l2(got, ref)
        print(f"full U-Net step t={tv}: rel-L2 = {err:.3e}")
        assert err < EPS_TOL, (tv, err)


@pytest.mark.timeout(600)
def test_full_vae_decode_and_encode_vs_fp32_oracle(cuda_dev, bench_model):
    m, cfg = bench_model
    sd = {k[len("vae."):]: v for k, v in m.state_dict().items() if k.startswith("vae.")}
    g = torch.Generator().manual_seed(6)
    z = torch.randn((1, 8, 48, 48, 48), generator=g).to(cuda_dev)
    got = m.vae.decode(z)
    with torch.no_grad():
        ref = R.vae_decode(sd, z, 1.0)
    assert got.shape == (1, 1, 48, 192, 192)
    err = rel_l2(got, ref)
    psnr = R.psnr((got.clamp(-1, 1) + 1) / 2, (ref.clamp(-1, 1) + 1) / 2)
    print(f"VAE decode: rel-L2 = {err:.3e}, PSNR(new, ref) = {psnr:.1f} dB")
    assert err < EPS_TOL, err
    del ref
    v = (torch.rand((1, 1, 8, 192, 192), generator=g) * 2 - 1).to(cuda_dev)
    ze = m.vae.encode(v)
    with torch.no_grad():
        zr = R.vae_encode(sd, v, 1.0)
    assert ze.shape == (1, 8, 8, 48, 48)
    assert rel_l2(ze, zr) < EPS_TOL, rel_l2(ze, zr)


def _oracle_generate_parts(sd, cfg, v_thick, steps, seed):
    """the reference's generate() (models/model.py:230-343) restated piece by piece on the GPU in true fp32, keeping the
    intermediates the teacher-forced gates need: conditioning, per-step (z_t, t, eps) records, z_0, decoded volume"""
    import torch.nn.functional as F
    vae_cfg, unet_cfg, _ = R.resolve_config(cfg)
    dev = v_thick.device
    torch.manual_seed(seed)
    with torch.no_grad():
        z_in = R.vae_encode(sd, v_thick, vae_cfg["scaling_factor"], "vae.")
        cond = F.interpolate(z_in, size=(48, z_in.shape[3], z_in.shape[4]), mode="trilinear", align_corners=False)
        torch.randn(tuple(cond.shape), device=dev)  # discarded draw (models/model.py:303)
        buffers = {k[len("diffusion."):]: v for k, v in sd.items() if k.startswith("diffusion.")}
        rec = []
        model = lambda z, t, c: R.unet_forward(sd, unet_cfg, z, t, c, "unet.")  # noqa: E731
        z0 = R.ddim_sample(model, buffers, tuple(cond.shape), cond, steps, dev, record=rec)
        out = R.vae_decode(sd, z0, vae_cfg["scaling_factor"], "vae.")
    return cond, rec, z0, out


@pytest.mark.timeout(2400)
@pytest.mark.parametrize("batch", [1, 4], ids=["config1_b1", "config2_b4"])
def test_full_generate_ddim50_all_gates(cuda_dev, bench_model, batch):
    """BASELINE configs[0] (batch 1) and configs[1] (batch 4, the benchmarked one) end to end:
    (B,1,8,192,192) -> (B,1,48,192,192), DDIM-50 = 51 U-Net evaluations.  Gates:
      1. teacher-forced on the fp32 reference trajectory: eps rel-L2 <= 1e-2 at EVERY one of the 51 steps;
      2. teacher-forced decode of the reference z_0: PSNR(new, ref) >= 60 dB;
      3. free-running generate(): |PSNR(new, target) - PSNR(ref, target)| <= 0.05 dB (direct PSNR(new, ref) reported);
      4. a second run of generate() with the same seed returns the same bits."""
    m, cfg = bench_model
    sd = _sd(m, cuda_dev)
    g = torch.Generator().manual_seed(1234)
    v_thick = (torch.rand((batch, 1, 8, 192, 192), generator=g) * 2 - 1).to(cuda_dev)
    target = (torch.rand((batch, 1, 48, 192, 192), generator=g) * 2 - 1).to(cuda_dev)
    cond, rec, z0_ref, ref = _oracle_generate_parts(sd, cfg, v_thick, 50, seed=42)
    assert len(rec) == 51 and [r[1] for r in rec][:3] == [999, 980, 960] and rec[-1][1] == 0
    # gate 1: every step of the reference trajectory
    errs = []
    for z_t, t_idx, eps_ref in rec:
        t = torch.full((batch,), t_idx, device=cuda_dev, dtype=torch.long)
        errs.append(rel_l2(m.unet(z_t, t, cond), eps_ref))
    print(f"batch {batch}: teacher-forced eps rel-L2 over 51 steps: max {max(errs):.3e} (t={rec[errs.index(max(errs))][1]}), "
          f"mean {sum(errs) / len(errs):.3e}")
    assert max(errs) <= EPS_TOL, errs
    n = lambda v: (v.clamp(-1, 1) + 1) / 2  # noqa: E731
    # gate 2: decode of the reference latent
    dec = m.vae.decode(z0_ref)
    p_dec = R.psnr(n(dec), n(ref))
    print(f"batch {batch}: teacher-forced decode PSNR(new, ref) = {p_dec:.1f} dB, rel-L2 {rel_l2(dec, ref):.3e}")
    assert p_dec >= 60.0, p_dec
    # the encoder + depth upsample feeding the loop
    assert rel_l2(m.vae.encode(v_thick), R.vae_encode(sd, v_thick, 1.0, "vae.")) < EPS_TOL
    # gate 3: free-running
    torch.manual_seed(42)
    got = m.generate(v_thick, "ddim", 50, target_depth=48)
    assert got.shape == ref.shape == (batch, 1, 48, 192, 192) and torch.isfinite(got).all()
    p_new, p_ref, p_direct = R.psnr(n(got), n(target)), R.psnr(n(ref), n(target)), R.psnr(n(got), n(ref))
    print(f"batch {batch}: generate DDIM-50 PSNR(new,target)={p_new:.4f} PSNR(ref,target)={p_ref:.4f} "
          f"PSNR(new,ref)={p_direct:.2f} dB")
    assert abs(p_new - p_ref) <= 0.05, (p_new, p_ref)
    assert m.last_nan_flag.item() == 0
    # gate 4: reproducibility
    torch.manual_seed(42)
    again = m.generate(v_thick, "ddim", 50, target_depth=48)
    assert torch.equal(again, got)


@pytest.mark.timeout(900)
def test_full_shape_batch4_unet_and_vae_vs_fp32_oracle(cuda_dev, bench_model):
    """the benchmarked plans (batch 4: no split-K, partially filled 6x6 level, CTA-pair units, attention depth splits)
    on N(0,1) inputs at t = 999 / 500 / 0, and the VAE at batch 4"""
    m, cfg = bench_model
    _, unet_cfg, _ = R.resolve_config(cfg)
    sd = {k[len("unet."):]: v for k, v in m.state_dict().items() if k.startswith("unet.")}
    g = torch.Generator().manual_seed(15)
    x = torch.randn((4, 8, 48, 48, 48), generator=g).to(cuda_dev)
    c = torch.randn((4, 8, 48, 48, 48), generator=g).to(cuda_dev)
    for tv in ([999] * 4, [500] * 4, [0, 250, 750, 999]):  # incl. per-sample timesteps (training forward)
        t = torch.tensor(tv, device=cuda_dev)
        got = m.unet(x, t, c)
        with torch.no_grad():
            ref = R.unet_forward(sd, unet_cfg, x, t, c)
        errs = [rel_l2(got[i], ref[i]) for i in range(4)]
        print(f"batch-4 U-Net step t={tv}: per-sample rel-L2 {['%.2e' % e for e in errs]}")
        assert max(errs) < EPS_TOL, (tv, errs)
    del ref, got
    vsd = {k[len("vae."):]: v for k, v in m.state_dict().items() if k.startswith("vae.")}
    z = torch.randn((4, 8, 48, 48, 48), generator=g).to(cuda_dev)
    got = m.vae.decode(z)
    with torch.no_grad():
        ref = R.vae_decode(vsd, z, 1.0)
    n = lambda a: (a.clamp(-1, 1) + 1) / 2  # noqa: E731
    print(f"batch-4 VAE decode: rel-L2 {rel_l2(got, ref):.3e}, PSNR {R.psnr(n(got), n(ref)):.1f} dB")
    assert rel_l2(got, ref) < EPS_TOL and R.psnr(n(got), n(ref)) >= 60.0
    del ref, got
    v = (torch.rand((4, 1, 8, 192, 192), generator=g) * 2 - 1).to(cuda_dev)
    with torch.no_grad():
        zr = R.vae_encode(vsd, v, 1.0)
    assert rel_l2(m.vae.encode(v), zr) < EPS_TOL


def test_stitching_same_depth_and_volume_generation(cuda_dev):
    """sample_with_stitching mirrors the reference (same-depth windows); generate_volume adds the depth upsample the
    reference stitcher lacks and equals a manual blend of per-window generate() calls"""
    from v2v_b200.inference import DDIMSampler
    from v2v_b200.inference.volume import generate_volume, window_starts
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(g["config"]).eval().to(cuda_dev)
    gen = torch.Generator().manual_seed(31)
    vol = (torch.rand((1, 1, 4, 32, 32), generator=gen) * 2 - 1).to(cuda_dev)
    smp = DDIMSampler(m.diffusion, m.unet)
    torch.manual_seed(5)
    out = smp.sample_with_stitching(vol, m.vae, num_inference_steps=2, patch_size=(4, 16, 16),
                                    target_patch_size=(4, 16, 16), stride=(4, 8, 8), device=cuda_dev, progress=False)
    assert out.shape == vol.shape and torch.isfinite(out).all()
    with pytest.raises(RuntimeError):  # 2 -> 6 slices through the reference-style stitcher fails, as in the reference
        smp.sample_with_stitching(vol, m.vae, 2, (2, 16, 16), (6, 16, 16), (2, 8, 8), cuda_dev, progress=False)
    # thick -> thin volume: 4 thick slices -> 12 thin, 3x3 windows in-plane, 2 along depth
    kw = dict(patch_size=(2, 16, 16), target_patch_size=(6, 16, 16), stride=(2, 8, 8))
    torch.manual_seed(6)
    full = generate_volume(m, vol, "ddim", 2, batch=1, **kw)
    assert full.shape == (1, 1, 12, 32, 32) and torch.isfinite(full).all() and full.abs().max() <= 1.0
    torch.manual_seed(6)
    patches, starts = [], []
    for (d0, h0, w0) in window_starts(4, 32, 32, kw["patch_size"], kw["stride"]):
        patches.append(m.generate(vol[:, :, d0:d0 + 2, h0:h0 + 16, w0:w0 + 16].contiguous(), "ddim", 2, target_depth=6))
        starts.append((d0 * 3, h0, w0))
    ref = R.stitch(patches, starts, (1, 1, 12, 32, 32))
    # same windows, same seeds; not bitwise because two free-running DDIM loops differ at the fp16 floor
    assert rel_l2(full, ref) < 0.1, rel_l2(full, ref)
    # sharded over two ranks: the partial (accumulator, weight) pairs, summed and normalised, ARE the single-rank volume.
    # Every window gets the same initial noise (randn patched to a fixed per-sample tensor) and windows run one per
    # launch, so a window's decoded patch does not depend on which rank / batch it falls into -- what is left is the
    # fp32 association order of the blend ((a+b)+c on one rank vs a+(b+c) across ranks).
    from v2v_b200 import ops
    orig, cache = torch.randn, {}

    def same_noise_for_every_window(shape, **k):
        key = tuple(shape[1:])
        if key not in cache:
            cache[key] = orig((1,) + key, generator=torch.Generator().manual_seed(99)).to(cuda_dev)
        return cache[key].expand(tuple(shape)).clone()

    try:
        torch.randn = same_noise_for_every_window
        single = generate_volume(m, vol, "ddim", 2, batch=1, **kw)
        parts = [generate_volume(m, vol, "ddim", 2, batch=1, rank=r, world=2, **kw) for r in range(2)]
    finally:
        torch.randn = orig
    n_items = len(window_starts(4, 32, 32, kw["patch_size"], kw["stride"]))
    assert n_items == 18  # 2 x 3 x 3 windows -> 9 per rank
    assert all((p[1] > 0).any() for p in parts) and not torch.equal(parts[0][1], parts[1][1])
    acc, wsum = parts[0][0] + parts[1][0], parts[0][1] + parts[1][1]
    assert (wsum > 0).all()
    merged = ops.stitch_normalize(acc, wsum)
    assert single.shape == merged.shape == (1, 1, 12, 32, 32)
    assert torch.allclose(merged, single, rtol=0, atol=2e-6), (merged - single).abs().max().item()


@pytest.mark.timeout(1200)
def test_config3_full_512_volume(cuda_dev, bench_model):
    """BASELINE config 3: one full 512x512 slab, 8 thick -> 48 thin: latent (1,8,48,128,128).
    U-Net step parity at that shape, then encode + DDIM (3 evaluations) + decode end to end vs the fp32 oracle."""
    m, cfg = bench_model
    _, unet_cfg, _ = R.resolve_config(cfg)
    sd = _sd(m, cuda_dev)
    usd = {k[len("unet."):]: v for k, v in sd.items() if k.startswith("unet.")}
    g = torch.Generator().manual_seed(8)
    x = torch.randn((1, 8, 48, 128, 128), generator=g).to(cuda_dev)
    c = torch.randn((1, 8, 48, 128, 128), generator=g).to(cuda_dev)
    t = torch.tensor([500], device=cuda_dev)
    got = m.unet(x, t, c)
    with torch.no_grad():
        ref = R.unet_forward(usd, unet_cfg, x, t, c)
    err = rel_l2(got, ref)
    print(f"512^2 U-Net step: rel-L2 = {err:.3e}")
    assert err < EPS_TOL, err
    del ref, got, x, c
    v = (torch.rand((1, 1, 8, 512, 512), generator=g) * 2 - 1).to(cuda_dev)
    target = (torch.rand((1, 1, 48, 512, 512), generator=g) * 2 - 1).to(cuda_dev)
    torch.manual_seed(42)
    out = m.generate(v, "ddim", 2, target_depth=48)
    torch.manual_seed(42)
    with torch.no_grad():
        ref = R.generate(sd, cfg, v, "ddim", 2, target_depth=48)
    assert out.shape == ref.shape == (1, 1, 48, 512, 512)
    n = lambda a: (a.clamp(-1, 1) + 1) / 2  # noqa: E731
    p_new, p_ref = R.psnr(n(out), n(target)), R.psnr(n(ref), n(target))
    print(f"512^2 generate: PSNR(new,target)={p_new:.4f} PSNR(ref,target)={p_ref:.4f} PSNR(new,ref)={R.psnr(n(out), n(ref)):.2f}")
    assert abs(p_new - p_ref) <= 0.05


@pytest.mark.timeout(900)
def test_config4_ddpm_batch4_full_shape_short_schedule(cuda_dev):
    """BASELINE config 4 (DDPM, batch 32 sharded over 8 GPUs = 4 patches per rank) at the per-rank shape
    (4,8,48,48,48): the whole-loop C entry on a 4-step schedule against the fp32 oracle's ancestral loop fed the same
    torch noise stream (the 1000-step loop is this step 1000 times; the update itself is bit-exact, see
    test_ddpm_update_bit_exact_and_teacher_forced_step)."""
    import os
    import yaml
    from v2v_b200.models import VideoToVideoDiffusion
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "golden", "slice_interpolation_full_medium.yaml")))
    cfg = dict(cfg, diffusion_timesteps=4)
    torch.manual_seed(0)
    m = VideoToVideoDiffusion(cfg).eval().to(cuda_dev)
    sd = _sd(m, cuda_dev)
    _, unet_cfg, diff_cfg = R.resolve_config(cfg)
    g = torch.Generator().manual_seed(44)
    cond = torch.randn((4, 8, 48, 48, 48), generator=g).to(cuda_dev)
    model = lambda z, t, c: R.unet_forward(sd, unet_cfg, z, t, c, "unet.")  # noqa: E731
    torch.manual_seed(9)
    with torch.no_grad():
        ref = R.ddpm_sample(model, R.diffusion_buffers(**diff_cfg), tuple(cond.shape), cond, cuda_dev)
    torch.manual_seed(9)
    got = m.diffusion.p_sample_loop(m.unet, tuple(cond.shape), cond, cuda_dev, progress=False)
    errs = [rel_l2(got[i], ref[i]) for i in range(4)]
    print(f"config-4 shape, DDPM 4-step schedule, batch 4: per-sample rel-L2 {['%.2e' % e for e in errs]}")
    assert max(errs) < 2e-2, errs
    torch.manual_seed(9)
    assert torch.equal(m.diffusion.p_sample_loop(m.unet, tuple(cond.shape), cond, cuda_dev, progress=False), got)


@pytest.mark.timeout(1200)
def test_config5_stitched_volume_vs_oracle(cuda_dev, bench_model):
    """BASELINE config 5 in miniature at the real window shape: a (1,1,8,288,288) slab -> 2 x 2 windows of 192 x 192
    (stride 96), each encode -> 8->48 depth upsample -> DDIM -> decode, Gaussian-blended -- generate_volume (windows
    batched 4 at a time) against the oracle's per-window generate() + the reference's blend (inference/sampler.py:379-451)"""
    from v2v_b200.inference.volume import generate_volume, window_starts
    m, cfg = bench_model
    sd = _sd(m, cuda_dev)
    g = torch.Generator().manual_seed(55)
    vol = (torch.rand((1, 1, 8, 288, 288), generator=g) * 2 - 1).to(cuda_dev)
    starts = window_starts(8, 288, 288)
    assert starts == [(0, 0, 0), (0, 0, 96), (0, 96, 0), (0, 96, 96)]
    # the same initial noise for every window on both sides (the oracle runs windows one at a time, ours in one batch)
    orig, cache = torch.randn, {}

    def same_noise(shape, **k):
        key = tuple(shape[1:])
        if key not in cache:
            cache[key] = orig((1,) + key, generator=torch.Generator().manual_seed(7)).to(cuda_dev)
        return cache[key].expand(tuple(shape)).clone()

    try:
        torch.randn = same_noise
        got = generate_volume(m, vol, "ddim", 2, batch=4)
        with torch.no_grad():
            patches = [R.generate(sd, cfg, vol[:, :, :, h0:h0 + 192, w0:w0 + 192].contiguous(), "ddim", 2, target_depth=48)
                       for (_, h0, w0) in starts]
    finally:
        torch.randn = orig
    ref = R.stitch(patches, [(0, h0, w0) for (_, h0, w0) in starts], (1, 1, 48, 288, 288))
    n = lambda a: (a.clamp(-1, 1) + 1) / 2  # noqa: E731
    err, p = rel_l2(got, ref), R.psnr(n(got), n(ref))
    print(f"config-5 stitched (1,1,48,288,288): rel-L2 {err:.3e}, PSNR(new, ref) {p:.1f} dB")
    assert got.shape == ref.shape and err < 0.1 and p > 35.0, (err, p)  # free-running DDIM-2, like the config-3 test


def test_unet_192_channels_group_widths_that_are_not_powers_of_two(cuda_dev):
    """model_channels = 192 (a multiple of 64 the reference accepts): 24 and 6 channels per GroupNorm group, so the conv
    epilogue's power-of-two statistics fold does not apply and the planner inserts a statistics pass; Cout = 192 / 384
    run on the BN = 64 / 128 single-CTA conv kernels (such configs used to fail at program-build time)"""
    from v2v_b200.models import UNet3D
    cfg = dict(latent_dim=4, model_channels=192, num_res_blocks=1, attention_levels=[1], channel_mult=(1, 2), num_heads=2,
               time_embed_dim=128)
    torch.manual_seed(3)
    m = UNet3D(**cfg).eval().to(cuda_dev)
    sd = _sd(m, cuda_dev)
    g = torch.Generator().manual_seed(12)
    x = torch.randn((2, 4, 5, 8, 12), generator=g).to(cuda_dev)
    c = torch.randn((2, 4, 5, 8, 12), generator=g).to(cuda_dev)
    t = torch.tensor([999, 17], device=cuda_dev)
    with torch.no_grad():
        ref = R.unet_forward(sd, cfg, x, t, c)
    got = m(x, t, c)
    err = rel_l2(got, ref)
    print(f"U-Net with model_channels 192: rel-L2 = {err:.3e}")
    assert err < EPS_TOL, err
    assert torch.equal(m(x, t, c), got)


def test_ragged_shapes_empty_batch_and_nan_input(cuda_dev):
    """edge cases: odd depth / non-square latent (partial TMA boxes everywhere), batch 3, empty batch, NaNs in v_in"""
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(g["config"]).eval().to(cuda_dev)
    sd = _sd(m, cuda_dev)
    _, unet_cfg, _ = R.resolve_config(g["config"])
    usd = {k[5:]: w for k, w in sd.items() if k.startswith("unet.")}
    gen = torch.Generator().manual_seed(77)
    for shape in [(3, 4, 5, 6, 10), (1, 4, 1, 2, 2), (2, 4, 7, 14, 6)]:
        x = torch.randn(shape, generator=gen).to(cuda_dev)
        c = torch.randn(shape, generator=gen).to(cuda_dev)
        t = torch.randint(0, 1000, (shape[0],), generator=gen).to(cuda_dev)
        with torch.no_grad():
            ref = R.unet_forward(usd, unet_cfg, x, t, c)
        assert rel_l2(m.unet(x, t, c), ref) < EPS_TOL, shape
    vsd = {k[4:]: w for k, w in sd.items() if k.startswith("vae.")}
    v = (torch.rand((2, 1, 3, 20, 28), generator=gen) * 2 - 1).to(cuda_dev)
    with torch.no_grad():
        zr = R.vae_encode(vsd, v, 0.5)
        rr = R.vae_decode(vsd, zr, 0.5)
    assert rel_l2(m.vae.encode(v), zr) < EPS_TOL and rel_l2(m.vae.decode(zr), rr) < EPS_TOL
    # empty batch
    e = torch.empty((0, 4, 4, 8, 8), device=cuda_dev)
    assert m.unet(e, torch.empty((0,), dtype=torch.long, device=cuda_dev), e).shape == (0, 4, 4, 8, 8)
    assert m.vae.decode(e).shape == (0, 1, 4, 32, 32)
    assert m.generate(torch.empty((0, 1, 2, 16, 16), device=cuda_dev), "ddim", 2, target_depth=6).shape == (0, 1, 6, 16, 16)
    # NaNs in the input are zeroed like the reference does (models/model.py:262-264)
    vin = g["v_in"].to(cuda_dev).clone()
    vin[0, 0, 0, 0, :4] = float("nan")
    torch.manual_seed(1)
    out = m.generate(vin, "ddim", 2, target_depth=6)
    assert torch.isfinite(out).all()
