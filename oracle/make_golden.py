"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported read-only from /root/reference) on
small seeded cases.  Run once in the build container:  python oracle/make_golden.py
The fixtures pin oracle/ref_port.py (and, through it, the CUDA path) to the reference's own outputs.
Weights are not stored: every case records the seed and a SHA-256 of the reference's state_dict; the mirror
modules reproduce the same tensors from the seed (checked by hash in tests/test_oracle_golden.py).
"""
import contextlib
import hashlib
import io
import os
import sys

import torch

REF = os.environ.get("B2V_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


TINY = {  # flat config: the reference reads U-Net keys from the top level
    "in_channels": 1, "latent_dim": 4, "vae_base_channels": 64, "vae_scaling_factor": 0.5,
    "unet_model_channels": 64, "unet_num_res_blocks": 1, "unet_attention_levels": [1], "unet_channel_mult": [1, 2],
    "unet_num_heads": 2, "unet_time_embed_dim": 128, "noise_schedule": "cosine", "diffusion_timesteps": 1000,
}


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    from models.unet3d import UNet3D
    from models.vae import SliceInterpolationVAE
    from models.diffusion import GaussianDiffusion
    from models.model import VideoToVideoDiffusion
    from inference.sampler import DDIMSampler, DDPMSampler
    import yaml

    # ---- 1. U-Net forward, tiny config (exercises res blocks with/without residual conv, attention, down/up, skip)
    torch.manual_seed(0)
    unet = UNet3D(latent_dim=4, model_channels=64, num_res_blocks=1, attention_levels=[1], channel_mult=(1, 2),
                  num_heads=2, time_embed_dim=128).eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 4, 4, 8, 8), generator=g)
    c = torch.randn((2, 4, 4, 8, 8), generator=g)
    t = torch.tensor([500, 37])
    with torch.no_grad():
        eps = unet(x, t, c)
    torch.save({"seed": 0, "sd_hash": sd_hash(unet.state_dict()), "x": x, "c": c, "t": t, "eps": eps,
                "cfg": dict(latent_dim=4, model_channels=64, num_res_blocks=1, attention_levels=[1],
                            channel_mult=(1, 2), num_heads=2, time_embed_dim=128)},
               os.path.join(OUT, "unet_tiny.pt"))

    # ---- 2. VAE encode / decode, tiny
    torch.manual_seed(1)
    vae = quiet(SliceInterpolationVAE, in_channels=1, latent_dim=4, base_channels=64, scaling_factor=0.5).eval()
    g = torch.Generator().manual_seed(12)
    v = torch.rand((1, 1, 3, 16, 16), generator=g) * 2 - 1
    with torch.no_grad():
        z = vae.encode(v)
        rec = vae.decode(z)
    torch.save({"seed": 1, "sd_hash": sd_hash(vae.state_dict()), "v": v, "z": z, "recon": rec,
                "cfg": dict(in_channels=1, latent_dim=4, base_channels=64, scaling_factor=0.5)},
               os.path.join(OUT, "vae_tiny.pt"))

    # ---- 3. schedule known answers
    diff = GaussianDiffusion("cosine", 1000)
    samp = DDIMSampler(diff, None)
    lin = GaussianDiffusion("linear", 100, 1e-4, 0.02)
    torch.save({"alphas_cumprod": diff.alphas_cumprod.clone(), "betas": diff.betas.clone(),
                "posterior_log_variance_clipped": diff.posterior_log_variance_clipped.clone(),
                "posterior_mean_coef1": diff.posterior_mean_coef1.clone(),
                "posterior_mean_coef2": diff.posterior_mean_coef2.clone(),
                "ts50": torch.tensor(samp._get_timesteps(50).copy()), "ts20": torch.tensor(samp._get_timesteps(20).copy()),
                "ts7": torch.tensor(samp._get_timesteps(7).copy()),
                "linear100_alphas_cumprod": lin.alphas_cumprod.clone()},
               os.path.join(OUT, "schedule.pt"))

    # ---- 4. DDIM loop (5 steps -> 6 evaluations), eta = 0 and eta = 0.5, teacher-forcing records
    cond = torch.randn((1, 4, 4, 8, 8), generator=torch.Generator().manual_seed(13))
    rec_steps = []
    orig = unet.forward

    def spy(z_, t_, c_):
        e = orig(z_, t_, c_)
        rec_steps.append((z_.clone(), int(t_[0]), e.clone()))
        return e
    unet.forward = spy
    torch.manual_seed(42)
    z5 = DDIMSampler(diff, unet).sample((1, 4, 4, 8, 8), cond, 5, "cpu", progress=False)
    steps0 = list(rec_steps)
    rec_steps.clear()
    torch.manual_seed(42)
    z5e = DDIMSampler(diff, unet).sample((1, 4, 4, 8, 8), cond, 5, "cpu", eta=0.5, progress=False)
    unet.forward = orig
    torch.save({"cond": cond, "seed": 42, "z_final": z5, "z_final_eta05": z5e,
                "steps": [{"z": a, "t": b, "eps": e} for a, b, e in steps0]}, os.path.join(OUT, "ddim_tiny.pt"))

    # ---- 5. DDPM ancestral loop on a short schedule (12 training steps)
    diff12 = GaussianDiffusion("cosine", 12)
    torch.manual_seed(43)
    zp = DDPMSampler(diff12, unet).sample((1, 4, 4, 8, 8), cond, "cpu", progress=False)
    torch.save({"cond": cond, "seed": 43, "timesteps": 12, "z_final": zp}, os.path.join(OUT, "ddpm_tiny.pt"))

    # ---- 6. end-to-end generate(), tiny model, 2 thick -> 6 thin slices
    torch.manual_seed(2)
    model = quiet(VideoToVideoDiffusion, TINY).eval()
    v_in = torch.rand((1, 1, 2, 16, 16), generator=torch.Generator().manual_seed(14)) * 2 - 1
    torch.manual_seed(44)
    with torch.no_grad():
        v_out = quiet(model.generate, v_in, "ddim", 4, 1.0, 6)
    torch.save({"seed": 2, "sample_seed": 44, "sd_hash": sd_hash(model.state_dict()), "config": TINY, "v_in": v_in,
                "v_out": v_out, "steps": 4, "target_depth": 6}, os.path.join(OUT, "generate_tiny.pt"))

    # ---- 7. the benchmark model: shipped YAML, resolved the reference's way; hash + counts only
    cfg = yaml.safe_load(open(os.path.join(REF, "config", "slice_interpolation_full_medium.yaml")))
    torch.manual_seed(0)
    full = quiet(VideoToVideoDiffusion, cfg)
    counts = full.count_parameters()
    torch.save({"seed": 0, "sd_hash": sd_hash(full.state_dict()), "n_keys": len(full.state_dict()),
                "counts": counts,
                "resolved": dict(unet_model_channels=full.unet.model_channels, channel_mult=tuple(full.unet.channel_mult),
                                 latent_dim=full.vae.latent_dim, scaling_factor=full.vae.scaling_factor)},
               os.path.join(OUT, "full_model.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
