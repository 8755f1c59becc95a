"""CPU suite, part 1: the oracle (oracle/ref_port.py) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), and the mirror modules' initialisation against the reference's state_dict
hashes.  Nothing here reads /root/reference."""
import pytest
import torch

from helpers import TINY_UNET, golden, rel_l2, sd_hash, tiny_unet, tiny_vae
from oracle import ref_port as R

TOL = 2e-5  # fp32 CPU, possibly different thread counts / oneDNN paths than the generating machine


def test_unet_mirror_init_matches_reference_hash():
    g = golden("unet_tiny.pt")
    assert sd_hash(tiny_unet(g["seed"]).state_dict()) == g["sd_hash"]


def test_unet_forward_port_matches_reference():
    g = golden("unet_tiny.pt")
    sd = tiny_unet(g["seed"]).state_dict()
    with torch.no_grad():
        eps = R.unet_forward(sd, g["cfg"], g["x"], g["t"], g["c"])
    assert eps.shape == g["eps"].shape
    assert rel_l2(eps, g["eps"]) < TOL


def test_attention_is_the_degenerate_sum_over_depth():
    """the reference block equals x + proj_out(sum_T V) broadcast over T"""
    sd = tiny_unet(0).state_dict()
    p = "down_blocks.1.0.1"
    x = torch.randn(2, 128, 4, 4, 4, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        lit = R.unet_attention(sd, p, x, heads=2)
        xn = torch.nn.functional.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
        v = torch.nn.functional.conv3d(xn, sd[p + ".qkv.weight"][256:], sd[p + ".qkv.bias"][256:])
        fold = torch.nn.functional.conv3d(v.sum(2, keepdim=True).expand_as(v), sd[p + ".proj_out.weight"],
                                          sd[p + ".proj_out.bias"]) + x
    assert (lit - fold).abs().max().item() < 1e-4


def test_vae_port_matches_reference():
    g = golden("vae_tiny.pt")
    vae = tiny_vae(g["seed"])
    assert sd_hash(vae.state_dict()) == g["sd_hash"]
    sd = vae.state_dict()
    with torch.no_grad():
        z = R.vae_encode(sd, g["v"], g["cfg"]["scaling_factor"])
        rec = R.vae_decode(sd, g["z"], g["cfg"]["scaling_factor"])
    assert rel_l2(z, g["z"]) < TOL and rel_l2(rec, g["recon"]) < TOL


def test_schedule_known_answers():
    g = golden("schedule.pt")
    b = R.diffusion_buffers("cosine", 1000)
    for k in ("alphas_cumprod", "betas", "posterior_log_variance_clipped", "posterior_mean_coef1",
              "posterior_mean_coef2"):
        assert torch.allclose(b[k], g[k], rtol=1e-6, atol=0), k
    assert torch.allclose(R.diffusion_buffers("linear", 100, 1e-4, 0.02)["alphas_cumprod"],
                          g["linear100_alphas_cumprod"], rtol=1e-6)
    ac = b["alphas_cumprod"]
    # SURVEY section 8(c) known answers
    for idx, val in ((0, 0.99990), (20, 0.9977304), (500, 0.4920960), (980, 8.761887e-4), (998, 2.427987e-6)):
        assert abs(ac[idx].item() - val) <= 2e-6 * max(1.0, val / 1e-3) or abs(ac[idx].item() / val - 1) < 1e-4
    assert abs(ac[999].item() / 2.428390e-10 - 1) < 1e-3
    assert (b["betas"] == 0.9999).sum().item() == 1 and (b["betas"] == 0.0001).sum().item() == 13
    assert abs(b["posterior_log_variance_clipped"][0].item() + 46.0517) < 1e-3
    assert abs(1.0 / (torch.sqrt(ac[999] + 1e-8) + 1e-8).item() - 9879.8) < 1.0
    for n, key in ((50, "ts50"), (20, "ts20"), (7, "ts7")):
        ts = R.ddim_timesteps(1000, n)
        assert ts.tolist() == g[key].tolist()
    assert len(R.ddim_timesteps(1000, 50)) == 51 and len(R.ddim_timesteps(1000, 20)) == 21
    with pytest.raises(ValueError):
        R.diffusion_buffers("sigmoid")


def test_ddim_port_matches_reference_trajectory():
    g = golden("ddim_tiny.pt")
    sd = tiny_unet(0).state_dict()
    model = lambda z, t, c: R.unet_forward(sd, TINY_UNET, z, t, c)  # noqa: E731
    buf = R.diffusion_buffers("cosine", 1000)
    with torch.no_grad():
        torch.manual_seed(g["seed"])
        rec = []
        z = R.ddim_sample(model, buf, (1, 4, 4, 8, 8), g["cond"], 5, "cpu", record=rec)
        torch.manual_seed(g["seed"])
        ze = R.ddim_sample(model, buf, (1, 4, 4, 8, 8), g["cond"], 5, "cpu", eta=0.5)
    assert len(rec) == len(g["steps"]) == 6
    assert [r[1] for r in rec] == [s["t"] for s in g["steps"]] == [999, 800, 600, 400, 200, 0]
    assert torch.equal(rec[0][0], g["steps"][0]["z"])  # same initial noise from the same seed
    # teacher-forced: feed the reference's z_t, compare eps (free-running drifts through the +-10 clamp, F9)
    for s in g["steps"]:
        with torch.no_grad():
            e = model(s["z"], torch.tensor([s["t"]]), g["cond"])
        assert rel_l2(e, s["eps"]) < TOL
    assert rel_l2(z, g["z_final"]) < 1e-3 and rel_l2(ze, g["z_final_eta05"]) < 1e-3


def test_ddim_step_formula_against_reference_steps():
    """one update from the reference's recorded (z_t, eps) must land on the reference's next z_t exactly"""
    g = golden("ddim_tiny.pt")
    ac = R.diffusion_buffers("cosine", 1000)["alphas_cumprod"]
    st = g["steps"]
    for i in range(len(st) - 1):
        z_next = R.ddim_step(st[i]["z"], st[i]["eps"], ac[st[i]["t"]], ac[st[i + 1]["t"]])
        assert torch.allclose(z_next, st[i + 1]["z"], rtol=1e-6, atol=1e-6)
    z_last = R.ddim_step(st[-1]["z"], st[-1]["eps"], ac[0], torch.tensor(1.0))
    assert torch.allclose(z_last, g["z_final"], rtol=1e-6, atol=1e-6)


def test_ddpm_port_matches_reference():
    g = golden("ddpm_tiny.pt")
    sd = tiny_unet(0).state_dict()
    model = lambda z, t, c: R.unet_forward(sd, TINY_UNET, z, t, c)  # noqa: E731
    with torch.no_grad():
        torch.manual_seed(g["seed"])
        z = R.ddpm_sample(model, R.diffusion_buffers("cosine", g["timesteps"]), (1, 4, 4, 8, 8), g["cond"], "cpu")
    assert rel_l2(z, g["z_final"]) < 1e-4


def test_generate_port_matches_reference():
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("generate_tiny.pt")
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(g["config"]).eval()
    assert sd_hash(m.state_dict()) == g["sd_hash"]
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        v = R.generate(m.state_dict(), g["config"], g["v_in"], "ddim", g["steps"], target_depth=g["target_depth"])
    assert v.shape == g["v_out"].shape == (1, 1, 6, 16, 16)
    assert rel_l2(v, g["v_out"]) < 1e-3
    with pytest.raises(ValueError):
        R.generate(m.state_dict(), g["config"], g["v_in"], "euler")


@pytest.mark.timeout(300)
def test_benchmark_model_resolution_and_init_hash():
    """the shipped YAML resolves to U-Net defaults and our mirror builds the reference's exact tensors"""
    import yaml
    import os
    from v2v_b200.models import VideoToVideoDiffusion
    g = golden("full_model.pt")
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "golden", "slice_interpolation_full_medium.yaml")))
    vae_cfg, unet_cfg, diff_cfg = R.resolve_config(cfg)
    assert unet_cfg["model_channels"] == 128 and unet_cfg["num_heads"] == 4 and unet_cfg["time_embed_dim"] == 512
    assert vae_cfg == dict(in_channels=1, base_channels=128, latent_dim=8, scaling_factor=1.0)
    torch.manual_seed(g["seed"])
    m = VideoToVideoDiffusion(cfg)
    sd = m.state_dict()
    assert len(sd) == g["n_keys"] == 458
    assert m.count_parameters() == g["counts"]
    assert g["counts"]["unet"] == 264658184 and g["counts"]["vae"] == 90301593
    assert sd_hash(sd) == g["sd_hash"]


def test_philox_restatement_against_random123_known_answers():
    """the device generator of b2v_ddpm_sample (noise == NULL) is Philox4x32-10; the numpy restatement the GPU test
    compares against is pinned here with Random123's published known-answer vectors (kat_vectors: philox4x32 10)"""
    import numpy as np
    from oracle import philox as P
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        got = P.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == list(out)
    x = P.normal(400_000, 1234, 7)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3 and np.isfinite(x).all()
    assert not np.array_equal(P.normal(64, 1234, 7), P.normal(64, 1234, 8))  # the step is part of the counter
