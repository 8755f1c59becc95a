import hashlib
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


TINY_UNET = dict(latent_dim=4, model_channels=64, num_res_blocks=1, attention_levels=[1], channel_mult=(1, 2),
                 num_heads=2, time_embed_dim=128)


def tiny_unet(seed=0):
    from v2v_b200.models import UNet3D
    torch.manual_seed(seed)
    return UNet3D(**TINY_UNET).eval()


def tiny_vae(seed=1):
    from v2v_b200.models import VideoVAE
    torch.manual_seed(seed)
    return VideoVAE(in_channels=1, latent_dim=4, base_channels=64, scaling_factor=0.5).eval()
