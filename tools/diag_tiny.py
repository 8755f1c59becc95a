"""diagnostic: tiny-config parity at several sizes (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import TINY_UNET, tiny_unet, tiny_vae, rel_l2
from oracle import ref_port as R
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
vae = tiny_vae(1).to(dev)
sd = {k: v.to(dev) for k, v in vae.state_dict().items()}
for (T, H) in [(3, 16), (4, 32), (8, 64), (2, 96)]:
    g = torch.Generator().manual_seed(T * H)
    v = (torch.rand((1, 1, T, H, H), generator=g) * 2 - 1).to(dev)
    z = vae.encode(v)
    with torch.no_grad():
        zr = R.vae_encode(sd, v, 0.5)
        rr = R.vae_decode(sd, zr, 0.5)
    rec = vae.decode(zr)
    print(f"vae T={T} H={H}: enc rel={rel_l2(z, zr):.3e} dec rel={rel_l2(rec, rr):.3e} |z|={zr.abs().max().item():.3f}")
from v2v_b200.models import VideoVAE
for base, L in [(64, 8), (128, 4), (128, 8)]:
    torch.manual_seed(1)
    v2 = VideoVAE(1, L, base, 1.0).eval().to(dev)
    sd2 = {k: v.to(dev) for k, v in v2.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    v = (torch.rand((1, 1, 4, 32, 32), generator=g) * 2 - 1).to(dev)
    with torch.no_grad():
        zr = R.vae_encode(sd2, v, 1.0); rr = R.vae_decode(sd2, zr, 1.0)
    print(f"vae base={base} L={L}: enc rel={rel_l2(v2.encode(v), zr):.3e} dec rel={rel_l2(v2.decode(zr), rr):.3e}")
m = tiny_unet(0).to(dev)
sdu = {k: v.to(dev) for k, v in m.state_dict().items()}
for (T, h) in [(4, 8), (8, 16), (6, 24)]:
    g = torch.Generator().manual_seed(T + h)
    x = torch.randn((2, 4, T, h, h), generator=g).to(dev); c = torch.randn((2, 4, T, h, h), generator=g).to(dev)
    t = torch.tensor([500, 37], device=dev)
    with torch.no_grad():
        ref = R.unet_forward(sdu, TINY_UNET, x, t, c)
    print(f"unet T={T} h={h}: rel={rel_l2(m(x, t, c), ref):.3e}")
from v2v_b200.models import UNet3D
for kw in [dict(latent_dim=8, model_channels=64, num_res_blocks=1, attention_levels=[], channel_mult=(1, 2), num_heads=2, time_embed_dim=128),
           dict(latent_dim=4, model_channels=128, num_res_blocks=1, attention_levels=[1], channel_mult=(1, 2), num_heads=2, time_embed_dim=128),
           dict(latent_dim=4, model_channels=64, num_res_blocks=1, attention_levels=[0], channel_mult=(1,), num_heads=2, time_embed_dim=128)]:
    torch.manual_seed(0)
    mm = UNet3D(**kw).eval().to(dev)
    sdm = {k: v.to(dev) for k, v in mm.state_dict().items()}
    L = kw["latent_dim"]
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1, L, 4, 8, 8), generator=g).to(dev); c = torch.randn((1, L, 4, 8, 8), generator=g).to(dev)
    t = torch.tensor([500], device=dev)
    with torch.no_grad():
        ref = R.unet_forward(sdm, kw, x, t, c)
    print(f"unet {kw['model_channels']}ch L={L} att={kw['attention_levels']} mult={kw['channel_mult']}: rel={rel_l2(mm(x, t, c), ref):.3e}")
