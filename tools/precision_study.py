"""per-step U-Net parity of the current operand mode (B2V_OPERANDS=bf16 for the BF16-operand emulation) at the
BASELINE patch shape, against the fp32 oracle on the same GPU"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import ref_port as R
from v2v_b200.models import VideoToVideoDiffusion
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
cfg = bench.load_cfg()
torch.manual_seed(0)
m = VideoToVideoDiffusion(cfg).eval().to(dev)
sd = {k: v.to(dev) for k, v in m.state_dict().items()}
_, unet_cfg, _ = R.resolve_config(cfg)
usd = {k[5:]: w for k, w in sd.items() if k.startswith("unet.")}
vsd = {k[4:]: w for k, w in sd.items() if k.startswith("vae.")}
g = torch.Generator().manual_seed(5)
x = torch.randn((1, 8, 48, 48, 48), generator=g).to(dev); c = torch.randn((1, 8, 48, 48, 48), generator=g).to(dev)
mode = os.environ.get("B2V_OPERANDS", "fp16")
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
for tv in (999, 500, 0):
    t = torch.tensor([tv], device=dev)
    with torch.no_grad():
        ref = R.unet_forward(usd, unet_cfg, x, t, c)
    print(f"operands={mode}: U-Net step t={tv}: rel-L2 = {rel(m.unet(x, t, c), ref):.3e}")
z = torch.randn((1, 8, 48, 48, 48), generator=g).to(dev)
with torch.no_grad():
    ref = R.vae_decode(vsd, z, 1.0)
got = m.vae.decode(z)
n = lambda v: (v.clamp(-1, 1) + 1) / 2
print(f"operands={mode}: VAE decode rel-L2 = {rel(got, ref):.3e}  PSNR(new,ref) = {R.psnr(n(got), n(ref)):.1f} dB")
