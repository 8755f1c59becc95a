"""small ncu target: ONE U-Net step (batch 4, eager launches) + ONE VAE decode at the bench shape"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2V_EAGER"] = "1"
import torch
import bench
from v2v_b200.models import VideoToVideoDiffusion
B = 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VideoToVideoDiffusion(bench.load_cfg()).eval().to(dev)
x = torch.randn((B, 8, 48, 48, 48), device=dev); c = torch.randn_like(x); t = torch.full((B,), 500, device=dev)
for _ in range(2):
    e = m.unet(x, t, c)
v = m.vae.decode(x)
torch.cuda.synchronize()
print("ok", float(e.abs().mean()), float(v.abs().mean()))
