"""small end-to-end target for compute-sanitizer: every kernel of the library at tiny shapes"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["B2V_EAGER"] = "1"
import torch
from helpers import golden
from v2v_b200 import ops
from v2v_b200.models import VideoToVideoDiffusion
from v2v_b200.utils import calculate_video_metrics
from v2v_b200.utils.inputs import extract_thick_patch
from v2v_b200.inference.volume import generate_volume
dev = torch.device("cuda:0")
g = golden("generate_tiny.pt")
torch.manual_seed(g["seed"])
m = VideoToVideoDiffusion(g["config"]).eval().to(dev)
torch.manual_seed(1)
out = m.generate(g["v_in"].to(dev), "ddim", 2, target_depth=6)
out2 = m.generate(g["v_in"].to(dev), "ddpm" if False else "ddim", 1, target_depth=6)
vol = (torch.rand((1, 1, 2, 32, 32), device=dev) * 2 - 1)
full = generate_volume(m, vol, "ddim", 1, patch_size=(2, 16, 16), target_patch_size=(6, 16, 16), stride=(2, 16, 16), batch=2)
print(calculate_video_metrics((out + 1) / 2, (out2 + 1) / 2)["psnr"])
print(extract_thick_patch(torch.randn((1, 10, 32, 32), device=dev), 0, 12, 60, 4, 4, 4, (16, 16)).shape)
gen = torch.Generator().manual_seed(3)
for (kind, cin0, cin1, cout, N, D, H, W) in [(0, 128, 0, 128, 2, 2, 8, 8), (0, 64, 64, 256, 1, 2, 6, 6), (2, 64, 0, 128, 1, 2, 8, 8),
                                              (3, 128, 0, 128, 1, 2, 4, 4), (0, 512, 0, 512, 1, 3, 6, 6), (0, 128, 0, 1, 1, 2, 8, 8)]:
    cin = cin0 + cin1
    shape = {0: (cout, cin, 3, 3, 3), 2: (cout, cin, 3, 4, 4), 3: (cin, cout, 3, 4, 4)}[kind]
    conv = ops.Conv(kind, torch.randn(shape, generator=gen) / 40, torch.zeros(cout), cin0, cin1, cout)
    x = torch.randn((N, D, H, W, cin), device=dev, dtype=torch.float16)
    x0, x1 = x[..., :cin0].contiguous(), (x[..., cin0:].contiguous() if cin1 else None)
    o, st = conv(x0, x1, out_fp32=(cout <= 16), groups=0 if cout <= 16 else 8)
torch.cuda.synchronize()
print("sanitize target ok", float(out.abs().mean()), float(full.abs().mean()))
