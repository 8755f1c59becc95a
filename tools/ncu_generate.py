"""ncu target for the launch list: ONE generate() at the bench shape (batch 4) with DDIM-2 = 3 of the 51 U-Net
evaluations -- VAE encode, depth upsample, 3 sampler steps, VAE decode, all guards -- eager launches (B2V_EAGER) so every
kernel is its own ncu record.  Shares of the full DDIM-50 job follow by weighting the U-Net steps 51 / 3."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2V_EAGER"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from v2v_b200 import _lib  # noqa: E402
from v2v_b200.models import VideoToVideoDiffusion  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VideoToVideoDiffusion(bench.load_cfg()).eval().to(dev)
v = bench.synthetic_input(bench.BATCH).to(dev)
torch.manual_seed(42)
n0 = _lib.launch_count()
out = m.generate(v, "ddim", 2, target_depth=bench.T_OUT)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()), "launches", _lib.launch_count() - n0)
