"""per-op CUDA-event profile of one U-Net step and one VAE decode/encode at the bench shape -> table + JSON"""
import sys, os, json, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from v2v_b200 import _lib
from v2v_b200.models import VideoToVideoDiffusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "ops_profile.json")
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VideoToVideoDiffusion(bench.load_cfg()).eval().to(dev)
x = torch.randn((B, 8, 48, 48, 48), device=dev); c = torch.randn_like(x); t = torch.full((B,), 500, device=dev)
m.unet(x, t, c); v = m.vae.decode(x); m.vae.encode(torch.rand((B, 1, 8, 192, 192), device=dev) * 2 - 1)
torch.cuda.synchronize()
L = _lib.lib(); buf = ctypes.create_string_buffer(1 << 20); res = {}
_lib.check(L.b2v_unet_profile(m.unet.native(dev), 5, buf, len(buf), _lib.stream()), "p"); res["unet"] = json.loads(buf.value.decode())
_lib.check(L.b2v_vae_profile(m.vae.native(dev), 1, 3, buf, len(buf), _lib.stream()), "p"); res["vae_decode"] = json.loads(buf.value.decode())
_lib.check(L.b2v_vae_profile(m.vae.native(dev), 0, 3, buf, len(buf), _lib.stream()), "p"); res["vae_encode"] = json.loads(buf.value.decode())
json.dump(res, open(out, "w"))
for part, ops in res.items():
    tot = sum(o["ms"] for o in ops)
    print(f"== {part}: {tot:.3f} ms, {len(ops)} ops, batch {B}")
    for o in ops:
        tf = o["flops"] / o["ms"] / 1e9 if o["flops"] else 0
        gb = o["bytes"] / o["ms"] / 1e6 if o["bytes"] else 0
        print(f"  {o['name']:34s} {o['ms']:8.4f} ms {100*o['ms']/tot:5.1f}%  {tf:7.1f} TF/s  {gb:7.1f} GB/s")
