"""find the first op of a program whose output is not bitwise reproducible run to run"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import tiny_vae, tiny_unet
from v2v_b200 import _lib
dev = torch.device("cuda:0")
L = _lib.lib()
g = torch.Generator().manual_seed(1)

def bisect(handle, prog, label):
    buf = [torch.zeros(64 << 20, dtype=torch.uint8, device=dev) for _ in range(3)]
    i = 0
    while True:
        name = L.b2v_debug_op_name(handle, prog, i).decode()
        if not name:
            break
        ns = []
        for r in range(3):
            n = L.b2v_debug_op_output(handle, prog, i + 1, i, ctypes.c_void_p(buf[r].data_ptr()), buf[r].numel(), _lib.stream())
            ns.append(n)
        n = ns[0]
        if n > 0:
            a, b, c = (x[:n].view(torch.float16) if 'conv_out' not in name and 'tanh' not in name else x[:n].view(torch.float32) for x in buf)
            same = torch.equal(buf[0][:n], buf[1][:n]) and torch.equal(buf[1][:n], buf[2][:n])
            nd = (buf[0][:n] != buf[1][:n]).sum().item()
            rel = ((a.float() - b.float()).norm() / (a.float().norm() + 1e-30)).item()
            print(f"{label} op {i:3d} {name:32s} bytes={n:9d} reproducible={same} differing_bytes={nd} rel={rel:.2e} finite={torch.isfinite(a.float()).all().item()}")
        i += 1

vae = tiny_vae(1).to(dev)
v = (torch.rand((1, 1, 3, 16, 16), generator=g) * 2 - 1).to(dev)
vae.encode(v); torch.cuda.synchronize()
bisect(vae.native(dev), 1, "enc")
