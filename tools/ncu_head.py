import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from v2v_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cout = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randn((1, 128, 48, 96, 96), generator=g).to(dev)
w = torch.randn((cout, 128, 3, 3, 3), generator=g) / 60.0
conv = ops.Conv(0, w, torch.zeros(cout), 128, 0, cout)
x16 = ops.to_cl16(x)
for _ in range(3):
    out, _ = conv(x16, out_fp32=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out, _ = conv(x16, out_fp32=True)
e1.record(); torch.cuda.synchronize()
print("head conv cout", cout, "ms", e0.elapsed_time(e1) / 5)
