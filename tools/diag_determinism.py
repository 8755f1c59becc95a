import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import tiny_vae, tiny_unet, rel_l2
from v2v_b200 import ops
dev = torch.device("cuda:0")
print("EAGER" if os.environ.get("B2V_EAGER") else "GRAPH")
g = torch.Generator().manual_seed(1)
# op level: same conv twice, bitwise
for (kind, cin, cout, D, H, W) in [(0, 64, 64, 3, 16, 16), (0, 128, 128, 4, 12, 12), (2, 64, 128, 3, 16, 16), (3, 128, 64, 3, 4, 4), (0, 256, 8, 3, 4, 4)]:
    x = torch.randn((1, cin, D, H, W), generator=g).to(dev)
    shape = {0: (cout, cin, 3, 3, 3), 2: (cout, cin, 3, 4, 4), 3: (cin, cout, 3, 4, 4)}[kind]
    w = torch.randn(shape, generator=g) / (cin * 27) ** 0.5
    conv = ops.Conv(kind, w, torch.zeros(cout), cin, 0, cout)
    x16 = ops.to_cl16(x)
    outs = [conv(x16, out_fp32=(cout <= 16), groups=0 if cout <= 16 else 8) for _ in range(3)]
    torch.cuda.synchronize()
    print("conv", kind, cin, cout, "bitwise-equal:", torch.equal(outs[0][0], outs[1][0]), torch.equal(outs[1][0], outs[2][0]),
          "stats maxdiff", 0 if outs[0][1] is None else (outs[0][1] - outs[1][1]).abs().max().item())
    # gn apply twice
    if cout > 16:
        gamma = torch.ones(cout, device=dev); beta = torch.zeros(cout, device=dev)
        a = ops.gn_apply(outs[0][0], outs[0][1], gamma, beta, 8)[0]; b = ops.gn_apply(outs[0][0], outs[0][1], gamma, beta, 8)[0]
        print("  gn_apply bitwise-equal:", torch.equal(a, b))
vae = tiny_vae(1).to(dev)
v = (torch.rand((1, 1, 3, 16, 16), generator=g) * 2 - 1).to(dev)
zs = [vae.encode(v).clone() for _ in range(4)]
print("vae.encode run-to-run:", [f"{rel_l2(zs[i], zs[0]):.2e}" for i in range(1, 4)])
z = zs[0]
rs = [vae.decode(z).clone() for _ in range(4)]
print("vae.decode run-to-run:", [f"{rel_l2(rs[i], rs[0]):.2e}" for i in range(1, 4)])
m = tiny_unet(0).to(dev)
x = torch.randn((2, 4, 4, 8, 8), generator=g).to(dev); c = torch.randn((2, 4, 4, 8, 8), generator=g).to(dev); t = torch.tensor([500, 37], device=dev)
es = [m(x, t, c).clone() for _ in range(4)]
print("unet run-to-run:", [f"{rel_l2(es[i], es[0]):.2e}" for i in range(1, 4)])
