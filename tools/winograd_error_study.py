"""CPU study (fp64 emulation): per-layer error of an fp16-operand in-plane Winograd F(2x2,3x3) convolution against the
direct fp16-operand convolution the kernels run today.  python tools/winograd_error_study.py"""
import torch, torch.nn.functional as F
torch.manual_seed(0)
def h(x): return x.half().double()
BT=torch.tensor([[1,0,-1,0],[0,1,1,0],[0,-1,1,0],[0,1,0,-1]],dtype=torch.float64)
G=torch.tensor([[1,0,0],[.5,.5,.5],[.5,-.5,.5],[0,0,1]],dtype=torch.float64)
AT=torch.tensor([[1,1,1,0],[0,1,-1,-1]],dtype=torch.float64)
def wino2d_inplane(x,w):
    # x: (N,C,D,H,W) fp64 (already fp16-representable), w: (O,C,3,3,3) fp64 exact weights
    N,C,D,H,W=x.shape; O=w.shape[0]
    xp=F.pad(x,(1,1,1,1,1,1))
    U=torch.einsum('ij,ocdjk,lk->ocdil',G,w,G)            # (O,C,3,4,4) fp32-ish
    U=h(U.float())                                          # operand rounding
    out=torch.zeros(N,O,D,H,W,dtype=torch.float64)
    for th in range(0,H,2):
        for tw in range(0,W,2):
            d=xp[:,:,:,th:th+4,tw:tw+4]                    # (N,C,D+2,4,4)
            V=torch.einsum('ij,ncdjk,lk->ncdil',BT,d,BT)
            V=h(V.float())
            M=torch.zeros(N,O,D,4,4,dtype=torch.float64)
            for kd in range(3):
                M+=torch.einsum('ocil,ncdil->nodil',U[:,:,kd],V[:,:,kd:kd+D]).float().double()
            Y=torch.einsum('ij,nodjk,lk->nodil',AT,M,AT)
            out[:,:,:,th:th+2,tw:tw+2]=Y
    return out
for C,O in [(128,128),(256,256)]:
    N,D,H,W=1,3,8,8
    # realistic activation: silu(unit normal) + small temb
    x=F.silu(torch.randn(N,C,D,H,W,dtype=torch.float64))+0.1*torch.randn(1,C,1,1,1,dtype=torch.float64)
    w=(torch.rand(O,C,3,3,3,dtype=torch.float64)*2-1)/ (C*27)**0.5   # kaiming-uniform-like default init
    ref=F.conv3d(x,w,padding=1)
    direct=F.conv3d(h(x),h(w),padding=1)
    wi=wino2d_inplane(h(x),w)
    rl=lambda a,b:( (a-b).norm()/b.norm()).item()
    print(C,O,'direct fp16-operand rel-L2 %.3e'%rl(direct,ref),' winograd F(2x2,3x3) fp16-operand rel-L2 %.3e'%rl(wi,ref))
