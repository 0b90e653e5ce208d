"""GPU diagnostic for the tcgen05 contraction kernels: prints error structure, not just pass/fail."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from eosvos_b200 import kernels as k

dev = torch.device("cuda:0")
torch.manual_seed(0)

def bf(x): return x.to(k.ACT_DTYPE)

def report(name, got, ref):
    d = (got.double() - ref.double())
    rel = (d.norm() / (ref.double().norm() + 1e-30)).item()
    print(f"{name}: rel={rel:.3e} max|d|={d.abs().max().item():.3e} ref_norm={ref.norm().item():.3e} "
          f"got_norm={got.norm().item():.3e} nan={torch.isnan(got).sum().item()}", flush=True)
    return rel

def conv_case(N, H, W, Cin, Cout, ks, s, p, bn=0):
    x = bf(torch.randn(N, Cin, H, W)).float(); w = bf(torch.randn(Cout, Cin, ks, ks) / math.sqrt(Cin*ks*ks)).float()
    ref = F.conv2d(x, w, None, s, p)
    try:
        y = k.conv2d_fprop(bf(x.permute(0,2,3,1).contiguous()).to(dev), bf(w.permute(0,2,3,1).contiguous()).to(dev), stride=s, pad=p, bn_hint=bn)
        torch.cuda.synchronize()
    except Exception as e:
        print("EXC", (N,H,W,Cin,Cout,ks,s,p,bn), e, flush=True); return
    got = y.float().cpu().permute(0,3,1,2)
    rel = report(f"fprop {(N,H,W,Cin,Cout,ks,s,p,bn)}", got, ref)
    if rel > 1e-2:
        d = (got - ref).abs()
        print("   err by out-channel block of 32:", [round(d[:, c:c+32].mean().item(), 4) for c in range(0, Cout, 32)][:8])
        print("   err by row:", [round(d[0, :, r].mean().item(), 4) for r in range(min(H, 12))])
        print("   err by col:", [round(d[0, :, :, c].mean().item(), 4) for c in range(min(W, 12))])
        # linear fit got ~ a*ref
        a = (got*ref).sum() / (ref*ref).sum()
        print("   scale got/ref:", a.item())

if __name__ == "__main__":
    conv_case(1, 16, 8, 64, 64, 1, 1, 0)
    conv_case(1, 16, 8, 64, 64, 1, 1, 0, 128)
    conv_case(1, 16, 16, 128, 128, 1, 1, 0)
    conv_case(1, 16, 16, 64, 64, 3, 1, 1)
    conv_case(1, 24, 42, 256, 256, 3, 1, 1)
    conv_case(1, 24, 42, 128, 128, 3, 2, 1)
    conv_case(1, 48, 84, 64, 256, 3, 1, 1, 256)
    # wgrad
    for (N,H,W,Cin,Cout,ks,s,p) in [(1,16,8,64,64,1,1,0),(1,16,16,64,128,3,1,1),(1,24,42,128,128,3,2,1),(2,12,21,256,256,3,1,1)]:
        x = bf(torch.randn(N,Cin,H,W)).float(); w = (bf(torch.randn(Cout,Cin,ks,ks))/8).float().requires_grad_(True)
        y = F.conv2d(x, w, None, s, p); dy = bf(torch.randn_like(y)).float()
        (ref,) = torch.autograd.grad(y, w, dy)
        try:
            dw = k.conv2d_wgrad(bf(x.permute(0,2,3,1).contiguous()).to(dev), bf(dy.permute(0,2,3,1).contiguous()).to(dev), (ks,ks), stride=s, pad=p)
            torch.cuda.synchronize()
            rel = report(f"wgrad {(N,H,W,Cin,Cout,ks,s,p)}", dw.cpu(), ref)
            if rel > 1e-2:
                d = (dw.cpu()-ref).abs()
                print("   err by tap:", d.mean((0,1)).flatten().tolist())
                print("   err by co blk:", [round(d[c:c+32].mean().item(),4) for c in range(0,Cout,32)])
                print("   err by ci blk:", [round(d[:,c:c+32].mean().item(),4) for c in range(0,Cin,32)])
        except Exception as e:
            print("EXC wgrad", e, flush=True)
