"""Per-kernel throughput on the reference's layer shapes. CUDA-event timed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k

dev = torch.device("cuda:0")

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

def main():
    B = int(os.environ.get("B", "1"))
    shapes = [  # name, H, W, Cin, Cout, k, stride
        ("fpn_out_p2 3x3 256->256", 192, 336, 256, 256, 3, 1),
        ("fpn_out_p3 3x3 256->256", 96, 168, 256, 256, 3, 1),
        ("l1 3x3 64->64", 192, 336, 64, 64, 3, 1),
        ("l1 1x1 256->64", 192, 336, 256, 64, 1, 1),
        ("l1 1x1 64->256", 192, 336, 64, 256, 1, 1),
        ("l2 3x3 128->128", 96, 168, 128, 128, 3, 1),
        ("l2 1x1 128->512", 96, 168, 128, 512, 1, 1),
        ("l3 3x3 256->256", 48, 84, 256, 256, 3, 1),
        ("l3 1x1 1024->256", 48, 84, 1024, 256, 1, 1),
        ("l4 3x3 512->512", 24, 42, 512, 512, 3, 1),
        ("l4 1x1 512->2048", 24, 42, 512, 2048, 1, 1),
        ("l2.0 3x3/2 128->128", 192, 336, 128, 128, 3, 2),
        ("fpn_lat_c2 1x1 256->256", 192, 336, 256, 256, 1, 1),
    ]
    res = []
    for name, H, W, Cin, Cout, ks, s in shapes:
        pad = ks // 2
        x = torch.randn(B, H, W, Cin, device=dev).to(k.ACT_DTYPE)
        w = (torch.randn(Cout, ks, ks, Cin, device=dev) * 0.05).to(k.ACT_DTYPE)
        wt = (torch.randn(Cin, ks, ks, Cout, device=dev) * 0.05).to(k.ACT_DTYPE)
        Ho, Wo = (H + 2 * pad - ks) // s + 1, (W + 2 * pad - ks) // s + 1
        dy = torch.randn(B, Ho, Wo, Cout, device=dev).to(k.ACT_DTYPE)
        flops = 2.0 * B * Ho * Wo * Cout * Cin * ks * ks
        line = {"name": name, "gflop": flops / 1e9}
        for bn in (64, 128, 256, 512):
            if Cout % min(bn, 256): continue
            if bn == 512 and (B * Ho * Wo) % 128: continue
            try:
                t = timeit(lambda: k.conv2d_fprop(x, w, stride=s, pad=pad, bn_hint=bn))
            except Exception as e:
                line[f"fprop_bn{bn}_tflops"] = str(e)[-60:]
                continue
            line[f"fprop_bn{bn}_tflops"] = round(flops / t / 1e12, 1)
        for bn in (64, 128, 256):
            if Cin % bn: continue
            t = timeit(lambda: k.conv2d_dgrad(dy, wt, (H, W), stride=s, pad=pad, bn_hint=bn))
            line[f"dgrad_bn{bn}_tflops"] = round(flops / t / 1e12, 1)
        dw = torch.zeros(Cout, Cin, ks, ks, device=dev)
        for bn in (64, 128, 256):
            if Cin % bn: continue
            t = timeit(lambda: k.conv2d_wgrad(x, dy, (ks, ks), stride=s, pad=pad, bn_hint=bn, out=dw))
            line[f"wgrad_bn{bn}_tflops"] = round(flops / t / 1e12, 1)
        # GroupNorm at this output shape
        if Cout >= 64:
            y = torch.randn(B, Ho, Wo, Cout, device=dev).to(k.ACT_DTYPE)
            gam = torch.ones(Cout, device=dev); bet = torch.zeros(Cout, device=dev)
            t1 = timeit(lambda: k.gn_stats(y)); sums = k.gn_stats(y)
            t2 = timeit(lambda: k.gn_apply(y, sums, gam, bet, relu=True))
            nbytes = y.numel() * 2
            line["gn_stats_GBs"] = round(nbytes / t1 / 1e9); line["gn_apply_GBs"] = round(2 * nbytes / t2 / 1e9)
        print(json.dumps(line), flush=True)
        res.append(line)
    # meta update at full size
    n = 43_975_515
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev); lr = torch.rand(n // 49 + 1, device=dev)
    ps = [p[: (n // 49) * 49].view(-1, 49)]; gs = [g[: (n // 49) * 49].view(-1, 49)]; ls = [lr[: n // 49].view(-1, 1)]
    os_ = [torch.empty_like(ps[0])]
    plan = k.MetaUpdatePlan(ps, gs, ls, os_)
    t = timeit(lambda: k.meta_update(plan))
    print(json.dumps({"name": "meta_update flat", "GBs": round(12 * ps[0].numel() / t / 1e9), "us": round(t * 1e6, 1)}))

if __name__ == "__main__":
    main()
