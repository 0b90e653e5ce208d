import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
from tools.bench_kernels import timeit
dev = torch.device("cuda:0")
B = 3
shapes = [(192, 336, 64, 256, 1), (192, 336, 256, 64, 1), (192, 336, 64, 64, 3), (96, 168, 128, 512, 1), (96, 168, 512, 128, 1),
          (96, 168, 128, 128, 3), (48, 84, 256, 1024, 1), (48, 84, 1024, 256, 1), (48, 84, 256, 256, 3), (24, 42, 512, 2048, 1),
          (24, 42, 2048, 512, 1), (24, 42, 512, 512, 3), (96, 168, 256, 512, 1), (48, 84, 512, 1024, 1), (24, 42, 1024, 2048, 1)]
for (H, W, Cin, Cout, ks) in shapes:
    x = torch.randn(B, H, W, Cin, device=dev).to(k.ACT_DTYPE)
    w = (torch.randn(Cout, ks, ks, Cin, device=dev) * 0.05).to(k.ACT_DTYPE)
    gn = torch.zeros(B, 32, 2, device=dev)
    t0 = timeit(lambda: k.conv2d_fprop(x, w, stride=1, pad=ks // 2))
    t1 = timeit(lambda: k.conv2d_fprop(x, w, stride=1, pad=ks // 2, gn_sum=gn))
    y = k.conv2d_fprop(x, w, stride=1, pad=ks // 2)
    t2 = timeit(lambda: k.gn_stats(y))
    print((H, W, Cin, Cout, ks), f"K={Cin*ks*ks}: plain {t0*1e6:.1f} us, with gn_sum {t1*1e6:.1f} us, separate stats {t2*1e6:.1f} us "
          f"-> fused {'wins' if t1 < t0 + t2 else 'LOSES'} by {abs(t0 + t2 - t1)*1e6:.1f} us")
