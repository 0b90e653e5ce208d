import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
from tools.bench_kernels import timeit
dev = torch.device("cuda:0")
B = 3
for (H, W, Cin, Cout, ks) in [(192, 336, 64, 256, 1), (192, 336, 256, 64, 1), (192, 336, 64, 64, 3), (96, 168, 128, 512, 1), (48, 84, 256, 256, 3)]:
    x = torch.randn(B, H, W, Cin, device=dev).to(k.ACT_DTYPE)
    w = (torch.randn(Cout, ks, ks, Cin, device=dev) * 0.05).to(k.ACT_DTYPE)
    gn = torch.zeros(B, 32, 2, device=dev)
    t0 = timeit(lambda: k.conv2d_fprop(x, w, stride=1, pad=ks // 2))
    t1 = timeit(lambda: k.conv2d_fprop(x, w, stride=1, pad=ks // 2, gn_sum=gn))
    print((H, W, Cin, Cout, ks), f"plain {t0*1e6:.1f} us, with gn_sum {t1*1e6:.1f} us")
