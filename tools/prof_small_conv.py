"""ncu target: latency-bound small convolutions (3x3 256->256 at 12x21 and 48x84, batch 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
dev = torch.device("cuda:0")
w = (torch.randn(256, 3, 3, 256, device=dev) * 0.05).to(k.ACT_DTYPE)
xs = [torch.randn(1, 12, 21, 256, device=dev).to(k.ACT_DTYPE), torch.randn(1, 48, 84, 256, device=dev).to(k.ACT_DTYPE)]
for _ in range(3):
    for x in xs:
        k.conv2d_fprop(x, w, stride=1, pad=1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for x in xs:
    k.conv2d_fprop(x, w, stride=1, pad=1)
    k.conv2d_fprop(x, w, stride=1, pad=1, bn_hint=256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
