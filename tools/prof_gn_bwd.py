"""ncu target: GroupNorm backward (reduce + apply) on two layer shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as K
dev = torch.device("cuda:0")
cases = []
for C in (256, 64):
    x = torch.randn(3, 192, 336, C, device=dev).to(K.ACT_DTYPE)
    dy = torch.randn(3, 192, 336, C, device=dev).to(K.ACT_DTYPE)
    cases.append((x, dy, K.gn_stats(x), torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1))
def once():
    for x, dy, s, g, b in cases:
        K.gn_backward(x, s, g, b, dy, mask_mode=1)
for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
