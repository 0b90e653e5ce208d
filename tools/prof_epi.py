"""ncu target: HBM-bound 1x1 conv (64->256 at layer1 resolution x3 images), with and without GroupNorm statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as K
dev = torch.device("cuda:0")
x = torch.randn(3, 192, 336, 64, device=dev).to(K.ACT_DTYPE)
w = (torch.randn(256, 1, 1, 64, device=dev) * 0.05).to(K.ACT_DTYPE)
gn = torch.zeros(3, 32, 2, device=dev)
def once():
    K.conv2d_fprop(x, w)
    K.conv2d_fprop(x, w, gn_sum=gn)
for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
