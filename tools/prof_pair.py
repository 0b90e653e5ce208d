"""ncu target: 3x3 256->256 conv at P2 x3 images, single-CTA (bn 256) and CTA-pair (bn 512) kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as K
dev = torch.device("cuda:0")
x = torch.randn(3, 192, 336, 256, device=dev).to(K.ACT_DTYPE)
w = (torch.randn(256, 3, 3, 256, device=dev) * 0.02).to(K.ACT_DTYPE)
for bn in (256, 512):
    for _ in range(3):
        K.conv2d_fprop(x, w, stride=1, pad=1, bn_hint=bn)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for bn in (256, 512):
    K.conv2d_fprop(x, w, stride=1, pad=1, bn_hint=bn)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
