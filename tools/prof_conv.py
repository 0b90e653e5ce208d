import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
dev = torch.device("cuda:0")
B = 3
x = torch.randn(B, 192, 336, 256, device=dev).to(k.ACT_DTYPE)
w = (torch.randn(256, 3, 3, 256, device=dev) * 0.05).to(k.ACT_DTYPE)
dy = torch.randn(B, 192, 336, 256, device=dev).to(k.ACT_DTYPE)
dw = torch.zeros(256, 256, 3, 3, device=dev)
for _ in range(3):
    k.conv2d_fprop(x, w, stride=1, pad=1, bn_hint=256)
    k.conv2d_wgrad(x, dy, (3, 3), stride=1, pad=1, bn_hint=256, out=dw)
# an HBM-bound 1x1 (64->256 at 192x336) and GN apply
x1 = torch.randn(B, 192, 336, 64, device=dev).to(k.ACT_DTYPE)
w1 = (torch.randn(256, 1, 1, 64, device=dev) * 0.05).to(k.ACT_DTYPE)
for _ in range(3):
    y1 = k.conv2d_fprop(x1, w1, bn_hint=128)
torch.cuda.synchronize()
