import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
dev = torch.device("cuda:0")
x = torch.randn(3, 192, 336, 64, device=dev).to(k.ACT_DTYPE)
w = (torch.randn(256, 1, 1, 64, device=dev) * 0.05).to(k.ACT_DTYPE)
gn = torch.zeros(3, 32, 2, device=dev)
for _ in range(3): k.conv2d_fprop(x, w, gn_sum=gn)
torch.cuda.synchronize()
