"""Tile-width sweep of the RoI-head contractions (graph-replayed, warm): fc6 / fc7 at inference (1000 RoIs) and
training (1536 RoIs), the one-RoI mask-head convolution."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eosvos_b200 import kernels as k
dev = torch.device("cuda:0")
CASES = [("fc6 inf", 1000, 1, 1, 12544, 1024, 1), ("fc6 train", 1536, 1, 1, 12544, 1024, 1), ("fc7 inf", 1000, 1, 1, 1024, 1024, 1),
         ("mask conv 1 roi", 1, 14, 14, 256, 256, 3), ("mask conv 32 roi", 32, 14, 14, 256, 256, 3)]
for name, N, H, W, Ci, Co, ks in CASES:
    x = torch.randn(N, H, W, Ci, device=dev).to(k.ACT_DTYPE)
    w = (torch.randn(Co, ks, ks, Ci, device=dev) * 0.02).to(k.ACT_DTYPE)
    row = []
    for hint in (0, 64, 128, 256):
        kw = dict(stride=1, pad=ks // 2)
        if hint:
            kw["bn_hint"] = hint
        for _ in range(3):
            k.conv2d_fprop(x, w, **kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                k.conv2d_fprop(x, w, **kw)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e3 / 20)
    fl = 2.0 * N * H * W * Co * Ci * ks * ks
    print(f"{name:18s} auto {row[0]:7.1f} us | bn64 {row[1]:7.1f} | bn128 {row[2]:7.1f} | bn256 {row[3]:7.1f}   "
          f"({fl / min(row) * 1e-6:.0f} TF/s best)", flush=True)
