import os, sys
os.environ["EOSVOS_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from eosvos_b200 import kernels as K
from eosvos_b200.util import evaluate as E, synthetic
dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
frames, labels = synthetic.make_video(5, 2, 480, 854, 1)
fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
gt0 = torch.from_numpy((labels[0] == 1).astype(np.float32))[None, None]
inp, gts = fr[0:1].to(dev).repeat(3, 1, 1, 1), gt0.to(dev).repeat(3, 1, 1, 1)
E.finetune(model, opt, lambda e: (inp, gts), 30, 1, 0)
rec = []
orig = K.gn_apply
def spy(x, sums, gamma, beta, res=None, relu=False, eps=1e-5):
    N, C = x.shape[0], x.shape[-1]; HW = x.numel() // (N * C); m = (C // 32) * HW
    mean = sums[..., 0] / m; var = (sums[..., 1] / m - mean * mean).clamp(min=0)
    rec.append((C, HW, (mean.abs() / (var + eps).sqrt()).max().item(), x.float().abs().max().item()))
    return orig(x, sums, gamma, beta, res, relu, eps)
K.gn_apply = spy
import eosvos_b200.ops as ops
ops.K.gn_apply = spy
model.eval()
with torch.no_grad(): model(fr[1:2].to(dev), gt0.to(dev))
rec.sort(key=lambda r: -r[2])
print("top |mean|/std ratios (C, HW, ratio, max|z|):", [(c, hw, round(r, 1), round(mx, 1)) for c, hw, r, mx in rec[:8]])
print("max |z| overall:", max(r[3] for r in rec))
