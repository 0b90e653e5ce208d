"""ncu targets (cudaProfilerStart/Stop window, use --profile-from-start off): the roofline conv on CTA pairs and on a
single CTA, wgrad of the same layer (channels-last gradient, vector REDs), GroupNorm backward on the largest
activation, and the fused MetaOptimizer update on the real 201-tensor set with channels-last filter gradients."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eosvos_b200 import kernels as k
dev = torch.device("cuda:0")
B = 3
x = torch.randn(B, 192, 336, 256, device=dev).to(k.ACT_DTYPE)
w = (torch.randn(256, 3, 3, 256, device=dev) * 0.05).to(k.ACT_DTYPE)
dy = torch.randn(B, 192, 336, 256, device=dev).to(k.ACT_DTYPE)
model, opt = bench.build_model(dev)
params = [p.detach() for *_, p in opt.meta_model.param_groups()]
grads = [torch.randn_like(p).contiguous(memory_format=torch.channels_last) if (p.dim() == 4 and p.shape[2] > 1)
         else torch.randn_like(p) for p in params]
lrs = [l.detach() for l in opt.state["log_lr"]]
outs = [torch.empty_like(p) for p in params]
plan = k.MetaUpdatePlan(params, grads, lrs, outs)
gamma, beta = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
sums = k.gn_stats(x)


def once():
    k.conv2d_fprop(x, w, stride=1, pad=1)                    # CTA pairs (auto)
    k.conv2d_fprop(x, w, stride=1, pad=1, bn_hint=256)       # single CTA
    k.conv2d_wgrad(x, dy, (3, 3), stride=1, pad=1)
    k.gn_backward(x, sums, gamma, beta, dy, mask_mode=1)
    k.meta_update(plan)


for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
