"""ncu targets: dominant conv kernel on the roofline shape, wgrad on the same layer, fused update on the real set."""
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
dw = torch.zeros(256, 256, 3, 3, device=dev)
model, opt = bench.build_model(dev)
params = [p.detach() for *_, p in opt.meta_model.param_groups()]
grads = [torch.randn_like(p) for p in params]
lrs = [l.detach() for l in opt.state["log_lr"]]
outs = [torch.empty_like(p) for p in params]
plan = k.MetaUpdatePlan(params, grads, lrs, outs)
for _ in range(3):
    k.conv2d_fprop(x, w, stride=1, pad=1)
    k.conv2d_wgrad(x, dy, (3, 3), stride=1, pad=1, out=dw)
    k.meta_update(plan)
torch.cuda.synchronize()
