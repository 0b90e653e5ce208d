"""ncu targets of round 2 (cudaProfilerStart/Stop window, use --profile-from-start off): the proposal / sampling /
loss kernels of csrc/rpn.cu at the sizes of a batch-3 854x480 iteration and of an inference frame."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchvision.models.detection.anchor_utils import AnchorGenerator
from eosvos_b200 import kernels as k, ops

dev = torch.device("cuda:0")
N, A = 3, 3
feat_shapes = [(192, 336), (96, 168), (48, 84), (24, 42), (12, 21)]
hw = [h * w for h, w in feat_shapes]
g = torch.Generator().manual_seed(0)
heads = [(torch.randn(N * n, 16, generator=g) * torch.tensor([2.0] * 3 + [0.3] * 12 + [0.0])).to(dev) for n in hw]
ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)


class IL:
    pass


il = IL()
il.tensors = torch.empty((N, 3, 768, 1344), device="meta")
il.image_sizes = [(749, 1333)] * N
anchors = ag(il, [torch.empty((N, 1, h, w), device=dev) for h, w in feat_shapes])[0].contiguous()
gt = torch.tensor([[300.0, 200.0, 700.0, 520.0], [350.0, 180.0, 800.0, 600.0], [100.0, 100.0, 500.0, 400.0]], device=dev)
gt_off = torch.tensor([0, 1, 2, 3], dtype=torch.int32, device=dev)
gl = torch.ones(3, dtype=torch.int64, device=dev)


def once():
    boxes, scores, valid, ks = k.rpn_select(heads, hw, A, N, anchors, il.image_sizes, 2000, math.log(1000 / 16), 1e-3, 0.0)
    offs = [0]
    for _ in range(N):
        for kk in ks:
            offs.append(offs[-1] + kk)
    seg = torch.tensor(offs, dtype=torch.int32, device=dev)
    keep = k.nms_segments(boxes.view(-1, 4), seg, N * len(ks), max(ks), 0.7)
    padded, _, count = k.rpn_postnms(hw, A, N, 2000, boxes, scores, valid, keep, 2000)
    all_boxes, labels, matched, counts = k.roi_match(padded, count, gt, gl, gt_off, 1, 0.5)
    al, am, ac = k.rpn_anchor_match(anchors, gt, gt_off, N, 0.7, 0.3)
    cnt = ac.tolist()
    perms = [(torch.randperm(a, device=dev), torch.randperm(b, device=dev)) for a, b in cnt]
    npos = [min(a, 128) for a, _ in cnt]
    nneg = [min(b, 256 - p) for (_, b), p in zip(cnt, npos)]
    inds, _ = k.roi_sample(al, perms, npos, nneg, 256, 128)
    sampled = torch.cat([inds[i, :npos[i] + nneg[i]] + i * anchors.shape[0] for i in range(N)])
    hs = [h.clone().requires_grad_(True) for h in heads]
    lo, lb = ops.rpn_loss(hs, hw, A, sampled, al, am, anchors, gt, gt_off)
    (lo + lb).backward()
    rnd = torch.rand(1, 1, 4, 500, device=dev)
    stats = torch.tensor([[[100, 50, 420, 330, 9000]]], dtype=torch.int32, device=dev)
    out = torch.zeros(1, 1000, 4, device=dev)
    k.extend_boxes(stats, None, rnd, 500, 1.56, 1.56, 1344, 768, 0.1, out, 500)
    head = torch.randn(1000, 16, device=dev)
    k.det_top1(head, out.view(-1, 4), 1, 1000, 2, (10.0, 10.0, 5.0, 5.0), math.log(1000 / 16), 0.05, 1e-2, 1333.0, 749.0, 0.64, 0.64)


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
