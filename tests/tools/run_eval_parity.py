"""GPU diagnostic: fine-tune on the GPU for a few iterations, copy the weights into the CPU oracle, then compare
the inference path (boxes, probability maps, thresholded masks) frame by frame from IDENTICAL state."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from eosvos_b200.util import evaluate as E, synthetic
from oracle import model_oracle as MO, ops_oracle as O
import numpy as np

dev = torch.device("cuda:0")
full = len(sys.argv) > 1 and sys.argv[1] == "full"
iters = int(os.environ.get("ITERS", "30"))
Hh, Ww = (480, 854) if full else (120, 214)
model, opt = bench.build_model(dev)
oracle = MO.build_oracle_model(seed=1)
oracle.roi_heads.detections_per_img = 1
if not full:
    for m in (model, oracle):
        m.transform.min_size, m.transform.max_size = (200,), 333
frames, labels = synthetic.make_video(5, 4, Hh, Ww, 1)
fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
gt0 = torch.from_numpy((labels[0] == 1).astype(np.float32))[None, None]
inp = fr[0:1].to(dev).repeat(3, 1, 1, 1); gts = gt0.to(dev).repeat(3, 1, 1, 1)
hist = []
t0 = time.time()
E.finetune(model, opt, lambda e: (inp, gts), iters, 1, 0, on_iter=lambda e, l: hist.append(l.item()))
torch.cuda.synchronize(); print(f"{iters} iters in {time.time()-t0:.2f}s; loss {hist[0]:.3f} -> {hist[-1]:.3f}", [round(h, 2) for h in hist[::5]])
# copy fine-tuned weights into the oracle
sd = {f"{a}.{c}" if a else c: p.detach().cpu() for a, _, c, p in opt.meta_model.param_groups()}
osd = oracle.state_dict()
for k in sd: assert k in osd, k
osd.update(sd); oracle.load_state_dict(osd)
model.eval(); oracle.eval()
if os.environ.get("THRESH"):
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = float(os.environ["THRESH"])
tgt = gt0.clone()
for f in range(1, 4):
    torch.manual_seed(100 + f)
    with torch.no_grad(): oprobs, oboxes = oracle(fr[f:f+1], tgt)
    torch.manual_seed(100 + f)
    with torch.no_grad(): probs, boxes = model(fr[f:f+1].to(dev), tgt.to(dev))
    pm, om = probs.cpu() >= 0.5, oprobs >= 0.5
    inter, union = (pm & om).sum().item(), (pm | om).sum().item()
    gtm = torch.from_numpy(labels[f] == 1)
    print(f"frame {f}: boxes oracle {[round(v,2) for v in oboxes.flatten().tolist()]} product {[round(v,2) for v in boxes.flatten().tolist()]}")
    print(f"   mask IoU(product, oracle) {inter/max(union,1):.5f} (fg oracle {om.sum().item()}, product {pm.sum().item()}, mismatched px {union-inter}); "
          f"max|dprob| {(probs.cpu()-oprobs).abs().max().item():.4f}; J vs GT: oracle {MO.jaccard(om[0,0], gtm):.3f} product {MO.jaccard(pm[0,0], gtm):.3f}")
    nxt = O.threshold_targets(oprobs)
    tgt = gt0 if nxt.sum().item() == 0 else nxt
