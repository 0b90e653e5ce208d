import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from eosvos_b200.util import evaluate as E
from eosvos_b200 import kernels as K
dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
E.finetune(model, opt, lambda e: db[e % 4], 3, 1, 1)
model.roi_heads.score_thresh = float(os.environ.get("THRESH", "0.5"))
frames = [fr[1 + (i % 3):2 + (i % 3)].to(dev) for i in range(40)]
tgt = gt0[None, None].to(dev)
E.run_frames(model, iter(frames[:5]), tgt)
acc = collections.defaultdict(float)
def wrap(obj, name, label=None):
    f = getattr(obj, name); label = label or name
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[label] += time.perf_counter() - t; return r
    setattr(obj, name, g)
for n in ["_build_targets", "_transform", "_prepare_operands", "_backbone", "_rpn", "_roi_heads", "_filter_proposals", "_mask_branch", "_postprocess_detections", "_nms_by_label"]:
    wrap(model, n)
wrap(K, "mask_paste_threshold")
torch.cuda.synchronize(); t0 = time.perf_counter()
E.run_frames(model, iter(frames), tgt)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print(f"frame wall {1e3*tot/len(frames):.2f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]): print(f"  {k:28s} {1e3*v/len(frames):7.2f} ms/frame")
