"""Host-side timeline of a fine-tune iteration: wall time spent inside each section (includes waiting at sync points)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from eosvos_b200.util import evaluate as E
from eosvos_b200.networks import mask_rcnn as MR
from eosvos_b200.meta_optim import meta_optim as MOPT
dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
acc = collections.defaultdict(float)
def wrap(obj, name, label=None):
    f = getattr(obj, name); label = label or name
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[label] += time.perf_counter() - t; return r
    setattr(obj, name, g)
for n in ["_build_targets", "_transform", "_prepare_operands", "_backbone", "_rpn", "_roi_heads", "_filter_proposals", "_mask_branch",
          "_rpn_early_targets", "_rpn_fast", "_sample_rois_fast", "_roi_heads_train_fast", "_box_loss_graphed", "_mask_branch_rois",
          "_rpn_cat_outputs", "_rpn_losses"]:
    wrap(model, n)
wrap(model.rpn, "assign_targets_to_anchors"); wrap(model.rpn, "compute_loss", "rpn.compute_loss")
wrap(model.rpn.box_coder, "decode", "rpn.decode"); wrap(model.roi_heads, "select_training_samples")
wrap(opt, "step", "meta_optim.step")
ag = torch.autograd.grad
def agw(*a, **k):
    t = time.perf_counter(); r = ag(*a, **k); acc["autograd.grad(enqueue)"] += time.perf_counter() - t; return r
torch.autograd.grad = agw
E.finetune(model, opt, lambda e: db[e % 4], 4, 1, 1)
torch.cuda.synchronize(); acc.clear()
N = 10
t0 = time.perf_counter()
E.finetune(model, opt, lambda e: db[e % 4], N, 1, 2)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"iteration wall {1e3*tot/N:.2f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]): print(f"  {k:28s} {1e3*v/N:7.2f} ms/iter")
