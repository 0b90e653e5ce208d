"""Gradients of one training forward with the CUDA-graphed trunk vs the eager trunk (same seeds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from eosvos_b200.util import evaluate as E
dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
inputs, gts = batches[0][0].to(dev), batches[0][1].to(dev)
names = [f"{n_m}.{n_p}" for n_m, _, n_p, _ in opt.meta_model.param_groups()]

def run(graphs, steps):
    model.use_cuda_graphs = graphs
    opt.reset()
    out = []
    for s in range(steps):
        E.set_random_seeds(5 + s)
        model.train_without_dropout()
        loss, _ = model(inputs, gts)
        rec = {}
        ag = torch.autograd.grad
        def agw(*a, **k):
            r = ag(*a, **k); rec["g"] = [g.clone() for g in r]; return r
        torch.autograd.grad = agw
        opt.set_train_loss(loss)
        opt.step(loss)
        torch.autograd.grad = ag
        out.append((loss.item(), rec["g"]))
        ps = {f"{a}.{c}": t for a, _, c, t in opt.meta_model.param_groups()}
        for n, g in zip(names, rec["g"]):
            if not torch.isfinite(g).all():
                print(f"   [graphs={graphs} step {s}] grad {n}: nonfinite {int((~torch.isfinite(g)).sum())} of {g.numel()}, "
                      f"strides {g.stride()}, param finite after update: {bool(torch.isfinite(ps[n]).all())}")
                break
        opt.meta_model.detach_param_groups()
    return out

a = run(False, 3)
b = run(True, 3)
for s, ((la, ga), (lb, gb)) in enumerate(zip(a, b)):
    print(f"step {s}: loss eager {la:.5f} graph {lb:.5f}")
    worst = []
    for n, x, y in zip(names, ga, gb):
        d = (x - y).norm().item() / (x.norm().item() + 1e-12)
        worst.append((d, n, x.norm().item(), y.norm().item()))
    bad = [n for n, y in zip(names, gb) if not torch.isfinite(y).all()]
    print("   non-finite:", len(bad), "of", len(names), [n for n in bad if "backbone" not in n][:12])
    for n, x, y in zip(names, ga, gb):
        if "rpn" in n:
            print("   %-40s eager %.4e graph %.4e  nan-count %d" % (n, x.norm().item(), y.norm().item(), int((~torch.isfinite(y)).sum())))
    worst.sort(reverse=True)
    for w in worst[:8]:
        print("   rel %.3e  %-50s |eager| %.3e |graph| %.3e" % w)
