import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from eosvos_b200.util import evaluate as E
dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
E.finetune(model, opt, lambda e: db[e % 4], 3, 1, 1)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
E.finetune(model, opt, lambda e: db[e % 4], 5, 1, 2)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); ps = pstats.Stats(pr, stream=s).sort_stats("cumulative"); ps.print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); ps = pstats.Stats(pr, stream=s).sort_stats("tottime"); ps.print_stats(25); print(s.getvalue()[:6000])
