"""One warm fine-tune iteration (batch 3, 854x480) + one look-ahead run of 5 inference frames (one batched graph for the
target-independent part + five per-frame graphs) inside a cudaProfilerStart/Stop window.
Use with:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... """
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eosvos_b200.util import evaluate as E

dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
frames = [fr[1 + i:2 + i].to(dev) for i in range(bench.FRAMES_PER_STEP)]
tgt = gt0[None, None].to(dev)
n_warm = int(os.environ.get("WARM", "2"))
E.finetune(model, opt, lambda e: db[e % 4], n_warm, 1, 1)
E.run_frames(model, iter(frames), tgt)         # every inference graph exists before the profiled window
torch.cuda.synchronize()
t0 = time.perf_counter()
torch.cuda.cudart().cudaProfilerStart()
ts = []
def tick(e, l):
    torch.cuda.synchronize(); ts.append(time.perf_counter())
E.finetune(model, opt, lambda e: db[e % 4], int(os.environ.get("ITERS", "1")), 1, 2, on_iter=tick if os.environ.get("TICK") else None)
torch.cuda.synchronize()
t1 = time.perf_counter()
if ts: print("per-iter ms:", [round(1e3 * (b - a), 1) for a, b in zip([t0] + ts[:-1], ts)])
E.run_frames(model, iter(frames), tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
t2 = time.perf_counter()
print(f"finetune {1e3*(t1-t0):.1f} ms, {len(frames)} inference frames {1e3*(t2-t1):.1f} ms")
