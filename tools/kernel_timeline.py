"""In-situ per-kernel GPU time of warm fine-tune iterations (+ frames) via torch.profiler (CUPTI activity records).
Unlike the ncu launch list (cold L2, serialised) these are the durations inside the real back-to-back stream.
Never a bench value (the profiler adds host overhead); use for SHARES and per-kernel durations.
  ITERS=3 FRAMES=3 python tools/kernel_timeline.py [out.json]"""
import os, sys, json, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from eosvos_b200.util import evaluate as E

dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
iters = int(os.environ.get("ITERS", "3"))
nfr = int(os.environ.get("FRAMES", "3"))
frames = [fr[1 + (i % 3):2 + (i % 3)].to(dev) for i in range(max(nfr, 1))]
tgt = gt0[None, None].to(dev)
E.finetune(model, opt, lambda e: db[e % 4], 3, 1, 1)
E.run_frames(model, iter(frames), tgt)
torch.cuda.synchronize()


def summarize(prof, label, div):
    acc = collections.defaultdict(lambda: [0.0, 0])
    span = [1e30, 0.0]
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name)[:72]
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        acc[name][0] += dur
        acc[name][1] += 1
        span[0] = min(span[0], ev.time_range.start)
        span[1] = max(span[1], ev.time_range.end)
    tot = sum(v[0] for v in acc.values())
    print(f"== {label}: GPU busy {tot/div/1e3:.2f} ms per unit, span {(span[1]-span[0])/div/1e3:.2f} ms per unit, "
          f"{sum(v[1] for v in acc.values())/div:.0f} launches per unit")
    rows = sorted(acc.items(), key=lambda kv: -kv[1][0])
    for k, v in rows[:int(os.environ.get("TOP", "40"))]:
        print(f"{100*v[0]/tot:5.1f}% {v[0]/div:8.0f} us {v[1]/div:7.1f}  {k}")
    return {k: [v[0] / div, v[1] / div] for k, v in rows}


out = {}
if iters:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        E.finetune(model, opt, lambda e: db[e % 4], iters, 1, 2)
        torch.cuda.synchronize()
    out["iteration"] = summarize(prof, f"fine-tune iteration (batch {db[0][0].shape[0]})", iters)
if nfr:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        E.run_frames(model, iter(frames), tgt)
        torch.cuda.synchronize()
    out["frame"] = summarize(prof, "inference frame", nfr)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=0)
