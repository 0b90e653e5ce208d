"""Where does the end-to-end inference block (5 frames from pinned host memory, probabilities read back) spend its time?
Times the same 5-frame run_frames call in four settings (wall clock around a synchronise, 20 repeats)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eosvos_b200.util import evaluate as E, augment

dev = torch.device("cuda:0")
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
db = [(a.to(dev), b.to(dev)) for a, b in batches]
E.finetune(model, opt, lambda e: db[e % 4], bench.FIRST_ROUND_ITERS, 1, 0)
N = bench.FRAMES_PER_STEP
dev_frames = [fr[1 + i:2 + i].to(dev) for i in range(N)]
pin_frames = [fr[1 + i:2 + i].pin_memory() for i in range(N)]
tgt = gt0[None, None].to(dev)


def run(host, read_back, aug=False, reps=20):
    ts = []
    for r in range(reps + 3):
        a = None
        if aug:
            a = augment.PrefetchingAugmenter(fr[0].contiguous().pin_memory().to(dev, non_blocking=True), gt0.numpy(),
                                             bench.BATCH, lambda e: 1 + e, first_epoch=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        get = (lambda i: pin_frames[i].to(dev, non_blocking=True)) if host else (lambda i: dev_frames[i])
        probs, boxes = E.run_frames(model, (get(i) for i in range(N)), tgt)
        if read_back:
            s = float(E.to_host(probs).sum())
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        if a is not None:
            a.close()
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


for name, kw in (("device frames", dict(host=False, read_back=False)), ("host frames (H2D)", dict(host=True, read_back=False)),
                 ("host frames + D2H + host sum", dict(host=True, read_back=True)),
                 ("... + augmenter thread of the next block", dict(host=True, read_back=True, aug=True))):
    ms = run(**kw)
    print(f"{name:45s} {ms:7.2f} ms per {N}-frame block  ({N / ms * 1e3:6.1f} frames/s)", flush=True)
t0 = time.perf_counter(); x = pin_frames[0].to(dev, non_blocking=True); torch.cuda.synchronize()
print("one 4.9 MB H2D:", round((time.perf_counter() - t0) * 1e3, 3), "ms")
