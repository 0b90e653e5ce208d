"""Sweep the column-tile width (bn_hint) of conv_fprop for every distinct forward / data-gradient shape of a batch-B
iteration: prints us per launch for BN = 64 / 128 / 256 (L2 flushed between launches).  B=3: fine-tuning, B=1: inference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from eosvos_b200 import kernels as K

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "3"))
model, opt = bench.build_model(dev)
fr, gt0, batches = bench.build_workload(1)
batch = (batches[0][0][:B].to(dev), batches[0][1][:B].to(dev))
calls = bench.record_contractions(model, opt, batch)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for key, count in calls.items():
    if key[0] not in ("fprop", "dgrad"):
        continue
    if key[0] == "fprop":
        _, xs, ws, stride, pad = key
        x = (torch.randn(*xs, device=dev) * 0.1).to(K.ACT_DTYPE)
        w = (torch.randn(*ws, device=dev) * 0.1).to(K.ACT_DTYPE)
        cout = ws[0]
        fn = lambda bn: K.conv2d_fprop(x, w, stride=stride, pad=pad, bn_hint=bn)
    else:
        _, ds, ws, in_hw, stride, pad = key
        dy = (torch.randn(*ds, device=dev) * 0.1).to(K.ACT_DTYPE)
        w = (torch.randn(*ws, device=dev) * 0.1).to(K.ACT_DTYPE)
        cout = ws[0]
        if stride != 1:
            continue
        fn = lambda bn: K.conv2d_dgrad(dy, w, in_hw, stride=stride, pad=pad, bn_hint=bn)
    res = {}
    for bn in (0, 64, 128, 256):
        if bn and bn > max(64, (cout + 63) // 64 * 64):
            continue
        try:
            res[bn] = bench._time_launch(lambda: fn(bn), flush, 5) * 1e6
        except Exception as e:
            res[bn] = float("nan")
    rows.append((key, count, res))
rows.sort(key=lambda r: -r[2][0] * r[1])
for key, count, res in rows:
    best = min((v, k) for k, v in res.items() if k and v == v)
    print(f"{key[0]:6s} {str(key[1]):24s} {str(key[2]):22s} x{count:2d}  auto {res[0]:6.1f}  " +
          "  ".join(f"bn{k} {v:6.1f}" for k, v in res.items() if k) + f"   best bn{best[1]} ({res[0] / best[0]:.2f}x)")
tot_auto = sum(r[2][0] * r[1] for r in rows)
tot_best = sum(min(v for k, v in r[2].items() if k and v == v) * r[1] for r in rows)
print(f"B={B}: auto {tot_auto:.0f} us, best-of-sweep {tot_best:.0f} us")
