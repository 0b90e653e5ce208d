"""wgrad of the 64-channel layers -- three shallow CTAs per SM (default) vs the single deep-ring CTA
(EOSVOS_WGRAD_LIGHT=0) -- and of the stride-2 1x1 projections (gathered input), each launch alone with an L2 flush in between (as bench.py times layers)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import bench
from eosvos_b200 import kernels as K
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for (N, H, W, Ci, Co, k, st) in [(3, 192, 336, 64, 64, 3, 1), (3, 192, 336, 64, 256, 1, 1), (3, 48, 84, 1024, 2048, 1, 2),
                                 (3, 96, 168, 512, 1024, 1, 2), (3, 192, 336, 256, 512, 1, 2)]:
    x = torch.randn(N, H, W, Ci, device=dev).to(K.ACT_DTYPE)
    dy = torch.randn(N, (H - 1) // st + 1, (W - 1) // st + 1, Co, device=dev).to(K.ACT_DTYPE)
    dw = torch.zeros((Co, k, k, Ci), device=dev).permute(0, 3, 1, 2) if k > 1 else torch.zeros((Co, Ci, 1, 1), device=dev)
    t = bench._time_launch(lambda: K.conv2d_wgrad(x, dy, (k, k), stride=st, pad=k // 2, out=dw), flush, 7)
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).float(), (Co, Ci, k, k), dy.permute(0, 3, 1, 2).float(), stride=st, padding=k // 2)
    dw.zero_(); K.conv2d_wgrad(x, dy, (k, k), stride=st, pad=k // 2, out=dw); torch.cuda.synchronize()
    err = ((dw - ref).norm() / ref.norm()).item()
    out.append("%%.1f us (rel err %%.1e)" %% (t * 1e6, err))
print("RESULT", " | ".join(out))
''' % ROOT
for light in ("1", "0"):
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, EOSVOS_WGRAD_LIGHT=light))
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print("light=" + light, line[0] if line else r.stderr[-500:])
