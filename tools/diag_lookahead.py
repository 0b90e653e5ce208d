"""Where does the batched look-ahead differ from per-frame inference?  Prints max |difference| stage by stage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from test_model_gpu import build_pair
from eosvos_b200.util import evaluate as E, synthetic
from eosvos_b200 import kernels as K

if os.environ.get("POLLUTE"):
    # freed blocks are handed out again without clearing: any read of never-written memory now sees NaN / huge values
    junk = [torch.full(((256 << 20) // 4,), float(os.environ["POLLUTE"]), device="cuda") for _ in range(24)]
    del junk
model, opt, _, _, dev, _ = build_pair(min_size=240, max_size=427)
frames, labels = synthetic.make_video(5, 8, 240, 427, 1)
fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
gt0 = torch.from_numpy((labels[0] == 1).astype(np.float32))[None, None]
inp, gts = fr[0:1].to(dev).repeat(3, 1, 1, 1), gt0.to(dev).repeat(3, 1, 1, 1)
E.finetune(model, opt, lambda e: (inp, gts), 8, 1, 0)
model.roi_heads.detections_per_img = 1
model.roi_heads.score_thresh = 0.0
model.eval()
NF = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fs = [fr[f:f + 1].to(dev) for f in range(1, 1 + NF)]
with torch.no_grad():
    assert model.prefetch_frames(fs, True)
    g = [t.clone() for t in model._lookahead["outs"]]
    theta = [m._parameters[n] for m, n in model._pre_slots]
    model._frame_cfg = model._lookahead["cfg"]
    eb = [t.clone() for t in model._frames_pre_functional(torch.cat(fs), *theta)]
    names = ["f0", "f1", "f2", "f3", "padded", "count"]
    print("graph vs eager (same batch):", {n: float((a.float() - b.float()).abs().max()) for n, a, b in zip(names, g, eb)})
    for i in range(NF):
        e1 = model._frames_pre_functional(fs[i], *theta)
        print(f"eager batched slice {i} vs eager F=1:",
              {n: float((a[i:i + 1].float() - b.float()).abs().max()) for n, a, b in zip(names, eb, e1)},
              "count", int(eb[5][i]), int(e1[5][0]))
    # tails
    stats = K.mask_to_bbox(gt0.to(dev), 1)
    cfg = model._lookahead["cfg"]
    torch.manual_seed(5)
    rnd = model._extend_rands(1, 1, cfg["n_aug"], dev)
    tt = [m._parameters[n] for m, n in model._tail_slots]
    for i in range(NF):
        e1 = model._frames_pre_functional(fs[i], *theta)
        a = model._frame_tail_functional(*[t[i:i + 1] for t in eb[:5]], stats, stats, rnd, *tt)
        b = model._frame_tail_functional(*[t.clone() for t in e1[:5]], stats, stats, rnd, *tt)
        print(f"tail on slice {i} vs tail on F=1 tensors: probs", float((a[0] - b[0]).abs().max()), "box",
              float((a[1] - b[1]).abs().max()), "rows", int(a[4]), int(b[4]))
    # graph-replayed tails (through forward) against the eager tail on the same slices, same stats, same uniforms
    print("--- graphed tails")
    for rep in range(2):
        assert model.prefetch_frames(fs, True)
        outs = model._lookahead["outs"]
        for i in range(NF):
            torch.manual_seed(100 + i)
            rnd = model._extend_rands(1, 1, cfg["n_aug"], dev).clone()
            model._frame_cfg = cfg
            ref = [t.clone() for t in model._frame_tail_functional(*[t[i:i + 1].clone() for t in outs[:5]], stats, stats, rnd, *tt)]
            torch.manual_seed(100 + i)
            probs, box = model._forward_eval_tail(i, fs[i], (stats, None))
            print(f"rep {rep} tail graph {i} vs eager: probs", float((probs - ref[0]).abs().max()), "box",
                  float((box.view(-1) - ref[1].view(-1)).abs().max()), "rows", int(model._last_det["row"][0]), int(ref[4][0]),
                  "padded", float((outs[4][i] - outs[4][i]).abs().max()))
