"""CPU-only tests: the C-ABI library loads and exports every symbol include/eosvos_b200.h declares, the host
side mirrors the reference API, the product fails loudly without a GPU, sharding / schedule logic, and a
world_size-2 gloo run of the multi-GPU plumbing (no compute calls: there is no GPU here)."""
import math
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import eosvos_b200  # noqa: F401
    from eosvos_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "eosvos_b200.h")).read()
    declared = set(re.findall(r"\b(eosvos_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.eosvos_version() == 100
    assert lib.eosvos_act_dtype() in (0, 1)
    # error path: no device here -> a negative code and a message, never an exception across the ABI
    assert lib.eosvos_device_check(0) != 0 and len(_lib.last_error()) > 0


def test_weight_prep_tables_cover_every_layout():
    """Host half of the tiled operand preparation: every layout spec maps to a contiguous [X][Y][Z] walk, and the
    table / tile list reproduce torch's permute when the kernel's index arithmetic is replayed in numpy."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200 import kernels as K
    from eosvos_b200 import ops
    g = torch.Generator().manual_seed(3)
    cases = [(torch.randn(96, 40, 3, 3, generator=g), "f"), (torch.randn(96, 40, 3, 3, generator=g), "t"),
             (torch.randn(130, 72, 1, 1, generator=g), "f"), (torch.randn(130, 72, 1, 1, generator=g), "t"),
             (torch.randn(64, 3, 7, 7, generator=g), "stem"), (torch.randn(72, 49 * 24, generator=g), ("lf", 24)),
             (torch.randn(72, 49 * 24, generator=g), ("lt", 24)), (torch.randn(40, 72, generator=g), ("lf", 0)),
             (torch.randn(40, 72, generator=g), ("lt", 0)), (torch.randn(24, 16, 2, 2, generator=g), "dc")]
    cap = K._lib.load().eosvos_weight_prep_tile_elems()
    for w, kind in cases:
        for shape, dims, ss, ds in ops._spec_for(kind)(w):
            xyz = K._xyz_of_spec(dims, ss, ds)
            assert xyz is not None, kind
            X, Y, Z, dx, dy, dz = xyz
            assert X * Y * Z == w.numel() and (dx == 1 or dy == 1)
            # reference: the element-wise definition of the spec
            n = int(np.prod(shape))
            ref = np.zeros(n, np.float32)
            src = w.reshape(-1).numpy()
            idx = np.indices(dims).reshape(4, -1)
            ref[(idx * np.array(ds)[:, None]).sum(0)] = src[(idx * np.array(ss)[:, None]).sum(0)]
            # replay of weight_prep_kernel from the host tables
            dst = torch.zeros(n)
            tab, tiles = K.weight_prep_table([(w, dst, X, Y, Z, dx, dy, dz)])
            TX, TY = int(tab[0, 8]), int(tab[0, 9])
            assert TX * TY * Z <= cap and TX <= 64
            out = np.zeros(n, np.float32)
            tiles_y = (Y + TY - 1) // TY
            assert tiles.shape[0] == ((X + TX - 1) // TX) * tiles_y
            for _, tile in tiles:
                x0, y0 = (tile // tiles_y) * TX, (tile % tiles_y) * TY
                xs, ys = np.arange(x0, min(x0 + TX, X)), np.arange(y0, min(y0 + TY, Y))
                xx, yy, zz = np.meshgrid(xs, ys, np.arange(Z), indexing="ij")
                out[xx * dx + yy * dy + zz * dz] = src[(xx * Y + yy) * Z + zz]
            assert np.array_equal(out, ref), kind


def test_no_cpu_fallback():
    import eosvos_b200  # noqa: F401
    from eosvos_b200 import _lib, kernels
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    with pytest.raises(_lib.EosvosError):
        kernels.relu_bwd(torch.zeros(8, dtype=kernels.ACT_DTYPE), torch.zeros(8, dtype=kernels.ACT_DTYPE))
    net = torch.nn.Linear(4, 3)
    opt = MetaOptimizer(net, 1e-2, True, False, 'NEURON', False, None)
    opt.reset()
    opt.eval()
    loss = net(torch.ones(2, 4)).sum()
    with pytest.raises(_lib.EosvosError):
        opt.step(loss)                      # the fused update is a CUDA kernel; on CPU it must raise, not fall back


def test_single_sync_roi_sampling_equals_torchvision():
    """MaskRCNN._select_training_samples (one host sync, nonzero_static) must return exactly what torchvision's
    select_training_samples (the reference's call, mask_rcnn.py:113) returns from the same RNG state -- it is pure
    torch, so the equality is checked on the CPU."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200.networks.mask_rcnn import MaskRCNN
    torch.manual_seed(0)
    model = MaskRCNN('resnet50', num_classes=2,
                     batch_norm={'accum_stats': False, 'learn_weight': False, 'learn_bias': False}, train_encoder=True,
                     roi_pool_output_sizes={'box': 7, 'mask': 28}, eval_augment_rpn_proposals_mode='EXTEND',
                     replace_batch_with_group_norms=True, box_nms_thresh=0.5, maskrcnn_loss='LOVASZ')
    g = torch.Generator().manual_seed(5)
    gts = [torch.tensor([[100.0, 80.0, 300.0, 260.0]]), torch.tensor([[400.0, 120.0, 640.0, 300.0]])]
    proposals, targets = [], []
    for gt in gts:
        ctr = torch.rand(900, 2, generator=g) * torch.tensor([1300.0, 740.0])
        wh = torch.rand(900, 2, generator=g) * 300 + 8
        far = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(min=0)
        near = gt + (torch.rand(40, 4, generator=g) - 0.5) * 60          # enough IoU >= 0.5 candidates: 25 % cap applies
        proposals.append(torch.cat([far, near.clamp(min=0)], 0))
        targets.append({"boxes": gt, "labels": torch.ones(1, dtype=torch.int64),
                        "masks": torch.zeros(1, 8, 8, dtype=torch.uint8)})
    torch.manual_seed(11)
    ref_p, ref_m, ref_l, ref_r = model.roi_heads.select_training_samples([p.clone() for p in proposals], targets)
    torch.manual_seed(11)
    got_p, got_m, got_l, got_r, pos_in = model._select_training_samples([p.clone() for p in proposals], targets)
    for a, b in zip(ref_p + ref_m + ref_l + list(ref_r), got_p + got_m + got_l + list(got_r)):
        assert torch.equal(a, b)
    for l, p in zip(got_l, pos_in):
        assert torch.equal(p, torch.nonzero(l > 0).squeeze(1)) and 0 < p.numel() <= 128
        assert l.numel() == 512
    # the static-shape restatement of fastrcnn_loss used inside the box-branch graph == torchvision's
    from torchvision.models.detection.roi_heads import fastrcnn_loss
    logits = torch.randn(1024, 2, generator=g, requires_grad=True)
    boxreg = torch.randn(1024, 8, generator=g, requires_grad=True)
    lc, lb = fastrcnn_loss(logits, boxreg, got_l, list(got_r))
    sc, sb = MaskRCNN._fastrcnn_loss_static(logits, boxreg, torch.cat(got_l), torch.cat(list(got_r)))
    assert torch.allclose(lc, sc, rtol=1e-6, atol=0) and torch.allclose(lb, sb, rtol=1e-5, atol=1e-8)
    g1 = torch.autograd.grad(lc + lb, (logits, boxreg))
    g2 = torch.autograd.grad(sc + sb, (logits, boxreg))
    assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-9) for a, b in zip(g1, g2))


def test_davis_metrics_and_png_writer(tmp_path):
    """J / F restated from the published DAVIS-2017 definitions (the reference's `davis` package is not vendored):
    analytic cases, per-sequence statistics, and the object-id PNG round trip."""
    import cv2
    import eosvos_b200  # noqa: F401
    from eosvos_b200.util import metrics as MT
    H, W = 120, 160
    a = np.zeros((H, W), bool)
    a[30:90, 40:120] = True
    assert MT.jaccard(a, a) == 1.0 and MT.f_measure(a, a) == 1.0
    assert MT.jaccard(np.zeros_like(a), np.zeros_like(a)) == 1.0 and MT.f_measure(np.zeros_like(a), np.zeros_like(a)) == 1.0
    b = np.zeros_like(a)
    b[30:90, 80:160] = True                                  # shifted by half its width: IoU = 1/3
    assert abs(MT.jaccard(a, b) - (60 * 40) / (60 * 120)) < 1e-12
    assert MT.f_measure(a, np.zeros_like(a)) == 0.0          # precision 0 (n_gt == 0) -> F = 0
    far = np.zeros_like(a)
    far[100:110, 5:15] = True
    assert MT.f_measure(a, far) == 0.0                       # no boundary pixel within the tolerance
    one = np.roll(a, 1, axis=1)                              # one-pixel shift: inside the 2-pixel tolerance band
    assert math.ceil(0.008 * np.linalg.norm((H, W))) == 2 and MT.f_measure(a, one) == 1.0
    three = np.roll(a, 6, axis=1)                            # 6-pixel shift: only the horizontal edges still match
    f3 = MT.f_measure(a, three)
    assert 0.3 < f3 < 0.8
    bm = MT.seg2bmap(a)
    assert bm.sum() == 2 * (60 + 80) and bm[29, 39] and not bm[60, 80]      # one-pixel-wide closed contour
    # statistics: mean / recall(>0.5) / decay(first quarter - last quarter)
    m, r, d = MT.sequence_statistics([1.0, 0.9, 0.8, 0.7, 0.4, 0.3, 0.2, 0.1])
    assert abs(m - 0.55) < 1e-12 and abs(r - 0.5) < 1e-12 and d > 0.5
    # sequence driver drops the first and the last frame
    lab = np.zeros((5, H, W), np.uint8)
    lab[:, 30:90, 40:120] = 1
    pred = lab.copy()
    pred[0] = 0
    pred[4] = 0
    res = MT.evaluate_sequence_jf(pred, lab, 1)
    assert res["J"][0][0] == 1.0 and res["F"][0][0] == 1.0
    paths = MT.save_predictions(pred, str(tmp_path), "seq")
    assert len(paths) == 5 and os.path.basename(paths[3]) == "00003.png"
    back = cv2.imread(paths[2], cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint8 and back.shape == (H, W) and np.array_equal(back, pred[2])


def test_helper_func_mirrors_reference(tmp_path):
    import eosvos_b200  # noqa: F401
    from eosvos_b200.util import helper_func as HF
    from eosvos_b200.networks.mask_rcnn import MaskRCNN
    split = tmp_path / "val.txt"
    split.write_text("bear\ncamel\n")
    ckpt = tmp_path / "parent.pth"
    torch.save({"w": torch.ones(2)}, str(ckpt))
    model, states = HF.init_parent_model(
        'MaskRCNN', 'resnet50', True, None, True, {'accum_stats': False, 'learn_weight': False, 'learn_bias': False},
        {'box': 7, 'mask': 28}, 'EXTEND', 0.5, 'LOVASZ', train={'paths': [str(ckpt)], 'val_split_files': [str(split)]})
    assert isinstance(model, MaskRCNN) and states['train']['splits'] == [['bear', 'camel']]
    assert torch.equal(states['train']['states'][0]['w'], torch.ones(2))
    with pytest.raises(NotImplementedError):
        HF.init_parent_model('DeepLabV3', 'resnet50', True, None, True, {}, {}, None, 0.5, 'LOVASZ')
    # early_stopping (helper_func.py:388-398)
    assert HF.early_stopping([1.0, 0.9], None, 0.0) is False
    assert HF.early_stopping([1.0, 0.9, 0.8], 3, 0.0) is False                 # not more than `patience` entries yet
    assert HF.early_stopping([1.0, 0.5, 0.5, 0.5, 0.5], 3, 0.01) is True       # no improvement over the last 3
    assert HF.early_stopping([1.0, 0.9, 0.8, 0.7, 0.2], 3, 0.01) is False


def test_davis_io_and_dataset_driver(tmp_path):
    """DAVIS-layout round trip (JPEG frames, palette-PNG labels, split file) and the per-dataset driver: sharding by
    whole sequences over ranks, PNG output, J/F -- with the CUDA fine-tune/propagate step replaced by a stub that
    returns the ground truth (the real one needs the GPU; it is covered by tests/test_model_gpu.py)."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200.util import davis_io as IO
    from eosvos_b200.util import synthetic
    root = str(tmp_path / "DAVIS-2017")
    seqs = {"alpha": (5, 1), "beta": (8, 2), "gamma": (4, 1)}
    truth = {}
    for i, (name, (T, K)) in enumerate(seqs.items()):
        fr, lab = synthetic.make_video(3 + i, T, 96, 128, K)
        IO.write_sequence(root, name, fr, lab)
        truth[name] = (fr, lab)
    IO.write_split(root, "val", list(seqs))
    assert IO.list_sequences(root, "val") == list(seqs)
    fr, lab, names, has = IO.load_sequence(root, "beta")
    assert fr.shape == (8, 3, 96, 128) and fr.dtype == torch.float32 and 0.0 <= float(fr.min()) and float(fr.max()) <= 1.0
    assert names[0] == "00000" and has.all() and torch.equal(lab, torch.from_numpy(truth["beta"][1]))     # labels lossless
    ref = torch.from_numpy(truth["beta"][0]).permute(0, 3, 1, 2).float() / 255.0
    assert float((fr - ref).abs().mean()) < 0.02                                                        # JPEG is lossy

    calls = []

    def stub(model, meta_optim, state, frames, first_label, **cfg):
        name = next(n for n, (f, l) in truth.items() if f.shape[0] == frames.shape[0])
        calls.append((name, cfg))
        return torch.from_numpy(truth[name][1]), {"time_per_frame": 0.01}
    out_dir = str(tmp_path / "preds")
    res0 = IO.evaluate_dataset(None, None, None, root, "val", save_dir=out_dir, rank=0, world_size=2, evaluate_fn=stub,
                               num_epochs_eval=10, online_adapt_step=5, online_adapt_epochs=4)
    res1 = IO.evaluate_dataset(None, None, None, root, "val", save_dir=out_dir, rank=1, world_size=2, evaluate_fn=stub,
                               num_epochs_eval=10, online_adapt_step=5, online_adapt_epochs=4)
    assert set(res0) | set(res1) == set(seqs) and not (set(res0) & set(res1))
    assert "beta" in res0                       # the costliest sequence (2 objects, 8 frames) goes to rank 0 first
    for name, r in {**res0, **res1}.items():
        assert r["num_objects"] == seqs[name][1] and len(r["J"]) == seqs[name][1]
        assert all(j[0] == 1.0 for j in r["J"]) and all(f[0] == 1.0 for f in r["F"])
        assert len(os.listdir(os.path.join(out_dir, name))) == seqs[name][0]
    assert calls[0][1]["online_adapt_step"] == 5


def test_reference_checkpoint_key_remap():
    """A MetaOptimizer checkpoint written under torchvision 0.4 module names loads into the optimizer built on the
    installed torchvision (and a model state dict maps onto the installed model's keys)."""
    import re
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer, remap_reference_checkpoint
    from eosvos_b200.networks.mask_rcnn import MaskRCNN
    torch.manual_seed(0)
    model = MaskRCNN('resnet50', num_classes=2,
                     batch_norm={'accum_stats': False, 'learn_weight': False, 'learn_bias': False}, train_encoder=True,
                     roi_pool_output_sizes={'box': 7, 'mask': 28}, eval_augment_rpn_proposals_mode='EXTEND',
                     replace_batch_with_group_norms=True, box_nms_thresh=0.5, maskrcnn_loss='LOVASZ')
    opt = MetaOptimizer(model, 1e-3, True, False, 'NEURON', False, None)

    def to_tv04(k, sep):
        q = re.escape(sep)
        k = re.sub(rf"mask_head{q}(\d+){q}0{q}", lambda m: f"mask_head{sep}mask_fcn{int(m.group(1)) + 1}{sep}", k)
        k = re.sub(rf"rpn{q}head{q}conv{q}0{q}0{q}", f"rpn{sep}head{sep}conv{sep}", k)
        return re.sub(rf"fpn{q}(inner_blocks|layer_blocks){q}(\d+){q}0{q}", rf"fpn{sep}\1{sep}\2{sep}", k)
    new = opt.state_dict()
    old = {to_tv04(k, "-"): v.clone() + 1.0 for k, v in new.items()}
    assert set(old) != set(new) and "model_init_roi_heads-mask_head-mask_fcn1-weight" in old
    assert "log_init_lr_rpn-head-conv-weight" in old and "model_init_backbone-fpn-inner_blocks-2-bias" in old
    assert set(remap_reference_checkpoint(old)) == set(new)
    assert set(remap_reference_checkpoint(dict(new))) == set(new)            # idempotent on current names
    opt.load_state_dict(old)                                                 # strict: every key must land
    for k, v in opt.state_dict().items():
        assert torch.equal(v, old[to_tv04(k, "-")])
    msd = model.state_dict()
    assert set(remap_reference_checkpoint({to_tv04(k, "."): v for k, v in msd.items()})) == set(msd)


def test_meta_optimizer_mirrors_reference_api():
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.GroupNorm(4, 8), torch.nn.Flatten(),
                              torch.nn.Linear(8 * 6 * 6, 5))
    for level, use_log in (("NEURON", False), ("NEURON", True), ("PARAM", False), ("TENSOR", False), ("SINGLE", False)):
        opt = MetaOptimizer(net, 1e-2, True, False, level, use_log, 0.1)
        keys = list(opt.state_dict().keys())
        assert any(k.startswith("model_init_0-weight") for k in keys)
        if level in ("NEURON", "PARAM"):
            assert "log_init_lr_0-weight" in keys and opt.state_dict()["log_init_lr_0-weight"].shape == (
                (8, 1, 1, 1) if level == "NEURON" else (8, 3, 3, 3))
            assert opt.state_dict()["log_init_lr_1-bias"].shape == (8,)
        opt.reset()
        opt.clamp_init_lr()
        assert opt.state["num_steps"] == 0
        assert float(torch.as_tensor(opt.init_lr).max()) <= 0.1 + 1e-6
        opt.meta_model.detach_param_groups()
        assert all(p.requires_grad and p.is_leaf for *_, p in opt.meta_model.param_groups())
        opt.reset()
        assert all(p is q for (*_, p), q in zip(opt.meta_model.param_groups(), opt._model_init.values()))
    with pytest.raises(NotImplementedError):
        o = MetaOptimizer(net, 1e-2, True, True, "NEURON", False, None)
        o.train()
        o.step(net(torch.ones(1, 3, 8, 8)).sum())


def test_run_frames_lookahead_protocol(monkeypatch):
    """Host logic of run_frames' look-ahead (no GPU: a stand-in model): frames are announced in runs of
    EOSVOS_FRAME_BATCH before the first of them is run, the next run is announced only after the last frame of the
    current one ran, every frame is run exactly once and in order, and the announcement is dropped at the end."""
    from eosvos_b200.util import evaluate as E

    class Rpn:
        _eval_augment_proposals_mode = None

    class Stub:
        def __init__(self):
            self.rpn, self.num_classes, self.events, self._lookahead = Rpn(), 2, [], None

        def eval(self):
            return self

        def lookahead_ok(self):
            return True

        def prefetch_frames(self, frames, has_target):
            self.events.append(("announce", [int(f[0, 0, 0, 0]) for f in frames], has_target))
            self._lookahead = object()
            return True

        def __call__(self, inputs, targets):
            self.events.append(("run", int(inputs[0, 0, 0, 0])))
            self.last_propagated_target = self.last_target_stats = None
            return torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4)

    for batch, want in (("3", [[0, 1, 2], [3, 4, 5], [6, 7]]), ("5", [[0, 1, 2, 3, 4], [5, 6, 7]]), ("1", [])):
        monkeypatch.setenv("EOSVOS_FRAME_BATCH", batch)
        m = Stub()
        frames = (torch.full((1, 3, 4, 4), float(i)) for i in range(8))
        seen = []
        probs, boxes = E.run_frames(m, frames, None, on_frame=lambda i, t, p, b: seen.append(i))
        assert probs.shape[0] == 8 and seen == list(range(8)) and m._lookahead is None
        assert [e[1] for e in m.events if e[0] == "run"] == list(range(8))
        assert [e[1] for e in m.events if e[0] == "announce"] == want
        assert all(e[2] is False for e in m.events if e[0] == "announce")
        for run in want:                       # announced before its first frame, after the previous run's last one
            k = m.events.index(("announce", run, False))
            assert m.events[k + 1] == ("run", run[0]) and (run[0] == 0 or m.events[k - 1] == ("run", run[0] - 1))


def test_ona_schedule_and_sharding():
    import eosvos_b200  # noqa: F401
    from eosvos_b200.util import shard, synthetic
    s = shard.ona_schedule(70, 100, 5, 10)          # e-OSVOS-100-OnA on a 70-frame video
    assert sum(x[1] for x in s) == 230 and sum(x[3] - x[2] for x in s) == 69 and len(s) == 14
    assert s[0] == (0, 100, 1, 6) and s[-1][3] == 70
    assert shard.ona_schedule(10, 10) == [(0, 10, 1, 10)]
    spec = synthetic.davis_val_shaped_set()
    units = [(v, o, T) for v, (T, K) in enumerate(spec) for o in range(K)]
    assert len(spec) == 30
    for world in (1, 2, 4, 8):
        shards, loads = shard.shard_units(units, world, num_epochs_eval=100, online_adapt_step=5)
        assert sorted(u for sh in shards for u in sh) == sorted(units)
        assert max(loads) <= 1.25 * (sum(loads) / world) + max(loads) / len(units)


def test_synthetic_and_augmentation():
    import random
    import eosvos_b200  # noqa: F401
    from eosvos_b200.util import augment, synthetic
    f1, l1 = synthetic.make_video(3, 3, 120, 214, 2)
    f2, l2 = synthetic.make_video(3, 3, 120, 214, 2)
    assert np.array_equal(f1, f2) and np.array_equal(l1, l2) and set(np.unique(l1)) == {0, 1, 2}
    assert (l1[0] == 1).mean() > 0.01
    random.seed(1)
    img, gt = augment.augment_first_frame(f1[0].astype(np.float32) / 255, (l1[0] == 1).astype(np.float32))
    assert img.shape == (120, 214, 3) and gt.shape == (120, 214) and gt.max() == 1.0


def test_two_rank_gloo_plumbing(tmp_path):
    """N > 1 path on CPU: gloo, world_size 2 -- disjoint shards, max-over-ranks timing reduction."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        import eosvos_b200
        from eosvos_b200.util import shard, synthetic
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        spec = synthetic.davis_val_shaped_set()
        units = [(v, o, T) for v, (T, K) in enumerate(spec) for o in range(K)]
        shards, loads = shard.shard_units(units, w, num_epochs_eval=100, online_adapt_step=5)
        mine = torch.zeros(len(units)); idx = {{u: i for i, u in enumerate(units)}}
        for u in shards[r]: mine[idx[u]] = 1
        dist.all_reduce(mine)
        t = torch.tensor([float(loads[r])], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert bool((mine == 1).all()), "shards must partition the units"
        assert abs(t.item() - max(loads)) < 1e-9
        if r == 0: print("OK", w)
        dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK 2" in out.stdout, out.stderr[-2000:]
