"""Hook-free parity of the CUDA path against the oracle: the product's OWN proposals, EXTEND boxes, detection, frame
propagation and per-sequence evaluation, and the drop-in claim (row b) through the
reference's own call sites.

How a comparison between a 16-bit-storage GPU implementation and an fp32 CPU one is made well-posed
--------------------------------------------------------------------------------------------------
The path contains DISCRETE choices (top-k membership, NMS, arg-max detection among ~500 jittered copies of one box,
`int64` truncation of the paste box).  Two correct implementations whose logits differ by rounding noise can make
different choices where the inputs are (nearly) tied, and the outputs then differ by much more than the noise.  The
tests therefore compare from IDENTICAL state per frame (same weights, same propagated target, same CPU random stream)
and separate the two effects:
  * continuous quantities (scores, box coordinates, probabilities) must agree within the stated tolerance;
  * a discrete choice must either be equal, or be a tie within that tolerance (|score difference| <= 0.02) -- the
    oracle is then asked for ITS output under the product's choice and the continuous comparison is made on that.
Mask tolerance (BASELINE.json north_star: IoU >= 0.999): asserted when both sides truncate the paste box to the same
integers; a flipped truncation (a coordinate within the box tolerance of an integer) rescales the pasted mask by one
pixel, for which IoU >= 0.98 is asserted and the event is counted and printed.
"""
import copy
import os
from unittest import mock

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import evaluate_oracle as EO  # noqa: E402
from oracle import model_oracle as MO  # noqa: E402
from oracle import ops_oracle as O  # noqa: E402
from tests.test_model_gpu import build_pair, det_randperm, frame, rel  # noqa: E402
from tests.conftest import statistical  # noqa: E402


def _sync_state(opt, oracle):
    """theta of the product (whatever is installed in module._parameters) -> the oracle model."""
    sd = {f"{a}.{c}": p.detach().cpu() for a, _, c, p in opt.meta_model.param_groups()}
    osd = oracle.state_dict()
    osd.update(sd)
    oracle.load_state_dict(osd)


def _match_boxes(a, b):
    """for every row of a: (index of the nearest row of b, max-abs coordinate distance)."""
    d = (a[:, None, :] - b[None, :, :]).abs().amax(dim=2)
    val, idx = d.min(dim=1)
    return idx, val


def _iou(a, b):
    u = (a | b).sum().item()
    return 1.0 if u == 0 else (a & b).sum().item() / u


# ---------------------------------------------------------------------------------------------- a4
@pytest.mark.parametrize("full_size", [False, True])
@statistical()
def test_own_proposals_match_oracle(full_size):
    """rpn_forward (mask_rcnn.py:217-344 -> tv filter_proposals) WITHOUT the fixed_proposals hook, training and
    evaluation mode: decode, per-level top-k, clip, small-box / score filter, per-level NMS, post-NMS top-n, and the
    EXTEND augmentation.  fp16 features perturb the objectness by ~1e-2 relative, so membership at the top-k / NMS
    boundaries may differ: >= 95 % of the oracle's proposals must be present (a product box within 1 px; measured
    98.7 %), the matched boxes must agree to 1 px (median <= 0.15 px; measured 0.04 .. 0.09), and the jittered EXTEND
    boxes -- CPU random stream, exact integer target boxes -- must be bit-equal."""
    if full_size:
        model, opt, oracle, _, dev, _ = build_pair(min_size=None)
        from eosvos_b200.util import synthetic
        frames, labels = synthetic.make_video(5, 2, 480, 854, 1)
        img = torch.from_numpy(frames[:1]).permute(0, 3, 1, 2).float().div(255.0).contiguous()
        tgt = torch.from_numpy((labels[0] == 1).astype(np.float32))[None, None]
    else:
        model, opt, oracle, _, dev, _ = build_pair()
        img, tgt = frame()
    for train in (True, False):
        if train:
            oracle.train_without_dropout()
            model.train_without_dropout()
        else:
            oracle.eval()
            model.eval()
        with mock.patch("torch.randperm", det_randperm(5)):
            torch.manual_seed(21)
            with torch.set_grad_enabled(train):
                oracle(img, tgt)
        with mock.patch("torch.randperm", det_randperm(5)):
            torch.manual_seed(21)
            with torch.set_grad_enabled(train):
                model(img.to(dev), tgt.to(dev))
        own = model.last_proposals[0].detach().cpu()          # the graphed production path: no capture hook set
        ref = oracle.last_proposals[0]
        n_ext = 0
        if not train:
            n_ext = 500                                   # post_nms_top_n(testing) // 2 jittered copies per target box
            assert torch.equal(own[-n_ext:], ref[-n_ext:]), "EXTEND boxes differ"
            own, ref = own[:-n_ext], ref[:-n_ext]
        idx, dist = _match_boxes(ref, own)
        present = (dist <= 1.0).float().mean().item()
        matched = dist[dist <= 1.0]
        print(f"{'train' if train else 'eval'} {'854x480' if full_size else 'small'}: own {own.shape[0]} oracle "
              f"{ref.shape[0]} RPN proposals (+{n_ext} EXTEND, bit-equal); present {present:.4f}; matched box distance "
              f"median {matched.median().item():.4f} max {matched.max().item():.4f} px")
        assert abs(own.shape[0] - ref.shape[0]) <= 0.05 * ref.shape[0] + 2
        assert present >= 0.95
        assert matched.median().item() <= 0.15 and matched.max().item() <= 1.0
        # order: proposals are ranked by objectness; a common box sits at (nearly) the same rank on both sides
        common = idx[dist <= 1.0].float()
        at = torch.nonzero(dist <= 1.0).squeeze(1).float()
        disp = ((common - at).abs() / max(ref.shape[0], 1)).mean().item()
        print(f"    mean rank displacement of common proposals {disp:.4f} of the list length")
        assert disp <= 0.03


# ---------------------------------------------------------------------------------------------- a7
def test_own_detections_match_oracle():
    """postprocess_detections (mask_rcnn.py:347-420) on the product's OWN class logits / box regression / proposals
    (moved to the CPU): decode, softmax, clip, background drop, score threshold, small-box removal, per-class NMS,
    detections_per_img -- same inputs, so boxes / scores / labels / chosen rows must agree to fp32 rounding, for
    detections_per_img = 100 and = 1 (multi_object 'single_id')."""
    model, opt, oracle, _, dev, _ = build_pair()
    img, tgt = frame()
    model.eval()
    oracle.eval()
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = 0.05
    for per_img in (100, 1):
        model.roi_heads.detections_per_img = oracle.roi_heads.detections_per_img = per_img
        model.capture = {}
        torch.manual_seed(4)
        with torch.no_grad():
            model(img.to(dev), tgt.to(dev))
        cap = model.capture
        model.capture = None
        det = cap["detections"][0]
        rows = model.last_detection_rows[0].cpu()
        oh, ow = cap["x8"].shape[1:3]
        sizes = [tuple(model._resized_size(img.shape[-2], img.shape[-1]))]
        ref = oracle.detections_stage(cap["class_logits"].float().cpu(), cap["box_regression"].float().cpu(),
                                      [p.float().cpu() for p in cap["proposals"]], sizes)[0]
        oracle_rows = oracle.last_candidates[-1]["rows"] if hasattr(oracle, "last_candidates") else None
        n = ref["boxes"].shape[0]
        print(f"detections_per_img {per_img}: {det['boxes'].shape[0]} own vs {n} oracle detections")
        assert det["boxes"].shape[0] == n and n >= 1
        assert torch.equal(det["labels"].cpu(), ref["labels"])
        assert (det["scores"].cpu() - ref["scores"]).abs().max().item() <= 1e-5
        assert (det["boxes"].cpu() - ref["boxes"]).abs().max().item() <= 1e-3
        if oracle_rows is not None:
            assert torch.equal(rows, oracle_rows)


# ---------------------------------------------------------------------------------------------- a9 / a10 lock-step
class LockStep:
    """Records what the product did per round / frame and replays the oracle from the same state."""

    def __init__(self, model, opt, oracle, dev):
        self.model, self.opt, self.oracle, self.dev = model, opt, oracle, dev
        self.rounds = {}
        self.cur = None

    # product-side hooks ------------------------------------------------------------------------
    def after_finetune(self, obj, k, model, meta_optim):
        self.rounds[(obj, k)] = {"state": {f"{a}.{c}": p.detach().cpu().clone()
                                           for a, _, c, p in meta_optim.meta_model.param_groups()}, "frames": []}

    def before_frames(self, obj, k, frame_ids, start_target):
        r = self.rounds.setdefault((obj, k), {"frames": []})
        r.update(frame_ids=frame_ids, start=start_target.clone(), rng=torch.get_rng_state().clone())

    def on_frame(self, obj, k, i, targets, probs, boxes):
        m = self.model
        self.rounds[(obj, k)]["frames"].append(dict(
            target=None if targets is None else targets.detach().cpu().clone(), probs=probs.detach().cpu().clone(),
            boxes=boxes.detach().cpu().clone(), rows=[r.cpu().clone() for r in m.last_detection_rows],
            proposals=m.last_proposals[0].detach().cpu().clone()))

    def hooks(self):
        return {"after_finetune": self.after_finetune, "before_frames": self.before_frames, "on_frame": self.on_frame}

    # oracle replay -----------------------------------------------------------------------------
    def replay(self, frames, labels=None, score_tie=0.02):
        """-> list of per-frame dicts with the comparison figures."""
        out = []
        oracle = self.oracle
        for (obj, k), r in sorted(self.rounds.items()):
            if "state" in r:
                osd = oracle.state_dict()
                osd.update(r["state"])
                oracle.load_state_dict(osd)
            oracle.eval()
            torch.set_rng_state(r["rng"])
            for i, (f, rec) in enumerate(zip(r["frame_ids"], r["frames"])):
                oracle.fixed_detection_rows = None
                rng_before = torch.get_rng_state().clone()
                with torch.no_grad():
                    oprobs, oboxes = oracle(frames[f:f + 1], rec["target"])
                cand = oracle.last_candidates[0]
                o_rows = cand["rows"]
                p_rows = rec["rows"][0]
                info = dict(obj=obj, k=k, frame=f, n_det=(int(p_rows.numel()), int(o_rows.numel())), tie=False)
                if p_rows.numel() != o_rows.numel() and cand["scores"].numel():
                    # one side's best score fell on the other side of roi_heads.score_thresh: a tie against the
                    # threshold when the oracle's best score is within `score_tie` of it
                    info["thresh_tie"] = abs(float(cand["scores"].max()) - float(oracle.roi_heads.score_thresh)) <= score_tie
                if p_rows.numel() and o_rows.numel():
                    # map the product's source row to the oracle's proposal list (EXTEND rows sit at the end of both)
                    o_props = oracle.last_proposals[0]
                    n_o, n_p = o_props.shape[0], rec["proposals"].shape[0]
                    n_ext = 500 if rec["target"] is not None else 0
                    pr = int(p_rows[0])
                    if pr >= n_p - n_ext:
                        orow = n_o - (n_p - pr)
                        assert torch.equal(o_props[orow], rec["proposals"][pr]), "EXTEND boxes differ"
                    else:                         # an RPN proposal: the oracle's nearest one, if it has it
                        j, d = _match_boxes(rec["proposals"][pr:pr + 1], o_props[:n_o - n_ext])
                        orow = int(j[0]) if float(d[0]) <= 1.0 else None
                        if orow is None:
                            # the product's winner is an RPN proposal the oracle's own list lacks (the two lists agree
                            # to ~99 %: test_own_proposals_and_detection_match_oracle; the rest sits at the NMS / top-n
                            # cut): judge the detection stage on the product's proposal list instead
                            rng_after = torch.get_rng_state().clone()
                            oracle.fixed_proposals = [rec["proposals"].clone()]
                            torch.set_rng_state(rng_before)
                            with torch.no_grad():
                                oprobs, oboxes = oracle(frames[f:f + 1], rec["target"])
                            oracle.fixed_proposals = None
                            torch.set_rng_state(rng_after)
                            cand = oracle.last_candidates[0]
                            o_rows, orow = cand["rows"], pr
                            info["foreign_proposal"] = True
                    info["same_choice"] = orow is not None and orow == int(o_rows[0])
                    if not info["same_choice"] and orow is not None:
                        gap = float(cand["scores"][int(o_rows[0])] - cand["scores"][orow])
                        info["tie"], info["score_gap"] = gap <= score_tie, gap
                        if info["tie"]:           # the oracle's output under the product's (tied) choice
                            rng_after = torch.get_rng_state().clone()
                            oracle.fixed_detection_rows = [torch.tensor([orow])]
                            if info.get("foreign_proposal"):
                                oracle.fixed_proposals = [rec["proposals"].clone()]
                            torch.set_rng_state(rng_before)
                            with torch.no_grad():
                                oprobs, oboxes = oracle(frames[f:f + 1], rec["target"])
                            oracle.fixed_detection_rows = oracle.fixed_proposals = None
                            torch.set_rng_state(rng_after)
                pm, om = rec["probs"] >= 0.5, oprobs >= 0.5
                # boundary-tolerant agreement: pixels farther than 1 px from the oracle mask's boundary
                import torch.nn.functional as Fn
                of = om.float()
                inner = -Fn.max_pool2d(-of, 3, 1, 1) > 0.5            # erosion
                outer = Fn.max_pool2d(of, 3, 1, 1) > 0.5              # dilation
                core = inner | ~outer
                info["iou_off_boundary"] = _iou(pm & core, om & core)
                info.update(dprob_max=(rec["probs"] - oprobs).abs().max().item(),
                            dprob_mean=(rec["probs"] - oprobs).abs().mean().item(),
                            dbox=(rec["boxes"] - oboxes).abs().max().item(), iou=_iou(pm, om),
                            px=(int(pm.sum()), int(om.sum())))
                if labels is not None:
                    g = (labels[f] == obj + 1)[None, None]
                    info["dJ"] = abs(_iou(pm, g) - _iou(om, g))
                if rec["boxes"].abs().sum() > 0 and oboxes.abs().sum() > 0:
                    from tests.test_model_gpu import _paste_int_box
                    info["same_int_box"] = bool(torch.equal(_paste_int_box(rec["boxes"][0, 0]),
                                                            _paste_int_box(oboxes[0, 0])))
                out.append(info)
        return out


def _check_lockstep(infos, box_tol=0.5, require_det=True):
    flips = ties = 0
    for it in infos:
        print(it)
        if it["n_det"][0] != it["n_det"][1]:
            assert it.get("thresh_tie"), f"detection on one side only and not a tie against the score threshold: {it}"
            ties += 1
            continue
        if it["n_det"][0] == 0:
            assert it["px"] == (0, 0)
            continue
        assert it.get("same_choice") or it["tie"], f"different detection and not a score tie: {it}"
        ties += int(it["tie"])
        assert it["dbox"] <= box_tol, it
        if it.get("same_int_box", True):
            # north_star: IoU >= 0.999; where the mask is soft (large areas with p ~ 0.5) the thresholded IoU is
            # ill-conditioned and the per-pixel bound + J against the ground truth take its place
            # (|dp| <= eps everywhere means every disagreeing pixel has an oracle probability within eps of 0.5; with
            # eps <= 0.01 the count of such pixels -- hence dJ -- is a property of the mask's softness, bounded looser)
            assert (it["iou"] >= 0.999 or (it["dprob_max"] <= 0.05 and it.get("dJ", 0.0) <= 1e-3)
                    or (it["dprob_max"] <= 0.01 and it.get("dJ", 0.0) <= 3e-3)), it
            assert it["dprob_mean"] <= 2e-3, it
        else:
            # the paste box truncates to other integers: the pasted mask is resampled one pixel larger / smaller, so
            # only pixels within 1 px of the mask boundary may change
            # (the raw IoU then drops by ~ perimeter / area: 0.95 .. 0.98 for these objects; printed, not bounded)
            flips += 1
            assert it["iou_off_boundary"] >= 0.995 and it.get("dJ", 0.0) <= 1e-2, it
    if require_det:
        assert any(it["n_det"][0] for it in infos), "no frame produced a detection"
    print(f"lock-step: {len(infos)} frames, {ties} tied detection choices, {flips} paste-box truncation flips")


def _video(seed, T, K=1, h=480, w=854):
    from eosvos_b200.util import synthetic
    frames, labels = synthetic.make_video(seed, T, h, w, K)
    fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
    return fr, torch.from_numpy(labels)


@statistical()
def test_run_frames_matches_oracle_run_loader():
    """run_loader (helper_func.py:67-159) over 7 frames at 854x480 after 30 fine-tune iterations, hook-free (own
    proposals, own detection), frame by frame from identical state; plus the two fallback branches: an all-zero start
    target (:90-93 -> EXTEND mode, no augmentation) and an empty prediction (:124-126 -> back to the start target)."""
    from eosvos_b200.util import evaluate as E
    model, opt, oracle, _, dev, _ = build_pair(min_size=None)
    fr, labels = _video(5, 8)
    gt0 = (labels[0] == 1).float()[None, None]
    model.roi_heads.detections_per_img = oracle.roi_heads.detections_per_img = 1
    inp, gts = fr[0:1].to(dev).repeat(3, 1, 1, 1), gt0.to(dev).repeat(3, 1, 1, 1)
    E.finetune(model, opt, lambda e: (inp, gts), 30, 1, 0)
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = 0.05
    ls = LockStep(model, opt, oracle, dev)
    ls.after_finetune(0, 0, model, opt)
    torch.manual_seed(77)
    ls.before_frames(0, 0, list(range(1, 8)), gt0)
    # 7 frames: with the default look-ahead a batched run of 5 and one of 2
    probs, boxes = E.run_frames(model, (fr[f:f + 1].to(dev) for f in range(1, 8)), gt0.to(dev),
                                on_frame=lambda i, t, p, b: ls.on_frame(0, 0, i, t, p, b))
    assert probs.shape == (7, 1, 480, 854)
    _check_lockstep(ls.replay(fr, labels))

    # (b) empty prediction -> the next frame runs with the START target and EXTEND mode (helper_func.py:124-126)
    model.rpn._eval_augment_proposals_mode = oracle.rpn._eval_augment_proposals_mode = 'REPLACE'
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = 0.9999
    used = []
    E.run_frames(model, (fr[f:f + 1].to(dev) for f in range(1, 4)), gt0.to(dev),
                 on_frame=lambda i, t, p, b: used.append((t.cpu().clone(), p.cpu().clone(), model.rpn._eval_augment_proposals_mode)))
    oused = []
    EO.run_loader(oracle, [fr[f:f + 1] for f in range(1, 4)], gt0,
                  on_frame=lambda i, t: oused.append((t.clone(), oracle.rpn._eval_augment_proposals_mode)) and None)
    for (t, p, mode), (ot, omode) in zip(used, oused):
        assert p.abs().sum() == 0 and torch.equal(t, ot)
    assert [u[2] for u in used] == ['EXTEND'] * 3 and model.rpn._eval_augment_proposals_mode == 'EXTEND'
    assert oracle.rpn._eval_augment_proposals_mode == 'EXTEND' and [o[1] for o in oused] == ['REPLACE', 'EXTEND', 'EXTEND']
    # (c) all-zero start target (:90-93): no target, EXTEND mode
    model.rpn._eval_augment_proposals_mode = 'REPLACE'
    used = []
    E.run_frames(model, (fr[f:f + 1].to(dev) for f in range(1, 3)), torch.zeros_like(gt0).to(dev),
                 on_frame=lambda i, t, p, b: used.append(t))
    assert used[0] is None and model.rpn._eval_augment_proposals_mode == 'EXTEND'
    model.rpn._eval_augment_proposals_mode = oracle.rpn._eval_augment_proposals_mode = 'EXTEND'


@pytest.mark.parametrize("case", ["cfg1", "ona"])
@statistical()
def test_evaluate_sequence_matches_oracle(case):
    """evaluate (evaluate.py:111-326) on BASELINE configs[0] -- e-OSVOS-10, batch 1, one 854x480 10-frame single-object
    video -- and on an online-adaptation schedule (e-OSVOS-12-OnA, step 3, 4 adaptation iterations, batch 3, 8 frames):
    the product runs free (own proposals, detections, propagation, online-adaptation pseudo labels); the oracle is
    replayed per round from the product's fine-tuned weights and per frame from the product's propagated target.
    Also checks the merged object-id masks against the oracle's merge of the product's probabilities, and J against
    the ground truth on both sides (|dJ| <= 1e-3 when no truncation flip occurred, <= 1e-2 otherwise)."""
    from eosvos_b200.util import evaluate as E
    model, opt, oracle, _, dev, _ = build_pair(min_size=None)
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = 0.05
    oracle.roi_heads.detections_per_img = 1
    if case == "cfg1":
        fr, labels = _video(3, 10)
        cfg = dict(num_epochs_eval=10, online_adapt_step=0, batch_size=1, random_train_transform=False)
    else:
        fr, labels = _video(4, 8)
        cfg = dict(num_epochs_eval=12, online_adapt_step=3, online_adapt_epochs=4, batch_size=3,
                   random_train_transform=True)
    state = copy.deepcopy(opt.state_dict())
    ls = LockStep(model, opt, oracle, dev)
    pred, stats = E.evaluate_sequence(model, opt, state, fr, labels[0], seed=1, hooks=ls.hooks(), **cfg)
    infos = ls.replay(fr, labels)
    assert len(infos) == fr.shape[0] - 1
    _check_lockstep(infos, require_det=False)
    # merge (evaluate.py:323-326) of the product's probabilities, restated by the oracle's threshold rule
    T = fr.shape[0]
    for f in range(1, T):
        want = O.threshold_targets(stats["masks"][f][None])[0, 0].to(torch.uint8)
        assert torch.equal(pred[f], want)
    assert torch.equal(pred[0], labels[0])
    print("train_loss_seq", stats["train_loss_seq"], "J", E.jaccard_per_object(pred, labels, 1))


# ---------------------------------------------------------------------------------------------- b: drop-in
def _ref_available():
    from oracle import ref_shims
    return ref_shims.available()


@pytest.mark.skipif(not _ref_available(), reason="unmodified reference copy (oracle/_ref) not present")
@statistical()
def test_reference_call_sites_run_on_product_classes():
    """INTEGRATION.md's claim, executed: the reference's OWN `evaluate` worker (oracle/_ref, unmodified:
    src/util/evaluate.py:20-439 with its data loaders, run_loader and fine-tune loop) runs with
    `eosvos_b200.MaskRCNN` / `MetaOptimizer` swapped in by rebinding the two imported names, on a synthetic DAVIS-2017
    tree, and is compared with the same worker on the reference's own classes on the same GPU (cuDNN / ATen), and
    with this repo's `evaluate` worker.  Trajectories diverge chaotically between ANY two arithmetics over 40 + 5 + 5
    fine-tuning iterations + propagation from a random initialisation (the tight, per-step comparisons are the
    lock-step tests above; run-to-run differences of ONE implementation are of the same size), so the end-to-end
    comparison is statistical: mean J over the objects within 0.2, the mean of the per-round final training losses
    within 50 %, and the structure of the results exactly."""
    import tempfile
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    from eosvos_b200.networks.mask_rcnn import MaskRCNN
    from eosvos_b200.util import evaluate as E
    from oracle import ref_harness as RH
    over = {"parent_model.train.val_split_files": [], "parent_model.val.val_split_files": [],
            "parent_model.test.val_split_files": [], "num_epochs.eval": 40, "eval_online_adapt.step": 3,
            "eval_online_adapt.num_epochs": 5, "parent_model.box_nms_thresh": 0.05}
    cfg = RH.compose_config(["DAVIS-2017", "e-OSVOS-OnA"], over)

    def small(name, obj):
        if name == "model":
            obj.transform.min_size, obj.transform.max_size = (320,), 534

    with tempfile.TemporaryDirectory() as wd:
        RH.make_davis_tree(wd, [("synth_a", 3, 7, 2), ("synth_b", 8, 6, 1)], split="val_seqs", height=192, width=342)
        ref_shared, ref_preds, _, _ = RH.run_reference_evaluate(cfg, "val", wd, "cuda:0", spy=small, save_dir="out_ref")
        own_shared, own_preds, _, cap = RH.run_reference_evaluate(cfg, "val", wd, "cuda:0", spy=small,
                                                                  model_cls=MaskRCNN, optim_cls=MetaOptimizer,
                                                                  save_dir="out_dropin")
        assert isinstance(cap["model"], MaskRCNN) and isinstance(cap["meta_optim"], MetaOptimizer)
        # this repo's worker, same signature and protocol
        with RH._cwd(wd), mock.patch("eosvos_b200.util.helper_func.init_parent_model", _small_factory(320, 534)):
            sd = copy.deepcopy(cap["meta_optim"].state_dict())
            shared = RH.OneShotSharedDict(meta_iter=None, best_mean_J=0.0)
            E.evaluate(0, "val", sd, {"meta_iter": 0, "meta_epoch": 0}, cfg, shared, "out_own", {}, True, RH._Log(),
                       once=True)
            wk_shared = dict(shared)
            assert os.path.exists(os.path.join("out_own", "best_eval_preds", "DAVIS-2017", "val_seqs", "synth_a", "00003.png"))
    keys = {"init_J_seq", "J_seq", "J_recall_seq", "J_decay_seq", "train_losses_seq", "train_loss_seq", "F_seq",
            "F_recall_seq", "F_decay_seq", "time_per_frame", "meta_iter", "best_mean_J"}
    for name, sh in (("reference", ref_shared), ("drop-in", own_shared), ("own worker", wk_shared)):
        print(name, "J", [round(j, 3) for j in sh["J_seq"]], "F", [round(j, 3) for j in sh["F_seq"]], "init_J",
              [round(j, 3) for j in sh["init_J_seq"]], "train_loss_seq", [round(l, 3) for l in sh["train_loss_seq"]],
              "time/frame", round(sh["time_per_frame"], 3))
        assert keys <= set(sh.keys()), keys - set(sh.keys())
        assert len(sh["J_seq"]) == 3 and len(sh["train_loss_seq"]) == len(ref_shared["train_loss_seq"])
        assert set(sh["train_losses_seq"][0].keys()) == set(ref_shared["train_losses_seq"][0].keys())
    for sh in (own_shared, wk_shared):
        assert abs(float(np.mean(sh["J_seq"])) - float(np.mean(ref_shared["J_seq"]))) <= 0.2, (sh["J_seq"], ref_shared["J_seq"])
        ma, mb = float(np.mean(sh["train_loss_seq"])), float(np.mean(ref_shared["train_loss_seq"]))
        assert abs(ma - mb) <= 0.5 * mb, (sh["train_loss_seq"], ref_shared["train_loss_seq"])
        assert all(np.isfinite(sh["train_loss_seq"]))
    for seq in ref_preds:
        assert own_preds[seq].shape == ref_preds[seq].shape
        assert np.array_equal(own_preds[seq][0], ref_preds[seq][0])          # frame 0 = the annotation on both sides


def _small_factory(min_size, max_size):
    from eosvos_b200.util import helper_func as HF
    real = HF.init_parent_model

    def f(**kw):
        model, states = real(**kw)
        model.transform.min_size, model.transform.max_size = (min_size,), max_size
        return model, states
    return f


def test_static_shape_pipeline_matches_list_pipeline():
    """The statically shaped proposal / sampling / detection kernels (csrc/rpn.cu: fast path) against the list-based
    restatement of the same torchvision logic, on the SAME head outputs (two forward passes differ by atomics-order
    rounding, which NMS amplifies): identical proposals, bit-identical RoI samples / labels / regression targets under
    the same sampler permutations, identical detection."""
    from eosvos_b200.util import evaluate as E
    from eosvos_b200 import kernels as K
    model, opt, oracle, _, dev, _ = build_pair(min_size=None)
    fr, labels = _video(6, 3)
    gt0 = (labels[0] == 1).float()[None, None]
    inp, gts = fr[0:1].to(dev).repeat(3, 1, 1, 1), gt0.to(dev).repeat(3, 1, 1, 1)
    model.roi_heads.detections_per_img = 1
    E.finetune(model, opt, lambda e: (inp, gts), 20, 1, 0)
    model.roi_heads.score_thresh = 0.05
    for train in (True, False):
        model.train_without_dropout() if train else model.eval()
        x = inp if train else fr[1:2].to(dev)
        B = x.shape[0]
        targets = model._build_targets(gts[:B] if train else gt0.to(dev), False, dev)
        model._prepare_operands()
        K.zero_pool.reset()
        with torch.no_grad():
            x8, targets_t, (oh, ow), (Hp, Wp) = model._transform(x, targets)
            feats = model._backbone(x8)
            feats, head_outs = feats[:5], feats[-5:]
            image_sizes, image_shape = [(oh, ow)] * B, (B, 3, Hp, Wp)
            post = model.rpn.post_nms_top_n()
            padded, count = model._rpn_fast(feats, image_shape, image_sizes, head_outs, post)
            fast_list = model._proposal_list(padded, count, post)
            objectness, deltas, feat_shapes, per_level = model._rpn_cat_outputs(feats, head_outs)
            anchors = model._anchors(image_shape, image_sizes, feat_shapes, dev)
            props = model._decode(deltas, anchors, model.rpn.box_coder).view(B, -1, 4)
            list_boxes, _ = model._filter_proposals(props, objectness, image_sizes, per_level)
            for a, b in zip(fast_list, list_boxes):
                assert a.shape == b.shape, (a.shape, b.shape)
                assert (a - b).abs().max().item() <= 2e-3
            if train:
                with mock.patch("torch.randperm", det_randperm(5)):
                    smp = model._sample_rois_fast(model._match_rois_fast(padded, count, targets_t))
                with mock.patch("torch.randperm", det_randperm(5)):
                    p_l, m_l, l_l, r_l, pos_l = model._select_training_samples(list_boxes, targets_t)
                S = 512
                for i in range(B):
                    sl = slice(i * S, (i + 1) * S)
                    assert (smp["rois5"][sl, 1:] - p_l[i]).abs().max().item() <= 2e-3
                    assert torch.equal(smp["labels"][sl], l_l[i]) and torch.equal(smp["matched"][sl], m_l[i])
                    assert torch.allclose(smp["reg"][sl], r_l[i], rtol=1e-4, atol=1e-4)
                    pos_i = smp["pos"][(smp["pos"] >= i * S) & (smp["pos"] < (i + 1) * S)]
                    assert torch.equal(pos_i - i * S, pos_l[i])
            else:
                # detection: arg-max kernel vs the list-based postprocess on the same box-head output
                rois5 = torch.cat([torch.zeros(padded.shape[1], 1, device=dev), padded[0]], 1)
                head = model._box_branch(feats[:4], rois5)
                det = K.det_top1(head, padded.view(-1, 4), 1, padded.shape[1], 2, model.roi_heads.box_coder.weights,
                                 model.roi_heads.box_coder.bbox_xform_clip, 0.05, 1e-2, float(ow), float(oh), 1.0, 1.0)
                bl, sl_, ll = model._postprocess_detections(head[:, :2], head[:, 2:10], [padded[0]], image_sizes)
                assert bl[0].shape[0] == 1 and int(det["row"][0]) == int(model.last_detection_rows[0][0])
                assert (det["box"][0] - bl[0][0]).abs().max().item() <= 1e-3
                assert abs(float(det["score"][0]) - float(sl_[0][0])) <= 1e-6


@statistical()
def test_sparse_rpn_backward_matches_dense(monkeypatch):
    """The sparse backward of the RPN head (ops.RpnSparseFn: sampled anchors only) against the dense autograd path
    (conv dgrad / wgrad over every level) from the same weights, batch and sampler permutations.  The RPN losses must
    agree to rounding; the gradients of the RPN head's own parameters to 2e-2.  Deeper gradients pass through ~50
    16-bit layers whose atomics-ordered reductions make even two runs of the SAME path differ, so they are held to
    three times that run-to-run noise floor (measured alongside) plus 2e-2."""
    from eosvos_b200.util import evaluate as E
    model, opt, oracle, _, dev, _ = build_pair(min_size=None)
    fr, labels = _video(6, 2)
    gt0 = (labels[0] == 1).float()[None, None]
    inp, gts = fr[0:1].to(dev).repeat(3, 1, 1, 1), gt0.to(dev).repeat(3, 1, 1, 1)
    E.finetune(model, opt, lambda e: (inp, gts), 10, 1, 0)
    names = [f"{a}.{c}" for a, _, c, _ in opt.meta_model.param_groups()]
    out = {}
    for tag, mode in (("sparse", "1"), ("dense", "0"), ("dense2", "0")):
        monkeypatch.setenv("EOSVOS_RPN_SPARSE", mode)
        if tag != "dense2":
            model._graphs.clear()
        model.train_without_dropout()
        with mock.patch("torch.randperm", det_randperm(5)):
            torch.manual_seed(21)
            loss, losses = model(inp, gts)
        rpn_loss = losses["loss_objectness"] + losses["loss_rpn_box_reg"]
        params = [p for *_, p in opt.meta_model.param_groups()]
        grads = torch.autograd.grad(rpn_loss, params, allow_unused=True)
        out[tag] = ({k: v.item() for k, v in losses.items()},
                    {n: (g.detach().float().clone() if g is not None else None) for n, g in zip(names, grads)})
    monkeypatch.delenv("EOSVOS_RPN_SPARSE")
    model._graphs.clear()
    (la, ga), (lb, gb), (_, gc) = out["sparse"], out["dense"], out["dense2"]
    for k in ("loss_objectness", "loss_rpn_box_reg"):
        assert abs(la[k] - lb[k]) <= 2e-3 * abs(lb[k]) + 1e-6, (k, la[k], lb[k])

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-20)).item()
    groups = {}
    for n in names:
        a, b, c = ga[n], gb[n], gc[n]
        if b is None or b.abs().max() == 0:
            assert a is None or a.abs().max().item() < 1e-6, n
            continue
        assert a is not None, n
        key = "rpn" if n.startswith("rpn.") else (n.split(".")[1] + "." + n.split(".")[2] if n.startswith("backbone.body.layer") else n.split(".")[1] if n.startswith("backbone.body") else "fpn")
        g = groups.setdefault(key, [0.0, 0.0])
        g[0], g[1] = max(g[0], rel(a, b)), max(g[1], rel(c, b))
    for k, (d, floor) in groups.items():
        print(f"{k:12s} sparse-vs-dense {d:.4f}   dense-vs-dense (noise floor) {floor:.4f}")
        assert d <= (2e-2 if k == "rpn" else 3 * floor + 2e-2), (k, d, floor)


@statistical()
def test_youtube_vos_shaped_sequence_with_late_object():
    """BASELINE configs[3] semantics on the GPU: a 1280x720 YouTube-VOS-shaped sequence (-> 1333x749, padded to
    1344x768: the trunk shapes of the DAVIS case) with a second object whose first annotation is a LATER frame
    (src/data/youtube.py:131-185; evaluate.py:143-146,156-170).  The product's evaluate_sequence runs free; the oracle
    is replayed in lock-step per round / frame.  Checks the per-object schedules (the late object is fine-tuned on its
    own first frame and only predicted after it) and the reference's quirk that the late object's propagation starts
    without a target (the test loader's label file is the sequence's first one: empty for that object)."""
    from eosvos_b200.util import davis_io
    from eosvos_b200.util import evaluate as E
    from oracle import ref_harness as RH
    model, opt, oracle, _, dev, _ = build_pair(min_size=None)
    model.roi_heads.score_thresh = oracle.roi_heads.score_thresh = 0.05
    oracle.roi_heads.detections_per_img = 1
    T, appear = 6, [0, 2]
    frames, labels = RH.synthetic_video(21, T, 720, 1280, 2, appear=appear)
    fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
    lab = torch.from_numpy(labels)
    annotated = [i in set(appear) for i in range(T)]
    seq = davis_io.VideoSequence(fr, lab * torch.tensor(annotated)[:, None, None].to(lab.dtype), annotated=annotated,
                                 objects=[(1, 0), (2, 2)], test_mode=True)
    assert seq.num_objects == 2 and seq.gt_frame_id(1) == (2, 1)
    assert float(seq.label(2, 1, None).sum()) == 0 and float(seq.label(2, 1, 1).sum()) > 0
    state = copy.deepcopy(opt.state_dict())
    ls = LockStep(model, opt, oracle, dev)
    timers = {}
    pred, stats = E.evaluate_sequence(model, opt, state, seq, num_epochs_eval=10, online_adapt_step=2, online_adapt_epochs=3,
                                      batch_size=3, seed=1, random_train_transform=True, hooks=ls.hooks(), timers=timers)
    # object 0: rounds over frames [1,3), [3,5), [5,6); object 1 (first annotated at 2): [3,5), [5,6)
    rounds = sorted((o, k, tuple(r["frame_ids"])) for (o, k), r in ls.rounds.items())
    assert rounds == [(0, 0, (1, 2)), (0, 1, (3, 4)), (0, 2, (5,)), (1, 0, (3, 4)), (1, 1, (5,))], rounds
    assert timers["finetune_iters"] == (10 + 3 + 3) + (10 + 3) and timers["infer_frames"] == 5 + 3
    assert ls.rounds[(1, 0)]["frames"][0]["target"] is None            # no start target for the late object
    # the late object is never predicted before its first annotation -- and, as in the reference, not ON it either:
    # `masks[frame][obj] = 2 * train_frame_gt` uses the test loader's (empty) label (evaluate.py:156-168)
    assert (pred[:3] == 2).sum() == 0
    infos = ls.replay(fr, lab)
    _check_lockstep(infos, require_det=False)
    assert pred.shape == (T, 720, 1280)
