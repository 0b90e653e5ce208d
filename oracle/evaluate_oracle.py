"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's per-sequence evaluation on IN-MEMORY frames:
  * `OracleSequence`    -- what one sequence of `data.DAVIS` / `data.YouTube` hands out under
                           multi_object='single_id' (src/data/vos_dataset.py:193-339, src/data/davis.py:30-66,
                           src/data/youtube.py:107-185): RGB float32 / 255, per-object binary label, the YouTube-VOS
                           rules for objects that are first annotated in a later frame
  * `train_batch`       -- RandomHorizontalFlip + RandomScaleNRotate + EpochSampler batch
                           (src/data/custom_transforms.py:9-89,188-211, src/util/helper_func.py:254-294,521-545)
  * `run_loader`        -- src/util/helper_func.py:67-159 (MaskRCNN branch)
  * `evaluate_sequence` -- src/util/evaluate.py:111-326: rounds of fine-tuning (HOT LOOP A :220-280), online
                           adaptation batches (:172-253), FIRST_STEP restore (:196-205,283-287), frame ranges
                           (:294-314), object merge (:323-326)
It drives any model / optimizer pair with the reference's API (the oracle classes, the reference's own classes, or the
product's), so the same loop serves as restatement and as lock-step checker.

Random streams.  The reference iterates torch DataLoaders; every `iter(loader)` draws one int64 from torch's default
generator (`_BaseDataLoaderIter.__init__`: `_base_seed`) and every pass over a `RandomSampler` draws another
(`RandomSampler.__iter__`).  `loader_draws` reproduces that consumption so that -- on the CPU, where the model's own
`torch.randperm` / `torch.rand` calls use the same generator -- the restatement follows the reference bit for bit.
Pinned by tests/test_oracle_pins.py::test_evaluate_sequence_matches_reference against tests/golden/evaluate_*.pt
(written by oracle/make_golden.py from the unmodified reference's `evaluate()` worker).
"""
import copy
import random

import cv2
import numpy as np
import torch


def set_random_seeds(seed):
    """helper_func.py:515-518"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def loader_draws(n):
    for _ in range(n):
        torch.empty((), dtype=torch.int64).random_()


class OracleSequence:
    """One sequence in memory.  frames uint8 [T,H,W,3] RGB; labels uint8 [T,H,W] object ids; `annotated` [T] bool
    (frames that have a label file); `objects`: None for DAVIS (ids 1..K present in the first annotation) or, for
    YouTube-VOS, a list of (label id, first annotated frame) sorted by id (youtube.py:125-143)."""

    def __init__(self, frames, labels, annotated=None, objects=None, test_mode=False):
        self.frames = np.asarray(frames)
        self.labels = np.asarray(labels)
        T = self.frames.shape[0]
        self.annotated = np.ones(T, bool) if annotated is None else np.asarray(annotated, bool)
        self.label_frames = [i for i in range(T) if self.annotated[i]]      # == sorted label files
        self.objects = objects
        self.test_mode = test_mode or objects is not None and not self.annotated.all()
        if objects is None:
            self.num_objects = len([l for l in np.unique(self.labels[self.label_frames[0]]) if l != 0])
        else:
            self.num_objects = len(objects)

    def __len__(self):
        return self.frames.shape[0]

    def image(self, idx):
        """vos_dataset.py:232-233,276-279 with normalize False: float32 RGB / 255, [H,W,3]."""
        return self.frames[idx].astype(np.float32) / 255.0

    def gt_frame_id(self, obj):
        """(frame id, label index) of the object's first annotation: DAVIS frame 0 (vos_dataset.py:193-194);
        YouTube-VOS youtube.py:131-143."""
        if self.objects is None:
            return 0, None
        f = self.objects[obj][1]
        return f, self.label_frames.index(f)

    def label(self, idx, obj, label_idx=None):
        """vos_dataset.py:236-245 (which label file), :288-339 (single-id selection)."""
        if label_idx is not None:
            raw = self.labels[self.label_frames[label_idx]]
        elif self.test_mode:
            raw = self.labels[self.label_frames[0]]
        else:
            raw = self.labels[idx]
        label = raw.astype(np.float32)
        if self.num_objects > 1:
            uniq = [l for l in np.unique(label) if l != 0.0]
            if uniq:
                want = float(obj + 1) if self.objects is None else float(self.objects[obj][0])
                if want in uniq:
                    return (label == want).astype(np.float32)
                return np.zeros(label.shape, np.float32)
            return label
        return np.where(label != 0.0, 1.0, 0.0).astype(np.float32)


def _rot_and_sc(a, rot, sc, label):
    """custom_transforms.py:40-51"""
    h, w = a.shape[:2]
    M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
    return cv2.warpAffine(a, M, (w, h), flags=cv2.INTER_NEAREST if label else cv2.INTER_CUBIC)


def augment(image, gt, rots=(-30, 30), scales=(.75, 1.25)):
    """RandomHorizontalFlip (custom_transforms.py:188-211) then RandomScaleNRotate (:53-89), `random` module."""
    if random.random() < 0.5:
        image, gt = cv2.flip(image, flipCode=1), cv2.flip(gt, flipCode=1)
    num_labels = len(np.unique(gt))
    while True:
        rot = (rots[1] - rots[0]) * random.random() - (rots[1] - rots[0]) / 2
        sc = (scales[1] - scales[0]) * random.random() - (scales[1] - scales[0]) / 2 + 1
        aug_gt = _rot_and_sc(gt, rot, sc, True)
        if not num_labels > 1 or len(np.unique(aug_gt)) == num_labels:
            break
    return _rot_and_sc(image, rot, sc, False), aug_gt


def _to_tensor(a):
    """custom_transforms.py:254-272"""
    if a.ndim == 2:
        a = a[:, :, np.newaxis]
    return torch.from_numpy(np.ascontiguousarray(a.transpose((2, 0, 1))))


def train_batch(seq, frame_id, obj, label_idx, batch_size, random_transform, propagate_gt=None):
    """One pass over the reference's train loader: EpochSampler puts `batch_size` visits of the ONE train frame into
    a batch (helper_func.py:521-545), each visit running the dataset transform.  -> (inputs [B,3,H,W], gts [B,1,H,W])."""
    loader_draws(1 + batch_size)          # iter(DataLoader) + one RandomSampler pass per visit
    imgs, gts = [], []
    for _ in range(batch_size):
        img = seq.image(frame_id)
        gt = propagate_gt if propagate_gt is not None else seq.label(frame_id, obj, label_idx)
        if random_transform:
            img, gt = augment(img, gt)
        imgs.append(_to_tensor(img))
        gts.append(_to_tensor(gt))
    return torch.stack(imgs), torch.stack(gts)


def run_loader(model, frames, start_targets, on_frame=None):
    """helper_func.py:67-159, MaskRCNN branch.  frames: list of [1,3,H,W] tensors on the model's device.
    on_frame(i, targets) -> optional replacement targets (lock-step checking: the caller may feed the frame another
    implementation's propagated target).  -> (probs [F,K,H,W], boxes [F,K,4])."""
    mode = model.rpn._eval_augment_proposals_mode
    targets = None
    if mode is not None:
        if start_targets is None:
            targets = start_targets
        elif start_targets.sum().item() == 0:
            start_targets = None
            model.rpn._eval_augment_proposals_mode = 'EXTEND'
            targets = start_targets
        else:
            targets = start_targets.clone()
    probs_all, boxes_all = [], []
    loader_draws(1)
    with torch.no_grad():
        for i, inputs in enumerate(frames):
            model.eval()
            if on_frame is not None:
                repl = on_frame(i, targets)
                if repl is not None:
                    targets = repl
            probs, boxes = model(inputs, targets)
            if mode is not None:
                bg = probs.max(dim=1, keepdim=True)[0].lt(0.5)
                targets = probs.argmax(dim=1, keepdim=True).float() + 1.0
                targets[bg] = 0.0
                model.rpn._eval_augment_proposals_mode = mode
                if targets.sum().item() == 0:
                    model.rpn._eval_augment_proposals_mode = 'EXTEND'
                    targets = start_targets
            probs_all.append(probs)
            boxes_all.append(boxes)
    return torch.cat(probs_all), torch.cat(boxes_all)


DEFAULT_CFG = dict(seed=1, num_epochs_eval=10, step=0, ona_epochs=10, min_prop=0.5, batch_size=1,
                   random_train_transform=False, reset_model_mode='FIRST_STEP', patience=None, min_loss_improv=1e-3)


def early_stopping(loss_hist, patience, min_loss_improv):
    """helper_func.py:388-398"""
    if patience is None or len(loss_hist) <= patience:
        return False
    best_loss = torch.tensor(loss_hist).min()
    prev_best_loss = torch.tensor(loss_hist[:-patience]).min()
    return not bool(torch.gt(best_loss.sub(prev_best_loss).abs(), min_loss_improv))


def evaluate_sequence(model, meta_optim, meta_optim_state_dict, seq, device="cpu", hooks=None, **cfg):
    """evaluate.py:111-326 for ONE sequence, every object (multi_object 'single_id': evaluate.py:106-107,132).
    hooks (all optional, for lock-step checks):
        after_finetune(obj, k, model, meta_optim)        -- e.g. overwrite the weights with another implementation's
        before_frames(obj, k, frame_ids, start_target)   -- e.g. record / set the RNG state
        on_frame(obj, k, i, targets) -> targets or None  -- see run_loader
    -> (pred uint8 [T,H,W], record) with record = {"masks": [T] of [K,H,W] probs, "boxes", "train_loss_seq",
       "rounds": [(obj, k, iters, frame_min, frame_max)]}."""
    c = dict(DEFAULT_CFG)
    c.update(cfg)
    hooks = hooks or {}
    T = len(seq)
    H, W = seq.frames.shape[1:3]
    model.roi_heads.detections_per_img = 1
    masks, boxes = [], [None] * T
    rec = {"train_loss_seq": [], "rounds": [], "train_losses_seq": []}
    for obj in range(seq.num_objects):
        train_frame_id, label_idx = seq.gt_frame_id(obj)
        # the TEST loader's view of the train frame (its own `_label_id` is never set: vos_dataset.py:236-245)
        train_frame_gt = _to_tensor(seq.label(train_frame_id, obj, None))
        if c["step"]:
            step = c["step"]
            meta_frame_iter = range(train_frame_id + 1, T, step)
        else:
            step = T
            meta_frame_iter = [None]
        range_min = range_max = None
        state_first = None
        for k, _ in enumerate(meta_frame_iter):
            if k == 0:
                for f in range(T):
                    z = torch.zeros(1, H, W)
                    if not obj:
                        masks.append(z)
                    else:
                        masks[f] = torch.cat([masks[f], z])
                masks[train_frame_id][obj, :, :] = 2 * train_frame_gt
                range_min = train_frame_id + 1
                range_max = range_min
            else:
                range_min = range_max
                propagate_frame_gt = masks[range_min - 1][obj:obj + 1].ge(c["min_prop"]).float()
                propagate_frame_gts = []
                for pid in range(1, c["step"]):
                    g = masks[range_min - pid][obj:obj + 1].ge(c["min_prop"]).float()
                    propagate_frame_gts.append(np.copy(np.transpose(g.cpu().numpy(), (1, 2, 0))))
            range_max += step
            if range_max > T:
                range_max = T

            if k == 0 or c["reset_model_mode"] == 'FULL':
                meta_optim.load_state_dict(meta_optim_state_dict)
                meta_optim.reset()
                meta_optim.eval()
            elif c["reset_model_mode"] == 'FIRST_STEP':
                meta_optim.load_state_dict(meta_optim_state_dict)
                model.load_state_dict(state_first)
                meta_optim.eval()

            iters = c["num_epochs_eval"] if k == 0 else c["ona_epochs"]
            model.train_without_dropout()
            hist = []
            train_loss = None
            for epoch in range(1, iters + 1):
                set_random_seeds(c["seed"] + epoch + k)
                inputs, gts = train_batch(seq, train_frame_id, obj, label_idx, c["batch_size"],
                                          c["random_train_transform"] and k == 0)
                if k:
                    inputs, gts = inputs[:1], gts[:1]
                    n_prop = min(c["step"], c["batch_size"])
                    for pid in range(c["step"] - n_prop + 1, c["step"]):
                        pg = propagate_frame_gts[pid - 1]
                        if (pg == 1.0).astype(float).sum().item() != 0:
                            pi, pgts = train_batch(seq, range_min - pid, obj, label_idx, c["batch_size"], False,
                                                   propagate_gt=pg)
                            inputs = torch.cat([inputs, pi[:1]])
                            gts = torch.cat([gts, pgts[:1]])
                inputs, gts = inputs.to(device), gts.to(device)
                train_loss, train_losses = model(inputs, gts)
                hist.append(train_loss.item())
                model.zero_grad()
                meta_optim.set_train_loss(train_loss)
                if c["reset_model_mode"] == 'FIRST_STEP':
                    meta_optim.only_box_head = k != 0
                meta_optim.step(train_loss)
                meta_optim.meta_model.detach_param_groups()
                if early_stopping(hist, c["patience"], c["min_loss_improv"]):
                    break
            rec["train_loss_seq"].append(train_loss.item())
            rec["train_losses_seq"].append({n: v.item() for n, v in train_losses.items()})
            if "after_finetune" in hooks:
                hooks["after_finetune"](obj, k, model, meta_optim)
            if k == 0:
                state_first = copy.deepcopy(model.state_dict())

            frame_ids = list(range(range_min, range_max))
            targets = (train_frame_gt if k == 0 else propagate_frame_gt).unsqueeze(dim=0)
            if "before_frames" in hooks:
                hooks["before_frames"](obj, k, frame_ids, targets)
            frames = [_to_tensor(seq.image(f))[None].to(device) for f in frame_ids]
            on_frame = (lambda i, t, obj=obj, k=k: hooks["on_frame"](obj, k, i, t)) if "on_frame" in hooks else None
            probs_r, boxes_r = run_loader(model, frames, targets.to(device), on_frame)
            probs_r, boxes_r = probs_r.cpu(), boxes_r.cpu()
            for f, p, b in zip(frame_ids, probs_r, boxes_r):
                boxes[f] = b if boxes[f] is None else torch.cat([boxes[f], b])
                masks[f][-1:, :, :] = p
            rec["rounds"].append((obj, k, iters, range_min, range_max))
            if range_max == T:
                break
    rec["masks"] = [m.clone() for m in masks]
    rec["boxes"] = boxes
    pred = torch.zeros(T, H, W, dtype=torch.uint8)
    for f in range(T):                                   # evaluate.py:323-326
        bg = masks[f].max(dim=0, keepdim=True)[0].lt(0.5)
        m = masks[f].argmax(dim=0, keepdim=True).float() + 1.0
        m[bg] = 0.0
        pred[f] = m[0].to(torch.uint8)
    return pred, rec
