"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's model path for boxes that have no /root/reference checkout:
  * OracleMaskRCNN  -- reference src/networks/mask_rcnn.py:423-775 (constructor, GN swap, forward),
                       :217-344 (rpn_forward incl. EXTEND/REPLACE proposal augmentation),
                       :95-214 (roi_heads_forward), :347-420 (postprocess_detections), :24-92 (mask losses)
  * OracleMetaOptimizer -- reference src/meta_optim/meta_optim.py:10-214 + meta_model.py:49-80
  * run_frames      -- reference src/util/helper_func.py:67-159 (run_loader's propagation loop)
The arithmetic itself is torch / torchvision CPU fp32 -- the very operators the reference calls
(README.md:23; torchvision is an un-vendored dependency).  Pinned against the real
reference in tests/test_oracle_pins.py via tests/golden/ (made by oracle/make_golden.py).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torchvision.models.detection import MaskRCNN as TvMaskRCNN
from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
from torchvision.models.detection.roi_heads import fastrcnn_loss, maskrcnn_inference
from torchvision.models.detection.rpn import concat_box_prediction_layers
from torchvision.ops import MultiScaleRoIAlign
from torchvision.ops import boxes as box_ops
from torchvision.ops.misc import FrozenBatchNorm2d

from . import ops_oracle as O


class OracleMaskRCNN(TvMaskRCNN):
    def __init__(self, backbone="resnet50", num_classes=2, roi_pool_output_sizes=None,
                 eval_augment_rpn_proposals_mode="EXTEND", replace_batch_with_group_norms=True, box_nms_thresh=0.5,
                 maskrcnn_loss="LOVASZ"):
        roi_pool_output_sizes = roi_pool_output_sizes or {"box": 7, "mask": 28}
        bb = resnet_fpn_backbone(backbone_name=backbone, weights=None, trainable_layers=5)
        box_pool = MultiScaleRoIAlign(["0", "1", "2", "3"], roi_pool_output_sizes["box"], 2)
        mask_pool = MultiScaleRoIAlign(["0", "1", "2", "3"], roi_pool_output_sizes["mask"], 2)
        super().__init__(bb, num_classes, box_roi_pool=box_pool, mask_roi_pool=mask_pool, mask_head=None,
                         box_score_thresh=box_nms_thresh)
        self.num_classes = num_classes
        # same attributes the reference's callers poke from outside (helper_func.py:72,92,123-125)
        self.rpn._eval_augment_proposals_mode = eval_augment_rpn_proposals_mode
        self.roi_heads._eval_augment_proposals_mode = eval_augment_rpn_proposals_mode
        self.loss_kind = maskrcnn_loss
        if replace_batch_with_group_norms:            # mask_rcnn.py:523-534
            for module in self.modules():
                for k, m in list(module._modules.items()):
                    if isinstance(m, (FrozenBatchNorm2d, nn.BatchNorm2d)):
                        gn = nn.GroupNorm(32, m.weight.shape[0])
                        gn.weight.data = m.weight
                        gn.bias.data = m.bias
                        module._modules[k] = gn
        self.backbone.requires_grad_(True)            # mask_rcnn.py:493-494
        self.fixed_proposals = None                   # test hook: bypass RPN proposals with given boxes
        self.fixed_detection_rows = None              # test hook: take THESE proposal rows as the detections

    @property
    def aug_mode(self):
        return self.rpn._eval_augment_proposals_mode

    @aug_mode.setter
    def aug_mode(self, v):
        self.rpn._eval_augment_proposals_mode = v

    # ---- mask_rcnn.py:582-714
    @staticmethod
    def build_targets(targets, flip_label=False):
        out = []
        if flip_label:
            targets = 1 - targets
        for mask in targets:
            ids = [v.item() for v in torch.unique(mask) if v.item() != 0.0 and v.item() != 255.0]
            obj_ids = torch.tensor(ids)
            masks = mask == obj_ids[:, None, None]
            masks[(mask == 255.0).expand_as(masks)] = True
            assert len(ids) >= 1, f"num_objs: {len(ids)}"
            boxes = []
            for i in range(len(ids)):
                pos = np.where(masks[i].numpy())
                boxes.append([np.min(pos[1]), np.min(pos[0]), np.max(pos[1]) + 1, np.max(pos[0]) + 1])
            boxes = torch.as_tensor(boxes, dtype=torch.float32)
            masks = masks.type(torch.uint8)
            if (mask == 255.0).any():
                masks[(mask == 255.0).expand_as(masks)] = 255
                masks[(mask == 0.0).expand_as(masks)] = 255
            if flip_label:
                masks = 1 - masks
            out.append({"boxes": boxes, "labels": obj_ids.type(torch.int64), "masks": masks})
        return out

    # ---- mask_rcnn.py:217-344
    def rpn_stage(self, images, features, targets):
        rpn = self.rpn
        feats = list(features.values())
        objectness, deltas = rpn.head(feats)
        anchors = rpn.anchor_generator(images, feats)
        n_img = len(anchors)
        per_level = [o[0].numel() for o in objectness]
        objectness, deltas = concat_box_prediction_layers(objectness, deltas)
        proposals = rpn.box_coder.decode(deltas.detach(), anchors).view(n_img, -1, 4)
        boxes, scores = rpn.filter_proposals(proposals, objectness, images.image_sizes, per_level)
        if not self.training and targets is not None and self.aug_mode is not None:
            share = 0.1
            post = rpn.post_nms_top_n()
            n_aug = post // 2 if self.aug_mode == "EXTEND" else post
            for i, target in enumerate(targets):
                ih, iw = images.tensors[i].shape[-2:]
                extra = []
                for box in target["boxes"]:
                    bw, bh = box[2] - box[0], box[3] - box[1]
                    x0 = box[0] - torch.rand((n_aug,)) * bw * share
                    y0 = box[1] - torch.rand((n_aug,)) * bh * share
                    x1 = box[2] + torch.rand((n_aug,)) * bw * share
                    y1 = box[3] + torch.rand((n_aug,)) * bh * share
                    extra.append(torch.stack([x0.clamp(0, iw), y0.clamp(0, ih), x1.clamp(0, iw), y1.clamp(0, ih)], 1))
                extra = torch.cat(extra, 0)
                if self.aug_mode == "EXTEND":
                    boxes[i] = torch.cat([boxes[i][:post // 2], extra], 0)
                elif self.aug_mode == "REPLACE":
                    boxes[i] = extra
                else:
                    raise NotImplementedError
        losses = {}
        if self.training:
            labels, matched = rpn.assign_targets_to_anchors(anchors, targets)
            reg = rpn.box_coder.encode(matched, anchors)
            lo, lb = rpn.compute_loss(objectness, deltas, labels, reg)
            losses = {"loss_objectness": lo, "loss_rpn_box_reg": lb}
        self.last_rpn_raw = (objectness, deltas)
        return boxes, losses

    # ---- mask_rcnn.py:347-420
    def detections_stage(self, class_logits, box_regression, proposals, image_shapes):
        rh = self.roi_heads
        ncls = class_logits.shape[-1]
        per_img = [len(p) for p in proposals]
        pred_boxes = rh.box_coder.decode(box_regression, proposals).split(per_img, 0)
        pred_scores = F.softmax(class_logits, -1).split(per_img, 0)
        res = []
        for boxes, scores, shape in zip(pred_boxes, pred_scores, image_shapes):
            boxes = box_ops.clip_boxes_to_image(boxes, shape)
            labels = torch.arange(ncls).view(1, -1).expand_as(scores)
            boxes, scores, labels = boxes[:, 1:].reshape(-1, 4), scores[:, 1:].flatten(), labels[:, 1:].flatten()
            all_boxes, all_scores = boxes, scores           # one candidate per (proposal row, foreground class)
            inds = torch.nonzero(scores > rh.score_thresh).squeeze(1)
            boxes, scores, labels = boxes[inds], scores[inds], labels[inds]
            keep = box_ops.remove_small_boxes(boxes, min_size=1e-2)
            boxes, scores, labels, inds = boxes[keep], scores[keep], labels[keep], inds[keep]
            keep = box_ops.batched_nms(boxes, scores, labels, rh.nms_thresh)[:rh.detections_per_img]
            rows = inds[keep]
            if self.fixed_detection_rows is not None:      # test hook: the caller's choice among the candidates
                rows = self.fixed_detection_rows[len(res)].to(torch.int64)
                lab_all = torch.arange(ncls).view(1, -1).expand(len(all_scores) // (ncls - 1), ncls)[:, 1:].flatten()
                res.append(dict(boxes=all_boxes[rows], scores=all_scores[rows], labels=lab_all[rows]))
            else:
                res.append(dict(boxes=boxes[keep], scores=scores[keep], labels=labels[keep]))
            self.last_candidates = getattr(self, "last_candidates", [])
            self.last_candidates.append(dict(boxes=all_boxes, scores=all_scores, rows=rows))
        return res

    # ---- mask_rcnn.py:95-214
    def roi_stage(self, features, proposals, image_shapes, targets):
        rh = self.roi_heads
        if self.training:
            proposals, matched_idxs, labels, reg_targets = rh.select_training_samples(proposals, targets)
        self.last_sampled_proposals = [p.detach().clone() for p in proposals]
        bf = rh.box_head(rh.box_roi_pool(features, proposals, image_shapes))
        class_logits, box_regression = rh.box_predictor(bf)
        self.last_box_raw = (class_logits, box_regression)
        result, losses = [], {}
        if self.training:
            lc, lb = fastrcnn_loss(class_logits, box_regression, labels, reg_targets)
            losses = dict(loss_classifier=lc, loss_box_reg=lb)
            mask_props, pos_idx = [], []
            for i in range(len(proposals)):
                pos = torch.nonzero(labels[i] > 0).squeeze(1)
                mask_props.append(proposals[i][pos])
                pos_idx.append(matched_idxs[i][pos])
        else:
            self.last_candidates = []
            result = self.detections_stage(class_logits, box_regression, proposals, image_shapes)
            self.last_detections = [{k: v.detach().clone() for k, v in r.items()} for r in result]
            mask_props = [r["boxes"] for r in result]
        mf = rh.mask_roi_pool(features, mask_props, image_shapes)
        if mf.shape[0] > 0:
            mf = rh.mask_head(mf)
        elif len(mask_props) > 1:
            raise NotImplementedError
        mask_logits = rh.mask_predictor(mf)
        self.last_mask_logits = mask_logits
        if self.training:
            gt_masks = [t["masks"] for t in targets]
            gt_labels = [t["labels"] for t in targets]
            lab = torch.cat([l[i] for l, i in zip(gt_labels, pos_idx)], 0)
            M = mask_logits.shape[-1]
            tg = torch.cat([O.mask_targets(m, torch.cat([i[:, None].to(p), p], 1), M)
                            for m, p, i in zip(gt_masks, mask_props, pos_idx)], 0)
            self.last_mask_targets = (lab, tg, [p.detach().clone() for p in mask_props], pos_idx)
            losses["loss_mask"] = O.mask_loss(mask_logits, lab, tg, self.loss_kind)
        else:
            probs = maskrcnn_inference(mask_logits, [r["labels"] for r in result])
            for p, r in zip(probs, result):
                r["masks"] = p
        return result, losses

    # ---- mask_rcnn.py:572-775 (+ tv generalized_rcnn.py forward)
    def forward(self, inputs, targets=None, box_coord_perm=None, flip_label=False):
        tdicts = self.build_targets(targets, flip_label) if targets is not None else None
        sizes = [tuple(i.shape[-2:]) for i in inputs]
        images, tt = self.transform([i for i in inputs], tdicts)
        self.last_images = images
        features = self.backbone(images.tensors)
        self.last_features = features
        proposals, rpn_losses = self.rpn_stage(images, features, tt)
        if self.fixed_proposals is not None:
            proposals = [p.clone() for p in self.fixed_proposals]
        self.last_proposals = [p.detach().clone() for p in proposals]
        detections, det_losses = self.roi_stage(features, proposals, images.image_sizes, tt)
        detections = self.transform.postprocess(detections, images.image_sizes, sizes)
        if self.training:
            raw = {}
            raw.update(det_losses)
            raw.update(rpn_losses)
            losses = {k: v for k, v in raw.items() if v.requires_grad}
            return sum(losses.values()), losses
        masks_out, boxes_out = [], []
        for det in detections:
            ms, bs = [], []
            for c in range(1, self.num_classes):
                hit = (det["labels"] == c).nonzero()
                if len(hit):
                    ms.append(det["masks"][hit[0]][0])
                    bs.append(det["boxes"][hit[0]])
                else:
                    ms.append(torch.zeros_like(inputs)[0, 0].unsqueeze(0))
                    bs.append(torch.zeros(1, 4))
            masks_out.append(torch.cat(ms, 0).unsqueeze(0))
            boxes_out.append(torch.cat(bs, 0).unsqueeze(0))
        return torch.cat(masks_out, 0), torch.cat(boxes_out, 0)

    def train_without_dropout(self):
        self.train()


def build_oracle_model(seed=1, maskrcnn_loss="LOVASZ", min_size=None, max_size=None):
    torch.manual_seed(seed)
    m = OracleMaskRCNN(maskrcnn_loss=maskrcnn_loss)
    if min_size is not None:
        m.transform.min_size = (min_size,)
        m.transform.max_size = max_size
    return m


class OracleMetaOptimizer:
    """NEURON-level learned-LR SGD: meta_optim.py:46-67 (LR init), :144-155 (reset), :177-214 (step);
    meta_model.py:49-80 (param groups live in module._parameters)."""

    def __init__(self, model, init_lr=1e-3, use_log_init_lr=False, **_unused):
        self.model = model
        self.meta_model = self                      # meta_optim.meta_model.detach_param_groups() (evaluate.py:274)
        self.use_log = use_log_init_lr
        self.only_box_head = False
        self.training = False
        self.lrs = []
        for _, p in model.named_parameters():
            if p.requires_grad:
                lr = torch.ones((p.shape[0],) + (1,) * (p.dim() - 1)).mul(init_lr)
                lr += torch.rand_like(lr).sub(0.5) * init_lr
                self.lrs.append(lr.log() if use_log_init_lr else lr)
        self.init = OrderedDict((n, p) for n, p in model.named_parameters() if p.requires_grad)

    def groups(self):
        for n_m, module in self.model.named_modules():
            for n_p, p in module._parameters.items():
                if p is not None and p.requires_grad:
                    yield n_m, module, n_p, p

    def reset(self, keep_state=False):
        if keep_state:
            self.detach_param_groups()
            return
        for n_m, module, n_p, _ in self.groups():
            module._parameters[n_p] = self.init[f"{n_m}.{n_p}"]

    # ---- the reference's optimizer surface used by evaluate.py:119-121,196-205,267-274
    def state_dict(self):
        names = list(self.init.keys())
        sd = OrderedDict((f"log_init_lr_{n.replace('.', '-')}", l) for n, l in zip(names, self.lrs))
        sd.update((f"model_init_{n.replace('.', '-')}", p) for n, p in self.init.items())
        return sd

    def load_state_dict(self, sd):
        with torch.no_grad():
            for n, l in zip(list(self.init.keys()), self.lrs):
                l.copy_(sd[f"log_init_lr_{n.replace('.', '-')}"])
            for n, p in self.init.items():
                p.copy_(sd[f"model_init_{n.replace('.', '-')}"])

    def eval(self):
        self.training = False

    def train(self):
        self.training = True

    def set_train_loss(self, loss):
        pass

    def detach_param_groups(self):
        for _, module, n_p, p in self.groups():
            d = p.detach()
            d.requires_grad = True
            module._parameters[n_p] = d

    def step(self, loss):
        groups = list(self.groups())
        grads = torch.autograd.grad(loss, [p for *_, p in groups])
        new = O.meta_update([p for *_, p in groups], grads, self.lrs, self.use_log)
        for (_, module, n_p, _), t in zip(groups, new):
            t = t.detach()
            t.requires_grad = True
            module._parameters[n_p] = t
        return grads


def run_frames(model, frames, start_target):
    """helper_func.py:67-159 for MaskRCNN with eval_augment_rpn_proposals_mode set: propagate the
    thresholded prediction as the next frame's target; empty prediction -> fall back to the start target."""
    probs_all, boxes_all = [], []
    targets = start_target.clone()
    model.eval()
    with torch.no_grad():
        for f in frames:
            probs, boxes = model(f[None], targets)
            nxt = O.threshold_targets(probs)
            targets = start_target if nxt.sum().item() == 0 else nxt
            probs_all.append(probs)
            boxes_all.append(boxes)
    return torch.cat(probs_all), torch.cat(boxes_all)


def jaccard(pred, gt):
    """Region similarity J (DAVIS): |A and S| / |A or S|, 1 when both are empty."""
    pred, gt = pred.bool(), gt.bool()
    union = (pred | gt).sum().item()
    return 1.0 if union == 0 else (pred & gt).sum().item() / union


def oracle_meta_gradients(model, lrs, train_batch, meta_batch, num_epochs=5, rng_states=None):
    """First-order BPTT of reference src/util/meta_run.py:124-214 in plain torch: theta_{k+1} = theta_k - lr * g_k with
    g_k detached (meta_optim.py:202-207, second_order False), meta loss after the last step, gradients w.r.t. theta_0
    and the learning rates.  `model` must hold leaf parameters theta_0; `lrs` leaf tensors (requires_grad).
    train_batch: one (inputs, gts) pair or a list with one pair per epoch; rng_states: optional torch RNG state to
    install before every forward (num_epochs train forwards + the meta forward).
    Pinned to the unmodified reference worker by tests/test_oracle_pins.py::test_meta_gradients_match_reference."""
    slots = [(n_m, module, n_p) for n_m, module in model.named_modules()
             for n_p, p in module._parameters.items() if p is not None and p.requires_grad]
    theta0 = [module._parameters[n_p] for _, module, n_p in slots]
    cur = theta0
    try:
        for e in range(num_epochs):
            model.train()
            if rng_states is not None:
                torch.set_rng_state(rng_states[e])
            loss, _ = model(*(train_batch[e] if isinstance(train_batch, list) else train_batch))
            grads = torch.autograd.grad(loss, cur)
            cur = [p - g.detach() * lr for p, g, lr in zip(cur, grads, lrs)]
            for (_, module, n_p), t in zip(slots, cur):
                module._parameters[n_p] = t
        if rng_states is not None:
            torch.set_rng_state(rng_states[num_epochs])
        meta_loss, _ = model(*meta_batch)
        out = torch.autograd.grad(meta_loss, list(theta0) + list(lrs), allow_unused=True)
    finally:
        for (_, module, n_p), t in zip(slots, theta0):
            module._parameters[n_p] = t
    n = len(theta0)
    return meta_loss.detach(), out[:n], out[n:]


def radam_reference(p, g, m, v, step, lr, wd, betas=(0.9, 0.999), eps=1e-8):
    """reference src/util/radam.py:28-94 for one tensor (degenerated_to_sgd=True)."""
    import math
    beta1, beta2 = betas
    v = v * beta2 + (1 - beta2) * g * g
    m = m * beta1 + (1 - beta1) * g
    beta2_t = beta2 ** step
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
    p = p.clone()
    if n_sma >= 5:
        ss = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - beta1 ** step)
        if wd != 0:
            p = p + (-wd * lr) * p
        p = p + (-ss * lr) * m / (v.sqrt() + eps)
    else:
        ss = 1.0 / (1 - beta1 ** step)
        if wd != 0:
            p = p + (-wd * lr) * p
        p = p + (-ss * lr) * m
    return p, m, v
