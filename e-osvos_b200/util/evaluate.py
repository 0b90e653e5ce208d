"""Per-video one-shot fine-tuning + (online-adapted) inference: the loop whose iterations/s and frames/s
are the headline metric.  Mirrors reference src/util/evaluate.py:111-326 (HOT LOOP A :220-280, OnA batch
assembly :172-253, state restore :196-205,283-287, inference ranges :294-314, object merge :323-326) and
src/util/helper_func.py:67-159 (run_loader's target propagation), on in-memory frames instead of the
reference's dataset objects.  Wall-clock accounting follows evaluate.py:152,319-320,436."""
import copy
import random
import time

import numpy as np
import torch

from .. import kernels as K
from . import augment
from .shard import ona_schedule


def set_random_seeds(seed):
    """helper_func.py:515-518"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def run_frames(model, frames_dev, start_target):
    """helper_func.py:67-159 (MaskRCNN branch).  frames_dev: iterable of [1,3,H,W] device tensors.
    Returns (probs [F,K,H,W], boxes [F,K,4])."""
    mode = model.rpn._eval_augment_proposals_mode
    targets = None
    if mode is not None:
        if start_target is None:
            targets = None
        elif start_target.sum().item() == 0:
            start_target = None
            model.rpn._eval_augment_proposals_mode = 'EXTEND'
            targets = None
        else:
            targets = start_target.clone()
    probs_all, boxes_all = [], []
    it = iter(frames_dev)
    nxt_inputs = next(it, None)
    prefetch = getattr(model, "prefetch_backbone", None)
    with torch.no_grad():
        model.eval()
        if nxt_inputs is not None and prefetch is not None:
            prefetch(nxt_inputs)
        while nxt_inputs is not None:
            inputs, nxt_inputs = nxt_inputs, next(it, None)
            model.eval()
            if prefetch is not None and nxt_inputs is not None:
                # look-ahead: the trunk of the NEXT frame is independent of this frame's result; enqueue it behind
                # this frame's trunk so the GPU stays busy while the host filters proposals / waits for detections
                prefetch(nxt_inputs)
            probs, boxes = model(inputs, targets)
            if mode is not None:
                # threshold / argmax + pixel count come out of the fused tail kernel (no extra passes)
                nxt, stats = model.last_propagated_target, model.last_target_stats
                model.rpn._eval_augment_proposals_mode = mode
                stats_cpu = stats.cpu()                      # the ONE host sync per frame (helper_func.py:124)
                if int(stats_cpu[..., 4].sum()) == 0:
                    model.rpn._eval_augment_proposals_mode = 'EXTEND'
                    targets = start_target
                else:
                    targets = nxt
                    # boxes / counts of the propagated target are already on the host: the next forward skips its
                    # own mask->box kernel and read-back
                    K.target_stats.put(targets, stats_cpu, torch.zeros(stats_cpu.shape[0], dtype=torch.int32))
            probs_all.append(probs)
            boxes_all.append(boxes)
    return torch.cat(probs_all), torch.cat(boxes_all)


def finetune(model, meta_optim, batch_fn, num_iters, seed, round_idx, reset_mode='FIRST_STEP', on_iter=None):
    """evaluate.py:213-281.  batch_fn(epoch) -> (inputs [B,3,H,W], gts [B,1,H,W]) on the device."""
    model.train_without_dropout()
    loss = None
    for epoch in range(1, num_iters + 1):
        set_random_seeds(seed + epoch + round_idx)
        inputs, gts = batch_fn(epoch)
        loss, losses = model(inputs, gts)
        model.zero_grad()
        meta_optim.set_train_loss(loss)
        if reset_mode == 'FIRST_STEP':
            meta_optim.only_box_head = round_idx != 0
        meta_optim.step(loss)
        meta_optim.meta_model.detach_param_groups()
        if on_iter is not None:
            on_iter(epoch, loss)
    return loss


def evaluate_sequence(model, meta_optim, meta_optim_state_dict, frames, first_label, *, num_epochs_eval,
                      online_adapt_step=0, online_adapt_epochs=10, min_prop=0.5, batch_size=3, seed=1,
                      random_train_transform=True, reset_model_mode='FIRST_STEP', device=None, timers=None,
                      device_augment=True):
    """frames: float32 [T,3,H,W] in [0,1] (host, ideally pinned); first_label: [H,W] object ids.
    Returns (pred uint8 [T,H,W], stats dict)."""
    device = device or next(model.parameters()).device
    T, _, H, W = frames.shape
    num_objects = int(first_label.max())
    model.roi_heads.detections_per_img = 1          # multi_object == 'single_id' (evaluate.py:106-107)
    masks = [torch.zeros(num_objects, H, W) for _ in range(T)]
    timers = timers if timers is not None else {}
    timers.setdefault("finetune_iters", 0)
    timers.setdefault("finetune_s", 0.0)
    timers.setdefault("infer_frames", 0)
    timers.setdefault("infer_s", 0.0)
    eval_time, num_frames = 0.0, 0
    frame0_np = frames[0].permute(1, 2, 0).contiguous().numpy()

    def to_dev(t):
        return t.to(device, non_blocking=True)

    for obj in range(num_objects):
        gt0 = (first_label == (obj + 1)).float()[None]                 # [1,H,W]
        gt0_np = gt0[0].numpy()
        masks[0][obj] = 2 * gt0[0]
        start = time.perf_counter()
        if online_adapt_step:
            step = online_adapt_step
            meta_frames = list(range(1, T, step))
        else:
            step = T
            meta_frames = [0]
        range_max = 1
        state_first = None
        for k, _ in enumerate(meta_frames):
            if k == 0:
                range_min = 1
                propagate_gt, propagate_gts = None, []
            else:
                range_min = range_max
                propagate_gt = masks[range_min - 1][obj:obj + 1].ge(min_prop).float()
                propagate_gts = [masks[range_min - pid][obj:obj + 1].ge(min_prop).float()
                                 for pid in range(1, online_adapt_step)]
            range_max = min(range_max + step, T)

            if k == 0 or reset_model_mode == 'FULL':
                meta_optim.load_state_dict(meta_optim_state_dict)
                meta_optim.reset()
                meta_optim.eval()
            elif reset_model_mode == 'FIRST_STEP':
                meta_optim.load_state_dict(meta_optim_state_dict)
                model.load_state_dict(state_first)
                meta_optim.eval()

            iters = num_epochs_eval if k == 0 else online_adapt_epochs

            if k == 0 and random_train_transform and device_augment:
                dev_aug = augment.PrefetchingAugmenter(to_dev(frames[0]), gt0_np, batch_size,
                                                       lambda e, k=k: seed + e + k)

                def batch_fn(epoch):
                    return dev_aug.get(epoch)
            elif k == 0:
                def batch_fn(epoch):
                    imgs, gts = [], []
                    for _ in range(batch_size):
                        if random_train_transform:
                            im, g = augment.augment_first_frame(frame0_np, gt0_np)
                        else:
                            im, g = frame0_np, gt0_np
                        imgs.append(torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))))
                        gts.append(torch.from_numpy(np.ascontiguousarray(g))[None])
                    return to_dev(torch.stack(imgs)), to_dev(torch.stack(gts))
            else:
                n_prop = min(online_adapt_step, batch_size)
                first_pid = online_adapt_step - n_prop + 1
                sel = [(range_min - pid, propagate_gts[pid - 1]) for pid in range(first_pid, online_adapt_step)
                       if propagate_gts[pid - 1].sum().item() != 0]
                inputs_k = to_dev(torch.stack([frames[0]] + [frames[f] for f, _ in sel]))
                gts_k = to_dev(torch.stack([gt0] + [g for _, g in sel]))

                def batch_fn(epoch, inputs_k=inputs_k, gts_k=gts_k):
                    return inputs_k, gts_k

            t0 = time.perf_counter()
            finetune(model, meta_optim, batch_fn, iters, seed, k, reset_model_mode)
            torch.cuda.synchronize(device)
            timers["finetune_s"] += time.perf_counter() - t0
            timers["finetune_iters"] += iters
            if k == 0:
                state_first = copy.deepcopy(model.state_dict())

            t0 = time.perf_counter()
            start_target = (gt0 if k == 0 else propagate_gt)[None]
            fr = (to_dev(frames[f:f + 1]) for f in range(range_min, range_max))
            probs, _ = run_frames(model, fr, to_dev(start_target))
            probs = probs.cpu()
            timers["infer_s"] += time.perf_counter() - t0
            timers["infer_frames"] += range_max - range_min
            for f, p in zip(range(range_min, range_max), probs):
                masks[f][obj:obj + 1] = p
            if range_max == T:
                break
        eval_time += time.perf_counter() - start
        num_frames += T

    pred = torch.zeros(T, H, W, dtype=torch.uint8)
    for f in range(T):                                   # evaluate.py:323-326
        bg = masks[f].max(dim=0)[0].lt(0.5)
        m = masks[f].argmax(dim=0) + 1
        m[bg] = 0
        pred[f] = m.to(torch.uint8)
    return pred, {"time_per_frame": eval_time / max(num_frames, 1), "eval_time": eval_time, "num_frames": num_frames}


def jaccard_per_object(pred, labels, num_objects):
    """Region similarity J per object, mean over frames excluding first and last (DAVIS protocol)."""
    out = []
    T = pred.shape[0]
    for k in range(1, num_objects + 1):
        js = []
        for f in range(1, max(T - 1, 2)):
            p, g = pred[f] == k, labels[f] == k
            union = (p | g).sum().item()
            js.append(1.0 if union == 0 else (p & g).sum().item() / union)
        out.append(sum(js) / len(js))
    return out
