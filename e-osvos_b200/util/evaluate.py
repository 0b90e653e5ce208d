"""Per-video one-shot fine-tuning + (online-adapted) inference: the loop whose iterations/s and frames/s
are the headline metric, and the reference's `evaluate` worker around it.

  * `run_frames`        -- reference src/util/helper_func.py:67-159 (run_loader's target propagation, MaskRCNN branch)
  * `finetune`          -- HOT LOOP A, reference src/util/evaluate.py:213-281 (incl. early stopping :275-279)
  * `evaluate_sequence` -- reference src/util/evaluate.py:111-326 for one sequence: rounds, online-adaptation batches
                           (:172-253), FIRST_STEP restore (:196-205,283-287), frame ranges (:294-314), merge (:323-326),
                           per-object first annotated frame (src/data/youtube.py:131-185)
  * `evaluate`          -- the worker with the reference's signature and result contract (src/util/evaluate.py:20-23,
                           427-439): results are returned through `shared_dict`
Frames live in memory (`davis_io.VideoSequence`) instead of the reference's dataset objects.  Wall-clock accounting
follows evaluate.py:152,319-320,436."""
import collections
import copy
import os
import random
import time

import numpy as np
import torch

from .. import kernels as K
from . import augment
from .helper_func import early_stopping


_pinned = {}


def to_host(t):
    """Device -> host through a reused PINNED buffer (a plain `.cpu()` lands in pageable memory: a staged, several
    times slower copy).  The returned tensor is overwritten by the next call with the same shape / dtype."""
    key = (tuple(t.shape), t.dtype)
    buf = _pinned.get(key)
    if buf is None:
        if len(_pinned) > 16:
            _pinned.clear()
        buf = _pinned[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf


def set_random_seeds(seed):
    """helper_func.py:515-518"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def run_frames(model, frames_dev, start_target, on_frame=None):
    """helper_func.py:67-159 (MaskRCNN branch).  frames_dev: iterable of [1,3,H,W] device tensors.
    on_frame(i, targets, probs, boxes): observation hook (the target the frame was run with and its outputs).
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
    prefetch = getattr(model, "prefetch_backbone", None)
    # Sync-free propagation (EXTEND mode): the next frame needs only the BOX of this frame's thresholded prediction
    # (mask_rcnn.py:251-285), which the fused tail kernel leaves on the device; an empty prediction falls back to the
    # start target's box inside the EXTEND kernel (helper_func.py:124-126: `targets = start_targets`, mode EXTEND --
    # the mode the frame runs in anyway).  Other modes need the host to know, and keep the per-frame read-back.
    sync_free = mode == 'EXTEND' and start_target is not None and hasattr(model, "_forward_eval_fast")
    start_stats = None
    if sync_free:
        Kc = max(model.num_classes - 1, 1)
        start_stats = K.mask_to_bbox(start_target.to(torch.float32).contiguous(), Kc)
        K.target_stats.put(targets, start_stats, None)
    i = 0
    # Look-ahead over a run of frames: whatever of a frame does not depend on its predecessor's result (transform,
    # trunk, RPN head, proposal selection) runs for up to `ahead` frames in one batched graph -- possible whenever every
    # frame of the run takes the same branch: device-resident propagation (EXTEND) or no proposal augmentation at all.
    ahead = 0
    if (sync_free or mode is None) and getattr(model, "lookahead_ok", None) is not None:
        model.eval()
        ahead = int(os.environ.get("EOSVOS_FRAME_BATCH", "5")) if model.lookahead_ok() else 0
    queue = collections.deque()

    def pull():
        if not queue:
            for _ in range(max(ahead, 1)):
                f = next(it, None)
                if f is None:
                    break
                queue.append(f)
            if ahead > 1 and len(queue) >= 2:
                model.prefetch_frames(list(queue), targets is not None)
        return queue.popleft() if queue else None
    with torch.no_grad():
        model.eval()
        nxt_inputs = pull()
        if nxt_inputs is not None and prefetch is not None and ahead <= 1:
            prefetch(nxt_inputs)
        while nxt_inputs is not None:
            inputs = nxt_inputs
            if ahead <= 1 or queue:
                nxt_inputs = pull()
                pulled = True
            else:
                pulled = False                # last frame of a batched run: fetch the next run after this frame ran
            model.eval()
            if prefetch is not None and nxt_inputs is not None and ahead <= 1:
                # look-ahead: the trunk of the NEXT frame is independent of this frame's result; enqueue it behind
                # this frame's trunk so the GPU stays busy while the host prepares the heads
                prefetch(nxt_inputs)
            used = targets
            probs, boxes = model(inputs, targets)
            if mode is not None:
                # threshold / argmax + pixel count come out of the fused tail kernel (no extra passes)
                nxt, stats = model.last_propagated_target, model.last_target_stats
                model.rpn._eval_augment_proposals_mode = mode
                if sync_free and getattr(model, "_last_det", None) is not None:
                    K.target_stats.put(nxt, stats, start_stats)
                    targets = nxt
                    if on_frame is not None and used is not None and int(K.target_stats.get(used)[0][..., 4].sum()) == 0:
                        used = start_target                   # what the reference's loop would have passed
                else:
                    stats_cpu = stats.cpu()                      # one host sync per frame (helper_func.py:124)
                    if int(stats_cpu[..., 4].sum()) == 0:
                        model.rpn._eval_augment_proposals_mode = 'EXTEND'
                        targets = start_target
                    else:
                        targets = nxt
                        # boxes / counts of the propagated target are already on the host: the next forward skips
                        # its own mask->box kernel and read-back
                        K.target_stats.put(targets, stats_cpu, torch.zeros(stats_cpu.shape[0], dtype=torch.int32))
            if on_frame is not None:
                on_frame(i, used, probs, boxes)
            probs_all.append(probs)
            boxes_all.append(boxes)
            i += 1
            if not pulled:
                nxt_inputs = pull()
    if ahead > 1:
        model._lookahead = None           # batched features are only valid for the weights and frames they ran on
    return torch.cat(probs_all), torch.cat(boxes_all)


def finetune(model, meta_optim, batch_fn, num_iters, seed, round_idx, reset_mode='FIRST_STEP', on_iter=None,
             early_stopping_cfg=None):
    """evaluate.py:213-281.  batch_fn(epoch) -> (inputs [B,3,H,W], gts [B,1,H,W]) on the device.
    early_stopping_cfg = {"patience": int | None, "min_loss_improv": float} (cfgs/meta.yaml:97-99, default off: the
    loss is then never read back inside the loop)."""
    model.train_without_dropout()
    loss, losses = None, None
    patience = (early_stopping_cfg or {}).get("patience")
    hist = []
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
        if patience is not None:
            hist.append(loss.item())                         # evaluate.py:263 (a host sync per iteration)
            if early_stopping(hist, **early_stopping_cfg):
                break
    finetune.last_losses = losses
    return loss


def _as_sequence(frames, first_label):
    from .davis_io import VideoSequence
    T, _, H, W = frames.shape
    labels = torch.zeros((T, H, W), dtype=torch.uint8)
    labels[0] = torch.as_tensor(first_label).to(torch.uint8)
    ann = np.zeros(T, bool)
    ann[0] = True
    return VideoSequence(frames, labels, annotated=ann)


def evaluate_sequence(model, meta_optim, meta_optim_state_dict, frames, first_label=None, *, num_epochs_eval,
                      online_adapt_step=0, online_adapt_epochs=10, min_prop=0.5, batch_size=3, seed=1,
                      random_train_transform=True, reset_model_mode='FIRST_STEP', device=None, timers=None,
                      device_augment=True, early_stopping_cfg=None, hooks=None):
    """frames: a `davis_io.VideoSequence`, or float32 [T,3,H,W] in [0,1] (host, ideally pinned) together with
    first_label [H,W] object ids.  hooks (optional observation points, used by the lock-step parity tests):
        after_finetune(obj, k, model, meta_optim), before_frames(obj, k, frame_ids, start_target),
        on_frame(obj, k, i, targets, probs, boxes), batch(obj, k, epoch, inputs, gts)
    Returns (pred uint8 [T,H,W], stats dict)."""
    from .davis_io import VideoSequence
    seq = frames if isinstance(frames, VideoSequence) else _as_sequence(frames, first_label)
    frames = seq.frames
    hooks = hooks or {}
    device = device or next(model.parameters()).device
    T, _, H, W = frames.shape
    num_objects = seq.num_objects
    model.roi_heads.detections_per_img = 1          # multi_object == 'single_id' (evaluate.py:106-107)
    masks = [torch.zeros(num_objects, H, W) for _ in range(T)]
    boxes = [None] * T
    timers = timers if timers is not None else {}
    for k_, v_ in (("finetune_iters", 0), ("finetune_s", 0.0), ("infer_frames", 0), ("infer_s", 0.0)):
        timers.setdefault(k_, v_)
    eval_time, num_frames = 0.0, 0
    train_loss_seq, train_losses_seq = [], []

    def to_dev(t):
        return t.to(device, non_blocking=True)

    for obj in range(num_objects):
        train_frame_id, label_idx = seq.gt_frame_id(obj)
        gt0 = seq.label(train_frame_id, obj, label_idx)[None]                   # [1,H,W]: what fine-tuning sees
        # the start target of propagation is the TEST loader's view of the train frame, whose `_label_id` is never
        # set (evaluate.py:156-157 with vos_dataset.py:236-245): for a late-appearing YouTube-VOS object that is the
        # first label file, i.e. possibly empty
        train_frame_gt = seq.label(train_frame_id, obj, None)[None]
        gt0_np = gt0[0].numpy()
        frame0 = frames[train_frame_id]
        masks[train_frame_id][obj] = 2 * train_frame_gt[0]
        start = time.perf_counter()
        if online_adapt_step:
            step = online_adapt_step
            meta_frames = list(range(train_frame_id + 1, T, step))
        else:
            step = T
            meta_frames = [0]
        range_max = train_frame_id + 1
        state_first = None
        for k, _ in enumerate(meta_frames):
            if k == 0:
                range_min = train_frame_id + 1
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
                dev_aug = augment.PrefetchingAugmenter(to_dev(frame0), gt0_np, batch_size, lambda e, k=k: seed + e + k)

                def batch_fn(epoch):
                    return dev_aug.get(epoch)
            elif k == 0:
                frame0_np = frame0.permute(1, 2, 0).contiguous().numpy()
                static = None
                if not random_train_transform:
                    static = (to_dev(frame0[None].repeat(batch_size, 1, 1, 1)), to_dev(gt0[None].repeat(batch_size, 1, 1, 1)))

                def batch_fn(epoch, static=static):
                    if static is not None:
                        return static
                    imgs, gts = [], []
                    for _ in range(batch_size):
                        im, g = augment.augment_first_frame(frame0_np, gt0_np)
                        imgs.append(torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))))
                        gts.append(torch.from_numpy(np.ascontiguousarray(g))[None])
                    return to_dev(torch.stack(imgs)), to_dev(torch.stack(gts))
            else:
                n_prop = min(online_adapt_step, batch_size)
                first_pid = online_adapt_step - n_prop + 1
                sel = [(range_min - pid, propagate_gts[pid - 1]) for pid in range(first_pid, online_adapt_step)
                       if propagate_gts[pid - 1].sum().item() != 0]
                inputs_k = to_dev(torch.stack([frame0] + [frames[f] for f, _ in sel]))
                gts_k = to_dev(torch.stack([gt0] + [g for _, g in sel]))

                def batch_fn(epoch, inputs_k=inputs_k, gts_k=gts_k):
                    return inputs_k, gts_k

            if "batch" in hooks:
                inner = batch_fn

                def batch_fn(epoch, inner=inner, k=k):
                    b = inner(epoch)
                    hooks["batch"](obj, k, epoch, *b)
                    return b

            t0 = time.perf_counter()
            loss = finetune(model, meta_optim, batch_fn, iters, seed, k, reset_model_mode,
                            early_stopping_cfg=early_stopping_cfg)
            torch.cuda.synchronize(device)
            if hasattr(meta_optim, "check_finite"):
                meta_optim.check_finite()         # loud failure if the scaled 16-bit backward overflowed
            timers["finetune_s"] += time.perf_counter() - t0
            timers["finetune_iters"] += iters
            train_loss_seq.append(loss)                                   # read back after the sequence
            train_losses_seq.append(finetune.last_losses)
            if "after_finetune" in hooks:
                hooks["after_finetune"](obj, k, model, meta_optim)
            if k == 0:
                state_first = copy.deepcopy(model.state_dict())

            t0 = time.perf_counter()
            start_target = (train_frame_gt if k == 0 else propagate_gt)[None]
            frame_ids = list(range(range_min, range_max))
            if "before_frames" in hooks:
                hooks["before_frames"](obj, k, frame_ids, start_target)
            fr = (to_dev(frames[f:f + 1]) for f in frame_ids)
            on_frame = (lambda i, t, p, b, k=k: hooks["on_frame"](obj, k, i, t, p, b)) if "on_frame" in hooks else None
            probs, bxs = run_frames(model, fr, to_dev(start_target), on_frame)
            probs, bxs = to_host(probs), bxs.cpu()
            timers["infer_s"] += time.perf_counter() - t0
            timers["infer_frames"] += range_max - range_min
            for f, p, b in zip(frame_ids, probs, bxs):
                masks[f][obj:obj + 1] = p
                boxes[f] = b if boxes[f] is None else torch.cat([boxes[f], b])
            if range_max == T:
                break
        eval_time += time.perf_counter() - start
        num_frames += T

    pred = torch.zeros(T, H, W, dtype=torch.uint8)
    for f in range(T if num_objects else 0):             # evaluate.py:323-326
        bg = masks[f].max(dim=0)[0].lt(0.5)
        m = masks[f].argmax(dim=0) + 1
        m[bg] = 0
        pred[f] = m.to(torch.uint8)
    return pred, {"time_per_frame": eval_time / max(num_frames, 1), "eval_time": eval_time, "num_frames": num_frames,
                  "train_loss_seq": [float(l) for l in train_loss_seq],
                  "train_losses_seq": [{n: float(v) for n, v in d.items()} for d in train_losses_seq],
                  "masks": masks, "boxes": boxes}


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


# ------------------------------------------------------------------------------------------------------------------
# the reference's worker
# ------------------------------------------------------------------------------------------------------------------
def evaluate(rank, dataset_key, shared_meta_optim_state_dict, shared_variables, _config, shared_dict, save_dir,
             vis_win_names, evaluate_only, _log, once=False):
    """Drop-in for the reference's evaluation worker (src/util/evaluate.py:20-439): same arguments, same loop
    protocol (waits until `shared_dict['meta_iter']` is None, evaluates the dataset `_config['datasets'][dataset_key]`
    on `cuda:{rank}` with a copy of the shared MetaOptimizer state, publishes the results under the same
    `shared_dict` keys and sets `shared_dict['meta_iter']` to signal completion), same artefacts (uint8 object-id PNGs
    under {save_dir}/best_eval_preds/<name>/<split>/<seq>/, last_/best_ checkpoints).  `once=True` returns after
    one evaluation instead of waiting for the parent to clear `meta_iter` (the reference's parent terminates the
    process, train_meta.py:249-252).  Data is read from `data/<DATASET>` relative to the working directory
    (helper_func.py:265-273).  Not mirrored: the Visdom/matplotlib debug renderings (evaluate.py:384-425)."""
    from ..meta_optim.meta_optim import MetaOptimizer
    from . import davis_io, metrics
    from .helper_func import init_parent_model
    seed = _config['seed']
    datasets = _config['datasets']
    data_cfg = copy.deepcopy(_config['data_cfg'])
    if data_cfg['multi_object'] not in ('single_id', False):
        raise NotImplementedError("the B200 path evaluates with multi_object 'single_id' (the e-OSVOS configs) or "
                                  "single-object datasets")
    while True:
        while shared_dict['meta_iter'] is not None:
            time.sleep(0.25)
        meta_optim_state_dict = copy.deepcopy(dict(shared_meta_optim_state_dict))
        meta_iter = shared_variables['meta_iter']
        meta_epoch = shared_variables['meta_epoch']
        set_random_seeds(seed)
        device = torch.device(f'cuda:{rank}')
        model, parent_states = init_parent_model(**_config['parent_model'])
        if dataset_key in parent_states and parent_states[dataset_key]['states']:
            if len(parent_states[dataset_key]['states']) > 1:
                raise NotImplementedError
            model.load_state_dict(parent_states[dataset_key]['states'][0])
        meta_optim = MetaOptimizer(model, **_config['meta_optim_cfg'])
        meta_optim.load_state_dict(meta_optim_state_dict)
        model.to(device)
        meta_optim.to(device)

        ds = datasets[dataset_key]
        dataset = davis_io.open_dataset(ds['name'], ds['split'], multi_object=data_cfg['multi_object'],
                                        full_resolution=data_cfg.get('full_resolution', False))
        preds_save_dir = None
        if save_dir is not None:
            preds_save_dir = os.path.join(save_dir, 'best_eval_preds', f"{ds['name']}", f"{ds['split']}")
            for seq_name in dataset.seq_names:
                os.makedirs(os.path.join(preds_save_dir, seq_name), exist_ok=True)

        ona = _config['eval_online_adapt']
        eval_time, num_frames = 0.0, 0
        out = {k: [] for k in ('init_J_seq', 'J_seq', 'J_recall_seq', 'J_decay_seq', 'train_loss_seq',
                               'train_losses_seq', 'F_seq', 'F_recall_seq', 'F_decay_seq')}
        if data_cfg['multi_object'] == 'single_id':
            model.roi_heads.detections_per_img = 1
        for seq_name in dataset.seq_names:
            seq = dataset.load(seq_name, pin=True)
            if seq.num_objects == 1:
                # J of the un-adapted initialisation (evaluate.py:116-127): plain inference, no proposal augmentation
                meta_optim.load_state_dict(meta_optim_state_dict)
                meta_optim.reset()
                meta_optim.eval()
                if seq.test_mode or dataset.all_frames:
                    out['init_J_seq'].extend([0.0])
                else:
                    fr = (seq.frames[f:f + 1].to(device, non_blocking=True) for f in range(len(seq)))
                    probs, _ = run_frames(model, fr, None)
                    init_pred = probs[:, 0].ge(0.5).to(torch.uint8)
                    jf = metrics.evaluate_sequence_jf_device(
                        init_pred, (seq.labels != 0).to(torch.uint8).to(device), 1, measures=("J",))
                    out['init_J_seq'].extend([s[0] for s in jf["J"]])
            pred, stats = evaluate_sequence(
                model, meta_optim, meta_optim_state_dict, seq, num_epochs_eval=_config['num_epochs']['eval'],
                online_adapt_step=ona['step'], online_adapt_epochs=ona['num_epochs'], min_prop=ona['min_prop'],
                batch_size=data_cfg['batch_sizes']['train'], seed=seed,
                random_train_transform=data_cfg['random_train_transform'], reset_model_mode=ona['reset_model_mode'],
                device=device, early_stopping_cfg=_config.get('train_early_stopping_cfg'))
            eval_time += stats["eval_time"]
            num_frames += stats["num_frames"]
            out['train_loss_seq'].extend(stats["train_loss_seq"])
            out['train_losses_seq'].extend(stats["train_losses_seq"])
            pred_np = pred.numpy()
            if preds_save_dir is not None:
                keep = [f for f in range(len(seq)) if not dataset.all_frames or seq.annotated[f]]
                metrics.save_predictions(pred_np[keep], preds_save_dir, seq_name, [seq.names[f] for f in keep])
            if seq.test_mode:
                jf = {"J": [(0.0, 0.0, 0.0)], "F": [(0.0, 0.0, 0.0)]}
            else:
                ann = np.where(seq.annotated)[0]
                n_eval = seq.num_objects if data_cfg['multi_object'] else 1
                lab = seq.labels if data_cfg['multi_object'] else (seq.labels != 0)
                jf = metrics.evaluate_sequence_jf_device(pred[ann].to(torch.uint8).to(device),
                                                         lab[ann].to(torch.uint8).to(device), n_eval)
            if evaluate_only and _log is not None:
                _log.info(f"{dataset_key}: {seq_name} {[s[0] for s in jf['J']]}")
            for m in ("J", "F"):
                out[f'{m}_seq'].extend([s[0] for s in jf[m]])
                out[f'{m}_recall_seq'].extend([s[1] for s in jf[m]])
                out[f'{m}_decay_seq'].extend([s[2] for s in jf[m]])

        test_mode = dataset.test_mode
        save_meta_run = {'meta_optim_state_dict': meta_optim.state_dict(), 'vis_win_names': vis_win_names,
                         'meta_iter': meta_iter, 'meta_epoch': meta_epoch}
        if save_dir is not None and not test_mode:
            torch.save(save_meta_run, os.path.join(save_dir, f"last_{dataset_key}_meta_iter.model"))
        mean_J = torch.tensor(out['J_seq']).mean().item()
        if test_mode or mean_J > shared_dict['best_mean_J']:
            shared_dict['best_mean_J'] = mean_J
            if save_dir is not None and not test_mode:
                torch.save(save_meta_run, os.path.join(save_dir, f"best_{dataset_key}_meta_iter.model"))
        for k, v in out.items():
            shared_dict[k] = v
        shared_dict['time_per_frame'] = eval_time / max(num_frames, 1)
        # set meta_iter here to signal the main process that the evaluation is finished (evaluate.py:438-439)
        shared_dict['meta_iter'] = meta_iter
        if once:
            return
