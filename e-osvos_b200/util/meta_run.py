"""The reference's meta-training worker (src/util/meta_run.py:14-243) on the B200 path: same signature, same
shared-memory protocol with the parent process of src/train_meta.py (waits while `shared_dict['sub_iter_done']`, loads
the shared MetaOptimizer state, runs its share of the meta batch, ADDS its meta-gradients into
`shared_meta_optim_grads[name]` (host tensors), publishes `seqs_metrics` / `vis_data_seqs` and sets `sub_iter_done`).

Per task (meta_run.py:96-226): `num_epochs.train` fine-tune steps on one randomly chosen annotated frame with the
fused update kept in the autograd graph, meta loss on another random frame every `bptt_epochs`, first-order BPTT into
theta_0 / lambda -- `meta_train.task_meta_gradients`.  Task sampling restates src/meta_optim/meta_tasksets.py:54-154
(random train / meta frame with the object visible, optional per-task colour jitter + flip + scale/rotate shared by
both frames) on `davis_io` sequences.  The random draws the reference's DataLoaders make (`loader_draws`) are kept, so
a given seed selects the same tasks and frames as the reference.

The parent of the reference sums the per-process gradients on the CPU and steps RAdam there (train_meta.py:361-373);
that protocol keeps working with this worker.  The all-GPU variant (one NCCL all-reduce + fused RAdam, no parent) is
`meta_train.meta_iteration`.  Not mirrored: `single_obj_seq_mode` AUGMENT_* (pasting a second sequence's object),
`random_box_coord_perm`, `random_object_id_sub_group`, `multi_step_bptt_loss` -- all off in every shipped config."""
import random
import time

import numpy as np
import torch

from . import augment, davis_io
from .evaluate import set_random_seeds
from .helper_func import early_stopping, init_parent_model


def loader_draws(n):
    """The int64 draws torch's DataLoader / RandomSampler take from the default generator per iteration."""
    for _ in range(n):
        torch.empty((), dtype=torch.int64).random_()


class TaskSet:
    """(sequence, object) tasks of one or several datasets (meta_tasksets.py:20-52)."""

    def __init__(self, dataset_cfg, data_cfg, single_obj_seq_mode):
        names, splits = dataset_cfg['name'], dataset_cfg['split']
        if not isinstance(names, list):
            names, splits = [names], [splits]
        self.datasets = [davis_io.open_dataset(n, s, multi_object=data_cfg['multi_object'],
                                               full_resolution=data_cfg.get('full_resolution', False))
                         for n, s in zip(names, splits)]
        self.tasks = []
        for d, ds in enumerate(self.datasets):
            for seq_name in ds.seq_names:
                T, K = davis_io.sequence_meta(ds.root if not ds.youtube else ds.root + "/" + ds.part, seq_name,
                                              ds.resolution)
                K = K if ds.multi_object else 1
                if K == 1 and single_obj_seq_mode == 'IGNORE':
                    continue
                if K != 1 and single_obj_seq_mode == 'ONLY':
                    continue
                self.tasks += [(d, seq_name, k) for k in range(K)]
        self._cache = {}

    def __len__(self):
        return len(self.tasks)

    def sequence(self, d, seq_name):
        key = (d, seq_name)
        if key not in self._cache:
            if len(self._cache) > 8:
                self._cache.clear()
            self._cache[key] = self.datasets[d].load(seq_name)
        return self._cache[key]


def _frame_has_object(seq, idx, obj):
    return bool(seq.label(idx, obj).any())


def _random_frame_with_label(seq, obj):
    """vos_dataset.py:103-140: uniform draws until the object is visible."""
    while True:
        idx = torch.randint(len(seq), (1,)).item()
        if _frame_has_object(seq, idx, obj):
            return idx


class _TaskTransform:
    """Per-task deterministic augmentation (meta_tasksets.py:109-131): one colour jitter and one flip decision for the
    task, one scale / rotation per frame file (custom_transforms.py:20-89,118-150,188-211 with deterministic=True)."""

    def __init__(self):
        import torchvision
        cj = torchvision.transforms.ColorJitter(brightness=.2, contrast=.2, hue=.1, saturation=.2)
        self.cj_params = None
        self._cj = cj
        self.do_flip = random.random() < 0.5          # RandomHorizontalFlip(deterministic=True).__init__
        self.rot_sc = {}

    def __call__(self, image, gt, key):
        import cv2
        import torchvision.transforms.functional as TF
        from PIL import Image
        if self.cj_params is None:
            self.cj_params = self._cj.get_params(self._cj.brightness, self._cj.contrast, self._cj.saturation, self._cj.hue)
        order, b, c, s, h = self.cj_params
        pil = Image.fromarray(np.uint8(image * 255))
        for fn_id in order:
            if fn_id == 0 and b is not None:
                pil = TF.adjust_brightness(pil, b)
            elif fn_id == 1 and c is not None:
                pil = TF.adjust_contrast(pil, c)
            elif fn_id == 2 and s is not None:
                pil = TF.adjust_saturation(pil, s)
            elif fn_id == 3 and h is not None:
                pil = TF.adjust_hue(pil, h)
        image = np.array(pil, dtype=np.float32) / 255
        if self.do_flip:
            image, gt = cv2.flip(image, flipCode=1), cv2.flip(gt, flipCode=1)
        num_labels = len(np.unique(gt))
        while True:
            if key in self.rot_sc:
                rot, sc = self.rot_sc[key]
            else:
                rot = 60 * random.random() - 30
                sc = 0.5 * random.random() - 0.25 + 1 - 0.25          # scales (.5, 1.0)
            aug_gt = augment._warp(gt, rot, sc, True)
            if not num_labels > 1 or len(np.unique(aug_gt)) == num_labels:
                break
            assert key not in self.rot_sc
        self.rot_sc[key] = (rot, sc)
        return augment._warp(image, rot, sc, False), aug_gt


def _sample(seq, idx, obj, transform, device):
    img = seq.frames[idx].permute(1, 2, 0).contiguous().numpy()
    gt = seq.label(idx, obj).numpy()
    if transform is not None:
        img, gt = transform(img, gt, seq.names[idx])
    x = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))[None]
    y = torch.from_numpy(np.ascontiguousarray(gt))[None, None]
    return x.to(device), y.to(device)


def meta_run(rank, init_model_state_dict, shared_meta_optim_state_dict, global_rng_state, _config, dataset, shared_dict,
             shared_variables, shared_meta_optim_grads, save_dir, num_meta_processes, once=False):
    from ..meta_optim.meta_optim import MetaOptimizer
    per_gpu = max(int(_config['num_meta_processes_per_gpu']), 1)
    gpu_rank = rank // per_gpu + (int(_config['num_eval_gpus'] or 0) if _config['eval_datasets'] else 0)
    device = torch.device(f'cuda:{gpu_rank}')
    for k in ('random_flip_label', 'random_no_label', 'random_box_coord_perm', 'random_object_id_sub_group',
              'multi_step_bptt_loss'):
        if _config.get(k):
            raise NotImplementedError(f"{k}: not on the B200 path (off in every shipped config)")
    if _config['data_cfg']['batch_sizes']['train'] != 1 or _config['data_cfg']['batch_sizes']['meta'] != 1:
        raise NotImplementedError("meta-training runs at train / meta batch size 1 (cfgs/meta.yaml:121-126)")

    set_random_seeds(_config['seed'] + rank)
    model, _ = init_parent_model(**_config['parent_model'])
    model.load_state_dict(init_model_state_dict)
    meta_optim = MetaOptimizer(model, **_config['meta_optim_cfg'])
    num_epochs = _config['num_epochs']['train']
    sub_meta_batch_size = _config['meta_batch_size'] // num_meta_processes
    tasks = TaskSet(dataset, _config['data_cfg'], _config['single_obj_seq_mode'])
    model.to(device)
    meta_optim.to(device)

    while True:
        # iter(DataLoader(shuffle=True)): one draw for the loader's base seed, one for the RandomSampler's own
        # generator, whose permutation orders the tasks (torch/utils/data/sampler.py)
        loader_draws(1)
        gen = torch.Generator()
        gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        order = torch.randperm(len(tasks), generator=gen).tolist()
        for b0 in range(0, len(order), sub_meta_batch_size):
            batch = order[b0:b0 + sub_meta_batch_size]
            while shared_dict['sub_iter_done']:
                time.sleep(0.25)
            meta_optim.load_state_dict(dict(shared_meta_optim_state_dict))
            meta_optim.zero_grad()
            metrics = {m: {} for m in ['train_loss', 'train_losses', 'meta_loss', 'meta_losses', 'loss', 'J', 'F']}
            vis = {}
            # MetaTaskset.__getitem__ of every task of the mini batch first (frame choice, flip decision), as the
            # DataLoader fetches the whole batch before the inner loops start
            picked = []
            for t in batch:
                d, seq_name, obj = tasks.tasks[t]
                seq = tasks.sequence(d, seq_name)
                train_idx = _random_frame_with_label(seq, obj)
                meta_idx = _random_frame_with_label(seq, obj)
                transform = _TaskTransform() if _config['random_frame_transform_per_task'] else None
                picked.append((seq_name, seq, obj, train_idx, meta_idx, transform))
            for seq_name, seq, obj, train_idx, meta_idx, transform in picked:
                for m in metrics.values():
                    m.setdefault(seq_name, [])
                vis.setdefault(seq_name, [])
                memo = {}

                def get(which, seq=seq, obj=obj, transform=transform, memo=memo, idx={"train": train_idx, "meta": meta_idx}):
                    # the transforms draw their per-file parameters at first use, i.e. inside the first epoch, after
                    # that epoch's seeding; afterwards they are deterministic, so the sample is built once
                    if which not in memo:
                        memo[which] = _sample(seq, idx[which], obj, transform, device)
                    return memo[which]
                res = _task(model, meta_optim, get, num_epochs, _config, rank, shared_variables)
                if res is None:
                    continue
                first_train, first_losses, meta_loss, meta_losses, vis_rows = res
                metrics['meta_loss'][seq_name].append(meta_loss)
                metrics['meta_losses'][seq_name].append(meta_losses)
                metrics['train_loss'][seq_name].append(first_train)
                metrics['train_losses'][seq_name].append(first_losses)
                vis[seq_name].append(vis_rows)
                for name, param in meta_optim.named_parameters():
                    if param.grad is not None:
                        shared_meta_optim_grads[name] += param.grad.cpu()
            shared_dict['seqs_metrics'] = metrics
            shared_dict['vis_data_seqs'] = vis
            shared_dict['global_rng_state'] = global_rng_state
            shared_dict['sub_iter_done'] = True
            if once:
                return


def _task(model, meta_optim, get_batch, num_epochs, _config, rank, shared_variables):
    """meta_run.py:109-226 for one task; .grad of theta_0 / lambda is accumulated.  -> None when the meta loss is NaN."""
    meta_optim.train()
    meta_optim.reset()
    meta_optim.zero_grad()
    hist, losses_hist, vis_rows = [], [], []
    meta_loss = meta_losses = None
    nan = False
    for epoch in range(1, num_epochs + 1):
        if _config['increase_seed_per_meta_run']:
            set_random_seeds(_config['seed'] + rank + epoch + shared_variables['meta_iter'])
        else:
            set_random_seeds(_config['seed'] + rank + epoch)
        model.train_without_dropout()
        loader_draws(2)                                  # `for train_batch in train_loader`
        train_loss, train_losses = model(*get_batch("train"))
        losses_hist.append({k: v.item() for k, v in train_losses.items()})
        hist.append(train_loss.item())
        meta_optim.set_train_loss(train_loss)
        meta_optim.step(train_loss)
        vis_rows.append([hist[-1], 0.0, meta_optim.state_lr.cpu().detach().numpy()])
        stop = early_stopping(hist, **_config['train_early_stopping_cfg']) or epoch == num_epochs
        if not epoch % _config['bptt_epochs'] or stop:
            loader_draws(1)                              # `for meta_batch in meta_loader`
            meta_loss_t, meta_losses_t = model(*get_batch("meta"))
            nan = bool(torch.isnan(meta_loss_t).any())
            if nan:
                stop = True
            meta_loss_t.backward()
            meta_loss, meta_losses = meta_loss_t.item(), {k: v.item() for k, v in meta_losses_t.items()}
            if not stop:
                meta_optim.reset(keep_state=True)
        if stop:
            meta_optim.reset()
            break
    if nan:
        return None
    return hist[0], losses_hist[0], meta_loss, meta_losses, vis_rows
