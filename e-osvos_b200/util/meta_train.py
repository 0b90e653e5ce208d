"""Meta-training step (BASELINE.json config 5): first-order BPTT through the fine-tune iterations, meta-gradient
exchange, outer RAdam update.  Mirrors reference src/util/meta_run.py:96-243 (per-task inner loop, bptt_loss.backward,
gradient hand-off) and src/train_meta.py:110-127,361-373 (per-group lr / weight decay, average over the meta batch,
clamp, RAdam step, clamp_init_lr).

Where the reference sums gradients through CPU shared memory with an unlocked `+=` (meta_run.py:237-238) and steps
RAdam on the CPU, here every rank packs its meta-gradients into ONE flat fp32 buffer, a single NCCL all-reduce (sum)
over NVLink combines them, and the outer update runs replicated on every rank as fused CUDA launches
(libeosvos_b200: eosvos_radam_step), so no parameter broadcast is needed afterwards.
"""
import math

import torch

from .. import kernels as K
from .evaluate import set_random_seeds


class FusedRAdam:
    """RAdam (reference src/util/radam.py:28-94, degenerated_to_sgd=True) over MetaOptimizer.named_parameters()."""

    def __init__(self, meta_optim, model_init_lr=1e-5, log_init_lr_lr=1e-5, lr=1e-3, model_init_weight_decay=1e-3,
                 betas=(0.9, 0.999), eps=1e-8):
        self.meta_optim = meta_optim
        self.betas, self.eps = betas, eps
        self.groups = []
        for n, p in meta_optim.named_parameters():
            wd = 0.0
            if 'model_init' in n:
                g_lr, wd = model_init_lr, model_init_weight_decay
            elif 'log_init_lr' in n:
                g_lr = log_init_lr_lr
            else:
                g_lr = lr
            self.groups.append({"name": n, "param": p, "lr": g_lr, "wd": wd, "m": None, "v": None})
        self.step_count = 0

    def _rectification(self):
        beta1, beta2 = self.betas
        t = self.step_count
        beta2_t = beta2 ** t
        n_sma_max = 2 / (1 - beta2) - 1
        n_sma = n_sma_max - 2 * t * beta2_t / (1 - beta2_t)
        if n_sma >= 5:
            step = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max
                             / (n_sma_max - 2)) / (1 - beta1 ** t)
            return True, step
        return False, 1.0 / (1 - beta1 ** t)

    def step(self, flat_grad, offsets, gscale=1.0, grad_clip=None, lr_clamp=None):
        """flat_grad: summed meta-gradients of all parameters (fp32, packed at `offsets`)."""
        self.step_count += 1
        rect, step_size = self._rectification()
        for g, off in zip(self.groups, offsets):
            p = g["param"].data
            n = p.numel()
            if g["m"] is None:
                g["m"] = torch.zeros_like(p)
                g["v"] = torch.zeros_like(p)
            clamp = lr_clamp if ('log_init_lr' in g["name"] and lr_clamp is not None) else None
            K.radam_step(p, flat_grad[off:off + n], g["m"], g["v"], gscale=gscale, clip=grad_clip,
                         beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, lr=g["lr"], wd=g["wd"],
                         step_size=step_size * g["lr"], rectified=rect, clamp=clamp)
        # the kernel wrote through raw pointers: autograd's version counters did not move, so the 16-bit operand
        # layouts cached per (tensor, version) for theta_0 are stale now
        from .. import ops
        ops.clear_prep_cache()


def task_meta_gradients(model, meta_optim, train_batch, meta_batch, num_epochs=5, bptt_epochs=5, seed=1, rank=0,
                        meta_iter=0):
    """One task of meta_run.py:109-226: `num_epochs` fine-tune steps on the train frame with the graph kept through
    the fused update, meta loss on the meta frame every `bptt_epochs`, backward into theta_0 / lambda (.grad is
    accumulated, as in the reference).  Returns (first train loss, last meta loss)."""
    meta_optim.train()
    meta_optim.reset()
    first_train, meta_val = None, None
    for epoch in range(1, num_epochs + 1):
        set_random_seeds(seed + rank + epoch + meta_iter)
        model.train_without_dropout()
        train_loss, _ = model(*train_batch)
        if first_train is None:
            first_train = train_loss.detach()
        meta_optim.set_train_loss(train_loss)
        meta_optim.step(train_loss)
        last = epoch == num_epochs
        if epoch % bptt_epochs == 0 or last:
            meta_loss, _ = model(*meta_batch)
            meta_val = meta_loss.detach()
            if not torch.isnan(meta_loss).any():
                meta_loss.backward()
            if not last:
                meta_optim.reset(keep_state=True)
    meta_optim.reset()
    return first_train, meta_val


def pack_meta_gradients(meta_optim):
    """-> (flat fp32 buffer with every parameter's .grad, offsets); missing grads count as zero."""
    params = [p for _, p in meta_optim.named_parameters()]
    offsets, total = [], 0
    for p in params:
        offsets.append(total)
        total += p.numel()
    flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
    for p, off in zip(params, offsets):
        if p.grad is not None:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
    return flat, offsets


def meta_iteration(model, meta_optim, radam, tasks, meta_batch_size, grad_clip=None, num_epochs=5, bptt_epochs=5,
                   seed=1, meta_iter=0, process_group=None):
    """One meta-iteration (train_meta.py:254-377): this rank's tasks -> flat gradient -> NCCL all-reduce(sum) ->
    1/meta_batch_size, clamp, fused RAdam, clamp_init_lr -- replicated on every rank."""
    import torch.distributed as dist
    rank = dist.get_rank(process_group) if (dist.is_available() and dist.is_initialized()) else 0
    meta_optim.zero_grad()
    losses = []
    for train_batch, meta_batch in tasks:
        losses.append(task_meta_gradients(model, meta_optim, train_batch, meta_batch, num_epochs, bptt_epochs, seed, rank,
                                          meta_iter))
    flat, offsets = pack_meta_gradients(meta_optim)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=process_group)      # the ONE exchange step of the path
    lo = -33.0 if meta_optim._use_log_init_lr else 0.0
    hi = float("inf")
    if meta_optim._max_lr is not None:
        hi = math.log(meta_optim._max_lr) if meta_optim._use_log_init_lr else float(meta_optim._max_lr)
    radam.step(flat, offsets, gscale=1.0 / meta_batch_size, grad_clip=grad_clip, lr_clamp=(lo, hi))
    meta_optim.zero_grad()
    meta_optim.reset()
    return losses
