"""MetaOptimizer with learned initialisation and learned per-neuron learning rates -- same constructor,
attributes, state-dict keys and methods as reference src/meta_optim/meta_optim.py:10-214 -- whose
`step` applies  theta <- theta - lr (.) grad  to ALL parameter tensors with ONE fused CUDA kernel
(libeosvos_b200: eosvos_meta_update) instead of >= 402 elementwise launches.

Meta-training keeps working: the fused update is an autograd Function whose backward gives
d/d theta = upstream and d/d lr = -rowsum(upstream (.) grad) (first-order BPTT, reference
src/util/meta_run.py:124-214).  Second-order meta-gradients (second_order_gradients=True while
training) would need double-backward through hand-written kernels and raise NotImplementedError.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .. import kernels as K
from .meta_model import MetaModel


class _FusedUpdateFn(torch.autograd.Function):
    """outs_i = p_i - lr_i * g_i for every tensor i in one launch."""

    @staticmethod
    def forward(ctx, use_log, n, home, keep_grads, *tensors):
        params, grads, lrs = tensors[:n], tensors[n:2 * n], tensors[2 * n:]
        ps = [p.detach().contiguous() for p in params]
        gs = [g.detach() if (g.is_contiguous() or (g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last)))
              else g.detach().contiguous() for g in grads]
        ls = [l.detach().contiguous() for l in lrs]
        # one arena for all updated tensors; every tensor starts on a 256-byte boundary so the kernel's float4 path
        # applies to all of them (an unaligned view would fall back to scalar accesses).  `home` = the model's own
        # stable arena (MaskRCNN.theta_home): the update then lands where the graphed trunk reads its parameters,
        # in place from the second step on (element i is read and written by the same thread).
        if home is not None:
            arena, offs = home
        else:
            offs, total = [], 0
            for p in ps:
                offs.append(total)
                total += (p.numel() + 63) // 64 * 64
            arena = torch.empty(total, device=ps[0].device, dtype=torch.float32)
        outs = [arena[o:o + p.numel()].view(p.shape) for o, p in zip(offs, ps)]
        K.meta_update(K.MetaUpdatePlan(ps, gs, ls, outs), use_log)
        ctx.use_log, ctx.n = use_log, n
        if keep_grads:
            # meta-training back-propagates through this update later (d/d lr = -rowsum(upstream * g)), after the
            # next forward/backward has overwritten the static gradient buffers of the CUDA-graphed trunk: keep copies
            gs = [g.clone() for g in gs]
        ctx.save_for_backward(*gs, *ls)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *douts):
        n = ctx.n
        saved = ctx.saved_tensors
        gs, ls = saved[:n], saved[n:]
        live = [i for i in range(n) if douts[i] is not None]
        dls = [None] * n
        want = [i for i in live if ctx.needs_input_grad[4 + 2 * n + i]]
        if want:
            # d L / d lr = -rowsum(upstream (.) grad) for every tensor in one launch (meta_update.cu::lr_grad_kernel)
            ds = [douts[i] if (douts[i].is_contiguous() or (douts[i].dim() == 4 and douts[i].is_contiguous(
                memory_format=torch.channels_last))) else douts[i].contiguous() for i in want]
            res = K.lr_grad(ds, [gs[i] for i in want], [ls[i] for i in want], ctx.use_log)
            for i, r in zip(want, res):
                dls[i] = r
        dps = [douts[i] for i in range(n)]
        return (None, None, None, None, *dps, *([None] * n), *dls)


def remap_reference_checkpoint(state_dict):
    """Renames the keys of a checkpoint written by the reference under torchvision 0.4 (README.md:23) to the module
    names of the installed torchvision: `mask_head.mask_fcn{i}` -> `mask_head.{i-1}.0`,
    `rpn.head.conv` -> `rpn.head.conv.0.0`, `fpn.inner_blocks.{i}` / `fpn.layer_blocks.{i}` -> `....{i}.0`.  Works on
    model state dicts ('.'-separated) and on MetaOptimizer state dicts, whose keys embed the parameter name with
    '-' separators (`model_init_<name>`, `log_init_lr_<name>`, meta_optim.py:65,78).  Keys already in the new naming
    are left alone."""
    import re
    out = type(state_dict)()
    for k, v in state_dict.items():
        for sep in (".", "-"):
            q = re.escape(sep)
            k = re.sub(rf"mask_head{q}mask_fcn(\d+){q}", lambda m: f"mask_head{sep}{int(m.group(1)) - 1}{sep}0{sep}", k)
            k = re.sub(rf"rpn{q}head{q}conv{q}(weight|bias)$", rf"rpn{sep}head{sep}conv{sep}0{sep}0{sep}\1", k)
            k = re.sub(rf"fpn{q}(inner_blocks|layer_blocks){q}(\d+){q}(weight|bias)$",
                       rf"fpn{sep}\1{sep}\2{sep}0{sep}\3", k)
        out[k] = v
    return out


class MetaOptimizer(nn.Module):

    def __init__(self, model, init_lr, learn_model_init, second_order_gradients, lr_hierarchy_level,
                 use_log_init_lr, max_lr):
        super(MetaOptimizer, self).__init__()
        self._optim = None
        self._train_loss = None
        self._device = 'cpu'
        self._learn_model_init = learn_model_init
        self._second_order_gradients = second_order_gradients
        self._use_log_init_lr = use_log_init_lr
        self._max_lr = max_lr
        self._lr_hierarchy_level = lr_hierarchy_level

        self.meta_model = MetaModel(model)

        def _maybe_log(t):
            return t.log() if self._use_log_init_lr else t

        # learning-rate parameters, meta_optim.py:28-69 (same RNG consumption: one rand_like per tensor)
        if lr_hierarchy_level == 'SINGLE':
            self.log_init_lr = torch.nn.Parameter(_maybe_log(torch.ones(1, 1).mul(init_lr)))
        elif lr_hierarchy_level == 'TENSOR':
            lr = torch.ones(self.meta_model.num_param_groups, 1).mul(init_lr)
            lr += torch.rand_like(lr).sub(0.5) * init_lr
            self.log_init_lr = torch.nn.Parameter(_maybe_log(lr))
        elif lr_hierarchy_level in ('PARAM', 'NEURON'):
            self.log_init_lr = []
            for name, param in model.named_parameters():
                if not param.requires_grad:
                    continue
                if lr_hierarchy_level == 'PARAM':
                    lr = torch.ones_like(param).mul(init_lr)
                else:
                    lr = torch.ones((param.shape[0],) + (1,) * (len(param.shape) - 1)).mul(init_lr)
                lr += torch.rand_like(lr).sub(0.5) * init_lr
                lr = torch.nn.Parameter(_maybe_log(lr))
                self.register_parameter(f"log_init_lr_{name.replace('.', '-')}", lr)
                self.log_init_lr.append(lr)
        else:
            raise NotImplementedError

        # learned initialisation theta_0, meta_optim.py:71-78
        self._model_init = OrderedDict()
        for name, param in model.named_parameters():
            if param.requires_grad:
                self._model_init[name] = param
        if self._learn_model_init:
            for name, param in self._model_init.items():
                self.register_parameter(f"model_init_{name.replace('.', '-')}", param)

        self.state = {}
        self._init_state()

    def load_state_dict(self, state_dict, strict=True, **kw):
        """As nn.Module.load_state_dict; checkpoints written by the reference under torchvision 0.4 module names are
        accepted (remap_reference_checkpoint)."""
        own = set(super(MetaOptimizer, self).state_dict().keys())
        if any(k not in own for k in state_dict):
            remapped = remap_reference_checkpoint(state_dict)
            if sum(k in own for k in remapped) > sum(k in own for k in state_dict):
                state_dict = remapped
        return super(MetaOptimizer, self).load_state_dict(state_dict, strict=strict, **kw)

    # ---- meta_optim.py:83-107
    def _lr_values(self, lrs):
        if isinstance(lrs, list):
            vals = [l.exp().mean() if self._use_log_init_lr else l.mean() for l in lrs]
            return vals
        return lrs.exp() if self._use_log_init_lr else lrs

    @property
    def init_lr(self):
        v = self._lr_values(self.log_init_lr)
        return torch.Tensor(v) if isinstance(v, list) else v

    @property
    def state_lr(self):
        v = self._lr_values(self.state["log_lr"])
        return torch.tensor(v) if isinstance(v, list) else v

    def init_zero_grad(self):
        output = 0.0
        for param in self.parameters():
            output = output + param.mean()
        output.backward()
        self.zero_grad()

    # ---- meta_optim.py:116-133
    def clamp_init_lr(self):
        min_clamp = -33 if self._use_log_init_lr else 0
        max_clamp = None
        if self._max_lr is not None:
            max_clamp = torch.log(torch.tensor(self._max_lr)) if self._use_log_init_lr else self._max_lr
        if isinstance(self.log_init_lr, list):
            self.log_init_lr = [l.data.clamp_(min_clamp, max_clamp) for l in self.log_init_lr]
        else:
            self.log_init_lr.data.clamp_(min_clamp, max_clamp)

    # ---- meta_optim.py:135-163
    def to(self, device):
        super(MetaOptimizer, self).to(device)
        self._device = device
        if isinstance(self.state['log_lr'], list):
            self.state['log_lr'] = [l.to(device) for l in self.state['log_lr']]
        else:
            self.state['log_lr'] = self.state['log_lr'].to(device)

    def reset(self, keep_state=False):
        if keep_state:
            if isinstance(self.log_init_lr, list):
                self.state['log_lr'] = [l.detach() for l in self.state['log_lr']]
            else:
                self.state['log_lr'] = self.state['log_lr'].detach()
            self.meta_model.detach_param_groups()
        else:
            self.meta_model.init_param_groups(self._model_init)
            self._init_state()

    def _init_state(self):
        if self._lr_hierarchy_level == 'SINGLE':
            self.state["log_lr"] = self.log_init_lr.repeat(self.meta_model.num_param_groups, 1)
        else:
            self.state["log_lr"] = self.log_init_lr
        self.state["num_steps"] = 0

    def set_train_loss(self, train_loss):
        if not self.training or not self._second_order_gradients:
            train_loss = train_loss.detach()
        self._train_loss = train_loss

    def _nonfinite_flag(self, device):
        f = getattr(self, "_nonfinite", None)
        if f is None or f.device != device:
            f = self._nonfinite = torch.zeros(1, dtype=torch.int32, device=device)
        return f

    def check_finite(self):
        """Raises FloatingPointError when a fused update since the last check produced Inf / NaN parameters -- the
        backward runs in 16-bit storage under the static loss scale `ops.GRAD_SCALE` (fp16: 1024), so an activation
        gradient above 65504 / scale overflows.  One small device->host read: call it where the host synchronises
        anyway (evaluate_sequence does after every fine-tuning round).  Lower `ops.GRAD_SCALE` (or build the bf16
        variant, EOSVOS_ACT=bf16) if it fires."""
        f = getattr(self, "_nonfinite", None)
        if f is not None and int(f.item()) != 0:
            f.zero_()
            from .. import ops
            raise FloatingPointError(
                f"non-finite parameters after the fused MetaOptimizer update: the 16-bit backward overflowed under "
                f"loss scale {ops.GRAD_SCALE}")

    def _step_eval(self, train_loss):
        """Evaluation-time step (no graph is kept through the update): theta <- theta - lr (.) grad in ONE kernel,
        in place in the model's parameter arena, without the autograd Function and with the pointer table cached
        while the addresses of parameters / gradients / learning rates stay the same.  Returns False (nothing done)
        when the model's arena does not match the current parameters."""
        model = self.meta_model.model
        slots = self.meta_model.grad_slots()
        params = [m._parameters[n] for m, n in slots]
        dev = params[0].device
        arena, offs, shapes, index = model.theta_home()
        hv = getattr(self, "_home_views", None)
        if hv is None or hv[0] is not arena:
            idx = [index.get((id(m), n)) for m, n in slots]
            if arena.device != dev or any(i is None or shapes[i] != tuple(p.shape) for i, p in zip(idx, params)):
                return False
            views = [arena[offs[i]:offs[i] + p.numel()].view(shapes[i]).requires_grad_(True) for i, p in zip(idx, params)]
            hv = self._home_views = (arena, views)
            self._plan_cache = {}
        outs = hv[1]
        lrs = self.state["log_lr"]
        if not isinstance(lrs, list):
            lrs = [lrs[i] for i in range(len(params))]
        if any(l.device != dev for l in lrs):
            return False
        grads = torch.autograd.grad(train_loss, params)
        key = (tuple(p.data_ptr() for p in params), tuple((g.data_ptr(), g.stride(1) if g.dim() == 4 else 0) for g in grads),
               tuple(l.data_ptr() for l in lrs))
        plan = self._plan_cache.get(key)
        if plan is None:
            gs = [g.detach() if (g.is_contiguous() or (g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last)))
                  else g.detach().contiguous() for g in grads]
            plan = K.MetaUpdatePlan([p.detach() for p in params], gs, [l.detach() for l in lrs], [o.detach() for o in outs])
            plan.keep = gs
            if all(a.data_ptr() == b.data_ptr() for a, b in zip(gs, grads)):
                if len(self._plan_cache) > 8:
                    self._plan_cache.clear()
                self._plan_cache[key] = plan
        K.meta_update(plan, bool(self._use_log_init_lr), self._nonfinite_flag(dev))
        if params[0] is not outs[0]:
            for (m, n), o in zip(slots, outs):
                m._parameters[n] = o
        # values changed behind the same tensor objects: the per-tensor operand cache must not survive
        from .. import ops
        ops.clear_prep_cache()
        self.state["num_steps"] += 1
        return True

    # ---- meta_optim.py:177-214, fused
    def step(self, train_loss):
        if not self.training and callable(getattr(self.meta_model.model, "theta_home", None)):
            if self._step_eval(train_loss):
                return
        if self.training and self._second_order_gradients:
            raise NotImplementedError(
                "second_order_gradients=True needs create_graph through hand-written backward kernels; "
                "the B200 path supports the reference default (first-order, cfgs/meta.yaml:40)")
        groups = list(self.meta_model.param_groups())
        params = [p for _, _, _, p in groups]
        # same tensors, same order as [p for p in model.parameters() if p.requires_grad] (meta_optim.py:202-204)
        grads = torch.autograd.grad(train_loss, params)
        lrs = self.state["log_lr"]
        if not isinstance(lrs, list):
            lrs = [lrs[i] for i in range(len(params))]
        dev = params[0].device
        lrs = [l if l.device == dev else l.to(dev) for l in lrs]
        grads = [g if g.device == dev else g.to(dev) for g in grads]
        n = len(params)
        home = None
        theta_home = getattr(self.meta_model.model, "theta_home", None)
        # The stable arena is written IN PLACE (theta_{t+1} over theta_t).  That is only sound when no autograd graph
        # still refers to theta_t: true at evaluation time (every step is followed by detach_param_groups), not in
        # meta-training, where forward graphs may be kept across steps (multi_step_bptt_loss, meta_run.py:158-178).
        if callable(theta_home) and not self.training:
            arena, offs, shapes, index = theta_home()
            idx = [index.get((id(module), n_p)) for _, module, n_p, _ in groups]
            if arena.device == dev and all(i is not None and shapes[i] == tuple(p.shape) for i, p in zip(idx, params)):
                home = (arena, [offs[i] for i in idx])
        new_params = _FusedUpdateFn.apply(bool(self._use_log_init_lr), n, home, bool(self.training), *params, *grads,
                                          *lrs)
        self.meta_model.set_param_groups(new_params)
        self.state["num_steps"] += 1
