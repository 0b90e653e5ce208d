"""MetaModel: bookkeeping of the fine-tuned parameters living in `module._parameters`
(same interface as reference src/meta_optim/meta_model.py:5-114)."""
import torch
import torch.nn as nn


class MetaModel:
    """Tracks the model whose parameters the MetaOptimizer rewrites in place of nn.Parameters."""

    def __init__(self, model):
        self.model = model

    # ---- iteration over (module name, module, parameter name, tensor) -- meta_model.py:49-56
    def param_groups(self):
        # The (module, parameter-name) slots are fixed once the model is built; walking named_modules() on every
        # call (the reference does) costs ~1.5 ms per call on R50-FPN, so the slot list is cached and only the
        # tensors currently installed in module._parameters are re-read.
        slots = getattr(self, "_slots", None)
        if slots is None:
            slots = [(n_m, module, n_p) for n_m, module in self.model.named_modules()
                     for n_p, p in module._parameters.items() if p is not None]
            self._slots = slots
        for n_m, module, n_p in slots:
            p = module._parameters[n_p]
            if p is not None and p.requires_grad:
                yield n_m, module, n_p, p

    @property
    def num_param_groups(self):
        return sum(1 for _ in self.param_groups())

    # ---- meta_model.py:62-71
    def detach_param_groups(self):
        for _, module, n_p, p in self.param_groups():
            if p.grad_fn is None and not isinstance(p, nn.Parameter):
                continue                    # already a plain leaf (the fused evaluation-time update installs those)
            d = p.detach()
            d.requires_grad = True
            module._parameters[n_p] = d

    def grad_slots(self):
        """(module, parameter name) of every fine-tuned tensor, in param_groups() order (cached: the set is fixed
        once the model is built)."""
        gs = getattr(self, "_grad_slots", None)
        if gs is None:
            gs = self._grad_slots = [(module, n_p) for _, module, n_p, _ in self.param_groups()]
        return gs

    def init_param_groups(self, group_inits):
        for n_m, module, n_p, _ in self.param_groups():
            key = f"{n_m}.{n_p}"
            if key in group_inits:
                module._parameters[n_p] = group_inits[key]

    # ---- meta_model.py:73-80 (the box-head variant is identical in the reference: `if True:`)
    def apply_param_groups_step(self, param_groups_step):
        for (_, module, n_p, p), step in zip(self.param_groups(), param_groups_step):
            module._parameters[n_p] = p - step

    apply_param_groups_step_box_head = apply_param_groups_step

    def set_param_groups(self, new_params):
        """Installs already-updated tensors (output of the fused update kernel)."""
        for (_, module, n_p, _), t in zip(self.param_groups(), new_params):
            module._parameters[n_p] = t

    # ---- helpers kept for API compatibility -- meta_model.py:15-47, 82-92
    def init_zero_grad(self):
        output = 0.0
        for param in self.model.parameters():
            if param.requires_grad:
                output = output + param.mean()
        output.backward()
        self.model.zero_grad()

    def deparameterize(self):
        for _, module, n_p, p in self.param_groups():
            module._parameters[n_p] = p.data
            module._parameters[n_p].grad = p.grad

    def get_flat_params(self):
        return torch.cat([p.view(-1) for p in self.model.parameters() if p.requires_grad]).unsqueeze(-1).detach()

    def set_flat_params(self, flat_params, keep_grads=True):
        offset = 0
        for _, module, n_p, p in self.param_groups():
            n = p.numel()
            flat = flat_params[offset:offset + n].view(*p.size())
            if keep_grads:
                module._parameters[n_p] = flat
            else:
                module._parameters[n_p].data.copy_(flat.data)
            offset += n
        self._invalidate_operand_cache()

    @staticmethod
    def _invalidate_operand_cache():
        """`.data` writes do not bump autograd's version counter, which (with identity and address) validates the
        cached 16-bit tensor-core operand layouts of a parameter (ops._prep_cache): drop them."""
        from .. import ops
        ops.clear_prep_cache()

    def copy_params_from(self, model: nn.Module):
        for dst, src in zip(self.model.parameters(), model.parameters()):
            dst.data.copy_(src.data)
        self._invalidate_operand_cache()

    def copy_params_to(self, model: nn.Module):
        for src, dst in zip(self.model.parameters(), model.parameters()):
            dst.data.copy_(src.data)
        self._invalidate_operand_cache()

    def copy_grads_from(self, model: nn.Module):
        for dst, src in zip(self.model.parameters(), model.parameters()):
            dst.grad.data.copy_(src.grad.data)
