"""torch.autograd.Function wrappers around the CUDA kernels.

They exist so that the reference's MetaOptimizer.step -- which calls torch.autograd.grad(loss, params)
on parameter tensors that live in module._parameters as NON-LEAF tensors (reference
src/meta_optim/meta_optim.py:202-204, src/meta_optim/meta_model.py:78-80) -- keeps working unchanged
on top of hand-written forward/backward kernels.  Activations flow between Functions as NHWC bf16;
parameters and their gradients are fp32 in torch's layouts.  Second-order (create_graph=True) is not
supported by hand-written backward kernels: Functions are marked once_differentiable, so asking for it
raises instead of silently falling back.
"""
import weakref

import torch
from torch.autograd.function import once_differentiable

from . import kernels as K

ACT = K.ACT_DTYPE
# Static loss scale of the 16-bit backward pass.  fp32 loss gradients enter the 16-bit domain only in
# HeadFn.backward (all three fp32 heads), where they are multiplied by GRAD_SCALE; every kernel that emits a
# parameter gradient multiplies by 1/GRAD_SCALE (`alpha`), so autograd sees exact-scale fp32 gradients.
GRAD_SCALE = 1024.0 if ACT == torch.float16 else 1.0


def _inv_scale():
    return 1.0 / GRAD_SCALE

# --------------------------------------------------------------------------------------------
# parameter -> tensor-core operand layouts (bf16), cached per live parameter tensor
# --------------------------------------------------------------------------------------------
_prep_cache = {}


def _cache_get(p, kind):
    key = (id(p), kind)
    hit = _prep_cache.get(key)
    if hit is not None:
        ref, version, ptr, val = hit
        if ref() is p and version == p._version and ptr == p.data_ptr():
            return val
    return None


def _cache_put(p, kind, val):
    key = (id(p), kind)

    def _drop(_, key=key):
        _prep_cache.pop(key, None)

    _prep_cache[key] = (weakref.ref(p, _drop), p._version, p.data_ptr(), val)
    return val


def clear_prep_cache():
    _prep_cache.clear()


def _contig_strides(dims):
    ds = [1, 1, 1, 1]
    for i in (2, 1, 0):
        ds[i] = ds[i + 1] * dims[i + 1]
    return ds


# kind -> (w) -> list of (out_shape, dims, src strides, dst strides): the tensor-core operand layouts of a parameter
def _spec_conv_f(w):
    """fp32 [O,I,KH,KW] -> [O,KH,KW,I] (fprop B operand)."""
    O, I, KH, KW = w.shape
    T = KH * KW
    d = (1, O, T, I)
    return [((O, KH, KW, I), d, (0, I * T, 1, T), _contig_strides(d))]


def _spec_conv_t(w):
    """fp32 [O,I,KH,KW] -> [I,KH,KW,O] (dgrad B operand)."""
    O, I, KH, KW = w.shape
    T = KH * KW
    d = (1, I, T, O)
    return [((I, KH, KW, O), d, (0, T, 1, I * T), _contig_strides(d))]


def _spec_linear_f(inner):
    def f(w):
        """fp32 [O,K] -> [O,1,1,K]; with inner=C the K axis is re-ordered (c, s) -> (s, c) (fc6 eats NHWC pooling)."""
        O, Kd = w.shape
        if inner:
            C, S = inner, Kd // inner
            d = (1, O, S, C)
            return [((O, 1, 1, Kd), d, (0, Kd, 1, S), _contig_strides(d))]
        d = (1, 1, O, Kd)
        return [((O, 1, 1, Kd), d, (0, 0, Kd, 1), _contig_strides(d))]
    return f


def _spec_linear_t(inner):
    def f(w):
        """fp32 [O,K] -> [K,1,1,O] (dgrad operand), same K re-ordering."""
        O, Kd = w.shape
        if inner:
            C, S = inner, Kd // inner
            d = (1, S, C, O)
            return [((Kd, 1, 1, O), d, (0, 1, S, Kd), _contig_strides(d))]
        d = (1, 1, Kd, O)
        return [((Kd, 1, 1, O), d, (0, 0, 1, Kd), _contig_strides(d))]
    return f


def _spec_deconv(w):
    """fp32 [I,O,2,2] -> fwd operand [(dy,dx,o)][i] and dgrad operand [i][(dy,dx,o)]."""
    I, O = w.shape[:2]
    d1, d2 = (1, 4, O, I), (1, I, 4, O)
    return [((4 * O, I), d1, (0, 1, 4, 4 * O), _contig_strides(d1)), ((I, 4 * O), d2, (0, 4 * O, 1, 4), _contig_strides(d2))]


def _spec_stem(w):
    """fp32 [O,3,7,7] -> [O,1,1,Kp] with k = t*3 + c, zero padded to a multiple of 64."""
    O, I, KH, KW = w.shape
    T = KH * KW
    Kp = ((T * I + 63) // 64) * 64
    return [((O, 1, 1, Kp), (1, O, T, I), (0, I * T, 1, T), (0, Kp, I, 1))]


def _spec_for(kind):
    if kind == "f":
        return _spec_conv_f
    if kind == "t":
        return _spec_conv_t
    if kind == "dc":
        return _spec_deconv
    if kind == "stem":
        return _spec_stem
    if kind[0] == "lf":
        return _spec_linear_f(kind[1])
    if kind[0] == "lt":
        return _spec_linear_t(kind[1])
    raise KeyError(kind)


def _prep_jobs(wd, kind):
    """-> (outputs, tiled jobs, element-wise jobs) for one parameter and operand kind."""
    outs, tiled, plain = [], [], []
    for shape, dims, ss, ds in _spec_for(kind)(wd):
        alloc = torch.zeros if kind == "stem" else torch.empty
        out = alloc(shape, device=wd.device, dtype=ACT)
        xyz = K._xyz_of_spec(dims, ss, ds)
        if xyz is not None:
            tiled.append((wd, out) + xyz)
        else:
            plain.append((wd, out, dims, ss, ds))
        outs.append(out)
    return outs, tiled, plain


def prep_many(requests):
    """requests: iterable of (parameter, kind).  Builds every missing operand layout with ONE kernel launch."""
    tiled, plain, pending = [], [], []
    for w, kind in requests:
        if _cache_get(w, kind) is not None:
            continue
        wd = w.detach()
        if not wd.is_contiguous():
            wd = wd.contiguous()
        outs, t, p = _prep_jobs(wd, kind)
        tiled += t
        plain += p
        pending.append((w, kind, outs[0] if len(outs) == 1 else tuple(outs)))
    K.weight_prep_multi(tiled)
    K.permute_cast_multi(plain)
    for w, kind, val in pending:
        _cache_put(w, kind, val)


# Operands prepared for the CUDA-graphed trunk: {(id(parameter tensor), kind): operand}.  Set by MaskRCNN while its
# trunk function runs (warm-up, capture); the operands live in persistent buffers filled by ONE captured launch.
_scope = None


def build_prep_plan(requests):
    """requests: list of (static parameter tensor, kind) -> (WeightPrepPlan, {(index, kind): operand}).  Every
    layout must be expressible as a tiled permutation (all conv / stem kinds are)."""
    jobs, vals = [], {}
    for i, (w, kind) in enumerate(requests):
        outs, t, p = _prep_jobs(w.detach(), kind)
        if p:
            raise K._lib.EosvosError(f"operand kind {kind} has no tiled layout")
        jobs += t
        vals[(i, kind)] = outs[0] if len(outs) == 1 else tuple(outs)
    return K.WeightPrepPlan(jobs, requests[0][0].device), vals


def _prep_direct(w, kind):
    """Capture-safe variant: per-tensor launches with by-value parameters (no host tables, no caching)."""
    wd = w.detach()
    if not wd.is_contiguous():
        wd = wd.contiguous()
    outs = []
    for shape, dims, ss, ds in _spec_for(kind)(wd):
        out = (torch.zeros if kind == "stem" else torch.empty)(shape, device=wd.device, dtype=ACT)
        K.permute_cast(wd, out, dims, ss, ds)
        outs.append(out)
    return outs[0] if len(outs) == 1 else tuple(outs)


def _prep_one(w, kind):
    if _scope is not None:
        hit = _scope.get((id(w), kind))
        if hit is not None:
            return hit
    if torch.cuda.is_current_stream_capturing():
        return _prep_direct(w, kind)
    hit = _cache_get(w, kind)
    if hit is None:
        prep_many([(w, kind)])
        hit = _cache_get(w, kind)
    return hit


def prep_conv_w(w):
    return _prep_one(w, "f")


def prep_conv_wt(w):
    return _prep_one(w, "t")


def prep_linear_w(w, inner=0):
    return _prep_one(w, ("lf", inner))


def prep_linear_wt(w, inner=0):
    return _prep_one(w, ("lt", inner))


def _fused_head_w(ws, pad_to):
    """Stacks several [O_i, K(,1,1)] fp32 weights into one bf16 [pad_to,1,1,K] (zero rows beyond sum O_i)
    and the transposed [K,1,1,64] operand (zero-padded columns) used by the backward GEMMs."""
    kind = ("head", pad_to)
    if _scope is not None:                                   # graphed trunk: built once per trunk call, reused by the
        hit = _scope.get((id(ws[0]), kind))                  # other pyramid levels and by the backward
        if hit is not None:
            return hit
    capturing = torch.cuda.is_current_stream_capturing()     # inside a CUDA graph the operands must be rebuilt from
    hit = None if capturing else _cache_get(ws[0], kind)     # the graph's static inputs on every replay: no caching
    if hit is not None and all(a is b() and a._version == v for a, (b, v) in zip(ws[1:], hit[2])):
        return hit[0], hit[1]
    flat = torch.cat([w.detach().reshape(w.shape[0], -1) for w in ws], 0)
    O, Kd = flat.shape
    wf = torch.zeros((pad_to, 1, 1, Kd), device=flat.device, dtype=ACT)
    wf.view(pad_to, Kd)[:O] = flat.to(ACT)
    wt = torch.zeros((Kd, 1, 1, 64), device=flat.device, dtype=ACT)
    wt.view(Kd, 64)[:, :O] = flat.t().to(ACT)
    if _scope is not None:
        _scope[(id(ws[0]), kind)] = (wf, wt)
    elif not capturing:
        _cache_put(ws[0], kind, (wf, wt, [(weakref.ref(w), w._version) for w in ws[1:]]))
    return wf, wt


def _gn_cpg_ok(C):
    return C % 32 == 0


# --------------------------------------------------------------------------------------------
# conv (+bias) (+residual, optionally on the 2x coarser grid) (+ReLU)
# --------------------------------------------------------------------------------------------
class Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, res, stride, pad, relu, res_half, fork=False):
        """fork=True (stride 1): also returns x itself; every OTHER consumer of x reads that alias, so in backward the
        gradient they produce arrives here and is added inside the dgrad epilogue (no separate sum kernel)."""
        wf = prep_conv_w(w)
        y = K.conv2d_fprop(x, wf, bias.detach() if bias is not None else None, res, stride=stride, pad=pad, relu=relu,
                           res_half=res_half)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (stride, pad, relu, res_half, bias is not None, res is not None)
        if fork:
            ctx.set_materialize_grads(False)
            return y, x
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dxa=None):
        x, w, y = ctx.saved_tensors
        stride, pad, relu, res_half, has_bias, has_res = ctx.cfg
        if dy is None:                       # only the alias was used
            return dxa, None, None, None, None, None, None, None, None
        dy = dy.contiguous()
        if relu:
            dy = K.relu_bwd(dy, y)
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = K.conv2d_dgrad(dy, prep_conv_wt(w), x.shape[1:3], stride=stride, pad=pad,
                                acc=dxa.contiguous() if dxa is not None else None)
        if ctx.needs_input_grad[1]:
            dw = K.conv2d_wgrad(x, dy, w.shape[2:], stride=stride, pad=pad, alpha=_inv_scale())
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy.view(-1, dy.shape[-1]), alpha=_inv_scale())
        if has_res and ctx.needs_input_grad[3]:
            dres = K.sum2x2(dy) if res_half else dy
        return dx, dw, db, dres, None, None, None, None, None


def conv2d(x, w, bias=None, res=None, stride=1, pad=0, relu=False, res_half=False, fork=False):
    if fork:
        assert stride == 1
    return Conv2dFn.apply(x, w, bias, res, stride, pad, relu, res_half, fork)


# --------------------------------------------------------------------------------------------
# conv -> GroupNorm(32) (+residual) (+ReLU): statistics come out of the conv epilogue
# --------------------------------------------------------------------------------------------
class ConvGnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, gamma, beta, res, stride, pad, relu, fork=False):
        wf = prep_conv_w(w)
        N = x.shape[0]
        # statistics ride in the conv epilogue when the K loop is long enough to hide them: measured on every
        # ResNet-50 shape (round-1 sweep), the fused form wins whenever K >= 4 * Cout and loses or ties
        # for the expanding 1x1 convs (K <= Cout / 2: the epilogue, whose length grows with Cout, is the bound there
        # and a separate pass over the mostly L2-resident output is cheaper)
        if w.shape[1] * w.shape[2] * w.shape[3] >= 2 * w.shape[0]:
            sums = K.zero_pool.take((N, 32, 2), x.device)
            z = K.conv2d_fprop(x, wf, stride=stride, pad=pad, gn_sum=sums)
        else:
            z = K.conv2d_fprop(x, wf, stride=stride, pad=pad)
            sums = K.gn_stats(z)
        g, b = gamma.detach(), beta.detach()
        y = K.gn_apply(z, sums, g, b, res, relu=relu)
        mode = 0 if not relu else (2 if res is not None else 1)
        ctx.save_for_backward(x, w, gamma, beta, z, sums, y if mode == 2 else None)
        ctx.cfg = (stride, pad, mode, res is not None)
        if fork:                             # see Conv2dFn.forward
            ctx.set_materialize_grads(False)
            return y, x
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dxa=None):
        x, w, gamma, beta, z, sums, y = ctx.saved_tensors
        stride, pad, mode, has_res = ctx.cfg
        if dy is None:
            return dxa, None, None, None, None, None, None, None, None
        dy = dy.contiguous()
        dz, dres, dgamma, dbeta = K.gn_backward(z, sums, gamma.detach(), beta.detach(), dy, yout=y, mask_mode=mode,
                                                want_dres=has_res and ctx.needs_input_grad[4], alpha=_inv_scale())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = K.conv2d_dgrad(dz, prep_conv_wt(w), x.shape[1:3], stride=stride, pad=pad,
                                acc=dxa.contiguous() if dxa is not None else None)
        elif dxa is not None:
            dx = dxa
        if ctx.needs_input_grad[1]:
            dw = K.conv2d_wgrad(x, dz, w.shape[2:], stride=stride, pad=pad, alpha=_inv_scale())
        return dx, dw, dgamma, dbeta, dres, None, None, None, None


def conv_gn(x, w, gamma, beta, res=None, stride=1, pad=0, relu=True, fork=False):
    if fork:
        assert stride == 1
    return ConvGnFn.apply(x, w, gamma, beta, res, stride, pad, relu, fork)


# --------------------------------------------------------------------------------------------
# stem: 7x7/2 conv (explicit im2col, K = 147 -> 192) -> GN -> ReLU.  No data gradient.
# --------------------------------------------------------------------------------------------
def _prep_stem_w(w):
    return _prep_one(w, "stem")


class StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x8, w, gamma, beta):
        O, I, KH, KW = w.shape
        wf = _prep_stem_w(w)
        Kp = wf.shape[-1]
        N = x8.shape[0]
        col, Ho, Wo = K.im2col_stem(x8, KH, KW, 2, 3, Kp)
        sums = K.zero_pool.take((N, 32, 2), x8.device)
        z = K.conv2d_fprop(col.view(N, 1, Ho * Wo, Kp), wf, gn_sum=sums).view(N, Ho, Wo, O)
        y = K.gn_apply(z, sums, gamma.detach(), beta.detach(), None, relu=True)
        ctx.save_for_backward(col, w, gamma, beta, z, sums)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        col, w, gamma, beta, z, sums = ctx.saved_tensors
        O, I, KH, KW = w.shape
        dz, _, dgamma, dbeta = K.gn_backward(z, sums, gamma.detach(), beta.detach(), dy.contiguous(), mask_mode=1,
                                             alpha=_inv_scale())
        dw = None
        if ctx.needs_input_grad[1]:
            dw = K.zero_pool.take(tuple(w.shape), w.device)
            T = KH * KW
            # column n = t*I + c  ->  torch offset c*T + t
            # col has K = T*I zero-padded to a multiple of 64: only the first T*I columns map into dw
            K.gemm_wgrad(col, dz.view(-1, O), dw, s_m=I * T, n_inner=I, s_n_inner=T, s_n_outer=1, alpha=_inv_scale(),
                         n_valid=T * I)
        return None, dw, dgamma, dbeta


def stem(x8, w, gamma, beta):
    return StemFn.apply(x8, w, gamma, beta)


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, arg = K.maxpool_fwd(x, 3, 2, 1)
        ctx.save_for_backward(arg)
        ctx.in_shape = tuple(x.shape)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        return K.maxpool_bwd(arg, dy.contiguous(), ctx.in_shape, 3, 2, 1)


def maxpool3x3s2(x):
    return MaxPoolFn.apply(x)


class Subsample2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.in_shape = tuple(x.shape)
        return K.subsample2(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return K.subsample2_bwd(dy.contiguous(), ctx.in_shape)


def subsample2(x):
    return Subsample2Fn.apply(x)


# --------------------------------------------------------------------------------------------
# Linear (+bias) (+ReLU) on [R, K] bf16 rows
# --------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, relu, inner):
        R, Kd = x.shape
        O = w.shape[0]
        wf = prep_linear_w(w, inner)
        y = K.conv2d_fprop(x.view(1, 1, R, Kd), wf, bias.detach() if bias is not None else None, relu=relu).view(R, O)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (relu, inner, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        relu, inner, has_bias = ctx.cfg
        R, Kd = x.shape
        O = w.shape[0]
        dy = dy.contiguous()
        if relu:
            dy = K.relu_bwd(dy, y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.conv2d_dgrad(dy.view(1, 1, R, O), prep_linear_wt(w, inner), (1, R)).view(R, Kd)
        if ctx.needs_input_grad[1]:
            if inner:
                # x columns are (position, channel)-ordered, the weight is (channel, position)-ordered: reduce in x's
                # order (contiguous columns -> vector REDs), then one transposing copy into torch's order
                tmp = K.zero_pool.take((O, Kd), x.device)
                K.gemm_wgrad(x, dy, tmp, s_m=Kd, alpha=_inv_scale())
                dw = torch.empty((O, Kd), device=x.device, dtype=torch.float32)
                pos = Kd // inner
                K.permute_cast(tmp, dw, (O, inner, pos, 1), (Kd, 1, inner, 0), (Kd, pos, 1, 0))
            else:
                dw = K.zero_pool.take((O, Kd), x.device)
                K.gemm_wgrad(x, dy, dw, s_m=Kd, alpha=_inv_scale())
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy, alpha=_inv_scale())
        return dx, dw, db, None, None


def linear(x, w, bias=None, relu=False, inner=0):
    return LinearFn.apply(x, w, bias, relu, inner)


# --------------------------------------------------------------------------------------------
# fused small-N heads: several 1x1 convs / Linears sharing one input, fp32 output [rows, 16]
# --------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    """y[rows, 16] (fp32) = x[rows, K] @ cat(ws)^T + cat(bs); columns beyond sum(O_i) are zero."""

    @staticmethod
    def forward(ctx, x, *wb):
        n = len(wb) // 2
        ws, bs = wb[:n], wb[n:]
        rows, Kd = x.shape
        wf, _ = _fused_head_w(ws, 16)
        bias = torch.zeros(16, device=x.device, dtype=torch.float32)
        o = 0
        for b in bs:
            bias[o:o + b.numel()] = b.detach()
            o += b.numel()
        y = K.conv2d_fprop(x.view(1, 1, rows, Kd), wf, bias, out_fp32=True).view(rows, 16)
        ctx.save_for_backward(x, *ws)
        ctx.n = n
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x = ctx.saved_tensors[0]
        ws = ctx.saved_tensors[1:]
        n = ctx.n
        rows, Kd = x.shape
        dyp = torch.zeros((rows, 64), device=x.device, dtype=ACT)
        dyp[:, :16] = dy * GRAD_SCALE        # fp32 -> 16-bit domain: apply the loss scale here
        _, wt = _fused_head_w(ws, 16)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = K.conv2d_dgrad(dyp.view(1, 1, rows, 64), wt, (1, rows)).view(rows, Kd)
        dwf = K.zero_pool.take((64, Kd), x.device)
        K.gemm_wgrad(x, dyp, dwf, s_m=Kd, alpha=_inv_scale())
        dbf = dy.sum(0)
        dws, dbs = [], []
        o = 0
        for w in ws:
            O = w.shape[0]
            dws.append(dwf[o:o + O].reshape(w.shape))
            dbs.append(dbf[o:o + O])
            o += O
        return (dx, *dws, *dbs)


def fused_heads(x, ws, bs):
    return HeadFn.apply(x, *ws, *bs)


# --------------------------------------------------------------------------------------------
# 2x2/2 transposed conv (+bias) (+ReLU)
# --------------------------------------------------------------------------------------------
def _prep_deconv(w):
    return _prep_one(w, "dc")


class Deconv2x2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, relu):
        wf, _ = _prep_deconv(w)
        y = K.deconv2x2_fprop(x, wf, bias.detach().repeat(4).contiguous() if bias is not None else None, relu=relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (relu, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        relu, has_bias = ctx.cfg
        dy = dy.contiguous()
        if relu:
            dy = K.relu_bwd(dy, y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.deconv2x2_dgrad(dy, _prep_deconv(w)[1])
        if ctx.needs_input_grad[1]:
            dw = K.deconv2x2_wgrad(x, dy, alpha=_inv_scale())
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy.view(-1, dy.shape[-1]), alpha=_inv_scale())
        return dx, dw, db, None


def deconv2x2(x, w, bias=None, relu=False):
    return Deconv2x2Fn.apply(x, w, bias, relu)


# --------------------------------------------------------------------------------------------
# multi-scale RoIAlign over the 4 FPN levels
# --------------------------------------------------------------------------------------------
class RoiAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, P, scales, f0, f1, f2, f3):
        feats = [f0, f1, f2, f3]
        out = K.roi_align_fwd(feats, scales, rois, P)
        ctx.save_for_backward(rois)
        ctx.cfg = (P, tuple(scales), [tuple(f.shape) for f in feats])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (rois,) = ctx.saved_tensors
        P, scales, shapes = ctx.cfg
        grads = K.roi_align_bwd(dout.contiguous(), shapes, scales, rois, P)
        return (None, None, None, *[g.to(ACT) for g in grads])


def roi_align(feats, scales, rois, P):
    return RoiAlignFn.apply(rois, P, scales, *feats)


# --------------------------------------------------------------------------------------------
# mask loss (forward and gradient produced together by one kernel)
# --------------------------------------------------------------------------------------------
class MaskLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, targets, kind):
        loss, dlogits = K.mask_loss(logits.contiguous(), labels, targets, kind)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None


def mask_loss(logits, labels, targets, kind):
    return MaskLossFn.apply(logits, labels, targets, kind)


class MaskLossWeightedFn(torch.autograd.Function):
    """sum_r weights[r] * lovasz_hinge(logits[r], targets[r]): the reference's mean over the positive RoIs
    (loss_lovasz.py:232-250 via mask_rcnn.py:56-92) when weights = valid / n_valid over a PADDED, fixed-size RoI list
    (what lets the mask branch run as a CUDA graph)."""

    @staticmethod
    def forward(ctx, logits, labels, targets, weights):
        per, dlogits = K.mask_loss_per_roi(logits.contiguous(), labels, targets)
        ctx.save_for_backward(dlogits.mul_(weights.view(-1, 1, 1, 1)))
        return (per * weights).sum()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None


def mask_loss_weighted(logits, labels, targets, weights):
    return MaskLossWeightedFn.apply(logits, labels, targets, weights)


# --------------------------------------------------------------------------------------------
# RPN losses on the sampled anchors (tv rpn.py compute_loss), straight from the per-level head outputs
# --------------------------------------------------------------------------------------------
class RpnLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sampled, labels, matched, anchors, gt_boxes, gt_off, hw, A, beta, *head_outs):
        heads = [h.detach() for h in head_outs]
        out = K.rpn_loss(heads, hw, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta)
        ctx.save_for_backward(sampled, labels, matched, anchors, gt_boxes, gt_off, *heads)
        ctx.cfg = (list(hw), A, beta)
        return out[0], out[1]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_obj, g_box):
        sampled, labels, matched, anchors, gt_boxes, gt_off, *heads = ctx.saved_tensors
        hw, A, beta = ctx.cfg
        dys = [K.zero_pool.take(tuple(h.shape), h.device) for h in heads]
        zero = None
        if g_obj is None or g_box is None:
            zero = torch.zeros((), device=labels.device)
        K.rpn_loss(heads, hw, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta, dys=dys,
                   g_obj=(g_obj if g_obj is not None else zero).contiguous(),
                   g_box=(g_box if g_box is not None else zero).contiguous())
        return (None,) * 9 + tuple(dys)


def rpn_loss(head_outs, hw, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta=1.0 / 9):
    """(loss_objectness, loss_rpn_box_reg) of tv rpn.py compute_loss.  sampled: int64 positions in the flattened
    (image, level, pixel, anchor) order, -1 = padding (the mean runs over `sampled.numel()`: pass exactly the samples)."""
    return RpnLossFn.apply(sampled, labels, matched, anchors, gt_boxes, gt_off, hw, A, beta, *head_outs)


class RpnSparseFn(torch.autograd.Function):
    """RPN losses (tv rpn.py compute_loss) as a function of the pyramid levels and the RPN head's parameters, given the
    head's forward results (shared-conv output t and fused head output o per level, computed without autograd inside
    the trunk graph).  The backward propagates through the 1x1 heads, the ReLU and the shared 3x3 conv for the
    sampled anchors only (csrc/rpn.cu, "sparse backward of the RPN head"): same gradients as the dense path."""

    @staticmethod
    def forward(ctx, sampled, labels, matched, anchors, gt_boxes, gt_off, A, beta, ts, os_, w_conv, b_conv, w_cls, b_cls,
                w_box, b_box, *feats):
        hw = [f.shape[1] * f.shape[2] for f in feats]
        out = K.rpn_loss(list(os_), hw, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta)
        ctx.save_for_backward(sampled, labels, matched, anchors, gt_boxes, gt_off, w_conv, w_cls, w_box, *ts, *os_, *feats)
        ctx.cfg = (A, beta, len(feats))
        return out[0], out[1]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_obj, g_box):
        A, beta, L = ctx.cfg
        sv = ctx.saved_tensors
        sampled, labels, matched, anchors, gt_boxes, gt_off, w_conv, w_cls, w_box = sv[:9]
        ts, os_, feats = sv[9:9 + L], sv[9 + L:9 + 2 * L], sv[9 + 2 * L:]
        dev = labels.device
        zero = torch.zeros((), device=dev) if (g_obj is None or g_box is None) else None
        g_obj = (g_obj if g_obj is not None else zero).contiguous()
        g_box = (g_box if g_box is not None else zero).contiguous()
        C = feats[0].shape[-1]
        dt, ev, xg, dw_cls, db_cls, dw_box, db_box = K.rpn_sparse_head(
            list(os_), list(ts), list(feats), A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta, g_obj, g_box,
            w_cls.detach().reshape(A, C).contiguous(), w_box.detach().reshape(4 * A, C).contiguous(), GRAD_SCALE)
        M = dt.shape[0]
        # shared 3x3 conv: weight gradient = dt^T X_g (channels-last filter gradient layout [Cout][KH][KW][Cin]),
        # bias gradient = column sums, input gradient rows G = dt W (columns (ci, tap)), scattered to the levels
        dw = K.zero_pool.take((C, 3, 3, C), dev)
        K.gemm_wgrad(xg, dt, dw.view(C, 9 * C), s_m=9 * C, alpha=_inv_scale())
        db = K.colsum(dt, alpha=_inv_scale())
        G = K.conv2d_fprop(dt.view(1, 1, M, C), prep_conv_wt(w_conv).view(9 * C, 1, 1, C)).view(M, 9 * C)
        dfs = [torch.zeros_like(f) for f in feats]
        K.rpn_sparse_scatter(dfs, ev, G)
        return (None,) * 10 + (dw.permute(0, 3, 1, 2), db, dw_cls.view_as(w_cls), db_cls, dw_box.view_as(w_box), db_box,
                               *dfs)


def rpn_loss_sparse(feats, ts, head_outs, head, sampled, labels, matched, anchors, gt_boxes, gt_off, beta=1.0 / 9):
    """(loss_objectness, loss_rpn_box_reg); `head` = torchvision RPNHead (conv, cls_logits, bbox_pred)."""
    from torch import nn
    conv = head.conv[0][0] if isinstance(head.conv, nn.Sequential) else head.conv
    A = head.cls_logits.weight.shape[0]
    return RpnSparseFn.apply(sampled, labels, matched, anchors, gt_boxes, gt_off, A, beta, tuple(ts), tuple(head_outs),
                             conv.weight, conv.bias, head.cls_logits.weight, head.cls_logits.bias, head.bbox_pred.weight,
                             head.bbox_pred.bias, *feats)
