"""Tensor-level wrappers over the C ABI (no autograd here; see ops.py).

PyTorch supplies device memory and streams only.  Activations are [N, H, W, C] ACT_DTYPE (fp16 in the
default build, bf16 with -DEOSVOS_ACT_BF16) contiguous,
parameters fp32 in the reference's (torch) layouts.
"""
import ctypes

import torch

from . import _lib
from ._lib import FLAG_OUT_FP32, FLAG_RELU, FLAG_RES_HALF, call

GN_EPS = 1e-5
# storage type of activations / tensor-core operands, fixed when the library was built
ACT_DTYPE = torch.float16 if _lib.load().eosvos_act_dtype() == 1 else getattr(torch, "bfloat16")


_raw_stream = torch._C._cuda_getCurrentRawStream


def _stream():
    # raw handle of torch's current stream on the current device (torch.cuda.current_stream() costs ~15 us)
    return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.EosvosError(f"{name} must live on a CUDA device (no CPU path exists)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.EosvosError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.EosvosError(f"{name} must be contiguous")
    idx = t.device.index
    if idx not in _lib._checked_devices:
        _lib.require_device(idx if idx is not None else torch.cuda.current_device())
        _lib._checked_devices.add(idx)
    return t


# Held by whoever captures a CUDA graph and by the augmentation worker thread around its CUDA section: allocations,
# synchronisations and frees issued by another thread while a capture is under way can invalidate it.
import threading  # noqa: E402
capture_lock = threading.RLock()


class PinnedStager:
    """Small host->device uploads (pointer tables, boxes, ids) through a ring of PINNED staging buffers with
    non_blocking copies.  A plain `torch.tensor(..., device='cuda')` / `.to(device)` from pageable memory makes the
    driver synchronise the stream before the copy -- i.e. the host would wait for all queued GPU work."""

    def __init__(self, slot_bytes=1 << 16, slots=16):
        self.slot_bytes, self.slots = slot_bytes, slots
        self.ring = None
        self.events = [None] * slots
        self.i = 0

    def put(self, array, device):
        """array: numpy array or CPU tensor -> device tensor of the same dtype / shape."""
        src = torch.from_numpy(array) if not isinstance(array, torch.Tensor) else array
        src = src.contiguous()
        nbytes = src.numel() * src.element_size()
        if nbytes == 0:
            return torch.empty(src.shape, dtype=src.dtype, device=device)
        if nbytes > self.slot_bytes:
            return src.pin_memory().to(device, non_blocking=True)
        if self.ring is None:
            self.ring = torch.empty((self.slots, self.slot_bytes), dtype=torch.uint8).pin_memory()
        j = self.i % self.slots
        self.i += 1
        if self.events[j] is not None:
            self.events[j].synchronize()          # the copy that last used this slot (16 uploads ago) is long done
        stage = self.ring[j, :nbytes].view(src.dtype).view(src.shape)
        stage.copy_(src)
        out = stage.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.events[j] = ev
        return out


stager = PinnedStager()


class TargetStatsCache:
    """Host-side (xmin, ymin, xmax, ymax, count) per object id of a label tensor, remembered for the EXACT tensor
    contents they were computed from: an entry is valid only while the same tensor object still has the same storage
    address and autograd version counter (any in-place edit bumps `_version`).  Producers that already know the
    boxes on the host (the augmentation thread, the fused inference tail) `put` them so MaskRCNN.forward skips its
    mask->box kernel and the D2H read-back."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self.entries = {}

    def put(self, t, stats, ign):
        import weakref
        key = id(t)
        if len(self.entries) >= self.capacity:
            for k in [k for k, v in self.entries.items() if v[0]() is None]:
                del self.entries[k]
            while len(self.entries) >= self.capacity:
                self.entries.pop(next(iter(self.entries)))
        self.entries[key] = (weakref.ref(t), t.data_ptr(), t._version, tuple(t.shape), stats, ign)

    def get(self, t):
        hit = self.entries.get(id(t))
        if hit is None:
            return None
        ref, ptr, version, shape, stats, ign = hit
        if ref() is t and ptr == t.data_ptr() and version == t._version and shape == tuple(t.shape):
            return stats, ign
        del self.entries[id(t)]
        return None

    def get_host(self, t):
        hit = self.get(t)
        return hit if hit is not None and not hit[0].is_cuda else None

    def get_device(self, t):
        """(stats int32 [B,K,5] ON THE DEVICE, fallback stats or None): the sync-free inference path -- the boxes of
        the propagated target never visit the host (put with a CUDA `stats` tensor; `ign` carries the fallback)."""
        hit = self.get(t)
        return hit if hit is not None and hit[0].is_cuda else None


target_stats = TargetStatsCache()


class ZeroPool:
    """Hands out zero-initialised fp32 views carved from ONE zeroed block per iteration (instead of ~200
    torch.zeros calls): accumulate-into outputs (split-K weight gradients, GroupNorm statistics, RoIAlign-bwd
    maps).  A block is freed by reference counting once every view of it is gone, so views never alias."""

    def __init__(self):
        self.block = None
        self.off = 0
        self.used = 0
        self.hint = 0
        self.cap_block = None
        self.cap_off = 0

    def reset(self):
        self.hint = max(self.hint, self.used)
        self.block, self.off, self.used = None, 0, 0

    def take(self, shape, device):
        if torch.cuda.is_current_stream_capturing():
            # inside CUDA-graph capture every buffer must come from the graph's private pool (fixed address,
            # memset recorded as a graph node): carve views from 32 MB zeroed blocks allocated during the capture
            n = 1
            for s in shape:
                n *= int(s)
            n_al = (n + 63) // 64 * 64
            if self.cap_block is None or self.cap_off + n_al > self.cap_block.numel():
                self.cap_block = torch.zeros(max(n_al, 8 << 20), device=device, dtype=torch.float32)
                self.cap_off = 0
            v = self.cap_block[self.cap_off:self.cap_off + n].view(tuple(int(s) for s in shape))
            self.cap_off += n_al
            return v
        self.cap_block = None          # not capturing: never hand out memory of a finished capture
        n = 1
        for s in shape:
            n *= int(s)
        n_al = (n + 63) // 64 * 64          # keep every view 256-byte aligned
        if self.block is None or self.block.device != device or self.off + n_al > self.block.numel():
            size = max(self.hint - self.used, n_al, 1 << 20)
            self.block = torch.zeros(size, device=device, dtype=torch.float32)
            self.off = 0
        v = self.block[self.off:self.off + n].view(shape)
        self.off += n_al
        self.used += n_al
        return v


zero_pool = ZeroPool()


# ------------------------------------------------------------------------------------------- K1
def conv2d_fprop(x, w, bias=None, res=None, *, stride=1, pad=0, relu=False, out_fp32=False, res_half=False,
                 gn_sum=None, bn_hint=0, out=None):
    """x [N,H,W,Cin] bf16, w [Cout,KH,KW,Cin] bf16 -> y [N,Ho,Wo,Cout]."""
    _chk(x, ACT_DTYPE, "x")
    _chk(w, ACT_DTYPE, "w")
    N, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2, (x.shape, w.shape)
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32 if out_fp32 else ACT_DTYPE)
    flags = (FLAG_RELU if relu else 0) | (FLAG_OUT_FP32 if out_fp32 else 0) | (FLAG_RES_HALF if res_half else 0)
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if res is not None:
        _chk(res, ACT_DTYPE, "res")
    if gn_sum is not None:
        _chk(gn_sum, torch.float32, "gn_sum")
    call("eosvos_conv2d_fprop", _ptr(x), _ptr(w), _ptr(bias), _ptr(res), _ptr(out), _ptr(gn_sum), N, H, W, Cin, Cout,
         KH, KW, stride, pad, flags, bn_hint, _stream())
    return out


def conv2d_dgrad(dy, wt, in_hw, *, stride=1, pad=0, out_fp32=False, bn_hint=0, acc=None):
    """dy [N,Ho,Wo,Cout] 16-bit, wt [Cin,KH,KW,Cout] 16-bit -> dx [N,H,W,Cin] (+ acc [N,H,W,Cin] 16-bit, stride 1)."""
    _chk(dy, ACT_DTYPE, "dy")
    _chk(wt, ACT_DTYPE, "wt")
    N, Ho, Wo, Cout = dy.shape
    Cin, KH, KW, Cout2 = wt.shape
    assert Cout == Cout2
    H, W = in_hw
    assert (H + 2 * pad - KH) // stride + 1 == Ho and (W + 2 * pad - KW) // stride + 1 == Wo
    dx = torch.empty((N, H, W, Cin), device=dy.device, dtype=torch.float32 if out_fp32 else ACT_DTYPE)
    if acc is not None:
        _chk(acc, ACT_DTYPE, "acc")
        assert tuple(acc.shape) == (N, H, W, Cin) and acc.is_contiguous()
    call("eosvos_conv2d_dgrad", _ptr(dy), _ptr(wt), _ptr(dx), _ptr(acc), N, H, W, Cin, Cout, KH, KW, stride, pad,
         FLAG_OUT_FP32 if out_fp32 else 0, bn_hint, _stream())
    return dx


def conv2d_wgrad(x, dy, ksize, *, stride=1, pad=0, alpha=1.0, bn_hint=0, split_hint=0, out=None, channels_last=None):
    """x [N,H,W,Cin], dy [N,Ho,Wo,Cout] 16-bit -> dw fp32, logical shape [Cout,Cin,KH,KW].
    channels_last (default for KH*KW > 1 when `out` is not given): memory order [Cout][KH][KW][Cin] = torch's
    channels_last strides, which lets the kernel reduce with 16-byte vector REDs; the MetaOptimizer update reads that
    layout directly (MetaUpdatePlan)."""
    _chk(x, ACT_DTYPE, "x")
    _chk(dy, ACT_DTYPE, "dy")
    KH, KW = ksize
    if KH == 1 and KW == 1 and stride == 2 and pad == 0:
        # ResNet's stride-2 1x1 projections: gather the even pixels once (a 16-bit copy of a quarter of x) and run the
        # flat pixel-major reduction on it, instead of strided TMA boxes that fetch every pixel to use one in four
        # (1024->2048 at 3x48x84: 146 us -> the cost of a plain 1x1 layer of that size)
        x, stride = subsample2(x), 1
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    if out is None:
        if channels_last is None:
            channels_last = KH * KW > 1
        if channels_last:
            out = zero_pool.take((Cout, KH, KW, Cin), x.device).permute(0, 3, 1, 2)
        else:
            out = zero_pool.take((Cout, Cin, KH, KW), x.device)
    else:
        cl = out.dim() == 4 and KH * KW > 1 and out.is_contiguous(memory_format=torch.channels_last)
        if channels_last is None:
            channels_last = cl and not out.is_contiguous()
        if (channels_last and not cl) or (not channels_last and not out.is_contiguous()):
            raise _lib.EosvosError("conv2d_wgrad: `out` strides do not match the requested gradient layout")
    call("eosvos_conv2d_wgrad", _ptr(x), _ptr(dy), _ptr(out), N, H, W, Cin, Cout, KH, KW, stride, pad, alpha, bn_hint,
         split_hint, 1 if channels_last else 0, _stream())
    return out


def gemm_wgrad(x, dy, out, *, s_m, n_inner=0, s_n_inner=1, s_n_outer=0, alpha=1.0, bn_hint=0, split_hint=0, n_valid=0):
    """out[m*s_m + (n//n_inner)*s_n_outer + (n%n_inner)*s_n_inner] += sum_r dy[r,m] * x[r,n]  for n < n_valid
    (0 = every column of x)."""
    _chk(x, ACT_DTYPE, "x")
    _chk(dy, ACT_DTYPE, "dy")
    _chk(out, torch.float32, "out")
    rows, n_cols = x.shape
    rows2, m_cols = dy.shape
    assert rows == rows2
    call("eosvos_gemm_wgrad", _ptr(x), _ptr(dy), _ptr(out), rows, n_cols, m_cols, s_m, n_inner, s_n_inner, s_n_outer,
         alpha, bn_hint, split_hint, n_valid, _stream())
    return out


def deconv2x2_fprop(x, wd, bias4=None, *, relu=False, bn_hint=0):
    """x [N,h,w,Cin] bf16, wd [4*Cout, Cin] bf16 ((dy,dx,co) rows) -> y [N,2h,2w,Cout] bf16."""
    _chk(x, ACT_DTYPE, "x")
    _chk(wd, ACT_DTYPE, "wd")
    N, h, w, Cin = x.shape
    Cout = wd.shape[0] // 4
    y = torch.empty((N, 2 * h, 2 * w, Cout), device=x.device, dtype=ACT_DTYPE)
    call("eosvos_deconv2x2_fprop", _ptr(x), _ptr(wd), _ptr(bias4), _ptr(y), N, h, w, Cin, Cout,
         FLAG_RELU if relu else 0, bn_hint, _stream())
    return y


def deconv2x2_dgrad(dy, wdt, *, bn_hint=0):
    """dy [N,2h,2w,Cout] bf16, wdt [Cin, 4*Cout] bf16 -> dx [N,h,w,Cin] bf16."""
    _chk(dy, ACT_DTYPE, "dy")
    _chk(wdt, ACT_DTYPE, "wdt")
    N, H2, W2, Cout = dy.shape
    Cin = wdt.shape[0]
    dx = torch.empty((N, H2 // 2, W2 // 2, Cin), device=dy.device, dtype=ACT_DTYPE)
    call("eosvos_deconv2x2_dgrad", _ptr(dy), _ptr(wdt), _ptr(dx), N, H2 // 2, W2 // 2, Cin, Cout, 0, bn_hint,
         _stream())
    return dx


def deconv2x2_wgrad(x, dy, *, alpha=1.0, bn_hint=0, split_hint=0):
    """-> dw fp32 [Cin, Cout, 2, 2] (torch ConvTranspose2d layout)."""
    _chk(x, ACT_DTYPE, "x")
    _chk(dy, ACT_DTYPE, "dy")
    N, h, w, Cin = x.shape
    Cout = dy.shape[-1]
    dw = zero_pool.take((Cin, Cout, 2, 2), x.device)
    call("eosvos_deconv2x2_wgrad", _ptr(x), _ptr(dy), _ptr(dw), N, h, w, Cin, Cout, alpha, bn_hint, split_hint, _stream())
    return dw


# ------------------------------------------------------------------------------------------- K2
def gn_stats(x):
    _chk(x, ACT_DTYPE, "x")
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    sums = torch.empty((N, 32, 2), device=x.device, dtype=torch.float32)
    call("eosvos_gn_stats", _ptr(x), _ptr(sums), N, HW, C, _stream())
    return sums


def gn_apply(x, sums, gamma, beta, res=None, relu=False, eps=GN_EPS):
    _chk(x, ACT_DTYPE, "x")
    _chk(sums, torch.float32, "sums")
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    y = torch.empty_like(x)
    call("eosvos_gn_apply", _ptr(x), _ptr(sums), _ptr(_chk(gamma, torch.float32)), _ptr(_chk(beta, torch.float32)),
         _ptr(res), _ptr(y), N, HW, C, eps, 1 if relu else 0, _stream())
    return y


def gn_backward(x, sums, gamma, beta, dy, yout=None, mask_mode=0, want_dres=False, eps=GN_EPS, alpha=1.0):
    _chk(x, ACT_DTYPE, "x")
    _chk(dy, ACT_DTYPE, "dy")
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    part = torch.empty((N, C, 2), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    dgamma = torch.empty((C,), device=x.device, dtype=torch.float32)
    dbeta = torch.empty((C,), device=x.device, dtype=torch.float32)
    call("eosvos_gn_backward", _ptr(x), _ptr(sums), _ptr(gamma), _ptr(beta), _ptr(dy), _ptr(yout), _ptr(part),
         _ptr(dx), _ptr(dres), _ptr(dgamma), _ptr(dbeta), N, HW, C, eps, mask_mode, alpha, _stream())
    return dx, dres, dgamma, dbeta


# ------------------------------------------------------------------------------------------- K4
def _level_args(feats_or_shapes):
    ptrs = (ctypes.c_void_p * 4)()
    Hs = (ctypes.c_int * 4)()
    Ws = (ctypes.c_int * 4)()
    return ptrs, Hs, Ws


def roi_align_fwd(feats, scales, rois, P, sampling=2):
    """feats: 4 NHWC bf16 levels; rois [R,5] fp32 (batch, x1, y1, x2, y2) -> [R,P,P,C] bf16."""
    ptrs, Hs, Ws = _level_args(feats)
    sc = (ctypes.c_float * 4)(*[float(s) for s in scales])
    for i, f in enumerate(feats):
        _chk(f, ACT_DTYPE, "feature level")
        ptrs[i] = f.data_ptr()
        Hs[i], Ws[i] = f.shape[1], f.shape[2]
    C = feats[0].shape[-1]
    _chk(rois, torch.float32, "rois")
    R = rois.shape[0]
    out = torch.empty((R, P, P, C), device=rois.device, dtype=ACT_DTYPE)
    call("eosvos_roi_align_fwd", ptrs, Hs, Ws, sc, _ptr(rois), _ptr(out), R, P, C, sampling, _stream())
    return out


def roi_align_bwd(dout, level_shapes, scales, rois, P, sampling=2):
    """-> list of 4 fp32 NHWC gradient maps (zero-initialised here, atomically accumulated)."""
    _chk(dout, ACT_DTYPE, "dout")
    _chk(rois, torch.float32, "rois")
    ptrs, Hs, Ws = _level_args(level_shapes)
    sc = (ctypes.c_float * 4)(*[float(s) for s in scales])
    outs = []
    for i, shp in enumerate(level_shapes):
        g = zero_pool.take(tuple(shp), dout.device)
        outs.append(g)
        ptrs[i] = g.data_ptr()
        Hs[i], Ws[i] = shp[1], shp[2]
    R, C = rois.shape[0], dout.shape[-1]
    call("eosvos_roi_align_bwd", ptrs, Hs, Ws, sc, _ptr(rois), _ptr(dout), R, P, C, sampling, _stream())
    return outs


def mask_targets(masks_u8, rois, M):
    """masks uint8 [G,H,W]; rois [R,5] (mask idx, box) -> fp32 [R,M,M]."""
    _chk(masks_u8, torch.uint8, "masks")
    _chk(rois, torch.float32, "rois")
    G, H, W = masks_u8.shape
    R = rois.shape[0]
    out = torch.empty((R, M, M), device=rois.device, dtype=torch.float32)
    call("eosvos_mask_targets", _ptr(masks_u8), _ptr(rois), _ptr(out), R, M, H, W, _stream())
    return out


# ------------------------------------------------------------------------------------------- K7
def mask_loss_per_roi(logits, labels, targets):
    """Lovasz hinge per RoI: -> (loss_per_roi [R], dlogits [R,Cc,M,M] = d loss_r / d logits, unscaled)."""
    _chk(logits, torch.float32, "logits")
    _chk(labels, torch.int64, "labels")
    _chk(targets, torch.float32, "targets")
    R, Cc = logits.shape[0], logits.shape[1]
    P = logits.shape[2] * logits.shape[3]
    loss = torch.zeros((), device=logits.device, dtype=torch.float32)
    per = torch.empty((R,), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    call("eosvos_mask_loss_lovasz", _ptr(logits), _ptr(labels), _ptr(targets), _ptr(loss), _ptr(per), _ptr(dlogits), R,
         Cc, P, _stream())
    return per, dlogits.mul_(float(R))       # the kernel folds the 1/R of the mean into the gradient


def mask_loss(logits, labels, targets, kind="LOVASZ"):
    """logits [R,Cc,M,M] fp32, labels [R] int64, targets [R,M,M] fp32 -> (loss scalar, dlogits)."""
    _chk(logits, torch.float32, "logits")
    _chk(labels, torch.int64, "labels")
    _chk(targets, torch.float32, "targets")
    R, Cc = logits.shape[0], logits.shape[1]
    P = logits.shape[2] * logits.shape[3]
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    if kind == "LOVASZ":
        call("eosvos_mask_loss_lovasz", _ptr(logits), _ptr(labels), _ptr(targets), _ptr(loss), None, _ptr(dlogits), R,
             Cc, P, _stream())
    elif kind == "BCE":
        call("eosvos_mask_loss_bce", _ptr(logits), _ptr(labels), _ptr(targets), _ptr(loss), _ptr(dlogits), R, Cc, P,
             _stream())
    else:
        raise NotImplementedError(kind)
    return loss, dlogits


# ------------------------------------------------------------------------------------------- K8
def mask_paste_threshold(logits, det_of_chan, det_label, det_box, B, K, H, W, thresh=0.5, want_target=True):
    """-> probs [B,K,H,W] fp32, target [B,1,H,W] fp32 ids, stats [B,K,5] int32 (xmin,ymin,xmax,ymax,count)."""
    dev = det_of_chan.device
    _chk(det_of_chan, torch.int32, "det_of_chan")
    D = logits.shape[0]
    if D > 0:
        _chk(logits, torch.float32, "logits")
        _chk(det_label, torch.int64, "det_label")
        _chk(det_box, torch.float32, "det_box")
    M = logits.shape[-1]
    Cc = logits.shape[1]
    probs = torch.empty((B, K, H, W), device=dev, dtype=torch.float32)
    target = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32) if want_target else None
    stats = torch.empty((B, K, 5), device=dev, dtype=torch.int32) if want_target else None
    call("eosvos_mask_paste_threshold", _ptr(logits) if D else None, _ptr(det_of_chan), _ptr(det_label) if D else None,
         _ptr(det_box) if D else None, _ptr(probs), _ptr(target), _ptr(stats), B, K, H, W, M, Cc, thresh, _stream())
    return probs, target, stats


def mask_to_bbox(target, K):
    """target [B,1,H,W] (or [B,H,W]) fp32 ids -> stats [B,K,5] int32."""
    _chk(target, torch.float32, "target")
    B = target.shape[0]
    H, W = target.shape[-2:]
    stats = torch.empty((B, K, 5), device=target.device, dtype=torch.int32)
    call("eosvos_mask_to_bbox", _ptr(target), _ptr(stats), B, K, H, W, _stream())
    return stats


def jf_counts(pred, gt, num_objects, radius):
    """pred, gt [T,H,W] uint8 object ids -> counts [T,num_objects,6] int32 (intersection, union, |boundary pred|,
    |boundary gt|, matched pred boundary, matched gt boundary).  Eight objects per launch pair."""
    _chk(pred, torch.uint8, "pred")
    _chk(gt, torch.uint8, "gt")
    if pred.shape != gt.shape or pred.dim() != 3:
        raise ValueError("jf_counts: pred and gt must both be [T,H,W]")
    T, H, W = pred.shape
    bmap = torch.empty((T, 2, H, W), device=pred.device, dtype=torch.uint8)
    parts = []
    for id0 in range(0, num_objects, 8):
        k = min(8, num_objects - id0)
        counts = torch.empty((T, k, 6), device=pred.device, dtype=torch.int32)
        call("eosvos_jf_counts", _ptr(pred), _ptr(gt), _ptr(bmap), _ptr(counts), T, id0, k, H, W, int(radius), _stream())
        parts.append(counts)
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


# ------------------------------------------------------------------------------------------- K5
def nms_segments(boxes_sorted, seg_offsets, num_segments, max_seg, thresh):
    """boxes [n,4] fp32 sorted by descending score inside each segment; seg_offsets int32 [S+1] (device).
    Returns keep flags uint8 [n]."""
    _chk(boxes_sorted, torch.float32, "boxes")
    _chk(seg_offsets, torch.int32, "seg_offsets")
    n = boxes_sorted.shape[0]
    keep = torch.zeros((n,), device=boxes_sorted.device, dtype=torch.uint8)
    if n == 0 or num_segments == 0:
        return keep
    nbytes = _lib.load().eosvos_nms_scratch_bytes(num_segments, max_seg)
    scratch = torch.empty((nbytes,), device=boxes_sorted.device, dtype=torch.uint8)
    call("eosvos_nms_segments", _ptr(boxes_sorted), _ptr(seg_offsets), num_segments, max_seg, float(thresh),
         _ptr(scratch), _ptr(keep), _stream())
    return keep


# ------------------------------------------------------------------------------------------- K5/K6 (rpn.cu)
def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def rpn_select(head_outs, hw, A, N, anchors, image_sizes, pre_nms_top_n, bbox_clip, min_size, score_thresh):
    """Per-(image, level) top-k of the objectness logits + decoding of the selected anchors only.
    head_outs: per level fp32 [N*hw_l, 16]; anchors fp32 [sum hw_l*A, 4]; image_sizes [(h, w)] * N.
    -> (boxes [N,C,4] (invalid ones zeroed), scores [N,C] (sigmoid), valid uint8 [N,C], per-level k list)."""
    L = len(head_outs)
    dev = head_outs[0].device
    for h in head_outs:
        _chk(h, torch.float32, "rpn head output")
    _chk(anchors, torch.float32, "anchors")
    ks = [min(pre_nms_top_n, n * A) for n in hw]
    C = sum(ks)
    hw_c = _int_array(hw)
    lib = _lib.load()
    nbytes = lib.eosvos_rpn_scratch_bytes(N, hw_c, L, A)
    zbytes = lib.eosvos_rpn_scratch_zero_bytes(N, L)
    scratch = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
    scratch[:zbytes].zero_()
    boxes = torch.empty((N, C, 4), device=dev, dtype=torch.float32)
    scores = torch.empty((N, C), device=dev, dtype=torch.float32)
    valid = torch.empty((N, C), device=dev, dtype=torch.uint8)
    ptrs = (ctypes.c_void_p * L)(*[h.data_ptr() for h in head_outs])
    sizes = (ctypes.c_float * (2 * N))(*[float(v) for s in image_sizes for v in s])
    call("eosvos_rpn_select", ptrs, hw_c, L, A, N, _ptr(anchors), sizes, int(pre_nms_top_n), float(bbox_clip),
         float(min_size), float(score_thresh), _ptr(scratch), _ptr(boxes), _ptr(scores), _ptr(valid), _stream())
    return boxes, scores, valid, ks


def rpn_postnms(hw, A, N, pre_nms_top_n, boxes, scores, valid, keep, post_n, out_boxes=None, out_offset=0,
                want_scores=False):
    """First post_n NMS survivors of every image in descending score order -> rows [out_offset, out_offset+post_n) of
    out_boxes [N, stride, 4] (padding rows zero), count int32 [N]."""
    dev = boxes.device
    if out_boxes is None:
        out_boxes = torch.empty((N, post_n, 4), device=dev, dtype=torch.float32)
    stride = out_boxes.shape[1]
    out_scores = torch.empty((N, stride), device=dev, dtype=torch.float32) if want_scores else None
    count = torch.empty((N,), device=dev, dtype=torch.int32)
    call("eosvos_rpn_postnms", _int_array(hw), len(hw), A, N, int(pre_nms_top_n), _ptr(boxes), _ptr(scores), _ptr(valid),
         _ptr(_chk(keep, torch.uint8, "keep")), int(post_n), int(stride), int(out_offset), _ptr(out_boxes),
         _ptr(out_scores), _ptr(count), _stream())
    return out_boxes, out_scores, count


def extend_boxes(stats, fallback, rnd, n_aug, ratio_w, ratio_h, img_w, img_h, share, out_boxes, out_offset):
    """Jittered copies of the target boxes (stats int32 [B,G,5], input-frame pixels) into out_boxes [B, stride, 4]
    at rows [out_offset, out_offset + G*n_aug); rnd fp32 [B,G,4,n_aug] (host-drawn uniforms, already on the device)."""
    _chk(stats, torch.int32, "stats")
    _chk(rnd, torch.float32, "rnd")
    B, G = stats.shape[0], stats.shape[1]
    call("eosvos_extend_boxes", _ptr(stats), _ptr(fallback), _ptr(rnd), B, G, int(n_aug), float(ratio_w), float(ratio_h),
         float(img_w), float(img_h), float(share), int(out_boxes.shape[1]), int(out_offset), _ptr(out_boxes), _stream())
    return out_boxes


def det_top1(head, proposals, B, R, num_classes, weights, bbox_clip, score_thresh, min_size, img_w, img_h, back_w, back_h):
    """Best valid (row, class) candidate per image from the fused box head output [B*R, 16] and proposals [B*R, 4].
    -> dict(box [B,4] input-frame coords, score [B], label int64 [B], row int32 [B] (-1 = none), roi [B,5], chan
    int32 [B*(ncls-1)])."""
    dev = head.device
    _chk(head, torch.float32, "box head output")
    _chk(proposals, torch.float32, "proposals")
    out = dict(box=torch.empty((B, 4), device=dev), score=torch.empty((B,), device=dev),
               label=torch.empty((B,), device=dev, dtype=torch.int64), row=torch.empty((B,), device=dev, dtype=torch.int32),
               roi=torch.empty((B, 5), device=dev), chan=torch.empty((B * (num_classes - 1),), device=dev, dtype=torch.int32))
    w = (ctypes.c_float * 4)(*[float(v) for v in weights])
    call("eosvos_det_top1", _ptr(head), _ptr(proposals), B, R, num_classes, w, float(bbox_clip), float(score_thresh),
         float(min_size), float(img_w), float(img_h), float(back_w), float(back_h), _ptr(out["box"]), _ptr(out["score"]),
         _ptr(out["label"]), _ptr(out["row"]), _ptr(out["roi"]), _ptr(out["chan"]), _stream())
    return out


def roi_match(proposals, count, gt_boxes, gt_labels, gt_off, max_gt, iou_thresh):
    """proposals [B,P,4] (first count[b] real) + the image's ground-truth boxes appended at rows P.. ->
    (all_boxes [B,P+max_gt,4], labels int64 (class / 0 background / -1 padding), matched int64, counts int32 [B,2])."""
    dev = proposals.device
    B, P = proposals.shape[0], proposals.shape[1]
    rows = P + max_gt
    all_boxes = torch.empty((B, rows, 4), device=dev, dtype=torch.float32)
    labels = torch.empty((B, rows), device=dev, dtype=torch.int64)
    matched = torch.empty((B, rows), device=dev, dtype=torch.int64)
    counts = torch.zeros((B, 2), device=dev, dtype=torch.int32)
    call("eosvos_roi_match", _ptr(_chk(proposals, torch.float32)), _ptr(_chk(count, torch.int32)),
         _ptr(_chk(gt_boxes, torch.float32)), _ptr(_chk(gt_labels, torch.int64)), _ptr(_chk(gt_off, torch.int32)), B, P,
         int(max_gt), float(iou_thresh), _ptr(all_boxes), _ptr(labels), _ptr(matched), _ptr(counts), _stream())
    return all_boxes, labels, matched, counts


def rpn_anchor_match(anchors, gt_boxes, gt_off, N, fg_iou, bg_iou):
    """tv rpn.py assign_targets_to_anchors (Matcher with low-quality matches) for all images: -> (labels int64
    [N, A_total] in {1, 0, -1}, matched int32 [N, A_total], counts int32 [N, 2] = (#fg, #bg))."""
    dev = anchors.device
    na = anchors.shape[0]
    zero = torch.zeros((gt_boxes.shape[0] + 2 * N,), device=dev, dtype=torch.int32)
    gt_best, counts = zero[:gt_boxes.shape[0]], zero[gt_boxes.shape[0]:].view(N, 2)
    labels = torch.empty((N, na), device=dev, dtype=torch.int64)
    matched = torch.empty((N, na), device=dev, dtype=torch.int32)
    call("eosvos_rpn_anchor_match", _ptr(_chk(anchors, torch.float32, "anchors")), na, _ptr(_chk(gt_boxes, torch.float32)),
         _ptr(_chk(gt_off, torch.int32)), N, float(fg_iou), float(bg_iou), _ptr(gt_best), _ptr(labels), _ptr(matched),
         _ptr(counts), _stream())
    return labels, matched, counts


def rpn_loss(head_outs, hw, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta, dys=None, g_obj=None,
             g_box=None):
    """tv rpn.py compute_loss on the sampled anchors straight from the per-level head outputs.  Forward (dys None) ->
    fp32 [2] = (loss_objectness, loss_rpn_box_reg); backward: scatters d loss / d head into the zeroed `dys`."""
    L = len(head_outs)
    ptrs = (ctypes.c_void_p * L)(*[h.data_ptr() for h in head_outs])
    dptr = (ctypes.c_void_p * L)(*[d.data_ptr() for d in dys]) if dys is not None else None
    out = torch.empty((2,), device=labels.device, dtype=torch.float32) if dys is None else None
    call("eosvos_rpn_loss", ptrs, dptr, _int_array(hw), L, int(A), _ptr(_chk(sampled, torch.int64, "sampled")),
         int(sampled.numel()), _ptr(_chk(labels, torch.int64, "labels")), _ptr(_chk(matched, torch.int32, "matched")),
         _ptr(anchors), _ptr(gt_boxes), _ptr(gt_off), float(beta), 0 if dys is None else 1, _ptr(out), _ptr(g_obj),
         _ptr(g_box), _stream())
    return out


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def rpn_sparse_head(head_outs, ts, feats, A, sampled, labels, matched, anchors, gt_boxes, gt_off, beta, g_obj, g_box,
                    w_cls, w_box, grad_scale):
    """Stage 1 of the sparse RPN-head backward -> (dt [M,C] ACT, ev_pix int32 [M,4], xg [M, 9*C] ACT, dw_cls [A,C],
    db_cls [A], dw_box [4A,C], db_box [4A]) (fp32 gradients unscaled)."""
    dev = sampled.device
    M = int(sampled.numel())
    C = feats[0].shape[-1]
    Hs, Ws = [f.shape[1] for f in feats], [f.shape[2] for f in feats]
    dt = torch.empty((M, C), device=dev, dtype=ACT_DTYPE)
    ev = torch.empty((M, 4), device=dev, dtype=torch.int32)
    xg = torch.empty((M, 9 * C), device=dev, dtype=ACT_DTYPE)
    acc = zero_pool.take((5 * A * C + 5 * A + 64,), dev)
    dw_cls, dw_box = acc[:A * C].view(A, C), acc[A * C:5 * A * C].view(4 * A, C)
    db_cls, db_box = acc[5 * A * C:5 * A * C + A], acc[5 * A * C + A:5 * A * C + 5 * A]
    call("eosvos_rpn_sparse_head", _ptr_array(head_outs), _ptr_array(ts), _ptr_array(feats), _int_array(Hs),
         _int_array(Ws), len(feats), int(A), int(C), _ptr(_chk(sampled, torch.int64, "sampled")), M,
         _ptr(_chk(labels, torch.int64, "labels")), _ptr(_chk(matched, torch.int32, "matched")), _ptr(anchors),
         _ptr(gt_boxes), _ptr(gt_off), float(beta), _ptr(g_obj), _ptr(g_box), _ptr(_chk(w_cls, torch.float32, "w_cls")),
         _ptr(_chk(w_box, torch.float32, "w_box")), float(grad_scale), _ptr(dt), _ptr(ev), _ptr(xg), _ptr(dw_cls),
         _ptr(db_cls), _ptr(dw_box), _ptr(db_box), _stream())
    return dt, ev, xg, dw_cls, db_cls, dw_box, db_box


def rpn_sparse_scatter(dfs, ev_pix, G):
    """df[level][pixel + tap][ci] += G[e][ci*9 + tap] into the zeroed ACT maps dfs (NHWC)."""
    C = dfs[0].shape[-1]
    call("eosvos_rpn_sparse_scatter", _ptr_array(dfs), _int_array([d.shape[1] for d in dfs]),
         _int_array([d.shape[2] for d in dfs]), len(dfs), int(C), _ptr(ev_pix), int(ev_pix.shape[0]),
         _ptr(_chk(G, ACT_DTYPE, "G")), _stream())


def roi_sample(labels, perms, num_pos, num_neg, S, Pmax):
    """BalancedPositiveNegativeSampler selection for all images in one launch.  labels int64 [B, rows]; perms =
    [(perm_pos, perm_neg)] per image (the reference's two device `torch.randperm` draws); num_pos / num_neg per image.
    -> (inds int64 [B,S] ascending rows, -1 padded; pos_in int64 [B,Pmax] positions of the foreground rows, -1 padded)."""
    import numpy as np
    dev = labels.device
    B, rows = labels.shape
    tab = np.empty((B, 4), dtype=np.int64)
    for b, ((pp, pn), a, c) in enumerate(zip(perms, num_pos, num_neg)):
        tab[b] = (pp.data_ptr(), pn.data_ptr(), a, c)
    table = stager.put(tab, dev)
    inds = torch.empty((B, S), device=dev, dtype=torch.int64)
    pos_in = torch.empty((B, Pmax), device=dev, dtype=torch.int64)
    scratch = torch.empty((_lib.load().eosvos_roi_sample_scratch_bytes(B, rows),), device=dev, dtype=torch.uint8)
    call("eosvos_roi_sample", _ptr(_chk(labels, torch.int64, "labels")), _ptr(table), B, rows, int(S), int(Pmax),
         _ptr(scratch), _ptr(inds), _ptr(pos_in), _stream())
    return inds, pos_in


def roi_encode(all_boxes, labels, matched, gt_boxes, gt_off, inds, weights):
    """Sampled rows inds int64 [B,S] -> (rois5 [B*S,5], labels [B*S], matched [B*S], regression targets [B*S,4])."""
    dev = all_boxes.device
    B, S = inds.shape
    rois5 = torch.empty((B * S, 5), device=dev, dtype=torch.float32)
    out_l = torch.empty((B * S,), device=dev, dtype=torch.int64)
    out_m = torch.empty((B * S,), device=dev, dtype=torch.int64)
    reg = torch.empty((B * S, 4), device=dev, dtype=torch.float32)
    w = (ctypes.c_float * 4)(*[float(v) for v in weights])
    call("eosvos_roi_encode", _ptr(all_boxes), _ptr(labels), _ptr(matched), _ptr(gt_boxes), _ptr(gt_off),
         _ptr(_chk(inds, torch.int64, "inds")), B, S, int(all_boxes.shape[1]), w, _ptr(rois5), _ptr(out_l), _ptr(out_m),
         _ptr(reg), _stream())
    return rois5, out_l, out_m, reg


# ------------------------------------------------------------------------------------------- K9
class MetaUpdatePlan:
    """Pointer/chunk tables of one (params, grads, lrs, outs) binding, cached on the device."""

    _chunk_cache = {}

    def __init__(self, params, grads, lrs, outs):
        import numpy as np
        chunk = _lib.load().eosvos_meta_update_chunk_elems()
        dev = params[0].device
        T = len(params)
        rows = np.empty((T, 8), dtype=np.int64)
        for t, (p, g, lr, o) in enumerate(zip(params, grads, lrs, outs)):
            n = p.numel()
            if g.numel() != n or o.numel() != n or n % lr.numel() != 0:
                raise _lib.EosvosError("meta_update: parameter / gradient / learning-rate sizes do not match")
            # gradients of KxK filters arrive in channels_last memory order [Cout][K*K][Cin] (conv2d_wgrad)
            taps, cin = 1, 1
            if not g.is_contiguous() and g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last):
                taps, cin = g.shape[2] * g.shape[3], g.shape[1]
            elif not g.is_contiguous():
                raise _lib.EosvosError("meta_update: gradient must be contiguous or channels_last")
            if not (p.is_cuda and g.is_cuda and lr.is_cuda and o.is_cuda) or p.dtype != torch.float32 or \
                    g.dtype != torch.float32 or lr.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.EosvosError("meta_update operands must be contiguous fp32 CUDA tensors (no CPU path exists)")
            rows[t] = (p.data_ptr(), g.data_ptr(), lr.data_ptr(), o.data_ptr(), n, n // lr.numel(), taps, cin)
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        key = (str(dev), tuple(int(n) for n in rows[:, 4]))
        hit = MetaUpdatePlan._chunk_cache.get(key)
        if hit is None:
            chunks = [[t, c] for t in range(T) for c in range((int(rows[t, 4]) + chunk - 1) // chunk)]
            hit = (torch.tensor(chunks, dtype=torch.int32).to(dev), len(chunks))      # once per shape signature
            MetaUpdatePlan._chunk_cache = {key: hit}
        self.chunks, self.num_chunks = hit
        self.table = stager.put(rows, dev)


def meta_update(plan, use_log=False, nonfinite=None):
    """nonfinite: optional int32 device tensor, OR-ed with 1 when an updated parameter is Inf / NaN."""
    call("eosvos_meta_update", _ptr(plan.table), _ptr(plan.chunks), plan.num_chunks, 1 if use_log else 0, _ptr(nonfinite),
         _stream())


_lr_work_cache = {}


def lr_grad(douts, grads, lrs, use_log=False):
    """d L / d lr_i = -rowsum(dout_i (.) grad_i) (x exp(lr_i) in log mode) for all tensors in ONE launch.
    douts / grads: fp32, contiguous or (4-D) channels_last; lrs: fp32, numel divides the tensor's numel.
    -> list of tensors shaped like lrs."""
    import numpy as np
    dev = douts[0].device
    T = len(douts)
    tab = np.empty((T, 10), dtype=np.int64)
    total = sum(l.numel() for l in lrs)
    flat = torch.empty((total,), device=dev, dtype=torch.float32)
    outs, o = [], 0

    def layout(x):
        if x.is_contiguous():
            return 1, 1
        if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
            return x.shape[2] * x.shape[3], x.shape[1]
        raise _lib.EosvosError("lr_grad: operands must be contiguous or channels_last")
    sig = []
    for t, (d, g, l) in enumerate(zip(douts, grads, lrs)):
        n = d.numel()
        if g.numel() != n or n % l.numel() != 0 or d.dtype != torch.float32 or g.dtype != torch.float32:
            raise _lib.EosvosError("lr_grad: sizes / dtypes do not match")
        gt, gc = layout(g)
        dt, dc = layout(d)
        out = flat[o:o + l.numel()]
        tab[t] = (d.data_ptr(), g.data_ptr(), l.data_ptr(), out.data_ptr(), n, n // l.numel(), gt, gc, dt, dc)
        outs.append(out.view(l.shape))
        o += l.numel()
        sig.append((n, l.numel()))
    key = (str(dev), tuple(sig))
    hit = _lr_work_cache.get(key)
    if hit is None:
        work = []
        for t, (n, rows) in enumerate(sig):
            if n // rows >= 256:
                work += [(t, r, 0) for r in range(rows)]
            else:
                work += [(t, r, 1) for r in range(0, rows, 256)]
        hit = (torch.tensor(work, dtype=torch.int32).to(dev), len(work))
        _lr_work_cache.clear()
        _lr_work_cache[key] = hit
    table = stager.put(tab, dev) if tab.nbytes <= stager.slot_bytes else torch.from_numpy(tab).to(dev)
    call("eosvos_lr_grad", _ptr(table), _ptr(hit[0]), hit[1], 1 if use_log else 0, _stream())
    return outs


def radam_step(p, g, m, v, *, gscale, clip, beta1, beta2, eps, lr, wd, step_size, rectified, clamp=None):
    for t in (p, g, m, v):
        _chk(t, torch.float32, "radam operand")
    lo, hi, do = (clamp[0], clamp[1], 1) if clamp is not None else (0.0, 0.0, 0)
    call("eosvos_radam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), gscale, clip if clip else 0.0, beta1,
         beta2, 1 - beta1, 1 - beta2, eps, lr, wd, step_size, 1 if rectified else 0, lo, hi, do, _stream())


# ------------------------------------------------------------------------------------------- misc
_DT = {torch.float32: 0, ACT_DTYPE: 1}


def permute_cast(src, dst, dims, sstride, dstride):
    """dst[i0,i1,i2,i3 . dstride] = src[i0,i1,i2,i3 . sstride] with dtype conversion (fp32/bf16)."""
    d = (ctypes.c_longlong * 4)(*dims)
    s = (ctypes.c_longlong * 4)(*sstride)
    t = (ctypes.c_longlong * 4)(*dstride)
    call("eosvos_permute_cast", _ptr(src), _ptr(dst), d, s, t, _DT[src.dtype], _DT[dst.dtype], _stream())
    return dst


_pm_chunk_cache = {}


def permute_cast_multi(jobs):
    """jobs: list of (src fp32 tensor, dst ACT tensor, dims[4], sstride[4], dstride[4]) -> one launch."""
    import numpy as np
    if not jobs:
        return
    chunk = _lib.load().eosvos_permute_cast_multi_chunk_elems()
    tab = np.empty((len(jobs), 14), dtype=np.int64)
    for t, (src, dst, dims, ss, ds) in enumerate(jobs):
        tab[t, 0], tab[t, 1] = src.data_ptr(), dst.data_ptr()
        tab[t, 2:6], tab[t, 6:10], tab[t, 10:14] = dims, ss, ds
    dev = jobs[0][0].device
    _chk(jobs[0][0], torch.float32, "permute_cast_multi source")
    totals = tab[:, 2] * tab[:, 3] * tab[:, 4] * tab[:, 5]
    key = (str(dev), totals.tobytes())
    hit = _pm_chunk_cache.get(key)
    if hit is None:                       # the chunk table only depends on the tensor sizes: build once
        counts = (totals + chunk - 1) // chunk
        tid = np.repeat(np.arange(len(jobs), dtype=np.int32), counts)
        cid = np.concatenate([np.arange(c, dtype=np.int32) for c in counts]) if len(jobs) else np.zeros(0, np.int32)
        hit = (torch.from_numpy(np.stack([tid, cid], 1).copy()).to(dev), int(counts.sum()))
        if len(_pm_chunk_cache) > 8:
            _pm_chunk_cache.clear()
        _pm_chunk_cache[key] = hit
    table = stager.put(tab, dev)
    call("eosvos_permute_cast_multi", _ptr(table), _ptr(hit[0]), hit[1], _stream())


def _xyz_of_spec(dims, ss, ds):
    """(dims, src strides, dst strides) of a permuting cast -> (X, Y, Z, dx, dy, dz) with the source contiguous as
    [X][Y][Z], or None when the source is not a contiguous 3-D walk or neither dx nor dy is 1."""
    ax = [(int(s), int(n), int(d)) for n, s, d in zip(dims, ss, ds) if int(n) > 1]
    ax.sort(key=lambda a: -a[0])
    if ax and ax[-1][0] == 1 and ax[-1][2] == 1:
        ax.append((0, 1, 0))            # fastest axis unchanged (plain cast): Z = 1, the run is the Y axis
    while len(ax) < 3:
        ax.insert(0, (0, 1, 0))
    if len(ax) != 3:
        return None
    (sx, X, dx), (sy, Y, dy), (sz, Z, dz) = ax
    if sz not in (1, 0) or (Y > 1 and sy != Z) or (X > 1 and sx != Y * Z):
        return None
    if Z == 1 and dy != 1 and dx != 1:
        return None
    if dy != 1 and dx != 1:
        return None
    return X, Y, Z, dx, dy, dz


def weight_prep_table(jobs):
    """jobs: list of (src fp32 tensor, dst ACT tensor, X, Y, Z, dx, dy, dz) -> (int64 table [n,10], int32 tiles [m,2])
    for eosvos_weight_prep_multi (host numpy arrays)."""
    import numpy as np
    cap = _lib.load().eosvos_weight_prep_tile_elems()
    tab = np.empty((len(jobs), 10), dtype=np.int64)
    tid, tix = [], []
    for t, (src, dst, X, Y, Z, dx, dy, dz) in enumerate(jobs):
        if dy == 1:
            TY = min(Y, max(1, min(128, cap // Z)))
            TX = max(1, min(X, 64, cap // (TY * Z)))
        else:
            TX = min(X, 32)
            TY = max(1, min(Y, cap // (TX * Z)))
        if TX * TY * Z > cap:
            raise _lib.EosvosError("weight_prep: innermost source dimension too long for one tile")
        tab[t] = (src.data_ptr(), dst.data_ptr(), X, Y, Z, dx, dy, dz, TX, TY)
        n = ((X + TX - 1) // TX) * ((Y + TY - 1) // TY)
        tid.append(np.full(n, t, dtype=np.int32))
        tix.append(np.arange(n, dtype=np.int32))
    tiles = np.stack([np.concatenate(tid), np.concatenate(tix)], 1).copy() if jobs else np.zeros((0, 2), np.int32)
    return tab, tiles


class WeightPrepPlan:
    """A fixed set of (parameter address -> operand buffer) conversions with its tables resident on the device:
    `launch()` is ONE kernel with static arguments, so it can be captured into a CUDA graph."""

    def __init__(self, jobs, device):
        tab, tiles = weight_prep_table(jobs)
        self.table = torch.from_numpy(tab).to(device)
        self.tiles = torch.from_numpy(tiles).to(device)
        self.num_tiles = int(tiles.shape[0])
        self.keep = [(j[0], j[1]) for j in jobs]

    def launch(self):
        call("eosvos_weight_prep_multi", _ptr(self.table), _ptr(self.tiles), self.num_tiles, _stream())


_wp_tile_cache = {}


def weight_prep_multi(jobs):
    """Eager variant: tables staged through pinned memory (tile list cached per shape signature)."""
    if not jobs:
        return
    dev = jobs[0][0].device
    key = (str(dev), tuple((j[2], j[3], j[4], j[5], j[6], j[7]) for j in jobs))
    tab, tiles = weight_prep_table(jobs)
    hit = _wp_tile_cache.get(key)
    if hit is None:
        hit = (torch.from_numpy(tiles).to(dev), int(tiles.shape[0]))
        if len(_wp_tile_cache) > 8:
            _wp_tile_cache.clear()
        _wp_tile_cache[key] = hit
    table = stager.put(tab, dev)
    call("eosvos_weight_prep_multi", _ptr(table), _ptr(hit[0]), hit[1], _stream())


def affine_warp_cubic(src_chw, minv, flip, B):
    """src [3,H,W] fp32, minv [B,6] fp32 (dst->src), flip [B] int32 -> [B,3,H,W] fp32 (bicubic, zero border)."""
    _chk(src_chw, torch.float32, "src")
    _chk(minv, torch.float32, "minv")
    _chk(flip, torch.int32, "flip")
    _, H, W = src_chw.shape
    out = torch.empty((B, 3, H, W), device=src_chw.device, dtype=torch.float32)
    call("eosvos_affine_warp_cubic", _ptr(src_chw), _ptr(minv), _ptr(flip), _ptr(out), B, H, W, _stream())
    return out


def label_warp_nearest(src_hw, minv64, flip, out=None):
    """src [H,W] fp32 ids, minv64 [B,6] float64 (OpenCV's inverted matrix), flip [B] int32 -> [B,1,H,W] fp32 ==
    cv2.warpAffine(flip(src), M, flags=cv2.INTER_NEAREST)."""
    _chk(src_hw, torch.float32, "label")
    _chk(minv64, torch.float64, "minv")
    _chk(flip, torch.int32, "flip")
    H, W = src_hw.shape
    B = minv64.shape[0]
    if out is None:
        out = torch.empty((B, 1, H, W), device=src_hw.device, dtype=torch.float32)
    call("eosvos_label_warp_nearest", _ptr(src_hw), _ptr(minv64), _ptr(flip), _ptr(out), B, H, W, _stream())
    return out


def nchw_to_nhwc_bf16(x):
    """fp32/bf16 [N,C,H,W] -> bf16 [N,H,W,C]."""
    N, C, H, W = x.shape
    x = x.contiguous()
    out = torch.empty((N, H, W, C), device=x.device, dtype=ACT_DTYPE)
    return permute_cast(x, out, (N, H, W, C), (C * H * W, W, 1, H * W), (H * W * C, W * C, C, 1))


def nhwc_to_nchw_fp32(x):
    """bf16/fp32 [N,H,W,C] -> fp32 [N,C,H,W]."""
    N, H, W, C = x.shape
    x = x.contiguous()
    out = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
    return permute_cast(x, out, (N, C, H, W), (H * W * C, 1, W * C, C), (C * H * W, H * W, W, 1))


def transform(img, oh, ow, Hp, Wp, mean, std, Cs=8):
    _chk(img, torch.float32, "image")
    B, _, h, w = img.shape
    out = torch.empty((B, Hp, Wp, Cs), device=img.device, dtype=ACT_DTYPE)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    call("eosvos_transform", _ptr(img), _ptr(out), B, h, w, oh, ow, Hp, Wp, Cs, m, s, _stream())
    return out


def mask_resize_nearest(masks_u8, oh, ow):
    _chk(masks_u8, torch.uint8, "masks")
    G, h, w = masks_u8.shape
    out = torch.empty((G, oh, ow), device=masks_u8.device, dtype=torch.uint8)
    call("eosvos_mask_resize_nearest", _ptr(masks_u8), _ptr(out), G, h, w, oh, ow, _stream())
    return out


def im2col_stem(x, KH=7, KW=7, stride=2, pad=3, Kp=192):
    _chk(x, ACT_DTYPE, "x")
    N, H, W, Cs = x.shape
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    col = torch.empty((N * Ho * Wo, Kp), device=x.device, dtype=ACT_DTYPE)
    call("eosvos_im2col_stem", _ptr(x), _ptr(col), N, H, W, Cs, KH, KW, stride, pad, Kp, _stream())
    return col, Ho, Wo


def maxpool_fwd(x, ksz=3, stride=2, pad=1, want_argmax=True):
    """-> (y, argmax uint8 [N,Ho,Wo,C]: winning window position kh*ksz+kw, first maximum in scan order)."""
    _chk(x, ACT_DTYPE, "x")
    N, H, W, C = x.shape
    Ho = (H + 2 * pad - ksz) // stride + 1
    Wo = (W + 2 * pad - ksz) // stride + 1
    y = torch.empty((N, Ho, Wo, C), device=x.device, dtype=ACT_DTYPE)
    arg = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.uint8) if want_argmax else None
    call("eosvos_maxpool_fwd", _ptr(x), _ptr(y), _ptr(arg), N, H, W, C, ksz, stride, pad, _stream())
    return y, arg


def maxpool_bwd(arg, dy, in_shape, ksz=3, stride=2, pad=1):
    N, H, W, C = in_shape
    dx = torch.empty(in_shape, device=dy.device, dtype=ACT_DTYPE)
    call("eosvos_maxpool_bwd", _ptr(_chk(arg, torch.uint8, "argmax")), _ptr(_chk(dy, ACT_DTYPE)), _ptr(dx), N, H, W, C,
         ksz, stride, pad, _stream())
    return dx


def subsample2(x):
    _chk(x, ACT_DTYPE, "x")
    N, H, W, C = x.shape
    y = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), device=x.device, dtype=ACT_DTYPE)
    call("eosvos_subsample2", _ptr(x), _ptr(y), N, H, W, C, 0, _stream())
    return y


def subsample2_bwd(dy, in_shape):
    _chk(dy, ACT_DTYPE, "dy")
    N, H, W, C = in_shape
    dx = torch.empty(in_shape, device=dy.device, dtype=ACT_DTYPE)
    call("eosvos_subsample2", _ptr(dy), _ptr(dx), N, H, W, C, 1, _stream())
    return dx


def sum2x2(dfine):
    _chk(dfine, ACT_DTYPE, "dfine")
    N, Hf, Wf, C = dfine.shape
    out = torch.empty((N, Hf // 2, Wf // 2, C), device=dfine.device, dtype=ACT_DTYPE)
    call("eosvos_sum2x2", _ptr(dfine), _ptr(out), N, Hf // 2, Wf // 2, C, _stream())
    return out


def relu_bwd(dy, y):
    _chk(dy, ACT_DTYPE, "dy")
    _chk(y, ACT_DTYPE, "y")
    out = torch.empty_like(dy)
    call("eosvos_relu_bwd", _ptr(dy), _ptr(y), _ptr(out), dy.numel(), _stream())
    return out


def colsum(dy2d, out=None, alpha=1.0):
    _chk(dy2d, ACT_DTYPE, "dy")
    M, C = dy2d.shape
    if out is None:
        out = zero_pool.take((C,), dy2d.device)
    call("eosvos_colsum", _ptr(dy2d), _ptr(out), M, C, alpha, _stream())
    return out
