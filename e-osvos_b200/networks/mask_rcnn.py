"""Drop-in for the reference's `networks.mask_rcnn.MaskRCNN` (reference src/networks/mask_rcnn.py:423-775)
whose forward/backward runs on the hand-written sm_100a kernels of libeosvos_b200.so.

Same constructor, same `forward(inputs, targets, box_coord_perm, flip_label)` contract, same module
tree (it subclasses torchvision's MaskRCNN exactly like the reference does, so `named_parameters()`,
`state_dict()`, `model.rpn._eval_augment_proposals_mode`, `model.roi_heads.detections_per_img` ...
are the reference's).  Parameters are read from the module tree on every call, so MetaModel's
replacement of `module._parameters[name]` by non-leaf tensors (reference meta_model.py:78-80) is seen.

What runs where:
  * convs / GroupNorm / ReLU / pooling / Linear / deconv / RoIAlign / mask loss / inference tail:
    CUDA kernels through the C ABI (ops.py);
  * anchor generation, box coding, IoU matching, samplers, top-k + NMS, the small RPN / box losses:
    the torchvision / torch operators the reference itself calls (kept in PyTorch so
    that RNG consumption -- device randperm, CPU torch.rand for EXTEND -- is identical).
There is no CPU path: tensors must be CUDA tensors on an sm_100 device.
"""
import os
import types
from collections import OrderedDict

import torch
from torch import nn
from torch.nn import functional as F
from torchvision.models.detection import MaskRCNN as _MaskRCNN
from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
from torchvision.models.detection.roi_heads import fastrcnn_loss
from torchvision.ops import MultiScaleRoIAlign
from torchvision.ops import boxes as box_ops
from torchvision.ops.misc import FrozenBatchNorm2d

from .. import kernels as K
from .. import ops


class _ImageListLike:
    """What AnchorGenerator needs from an ImageList: padded tensor shape + per-image sizes."""

    def __init__(self, shape, image_sizes, device):
        self.tensors = torch.empty(shape, device="meta")
        self.image_sizes = image_sizes
        self._device = device


class _ThreadLocalCapture:
    """While active, `torch.cuda.graph` (also the one `make_graphed_callables` opens) captures with
    cudaStreamCaptureModeThreadLocal instead of the global mode."""

    def __enter__(self):
        orig = torch.cuda.graph

        class graph(orig):
            def __init__(self, *args, **kwargs):
                kwargs.setdefault("capture_error_mode", "thread_local")
                super().__init__(*args, **kwargs)
        self.orig = orig
        torch.cuda.graph = graph
        return self

    def __exit__(self, *exc):
        torch.cuda.graph = self.orig
        return False


class MaskRCNN(_MaskRCNN):

    def __init__(self, backbone, num_classes, batch_norm=None, train_encoder=True,
                 roi_pool_output_sizes=None, eval_augment_rpn_proposals_mode=None,
                 replace_batch_with_group_norms=False, box_nms_thresh=0.5,
                 maskrcnn_loss='LOVASZ'):
        # construction order mirrors reference mask_rcnn.py:430-520 so that a given torch seed yields
        # the same random initialisation.
        self._num_groups = 32
        backbone_model = resnet_fpn_backbone(backbone_name=backbone, weights=None, trainable_layers=5)

        mask_roi_pool, box_roi_pool = None, None
        if roi_pool_output_sizes is not None:
            box_roi_pool = MultiScaleRoIAlign(featmap_names=['0', '1', '2', '3'],
                                              output_size=roi_pool_output_sizes['box'], sampling_ratio=2)
            mask_roi_pool = MultiScaleRoIAlign(featmap_names=['0', '1', '2', '3'],
                                               output_size=roi_pool_output_sizes['mask'], sampling_ratio=2)

        super(MaskRCNN, self).__init__(backbone_model, num_classes, box_roi_pool=box_roi_pool,
                                       mask_roi_pool=mask_roi_pool, mask_head=None,
                                       box_score_thresh=box_nms_thresh)

        self.num_classes = num_classes
        self.rpn._eval_augment_proposals_mode = eval_augment_rpn_proposals_mode
        self.roi_heads._eval_augment_proposals_mode = eval_augment_rpn_proposals_mode
        self.roi_heads.maskrcnn_loss = maskrcnn_loss
        self._roi_sizes = roi_pool_output_sizes or {'box': 7, 'mask': 14}

        if replace_batch_with_group_norms:
            self.replace_batch_with_group_norms()
        self._uses_group_norm = replace_batch_with_group_norms

        self._train_encoder = train_encoder
        if not train_encoder:
            self.requires_grad_(False)
            self.rpn.requires_grad_(True)
            self.roi_heads.box_head.requires_grad_(True)
            self.roi_heads.box_predictor.requires_grad_(True)
            self.roi_heads.mask_head.requires_grad_(True)
            self.roi_heads.mask_predictor.requires_grad_(True)
        else:
            self.backbone.requires_grad_(True)

        self._accum_batch_norm_stats = True
        if batch_norm is not None:
            self._accum_batch_norm_stats = batch_norm['accum_stats']
            for m in self.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.requires_grad = batch_norm['learn_weight']
                    m.bias.requires_grad = batch_norm['learn_bias']

        self._second_order_derivates_module_names = ['roi_heads']
        self.last_param_group_names = ['roi_heads.box_predictor.cls_score.weight',
                                       'roi_heads.box_predictor.cls_score.bias',
                                       'roi_heads.box_predictor.bbox_pred.weight',
                                       'roi_heads.box_predictor.bbox_pred.bias',
                                       'roi_heads.mask_predictor.mask_fcn_logits.weight',
                                       'roi_heads.mask_predictor.mask_fcn_logits.bias']
        self._anchor_cache = {}
        # test hooks (tests/test_model_gpu.py): bypass RPN proposals / capture stage outputs
        self.fixed_proposals = None
        self.fixed_detections = None
        self.capture = None
        # CUDA-graph the fixed-shape trunk (ResNet body + FPN, forward and backward): ~110 autograd Functions and
        # ~600 launches per iteration replay as two graphs with no host work (EOSVOS_CUDA_GRAPHS=0 disables)
        import os
        self.use_cuda_graphs = os.environ.get("EOSVOS_CUDA_GRAPHS", "1") != "0"
        self._graphs = {}
        self._trunk_slots = None
        self._side_stream = None
        self._prefetched = {}
        self._lookahead = None
        self._pf_slot = 0
        self._theta_home = None
        self._active_plan = None
        self._box_slots = None
        self._mask_slots = None

    # ---- reference API (mask_rcnn.py:523-570) ------------------------------------------------
    def replace_batch_with_group_norms(self):
        for module in self.modules():
            bn_keys = [k for k, m in module._modules.items()
                       if isinstance(m, FrozenBatchNorm2d) or isinstance(m, nn.BatchNorm2d)]
            for k in bn_keys:
                batch_norm = module._modules[k]
                group_norm = nn.GroupNorm(self._num_groups, batch_norm.weight.shape[0])
                group_norm.weight.data = batch_norm.weight
                group_norm.bias.data = batch_norm.bias
                module._modules[k] = group_norm

    def named_parameters_with_second_order_derivate(self, recurse=True):
        for name, param in self.named_parameters(recurse=recurse):
            if param.requires_grad and any(m in name for m in self._second_order_derivates_module_names):
                yield name, param

    def named_parameters_without_second_order_derivate(self, recurse=True):
        for name, param in self.named_parameters(recurse=recurse):
            if param.requires_grad and not any(m in name for m in self._second_order_derivates_module_names):
                yield name, param

    def train(self, mode=True):
        super(MaskRCNN, self).train(mode)
        if mode and getattr(self, "_prefetched", None):
            self._prefetched.clear()            # look-ahead features are only valid for the weights they ran on
        if mode:
            self._lookahead = None
        if not self._train_encoder:
            self.backbone.eval()
        if not self._accum_batch_norm_stats:
            for m in self.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.eval()
        return self

    def train_without_dropout(self):
        self.train()
        for m in self.modules():
            if isinstance(m, torch.nn.Dropout2d) or isinstance(m, torch.nn.Dropout):
                m.eval()

    # ---- targets (reference mask_rcnn.py:582-714), bbox on the device ---------------------------
    def _build_targets(self, targets, flip_label, device):
        if flip_label:
            targets = 1 - targets
        B = targets.shape[0]
        Kc = max(self.num_classes - 1, 1)
        t32 = targets.to(torch.float32).contiguous()
        pre = K.target_stats.get_host(targets) if not flip_label else None
        if pre is not None and tuple(pre[0].shape) == (B, Kc, 5):
            # fast path: the producer of `targets` already knows the per-id boxes / counts on the HOST (the
            # augmentation thread, or the previous frame's fused tail kernel via run_frames) -> no kernel, no sync
            stats, ign = pre
        else:
            stats_d = K.mask_to_bbox(t32, Kc)
            ign_d = (t32 == 255.0).flatten(1).any(dim=1).to(torch.int32)
            packed = torch.cat([stats_d.flatten(), ign_d]).cpu()   # ONE small D2H (the reference does several)
            stats = packed[:B * Kc * 5].view(B, Kc, 5)
            ign = packed[B * Kc * 5:]
            if not flip_label:      # static batches (OnA rounds re-use the same tensors every iteration) pay the sync
                K.target_stats.put(targets, stats, ign)     # once; keyed on (tensor, data_ptr, _version)
        out = []
        for b in range(B):
            mask = t32[b]                            # [1,H,W]
            # ids present = ids owning at least one pixel (reference: torch.unique minus {0, 255})
            ids = [k + 1 for k in range(Kc) if int(stats[b, k, 4]) > 0]
            num_objs = len(ids)
            assert num_objs >= 1, f"num_objs: {num_objs}"
            obj_ids = K.stager.put(torch.tensor(ids, dtype=torch.float32), device)
            masks = mask == obj_ids[:, None, None]
            has_ignore = bool(int(ign[b]))
            if has_ignore:
                masks = masks | (mask == 255.0)
            boxes = []
            for oid in ids:
                xmin, ymin, xmax, ymax, _ = stats[b, oid - 1].tolist()
                boxes.append([xmin, ymin, xmax + 1, ymax + 1])
            boxes = torch.as_tensor(boxes, dtype=torch.float32)
            labels = obj_ids.type(torch.int64)
            masks = masks.type(torch.uint8)
            if has_ignore:
                masks[(mask == 255.0).expand_as(masks)] = 255
                masks[(mask == 0.0).expand_as(masks)] = 255
            if flip_label:
                masks = 1 - masks
            area = (boxes[:, 3] - boxes[:, 1]) * (boxes[:, 2] - boxes[:, 0])
            out.append({"boxes": K.stager.put(boxes, device), "boxes_cpu": boxes, "labels": labels, "masks": masks,
                        "image_id": torch.zeros((1,), dtype=torch.int64, device=device),
                        "area": K.stager.put(area, device),
                        "iscrowd": torch.zeros((num_objs,), dtype=torch.int64, device=device)})
        return out

    def _prepare_operands(self, skip_box=False):
        """All 16-bit tensor-core operand layouts of the current parameters of the EAGER modules in ONE launch (cached
        per live tensor).  Graphed parts (trunk; box branch when `skip_box`) build theirs inside their graphs."""
        training = self.training
        graphed = self.use_cuda_graphs and self.capture is None
        key = (training, graphed, skip_box)
        plan = getattr(self, "_operand_plan", None)
        if plan is None or plan[0] != key:
            body = self.backbone.body
            ent = []
            if not graphed:
                ent.append((body.conv1, "weight", "stem"))
            trunk = set(id(m) for mod in (self.backbone, self.rpn.head) for m in mod.modules()) if graphed else set()
            for m in self.modules():
                if id(m) in trunk:
                    continue      # the graphed trunk builds its operands inside the graph from its static inputs
                if isinstance(m, nn.Conv2d) and m is not body.conv1 and m.weight.shape[0] >= 64:
                    ent.append((m, "weight", "f"))
                    if training:
                        ent.append((m, "weight", "t"))
            rh = self.roi_heads
            C = self.backbone.out_channels
            if not skip_box:
                ent.append((rh.box_head.fc6, "weight", ("lf", C)))
                ent.append((rh.box_head.fc7, "weight", ("lf", 0)))
                if training:
                    ent.append((rh.box_head.fc6, "weight", ("lt", C)))
                    ent.append((rh.box_head.fc7, "weight", ("lt", 0)))
            ent.append((rh.mask_predictor.conv5_mask, "weight", "dc"))
            plan = self._operand_plan = (key, ent)
        ops.prep_many([(m._parameters[n], kind) for m, n, kind in plan[1]])

    # ---- transform (tv transform.py:119-160, 25-84, 237-255) ------------------------------------
    def _resized_size(self, h, w):
        tr = self.transform
        if self.training:
            size = tr.torch_choice(tr.min_size)      # consumes CPU RNG exactly like torchvision
        else:
            size = tr.min_size[-1]
        # torchvision 0.26 eager path ("the normal way"): plain Python (double) arithmetic
        scale = min(size / min(h, w), tr.max_size / max(h, w))
        import math
        return int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))

    def _transform(self, inputs, targets, pre=None):
        tr = self.transform
        B, _, h, w = inputs.shape
        if pre is not None:
            x8, (oh, ow), (Hp, Wp) = pre
        else:
            oh, ow = self._resized_size(h, w)
            div = int(tr.size_divisible)
            Hp, Wp = (oh + div - 1) // div * div, (ow + div - 1) // div * div
            x8 = K.transform(inputs.to(torch.float32).contiguous(), oh, ow, Hp, Wp, tr.image_mean, tr.image_std, Cs=8)
        if targets is not None:
            # tv resize_boxes computes the ratios as fp32 tensors; fp32(oh)/fp32(h) is reproduced on the host
            rh = float(torch.tensor(oh, dtype=torch.float32) / torch.tensor(h, dtype=torch.float32))
            rw = float(torch.tensor(ow, dtype=torch.float32) / torch.tensor(w, dtype=torch.float32))
            new = []
            for t in targets:
                t = dict(t)
                xmin, ymin, xmax, ymax = t["boxes"].unbind(1)
                t["boxes"] = torch.stack((xmin * rw, ymin * rh, xmax * rw, ymax * rh), dim=1)
                cx0, cy0, cx1, cy1 = t["boxes_cpu"].unbind(1)
                t["boxes_cpu"] = torch.stack((cx0 * rw, cy0 * rh, cx1 * rw, cy1 * rh), dim=1)
                if self.training:
                    t["masks"] = K.mask_resize_nearest(t["masks"].contiguous(), oh, ow)
                new.append(t)
            targets = new
        return x8, targets, (oh, ow), (Hp, Wp)

    # ---- backbone (tv resnet.py Bottleneck v1.5 + FPN), NHWC bf16 --------------------------------
    def _norm_args(self, n):
        if not isinstance(n, nn.GroupNorm):
            raise NotImplementedError(
                "the B200 path implements the GroupNorm model the reference's configs build "
                "(cfgs/meta.yaml:76 replace_batch_with_group_norms: True)")
        return n.weight, n.bias

    def _bottleneck(self, blk, x):
        s = blk.conv2.stride[0]
        # conv1 forks its input: the identity branch reads the alias, so its gradient is added inside conv1's dgrad
        # epilogue instead of by a separate sum kernel over the block input (3 passes over up to 99 MB)
        fork = torch.is_grad_enabled() and x.requires_grad and blk.conv1.stride[0] == 1
        if fork:
            out, xa = ops.conv_gn(x, blk.conv1.weight, *self._norm_args(blk.bn1), None, 1, 0, True, fork=True)
        else:
            out = ops.conv_gn(x, blk.conv1.weight, *self._norm_args(blk.bn1), None, blk.conv1.stride[0], 0, True)
            xa = x
        out = ops.conv_gn(out, blk.conv2.weight, *self._norm_args(blk.bn2), None, s, 1, True)
        if blk.downsample is not None:
            ds = blk.downsample[0].stride[0]
            identity = ops.conv_gn(xa, blk.downsample[0].weight, *self._norm_args(blk.downsample[1]), None, ds, 0,
                                   False)
        else:
            identity = xa
        return ops.conv_gn(out, blk.conv3.weight, *self._norm_args(blk.bn3), identity, 1, 0, True)

    def theta_home(self):
        """Stable HBM home of every fine-tuned parameter: ONE fp32 arena, each tensor on a 256-byte boundary, in
        MetaModel.param_groups() order.  MetaOptimizer.step writes the updated parameters here (in place from the
        second step on), and the graphed trunk uses these very addresses as its static inputs, so a fine-tune
        iteration no longer copies 161 trunk tensors into the graph (the reference re-allocates every parameter
        on every step, meta_model.py:78-80).  -> (arena, offsets, shapes, {(id(module), name): index})"""
        dev = self.backbone.body.conv1.weight.device
        if self._theta_home is None or self._theta_home[0].device != dev:
            slots = [(m, n) for _, m in self.named_modules()
                     for n, p in m._parameters.items() if p is not None and p.requires_grad]
            offs, shapes, total = [], [], 0
            for m, n in slots:
                p = m._parameters[n]
                offs.append(total)
                shapes.append(tuple(p.shape))
                total += (p.numel() + 63) // 64 * 64
            arena = torch.empty(total, device=dev, dtype=torch.float32)
            self._theta_home = (arena, offs, shapes, {(id(m), n): i for i, (m, n) in enumerate(slots)})
            self._graphs.clear()
        return self._theta_home

    def _trunk_functional(self, x8, *theta):
        """The trunk as a pure function of tensors (what torch.cuda.make_graphed_callables captures)."""
        slots = self._trunk_slots
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            # every 16-bit tensor-core operand of the trunk in ONE launch (static tables -> capturable); the fused
            # small-head operands are rebuilt by the first pyramid level of each call
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            # the RPN head (shared 3x3 conv + objectness / box 1x1s on every level) has static shapes too: inside
            # the graph its ~15 forward and ~45 backward launches cost no host time on the critical path
            feats = self._backbone_eager(x8)
            lvl = int(os.environ.get("EOSVOS_GRAPH_HEAD", "2"))
            if lvl == 0:
                return tuple(feats)
            if lvl == 2 and os.environ.get("EOSVOS_RPN_SPARSE", "1") != "0":
                # RPN head forward only (no autograd nodes): its backward runs outside the graph, for the sampled
                # anchors only (ops.RpnSparseFn).  Outputs: levels, shared-conv outputs t, fused head outputs o.
                with torch.no_grad():
                    ts = self._rpn_head([f.detach() for f in feats], stage=1)
                    outs = self._rpn_head(ts, stage=2)
                return tuple(feats) + tuple(ts) + tuple(outs)
            alias = list(feats)
            if lvl == 1:
                outs = self._rpn_head(feats, stage=1, alias=alias)
            else:
                outs = self._rpn_head(feats, alias=alias)
            return tuple(alias) + tuple(outs)
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    @torch.no_grad()
    def prefetch_backbone(self, inputs):
        """Inference look-ahead: transform + trunk of a FUTURE frame, enqueued now.  The trunk of frame f+1 does not
        depend on frame f's result (only the proposal augmentation and the heads do, reference
        helper_func.py:100-126), so run_frames issues it before blocking on frame f's detections and the GPU works
        through the host-side section.  Two alternating graph instances keep frame f's features intact."""
        if self.training or not self.use_cuda_graphs or self.capture is not None:
            return
        if (os.environ.get("EOSVOS_FRAME_GRAPH", "1") != "0" and self._fast_ok() and self.num_classes == 2
                and self.roi_heads.detections_per_img == 1):
            return            # the whole frame replays as one graph: nothing to run ahead
        x8, _, sizes, padded = self._transform(inputs, None)
        slot = self._pf_slot
        self._pf_slot ^= 1
        feats = self._backbone(x8, slot=slot)
        if len(self._prefetched) >= 2:
            self._prefetched.clear()
        self._prefetched[id(inputs)] = (inputs, x8, sizes, padded, feats, inputs._version, inputs.data_ptr())

    def _backbone(self, x8, slot=0):
        if not self.use_cuda_graphs or self.capture is not None:
            return self._backbone_eager(x8)
        if self._trunk_slots is None:
            lvl = int(os.environ.get("EOSVOS_GRAPH_HEAD", "2"))
            mods = (self.backbone,) if lvl == 0 else ((self.backbone, self.rpn.head.conv) if lvl == 1
                                                      else (self.backbone, self.rpn.head))
            self._trunk_slots = [(m, n) for mod in mods for _, m in mod.named_modules()
                                 for n, p in m._parameters.items() if p is not None and p.requires_grad]
        grad_mode = torch.is_grad_enabled() and self.training
        key = (tuple(x8.shape), grad_mode, x8.device.index, slot)
        conv1 = self.backbone.body.conv1

        def kinds(m, n, t):
            if n != "weight" or not isinstance(m, nn.Conv2d):
                return ()
            if m is conv1:
                return ("stem",)
            if t.shape[0] >= 64:
                return ("f", "t") if grad_mode else ("f",)
            return ()
        return list(self._graphed_call(key, self._trunk_functional, [x8], self._trunk_slots, kinds, grad_mode))

    @staticmethod
    def _capture_inference_graph(fn, sample, nin):
        """Forward-only CUDA graph with a lean replay path (torch's make_graphed_callables wraps every call in an
        autograd Function and re-validates each argument; an inference frame is host-bound, so that matters):
        copy the inputs whose address differs from the static one, replay, hand out the static outputs."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(3):
                fn(*sample)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            outs = fn(*sample)
        ptrs = [t.data_ptr() for t in sample]

        def replay(*args):
            with torch.no_grad():
                for s, p, a in zip(sample, ptrs, args):
                    if a.data_ptr() != p:
                        s.copy_(a)
            graph.replay()
            return outs
        replay.graph, replay.static = graph, (sample, outs)     # keep the captured buffers alive
        return replay

    def _graphed_call(self, key, fn, inputs, slots, kinds, grad_mode, alias_inputs=0):
        """Runs fn(*inputs, *theta) -- theta = the tensors currently installed in `slots` -- as a CUDA graph (forward
        and, in grad mode, backward), capturing it on first use.  The graph's static parameter inputs are the
        parameters' home addresses (theta_home), its tensor-core operands live in persistent buffers filled by one
        captured launch (kinds(module, name, tensor) -> operand kinds of that parameter)."""
        theta = [m._parameters[n] for m, n in slots]
        ent = self._graphs.get(key)
        if ent is None:
            if os.environ.get("EOSVOS_DEBUG_GRAPHS"):
                print("[eosvos] capturing graph", key[:2], "have", len(self._graphs), flush=True)
            arena, offs, shapes, index = self.theta_home()
            # the first `alias_inputs` inputs already live at fixed addresses (static outputs of another graph): the
            # new graph reads them in place instead of receiving a copy on every call
            sample = [(t.detach() if i < alias_inputs else t.detach().clone()).requires_grad_(t.requires_grad and grad_mode)
                      for i, t in enumerate(inputs)]
            nin = len(sample)
            with torch.no_grad():
                for (m, n), t in zip(slots, theta):
                    i = index.get((id(m), n))
                    if i is None or shapes[i] != tuple(t.shape):
                        sample.append(t.detach().clone().requires_grad_(grad_mode))
                        continue
                    h = arena[offs[i]:offs[i] + t.numel()].view(shapes[i])
                    if h.data_ptr() != t.data_ptr():
                        h.copy_(t)
                    sample.append(h.requires_grad_(grad_mode))
            from .. import _lib
            try:      # graphed backward runs on the capture stream; the resulting AccumulateGrad stream note is benign
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
            # operand buffers + conversion tables of this graph instance (persistent: the graph bakes their addresses)
            reqs = [(t, kind) for (m, n), t in zip(slots, sample[nin:]) for kind in kinds(m, n, t)]
            plan, vals = ops.build_prep_plan(reqs) if reqs else (None, {})
            self._active_plan = plan
            ops._scope = {(id(reqs[i][0]), kind): v for (i, kind), v in vals.items()}
            c0 = _lib.launch_count()
            scope0 = dict(ops._scope)
            import gc
            gc_was_on = gc.isenabled()
            for attempt in (0, 1):
                try:
                    # thread-local capture mode + the capture lock: nothing another thread does (the augmentation
                    # worker allocates and synchronises on its own stream) may invalidate the capture; no cyclic
                    # garbage collection in between either (torch collects right before the capture begins): a
                    # collected model's graphs would release their memory pools in the middle of the capture
                    gc.disable()
                    with K.capture_lock, _ThreadLocalCapture():
                        if grad_mode:
                            with torch.enable_grad():
                                # (parameters whose gradient is produced outside the graph -- the RPN head under the
                                # sparse backward -- are unused inside)
                                graphed = torch.cuda.make_graphed_callables(fn, tuple(sample), allow_unused_input=True)
                        else:
                            graphed = self._capture_inference_graph(fn, sample, nin)
                    break
                except Exception as e:      # noqa: BLE001
                    # a capture invalidated from outside (e.g. the allocator releasing another graph's pool when the
                    # garbage collector runs mid-capture) is not an error of this graph: start it over once
                    if attempt or "capture" not in str(e).lower():
                        self._active_plan, ops._scope = None, None
                        raise
                    import sys
                    print(f"[eosvos] graph capture of {key[0]} restarted: {str(e).splitlines()[0][:160]}", file=sys.stderr)
                    self.capture_restarts = getattr(self, "capture_restarts", 0) + 1
                    torch.cuda.synchronize()
                    K.zero_pool.cap_block = None        # (carved from the abandoned capture's pool)
                    c0 = _lib.launch_count()
                    ops._scope = dict(scope0)
                    self._active_plan = plan
                finally:
                    if gc_was_on:
                        gc.enable()
            self._active_plan, ops._scope = None, None
            per_call = (_lib.launch_count() - c0) // 4      # 3 eager warm-up runs + 1 capture of the same kernels
            # the warm-up backward ran on uninitialised output gradients (torch's warm-up passes empty_like tensors):
            # never continue in a zero block it touched
            K.zero_pool.reset()
            ent = (graphed, per_call, plan, vals)
            if len(self._graphs) >= 40:         # bounded: graphs pin their activation pools
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = ent
        from .. import _lib
        _lib.add_replayed_launches(ent[1])
        return ent[0](*inputs, *theta)

    def _backbone_eager(self, x8):
        body = self.backbone.body
        x = ops.stem(x8, body.conv1.weight, *self._norm_args(body.bn1))
        x = ops.maxpool3x3s2(x)
        feats = []
        for name in ("layer1", "layer2", "layer3", "layer4"):
            for blk in getattr(body, name):
                x = self._bottleneck(blk, x)
            feats.append(x)
        fpn = self.backbone.fpn

        def conv_of(m):
            return m[0] if isinstance(m, nn.Sequential) else m

        inner = [conv_of(m) for m in fpn.inner_blocks]
        layer = [conv_of(m) for m in fpn.layer_blocks]
        last_inner = ops.conv2d(feats[3], inner[3].weight, inner[3].bias)
        results = [ops.conv2d(last_inner, layer[3].weight, layer[3].bias, pad=1)]
        for idx in (2, 1, 0):
            fh, fw = feats[idx].shape[1:3]
            if fh != 2 * last_inner.shape[1] or fw != 2 * last_inner.shape[2]:
                raise NotImplementedError("FPN levels must be exact 2x of each other (inputs are padded to /32)")
            last_inner = ops.conv2d(feats[idx], inner[idx].weight, inner[idx].bias, res=last_inner, res_half=True)
            results.insert(0, ops.conv2d(last_inner, layer[idx].weight, layer[idx].bias, pad=1))
        results.append(ops.subsample2(results[-1]))
        return results  # P2..P6

    # ---- RPN (reference mask_rcnn.py:217-344) ---------------------------------------------------
    def _anchors(self, image_shape, image_sizes, feat_shapes, device):
        """Anchors of one (batch, padded size, pyramid) signature, generated once and kept per signature (fine-tuning
        at batch 3 and inference at batch 1 alternate).  Generation is synchronised before the entry is published, so
        every stream (the RPN side stream in particular) may read them without further ordering."""
        key = (tuple(image_shape), tuple(image_sizes), tuple(feat_shapes), str(device))
        hit = self._anchor_cache.get(key)
        if hit is None:
            il = _ImageListLike(image_shape, list(image_sizes), device)
            fms = [torch.empty((image_shape[0], 1, h, w), device=device, dtype=torch.float32) for h, w in feat_shapes]
            hit = [a.detach() for a in self.rpn.anchor_generator(il, fms)]
            torch.cuda.current_stream(device).synchronize()     # once per signature
            if len(self._anchor_cache) >= 64:     # (well above the graph cache: captured graphs read these in place)
                self._anchor_cache.pop(next(iter(self._anchor_cache)))
            self._anchor_cache[key] = hit
        return hit

    def _rpn_head(self, feats, stage=0, alias=None):
        """tv RPNHead on every level: [pixels, 16] fp32 = (A objectness | 4A box deltas | padding) per level.
        stage 1: only the shared 3x3 conv; stage 2: only the 1x1 heads on its output (debug split)."""
        head = self.rpn.head
        conv = head.conv[0][0] if isinstance(head.conv, nn.Sequential) else head.conv
        outs = []
        for i, f in enumerate(feats):
            C = f.shape[-1]
            if stage == 2:
                t = f
            elif alias is not None and torch.is_grad_enabled() and f.requires_grad:
                # the level is also consumed outside (RoIAlign, coarser levels): hand out the alias so those
                # gradients are added inside this conv's dgrad epilogue
                t, fa = ops.conv2d(f, conv.weight, conv.bias, pad=1, relu=True, fork=True)
                alias[i] = fa
            else:
                t = ops.conv2d(f, conv.weight, conv.bias, pad=1, relu=True)
            if stage == 1:
                outs.append(t)
                continue
            outs.append(ops.fused_heads(t.reshape(-1, C), [head.cls_logits.weight, head.bbox_pred.weight],
                                        [head.cls_logits.bias, head.bbox_pred.bias]))
        return outs

    def _rpn_early_targets(self, feats, image_shape, image_sizes, targets):
        """Anchor labelling + sampling (tv rpn.py assign_targets_to_anchors, box_coder.encode, fg_bg_sampler) depend only
        on the anchors and the ground truth, not on the network: run them on a side stream now, so their host syncs
        (nonzero) do not wait for the trunk that is still executing on the main stream.  RNG order is unchanged (RPN
        sampler before RoI sampler)."""
        rpn = self.rpn
        fs = [(f.shape[1], f.shape[2]) for f in feats]
        anchors = self._anchors(image_shape, image_sizes, fs, feats[0].device)
        main = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=feats[0].device)
        with torch.cuda.stream(self._side_stream):
            # the ground-truth boxes are uploaded again on THIS stream (from their host copy), so nothing here
            # is ordered behind the main stream's queue
            side_targets = [{"boxes": K.stager.put(t["boxes_cpu"], feats[0].device)} for t in targets]
            labels, matched_gt_boxes = rpn.assign_targets_to_anchors(anchors, side_targets)
            regression_targets = rpn.box_coder.encode(matched_gt_boxes, anchors)
            pos, neg = rpn.fg_bg_sampler(labels)
            pos = torch.where(torch.cat(pos, dim=0))[0]
            neg = torch.where(torch.cat(neg, dim=0))[0]
            early = (torch.cat(labels, dim=0), torch.cat(regression_targets, dim=0), pos, neg)
            done = torch.cuda.Event()
            done.record(self._side_stream)
        for t in early:
            t.record_stream(main)
        return early + (done,)

    def _rpn_cat_outputs(self, feats, head_outs):
        head = self.rpn.head
        N = feats[0].shape[0]
        A = head.cls_logits.weight.shape[0]
        obj, dlt, feat_shapes = [], [], []
        for f, o in zip(feats, head_outs):
            _, H, W, C = f.shape
            obj.append(o[:, :A].reshape(N, H * W * A, 1))
            dlt.append(o[:, A:A + 4 * A].reshape(N, H * W * A, 4))
            feat_shapes.append((H, W))
        num_anchors_per_level = [o.shape[1] for o in obj]
        objectness = torch.cat(obj, dim=1).flatten(0, -2)
        pred_bbox_deltas = torch.cat(dlt, dim=1).flatten(0, -2)
        return objectness, pred_bbox_deltas, feat_shapes, num_anchors_per_level

    @staticmethod
    def _rpn_losses(early, objectness, pred_bbox_deltas):
        """tv rpn.py compute_loss with the indices sampled on the side stream."""
        labels_c, reg_c, pos, neg, done = early
        torch.cuda.current_stream().wait_event(done)
        sampled = torch.cat([pos, neg], dim=0)
        loss_rpn_box_reg = F.smooth_l1_loss(pred_bbox_deltas[pos], reg_c[pos], beta=1 / 9,
                                            reduction="sum") / (sampled.numel())
        loss_objectness = F.binary_cross_entropy_with_logits(objectness.flatten()[sampled], labels_c[sampled])
        return {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}

    def _segment_offsets(self, N, sizes, device):
        # kept per signature for the life of the model: captured graphs hold the address of the entry they saw
        key = (N, tuple(sizes), str(device))
        cache = self.__dict__.setdefault("_seg_cache", {})
        hit = cache.get(key)
        if hit is None:
            offs = [0]
            for _ in range(N):
                for k in sizes:
                    offs.append(offs[-1] + k)
            hit = cache[key] = torch.tensor(offs, dtype=torch.int32, device=device)
        return hit

    def _rpn_fast(self, feats, image_shape, image_sizes, head_outs, post_n, out=None, out_offset=0):
        """tv rpn.py filter_proposals on statically shaped buffers (csrc/rpn.cu): per-level top-k + decode of the
        selected anchors only, segmented NMS, post-NMS top-n -> (boxes [N, stride, 4] with the first count[n] of the
        post_n slots at out_offset real and the rest zero, count int32 [N] on the device).  No host synchronisation."""
        rpn = self.rpn
        N = feats[0].shape[0]
        device = feats[0].device
        A = rpn.head.cls_logits.weight.shape[0]
        feat_shapes = [(f.shape[1], f.shape[2]) for f in feats]
        hw = [h * w for h, w in feat_shapes]
        anchors = self._anchors(image_shape, image_sizes, feat_shapes, device)[0]
        pre = rpn.pre_nms_top_n()
        boxes_c, scores_c, valid_c, ks = K.rpn_select([o.detach() for o in head_outs], hw, A, N, anchors, image_sizes, pre,
                                                      rpn.box_coder.bbox_xform_clip, rpn.min_size, rpn.score_thresh)
        seg = self._segment_offsets(N, ks, device)
        keep = K.nms_segments(boxes_c.view(-1, 4), seg, N * len(ks), max(ks), rpn.nms_thresh)
        boxes, _, count = K.rpn_postnms(hw, A, N, pre, boxes_c, scores_c, valid_c, keep, post_n, out, out_offset)
        return boxes, count

    def _extend_rands(self, B, G, n_aug, device):
        """The CPU uniform numbers of reference mask_rcnn.py:270-273, drawn call by call in the reference's order
        (x_min, y_min, x_max, y_max per target box) into pinned memory, uploaded asynchronously."""
        ring = getattr(self, "_rand_ring", None)
        if ring is None or ring[0].shape != (B, G, 4, n_aug):
            ring = [torch.empty((B, G, 4, n_aug), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._rand_ring, self._rand_events, self._rand_i = ring, [None] * 4, 0
        j = self._rand_i % 4
        self._rand_i += 1
        if self._rand_events[j] is not None:
            self._rand_events[j].synchronize()
        buf = ring[j]
        for b in range(B):
            for g in range(G):
                for c in range(4):
                    torch.rand((n_aug,), out=buf[b, g, c])
        dev = buf.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._rand_events[j] = ev
        return dev

    def _proposal_list(self, padded, count, n_front):
        """padded proposals -> the reference's per-image lists (host sync; checkers and the general path only)."""
        cnt = count.tolist()
        return [torch.cat([padded[i, :min(c, n_front)], padded[i, n_front:]], dim=0) for i, c in enumerate(cnt)]

    def _rpn(self, feats, image_shape, image_sizes, targets, head_outs=None):
        rpn = self.rpn
        N = feats[0].shape[0]
        early = None
        if self.training:
            early = self._rpn_early_targets(feats, image_shape, image_sizes, targets)
        if head_outs is None:
            head_outs = self._rpn_head(feats)
        elif head_outs[0].dtype != torch.float32:          # debug split: only the shared conv ran inside the graph
            head_outs = self._rpn_head(head_outs, stage=2)
        objectness, pred_bbox_deltas, feat_shapes, num_anchors_per_level = self._rpn_cat_outputs(feats, head_outs)
        if self.capture is not None:
            self.capture.update(objectness=objectness.detach(), deltas=pred_bbox_deltas.detach())
        anchors = self._anchors(image_shape, image_sizes, feat_shapes, feats[0].device)
        proposals = self._decode(pred_bbox_deltas.detach(), anchors, rpn.box_coder).view(N, -1, 4)
        boxes, scores = self._filter_proposals(proposals, objectness, image_sizes, num_anchors_per_level)

        mode = rpn._eval_augment_proposals_mode
        if not self.training and targets is not None and mode is not None:
            # reference mask_rcnn.py:251-332: jittered copies of the previous-frame box (CPU torch.rand, same
            # call order => same random stream)
            random_share = 0.1
            post = rpn.post_nms_top_n()
            num_box_augs = post // 2 if mode == 'EXTEND' else post
            img_height, img_width = image_shape[-2:]
            for i, target in enumerate(targets):
                tb = target['boxes_cpu']
                target_boxes = []
                for box in tb:
                    bw, bh = box[2] - box[0], box[3] - box[1]
                    x_mins = box[0] - torch.rand((num_box_augs,)) * bw * random_share
                    y_mins = box[1] - torch.rand((num_box_augs,)) * bh * random_share
                    x_maxs = box[2] + torch.rand((num_box_augs,)) * bw * random_share
                    y_maxs = box[3] + torch.rand((num_box_augs,)) * bh * random_share
                    target_boxes.append(torch.stack([x_mins.clamp(0, img_width), y_mins.clamp(0, img_height),
                                                     x_maxs.clamp(0, img_width), y_maxs.clamp(0, img_height)], dim=1))
                target_boxes = K.stager.put(torch.cat(target_boxes, dim=0), scores[0].device)
                if mode == 'EXTEND':
                    boxes[i] = torch.cat([boxes[i][:post // 2], target_boxes], dim=0)
                elif mode == 'REPLACE':
                    boxes[i] = target_boxes
                else:
                    raise NotImplementedError

        losses = {}
        if self.training:
            losses = self._rpn_losses(early, objectness, pred_bbox_deltas)
        return boxes, losses

    @staticmethod
    def _decode(rel_codes, anchors, coder):
        """tv _utils.py BoxCoder.decode / decode_single restated with Python scalars (torchvision builds its 0.5
        constants with torch.tensor(..., device=cuda): a blocking H2D copy, i.e. a full stream sync per call)."""
        boxes = torch.cat(anchors, dim=0).to(rel_codes.dtype)
        wx, wy, ww, wh = coder.weights
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        dx = rel_codes[:, 0::4] / wx
        dy = rel_codes[:, 1::4] / wy
        dw = torch.clamp(rel_codes[:, 2::4] / ww, max=coder.bbox_xform_clip)
        dh = torch.clamp(rel_codes[:, 3::4] / wh, max=coder.bbox_xform_clip)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        c_to_c_h = 0.5 * pred_h
        c_to_c_w = 0.5 * pred_w
        out = torch.stack((pred_ctr_x - c_to_c_w, pred_ctr_y - c_to_c_h, pred_ctr_x + c_to_c_w,
                           pred_ctr_y + c_to_c_h), dim=2).flatten(1)
        return out.reshape(rel_codes.shape[0], -1, 4)

    def _filter_proposals(self, proposals, objectness, image_sizes, num_anchors_per_level):
        """tv rpn.py filter_proposals (called from reference mask_rcnn.py:249): per-level top-k, clip, drop small /
        low-score boxes, per-level NMS, keep post_nms_top_n by score.  NMS runs as ONE segmented launch over all
        (image, level) pairs (nms.cu); invalid boxes are zeroed in place (zero area: IoU 0, suppress nothing) so
        that segments keep their static top-k sizes."""
        rpn = self.rpn
        N = proposals.shape[0]
        device = proposals.device
        objectness = objectness.detach().reshape(N, -1)
        idx, sizes, offset = [], [], 0
        for ob in objectness.split(num_anchors_per_level, 1):
            k = min(rpn.pre_nms_top_n(), ob.shape[1])
            idx.append(ob.topk(k, dim=1)[1] + offset)          # sorted by descending score inside the level
            sizes.append(k)
            offset += ob.shape[1]
        top_n_idx = torch.cat(idx, dim=1)
        batch_idx = torch.arange(N, device=device)[:, None]
        prob = torch.sigmoid(objectness[batch_idx, top_n_idx])
        props = proposals[batch_idx, top_n_idx]
        Ktot = top_n_idx.shape[1]
        seg_offsets = self._segment_offsets(N, sizes, device)
        hw = K.stager.put(torch.tensor([[s[0], s[1]] for s in image_sizes], dtype=props.dtype), device)
        hmax, wmax = hw[:, 0:1], hw[:, 1:2]
        x1 = torch.minimum(props[..., 0].clamp(min=0), wmax)
        y1 = torch.minimum(props[..., 1].clamp(min=0), hmax)
        x2 = torch.minimum(props[..., 2].clamp(min=0), wmax)
        y2 = torch.minimum(props[..., 3].clamp(min=0), hmax)
        valid = ((x2 - x1) >= rpn.min_size) & ((y2 - y1) >= rpn.min_size) & (prob >= rpn.score_thresh)
        clipped = torch.stack([x1, y1, x2, y2], dim=-1)
        nms_in = (clipped * valid[..., None]).reshape(N * Ktot, 4).contiguous()
        flags = K.nms_segments(nms_in, seg_offsets, N * len(sizes), max(sizes), rpn.nms_thresh)
        flags = flags.view(N, Ktot).bool() & valid
        ranked = torch.where(flags, prob, torch.full_like(prob, -1.0))
        order = ranked.argsort(dim=1, descending=True, stable=True)
        counts = flags.sum(dim=1).tolist()                        # one host sync for the batch
        post = rpn.post_nms_top_n()
        final_boxes, final_scores = [], []
        for i in range(N):
            keep = order[i, :min(counts[i], post)]
            final_boxes.append(clipped[i, keep])
            final_scores.append(prob[i, keep])
        return final_boxes, final_scores

    def _nms_by_label(self, boxes, scores, labels, thresh):
        """batched_nms semantics (one NMS per label), result sorted by descending score."""
        n = boxes.shape[0]
        if n == 0:
            return torch.zeros((0,), dtype=torch.int64, device=boxes.device)
        o1 = scores.argsort(descending=True, stable=True)
        o2 = labels[o1].argsort(stable=True)
        order = o1[o2]                                            # grouped by label, descending score inside
        nlab = self.num_classes
        offs = torch.zeros(nlab + 1, dtype=torch.int32, device=boxes.device)
        offs[1:] = torch.bincount(labels, minlength=nlab)[:nlab].cumsum(0).to(torch.int32)
        flags = K.nms_segments(boxes[order].to(torch.float32).contiguous(), offs, nlab, n, thresh).bool()
        kept = order[flags]
        return kept[scores[kept].argsort(descending=True, stable=True)]

    # ---- RoI heads (reference mask_rcnn.py:95-214, 347-420) -------------------------------------
    _SCALES = (1 / 4, 1 / 8, 1 / 16, 1 / 32)

    @staticmethod
    def _rois5(boxes):
        return torch.cat([torch.cat([torch.full((b.shape[0], 1), i, dtype=b.dtype, device=b.device), b], dim=1)
                          for i, b in enumerate(boxes)], dim=0).to(torch.float32).contiguous()

    def _postprocess_detections(self, class_logits, box_regression, proposals, image_shapes):
        rh = self.roi_heads
        device = class_logits.device
        num_classes = class_logits.shape[-1]
        boxes_per_image = [len(b) for b in proposals]
        pred_boxes = rh.box_coder.decode(box_regression, proposals)
        pred_scores = F.softmax(class_logits, -1)
        pred_boxes = pred_boxes.split(boxes_per_image, 0)
        pred_scores = pred_scores.split(boxes_per_image, 0)
        all_boxes, all_scores, all_labels, all_rows = [], [], [], []
        for boxes, scores, image_shape in zip(pred_boxes, pred_scores, image_shapes):
            boxes = box_ops.clip_boxes_to_image(boxes, image_shape)
            labels = torch.arange(num_classes, device=device).view(1, -1).expand_as(scores)
            boxes, scores, labels = boxes[:, 1:], scores[:, 1:], labels[:, 1:]
            boxes, scores, labels = boxes.reshape(-1, 4), scores.flatten(), labels.flatten()
            inds = torch.nonzero(scores > rh.score_thresh).squeeze(1)
            boxes, scores, labels = boxes[inds], scores[inds], labels[inds]
            keep = box_ops.remove_small_boxes(boxes, min_size=1e-2)
            boxes, scores, labels, inds = boxes[keep], scores[keep], labels[keep], inds[keep]
            keep = self._nms_by_label(boxes, scores, labels, rh.nms_thresh)
            keep = keep[:rh.detections_per_img]
            all_boxes.append(boxes[keep])
            all_scores.append(scores[keep])
            all_labels.append(labels[keep])
            all_rows.append(inds[keep])
        # row (proposal index * (num_classes - 1) + class - 1) every kept detection came from: lets a checker tell
        # "another box" from "the same box, numerically different"
        self.last_detection_rows = all_rows
        return all_boxes, all_scores, all_labels

    def _mask_branch(self, feats, mask_proposals):
        return self._mask_branch_rois(feats, self._rois5(mask_proposals))

    def _mask_branch_rois(self, feats, rois5):
        rh = self.roi_heads
        P = self._roi_sizes['mask']
        x = ops.roi_align(feats[:4], self._SCALES, rois5, P)
        for blk in rh.mask_head:
            conv = blk[0] if isinstance(blk, nn.Sequential) else blk
            x = ops.conv2d(x, conv.weight, conv.bias, pad=1, relu=True)
        mp = rh.mask_predictor
        x = ops.deconv2x2(x, mp.conv5_mask.weight, mp.conv5_mask.bias, relu=True)
        n, M, _, C = x.shape
        o = ops.fused_heads(x.reshape(-1, C), [mp.mask_fcn_logits.weight], [mp.mask_fcn_logits.bias])
        ncls = mp.mask_fcn_logits.weight.shape[0]
        return o[:, :ncls].reshape(n, M, M, ncls).permute(0, 3, 1, 2).contiguous()   # [n, ncls, M, M] fp32

    def _select_training_samples(self, proposals, targets):
        """tv roi_heads.py:642-678 (select_training_samples) + _utils.py BalancedPositiveNegativeSampler with the same
        operations in the same order (same torch.randperm calls => same random stream), but ONE host sync (the
        candidate counts of all images) instead of four `torch.where` syncs per image: every index list is then a
        `nonzero_static` of known size.  Also returns the positions of the positives inside each image's sample."""
        rh = self.roi_heads
        dtype, device = proposals[0].dtype, proposals[0].device
        gt_boxes = [t["boxes"].to(dtype) for t in targets]
        gt_labels = [t["labels"] for t in targets]
        if any(g.numel() == 0 for g in gt_boxes):          # background image: keep torchvision's own special-casing
            p, m, l, r = rh.select_training_samples(proposals, targets)
            return p, m, l, r, [torch.nonzero(x > 0).squeeze(1) for x in l]
        proposals = rh.add_gt_proposals(proposals, gt_boxes)
        matched_idxs, labels = rh.assign_targets_to_proposals(proposals, gt_boxes, gt_labels)
        counts = torch.stack([torch.stack(((l >= 1).sum(), (l == 0).sum())) for l in labels]).tolist()
        sampler = rh.fg_bg_sampler
        out_p, out_m, out_l, out_g, pos_in = [], [], [], [], []
        for i, (l, (npos, nneg)) in enumerate(zip(labels, counts)):
            positive = torch.nonzero_static(l >= 1, size=npos).squeeze(1)
            negative = torch.nonzero_static(l == 0, size=nneg).squeeze(1)
            num_pos = min(npos, int(sampler.batch_size_per_image * sampler.positive_fraction))
            num_neg = min(nneg, sampler.batch_size_per_image - num_pos)
            perm1 = torch.randperm(npos, device=device)[:num_pos]
            perm2 = torch.randperm(nneg, device=device)[:num_neg]
            mask = torch.zeros_like(l, dtype=torch.bool)
            mask[positive[perm1]] = True
            mask[negative[perm2]] = True
            inds = torch.nonzero_static(mask, size=num_pos + num_neg).squeeze(1)
            p_i, l_i, m_i = proposals[i][inds], l[inds], matched_idxs[i][inds]
            out_p.append(p_i)
            out_l.append(l_i)
            out_m.append(m_i)
            out_g.append(gt_boxes[i][m_i])
            pos_in.append(torch.nonzero_static(l_i > 0, size=num_pos).squeeze(1))
        regression_targets = rh.box_coder.encode(out_g, out_p)
        return out_p, out_m, out_l, regression_targets, pos_in

    def _box_branch(self, feats4, rois5):
        """RoIAlign 7x7 -> fc6 -> fc7 -> class / box predictors (tv faster_rcnn.py:286-377) -> [R, 16] fp32."""
        rh = self.roi_heads
        Pb = self._roi_sizes['box']
        bx = ops.roi_align(feats4, self._SCALES, rois5, Pb)
        R, _, _, C = bx.shape
        h = ops.linear(bx.reshape(R, Pb * Pb * C), rh.box_head.fc6.weight, rh.box_head.fc6.bias, relu=True, inner=C)
        h = ops.linear(h, rh.box_head.fc7.weight, rh.box_head.fc7.bias, relu=True)
        bp = rh.box_predictor
        return ops.fused_heads(h, [bp.cls_score.weight, bp.bbox_pred.weight], [bp.cls_score.bias, bp.bbox_pred.bias])

    @staticmethod
    def _fastrcnn_loss_static(class_logits, box_regression, labels, reg_targets):
        """tv roi_heads.py:12-53 (fastrcnn_loss) on concatenated labels / targets, with the positive rows selected by
        a mask instead of `torch.where` (static shapes, no host sync): same terms, same 1/N normalisation."""
        loss_cls = F.cross_entropy(class_logits, labels)
        N, nc = class_logits.shape
        sel = box_regression.reshape(N, nc, 4).gather(1, labels.clamp(min=0)[:, None, None].expand(N, 1, 4))
        per = F.smooth_l1_loss(sel.squeeze(1), reg_targets, beta=1 / 9, reduction="none").sum(dim=1)
        return loss_cls, (per * (labels > 0).to(per.dtype)).sum() / N

    def _box_train_functional(self, f0, f1, f2, f3, rois5, labels, reg_targets, *theta):
        """Box branch + tv fastrcnn_loss (roi_heads.py:12-53) as a pure function with static shapes (R = 512 per
        image), what the second CUDA graph of a training iteration captures.  The loss is the reference's, with the
        positive-row selection written as a mask (`labels > 0`) instead of a data-dependent index list."""
        slots = self._box_slots
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            o = self._box_branch([f0, f1, f2, f3], rois5)
            nc = self.roi_heads.box_predictor.cls_score.weight.shape[0]
            class_logits, box_regression = o[:, :nc], o[:, nc:nc + 4 * nc]
            return self._fastrcnn_loss_static(class_logits, box_regression, labels, reg_targets)
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    _MASK_BUCKETS = (8, 32, 128)

    def _mask_train_functional(self, f0, f1, f2, f3, rois5, labels, gt_masks, tgt_rois, weights, *theta):
        """Mask branch + Lovasz loss on a padded, fixed-size list of positive RoIs (rows with weight 0 are padding)
        as a pure function with static shapes -- the third CUDA graph pair of a training iteration."""
        slots = self._mask_slots
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            logits = self._mask_branch_rois([f0, f1, f2, f3], rois5)
            tg = K.mask_targets(gt_masks, tgt_rois, logits.shape[-1])
            return ops.mask_loss_weighted(logits, labels, tg, weights)
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    def _mask_loss_graphed(self, feats, mask_proposals, pos_matched_idxs, targets):
        """Pads the positive RoIs of the batch to the next bucket size and replays the mask graph; None when the
        case is not covered (no positives, BCE, ragged ground truth)."""
        rh = self.roi_heads
        n = sum(p.shape[0] for p in mask_proposals)
        cap = len(mask_proposals) * int(rh.fg_bg_sampler.batch_size_per_image * rh.fg_bg_sampler.positive_fraction)
        sizes = [b for b in self._MASK_BUCKETS if b < cap] + [cap]
        if n == 0 or n > cap or rh.maskrcnn_loss != 'LOVASZ':
            return None
        R = next(b for b in sizes if b >= n)
        device = feats[0].device
        gt_masks = torch.cat([t["masks"] for t in targets], 0).contiguous()
        gt_labels = [t["labels"] for t in targets]
        lab = torch.cat([l[i] for l, i in zip(gt_labels, pos_matched_idxs)], dim=0)
        rois5 = self._rois5(mask_proposals)
        offs, tro = 0, []
        for t, p, i in zip(targets, mask_proposals, pos_matched_idxs):
            tro.append(torch.cat([(i + offs).to(p)[:, None], p], dim=1))
            offs += t["masks"].shape[0]
        tro = torch.cat(tro, 0).to(torch.float32)
        pad = R - n
        w = torch.full((R,), 1.0 / n, device=device, dtype=torch.float32)
        if pad:
            w[n:] = 0.0
            fill = rois5.new_zeros((pad, 5))
            fill[:, 0] = -1.0                       # negative image index = padding row (skipped by the kernels)
            rois5 = torch.cat([rois5, fill], 0)
            tro = torch.cat([tro, fill], 0)
            lab = torch.cat([lab, lab.new_ones(pad)], 0)
        if self._mask_slots is None:
            mods = (rh.mask_head, rh.mask_predictor)
            self._mask_slots = [(m, n_) for mod in mods for _, m in mod.named_modules()
                                for n_, p in m._parameters.items() if p is not None and p.requires_grad]

        def kinds(m, n_, t):
            if n_ != "weight":
                return ()
            if isinstance(m, nn.ConvTranspose2d):
                return ("dc",)
            if isinstance(m, nn.Conv2d) and t.shape[0] >= 64:
                return ("f", "t")
            return ()
        key = ("mask", R, tuple(gt_masks.shape), tuple(tuple(f.shape) for f in feats[:4]), device.index)
        return self._graphed_call(key, self._mask_train_functional,
                                  list(feats[:4]) + [rois5.contiguous(), lab.contiguous(), gt_masks, tro.contiguous(), w],
                                  self._mask_slots, kinds, True, alias_inputs=4)

    # ---- statically shaped fast paths (csrc/rpn.cu) -----------------------------------------------
    def _fast_ok(self):
        w = self.rpn.box_coder.weights
        return (os.environ.get("EOSVOS_FAST_PATH", "1") != "0" and self.capture is None and self.fixed_proposals is None
                and self.fixed_detections is None and tuple(float(x) for x in w) == (1.0, 1.0, 1.0, 1.0))

    def _pyramid_shapes(self, Hp, Wp):
        """Spatial sizes of P2..P6 for a padded input (strides 4..32, then the stride-2 sub-sampling of P5)."""
        shapes = [(Hp // s, Wp // s) for s in (4, 8, 16, 32)]
        shapes.append(((shapes[-1][0] - 1) // 2 + 1, (shapes[-1][1] - 1) // 2 + 1))
        return shapes

    def _anchor_match_async(self, image_shape, image_sizes, targets, device):
        """tv rpn.py assign_targets_to_anchors as two kernels (csrc/rpn.cu::anchor_match_kernel), queued BEFORE the
        trunk: labels depend only on the anchors and the ground truth.  The foreground / background counts travel to
        pinned memory asynchronously; `_sample_anchors_fast` picks them up after the trunk has been queued."""
        rpn = self.rpn
        N = image_shape[0]
        feat_shapes = self._pyramid_shapes(image_shape[2], image_shape[3])
        anchors = self._anchors(image_shape, image_sizes, feat_shapes, device)[0]
        gt_boxes = torch.cat([t["boxes"].to(torch.float32) for t in targets], 0).contiguous()
        offs = [0]
        for t in targets:
            offs.append(offs[-1] + t["boxes"].shape[0])
        gt_off = K.stager.put(torch.tensor(offs, dtype=torch.int32), device)
        m = rpn.proposal_matcher
        if not m.allow_low_quality_matches:
            raise NotImplementedError("RPN matcher without low-quality matches")
        labels, matched, counts = K.rpn_anchor_match(anchors, gt_boxes, gt_off, N, m.high_threshold, m.low_threshold)
        pin = getattr(self, "_anchor_count_pin", None)
        if pin is None or pin.shape[0] != N:
            pin = self._anchor_count_pin = torch.empty((N, 2), dtype=torch.int32).pin_memory()
        pin.copy_(counts, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return dict(labels=labels, matched=matched, anchors=anchors, gt_boxes=gt_boxes, gt_off=gt_off, pin=pin, ev=ev)

    def _sample_anchors_fast(self, am):
        """tv rpn.py fg_bg_sampler (BalancedPositiveNegativeSampler(256, 0.5)): the reference's two `torch.randperm`
        draws per image, then one selection kernel -> int64 positions of the sampled anchors in the flattened
        (image, level, pixel, anchor) order."""
        sampler = self.rpn.fg_bg_sampler
        labels = am["labels"]
        device = labels.device
        am["ev"].synchronize()            # queued before the trunk: long done
        S = sampler.batch_size_per_image
        Pmax = int(S * sampler.positive_fraction)
        perms, num_pos_list, num_neg_list = [], [], []
        for npos, nneg in am["pin"].tolist():
            num_pos = min(npos, Pmax)
            num_neg = min(nneg, S - num_pos)
            perms.append((torch.randperm(npos, device=device), torch.randperm(nneg, device=device)))
            num_pos_list.append(num_pos)
            num_neg_list.append(num_neg)
        inds, _ = K.roi_sample(labels, perms, num_pos_list, num_neg_list, S, Pmax)
        A_total = labels.shape[1]
        sizes = [a + b for a, b in zip(num_pos_list, num_neg_list)]
        if all(sz == S for sz in sizes):
            key = (len(sizes), A_total, str(device))
            if getattr(self, "_anchor_base_key", None) != key:
                self._anchor_base = (torch.arange(len(sizes), device=device, dtype=torch.int64) * A_total)[:, None]
                self._anchor_base_key = key
            return (inds + self._anchor_base).reshape(-1)
        return torch.cat([inds[i, :sz] + i * A_total for i, sz in enumerate(sizes)], 0)

    def _match_rois_fast(self, padded, count, targets):
        """Phase 1 of tv roi_heads.py:642-678 (add_gt_proposals + assign_targets_to_proposals) on the padded proposal
        buffer: one kernel, then an asynchronous read-back of the foreground / background counts into pinned memory.
        Nothing waits here, so the caller can do host work (RPN anchor targets) until `_sample_rois_fast` needs them."""
        rh = self.roi_heads
        device = padded.device
        gt_boxes = [t["boxes"].to(torch.float32) for t in targets]
        gt_labels = [t["labels"] for t in targets]
        Gs = [g.shape[0] for g in gt_boxes]
        gt_cat = torch.cat(gt_boxes, 0).contiguous()
        gl_cat = torch.cat(gt_labels, 0).contiguous()
        offs = [0]
        for g in Gs:
            offs.append(offs[-1] + g)
        gt_off = K.stager.put(torch.tensor(offs, dtype=torch.int32), device)
        m = rh.proposal_matcher
        if m.high_threshold != m.low_threshold or m.allow_low_quality_matches:
            raise NotImplementedError("RoI matcher with a between-threshold band")
        all_boxes, labels, matched, counts2 = K.roi_match(padded, count, gt_cat, gl_cat, gt_off, max(Gs), m.high_threshold)
        B = padded.shape[0]
        pin = getattr(self, "_count_pin", None)
        if pin is None or pin.shape[0] != B:
            pin = self._count_pin = torch.empty((B, 2), dtype=torch.int32).pin_memory()
        pin.copy_(counts2, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return dict(all_boxes=all_boxes, labels=labels, matched=matched, gt_cat=gt_cat, gt_off=gt_off, pin=pin, ev=ev)

    def _sample_rois_fast(self, mt):
        """Phase 2: BalancedPositiveNegativeSampler + box_coder.encode.  ONE host synchronisation (the counts the
        sampler's `torch.randperm(n)` calls need -- same calls, same order as the reference), then one selection kernel
        for all images and one gather / encode kernel.
        -> dict(rois5 [R,5], labels [R], matched [R], reg [R,4], sizes [per image], pos [positions of positives])."""
        rh = self.roi_heads
        labels, all_boxes, matched = mt["labels"], mt["all_boxes"], mt["matched"]
        device = labels.device
        mt["ev"].synchronize()                                   # the one host sync on the main stream
        cnt = mt["pin"].tolist()
        sampler = rh.fg_bg_sampler
        S = sampler.batch_size_per_image
        Pmax = int(S * sampler.positive_fraction)
        perms, num_pos_list, num_neg_list = [], [], []
        for npos, nneg in cnt:
            num_pos = min(npos, Pmax)
            num_neg = min(nneg, S - num_pos)
            perms.append((torch.randperm(npos, device=device), torch.randperm(nneg, device=device)))
            num_pos_list.append(num_pos)
            num_neg_list.append(num_neg)
        inds, pos_in = K.roi_sample(labels, perms, num_pos_list, num_neg_list, S, Pmax)
        sizes = [a + b for a, b in zip(num_pos_list, num_neg_list)]
        wts = rh.box_coder.weights
        if all(sz == S for sz in sizes):
            rois5, lab_c, m_c, reg_c = K.roi_encode(all_boxes, labels, matched, mt["gt_cat"], mt["gt_off"], inds, wts)
            pos = torch.cat([pos_in[i, :n] + i * S for i, n in enumerate(num_pos_list)], 0)
        else:
            parts, pos_parts, off = [], [], 0
            for i, sz in enumerate(sizes):
                r5, lc, mc, rc = K.roi_encode(all_boxes[i:i + 1], labels[i:i + 1], matched[i:i + 1], mt["gt_cat"],
                                              mt["gt_off"][i:i + 2], inds[i:i + 1, :sz].contiguous(), wts)
                r5[:, 0] = i
                parts.append((r5, lc, mc, rc))
                pos_parts.append(pos_in[i, :num_pos_list[i]] + off)
                off += sz
            rois5, lab_c, m_c, reg_c = [torch.cat(x, 0) for x in zip(*parts)]
            pos = torch.cat(pos_parts, 0)
        return dict(rois5=rois5, labels=lab_c, matched=m_c, reg=reg_c, sizes=sizes, pos=pos)

    def _roi_heads_train_fast(self, feats, mt, targets):
        rh = self.roi_heads
        smp = self._sample_rois_fast(mt)
        rois5, lab_c, reg_c = smp["rois5"], smp["labels"], smp["reg"]
        S = rh.fg_bg_sampler.batch_size_per_image
        graphed_box = (self.use_cuda_graphs and torch.is_grad_enabled() and os.environ.get("EOSVOS_GRAPH_BOX", "1") != "0"
                       and all(sz == S for sz in smp["sizes"]))
        if graphed_box:
            loss_classifier, loss_box_reg = self._box_loss_graphed(feats, rois5, lab_c, reg_c)
        else:
            o = self._box_branch(feats[:4], rois5)
            nc = rh.box_predictor.cls_score.weight.shape[0]
            loss_classifier, loss_box_reg = self._fastrcnn_loss_static(o[:, :nc], o[:, nc:nc + 4 * nc], lab_c, reg_c)
        losses = dict(loss_classifier=loss_classifier, loss_box_reg=loss_box_reg)
        pos = smp["pos"]
        n_mask = int(pos.shape[0])
        self.last_num_positives = n_mask
        kind = rh.maskrcnn_loss
        if kind not in ('BCE', 'LOVASZ'):
            raise NotImplementedError
        if n_mask == 0:
            losses["loss_mask"] = torch.zeros((), device=feats[0].device)
            return losses
        mask_rois5 = rois5[pos].contiguous()
        # mask targets: row (matched ground truth + offset of the image's masks in the concatenated list, box)
        g_off = [0]
        for t in targets:
            g_off.append(g_off[-1] + t["masks"].shape[0])
        img_off = K.stager.put(torch.tensor(g_off[:-1], dtype=torch.float32), feats[0].device)
        gt_row = smp["matched"][pos].to(torch.float32) + img_off[mask_rois5[:, 0].to(torch.int64)]
        tgt_rois = torch.cat([gt_row[:, None], mask_rois5[:, 1:]], dim=1).contiguous()
        gt_masks = torch.cat([t["masks"] for t in targets], 0).contiguous()
        if (kind == 'LOVASZ' and self.use_cuda_graphs and torch.is_grad_enabled()
                and os.environ.get("EOSVOS_GRAPH_MASK", "1") != "0"):
            lm = self._mask_loss_graphed_fast(feats, mask_rois5, lab_c[pos], tgt_rois, gt_masks)
            if lm is not None:
                losses["loss_mask"] = lm
                return losses
        mask_logits = self._mask_branch_rois(feats, mask_rois5)
        tg = K.mask_targets(gt_masks, tgt_rois, mask_logits.shape[-1])
        losses["loss_mask"] = ops.mask_loss(mask_logits, lab_c[pos].contiguous(), tg, kind)
        return losses

    _MASK_BUCKETS_FINE = (16, 32, 48, 64, 96, 128, 192, 256, 384)

    def _mask_loss_graphed_fast(self, feats, mask_rois5, lab, tgt_rois, gt_masks):
        """Mask branch (RoIAlign 28x28, 4 convs, deconv, logits), mask targets and the Lovasz loss -- forward and
        backward -- as one CUDA-graph pair on the positives padded to the next bucket size (padding rows: image index
        -1 = zeros through RoIAlign and no contribution backward, loss weight 0).  After the sampler's host sync the
        GPU is idle until the heads are queued: ~60 launches become two replays."""
        rh = self.roi_heads
        device = feats[0].device
        n = int(mask_rois5.shape[0])
        R = next((b for b in self._MASK_BUCKETS_FINE if b >= n), None)
        if R is None:
            return None
        pad = R - n
        fills = getattr(self, "_mask_fill_cache", None)
        if fills is None or fills[0] != device:
            fill = torch.zeros((max(self._MASK_BUCKETS_FINE), 5), device=device)
            fill[:, 0] = -1.0                       # negative image index = padding row (skipped by the kernels)
            fills = self._mask_fill_cache = (device, fill, torch.ones(max(self._MASK_BUCKETS_FINE), dtype=torch.int64, device=device))
        w = torch.zeros((R,), device=device, dtype=torch.float32)
        w[:n] = 1.0 / n
        if pad:
            mask_rois5 = torch.cat([mask_rois5, fills[1][:pad]], 0)
            tgt_rois = torch.cat([tgt_rois, fills[1][:pad]], 0)
            lab = torch.cat([lab, fills[2][:pad]], 0)
        if self._mask_slots is None:
            mods = (rh.mask_head, rh.mask_predictor)
            self._mask_slots = [(m, n_) for mod in mods for _, m in mod.named_modules()
                                for n_, p in m._parameters.items() if p is not None and p.requires_grad]

        def kinds(m, n_, t):
            if n_ != "weight":
                return ()
            if isinstance(m, nn.ConvTranspose2d):
                return ("dc",)
            if isinstance(m, nn.Conv2d) and t.shape[0] >= 64:
                return ("f", "t")
            return ()
        key = ("mask", R, tuple(gt_masks.shape), tuple(tuple(f.shape) for f in feats[:4]), device.index)
        return self._graphed_call(key, self._mask_train_functional,
                                  list(feats[:4]) + [mask_rois5.contiguous(), lab.contiguous(), gt_masks, tgt_rois.contiguous(), w],
                                  self._mask_slots, kinds, True, alias_inputs=4)

    def _box_loss_graphed(self, feats, rois5, lab_c, reg_c):
        """Box branch + tv fastrcnn_loss and their backward as one CUDA-graph pair (static shapes: 512 RoIs / image)."""
        rh = self.roi_heads
        if self._box_slots is None:
            mods = (rh.box_head, rh.box_predictor)
            self._box_slots = [(m, n) for mod in mods for _, m in mod.named_modules()
                               for n, p in m._parameters.items() if p is not None and p.requires_grad]
        C = feats[0].shape[-1]
        fc6 = rh.box_head.fc6

        def kinds(m, n, t):
            if n != "weight" or not isinstance(m, nn.Linear) or t.shape[0] < 64:
                return ()
            inner = C if m is fc6 else 0
            return (("lf", inner), ("lt", inner))
        key = ("box", tuple(rois5.shape), tuple(tuple(f.shape) for f in feats[:4]), feats[0].device.index)
        return self._graphed_call(key, self._box_train_functional, list(feats[:4]) + [rois5, lab_c, reg_c],
                                  self._box_slots, kinds, True, alias_inputs=4)

    def _box_eval_functional(self, f0, f1, f2, f3, rois5, *theta):
        slots = self._box_slots_eval
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            return self._box_branch([f0, f1, f2, f3], rois5)
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    def _mask_eval_functional(self, f0, f1, f2, f3, rois5, *theta):
        slots = self._mask_slots_eval
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            return self._mask_branch_rois([f0, f1, f2, f3], rois5)
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    def _heads_eval_fast(self, feats, padded, image_sizes, in_hw):
        """Box branch on the padded proposals (CUDA graph) -> best detection per image (det_top1_kernel:
        postprocess_detections with detections_per_img == 1) -> mask branch on that one RoI (CUDA graph).  Nothing
        here waits for the device: a frame without a detection carries a padding RoI (image index -1) and channel -1
        through the mask branch and the paste kernel, which then writes zeros."""
        rh = self.roi_heads
        device = feats[0].device
        B, R = padded.shape[0], padded.shape[1]
        rois5 = torch.cat([self._image_index(B, R, device), padded], dim=2).view(B * R, 5)
        C = feats[0].shape[-1]
        fc6 = rh.box_head.fc6
        if getattr(self, "_box_slots_eval", None) is None:
            mods = (rh.box_head, rh.box_predictor)
            self._box_slots_eval = [(m, n) for mod in mods for _, m in mod.named_modules()
                                    for n, p in m._parameters.items() if p is not None]
            mods = (rh.mask_head, rh.mask_predictor)
            self._mask_slots_eval = [(m, n) for mod in mods for _, m in mod.named_modules()
                                     for n, p in m._parameters.items() if p is not None]

        def box_kinds(m, n, t):
            if n != "weight" or not isinstance(m, nn.Linear) or t.shape[0] < 64:
                return ()
            return (("lf", C if m is fc6 else 0),)

        def mask_kinds(m, n, t):
            if n != "weight":
                return ()
            if isinstance(m, nn.ConvTranspose2d):
                return ("dc",)
            if isinstance(m, nn.Conv2d) and t.shape[0] >= 64:
                return ("f",)
            return ()
        fkey = (tuple(tuple(f.shape) for f in feats[:4]), feats[0].data_ptr(), device.index)
        head = self._graphed_call(("box_eval", B * R) + fkey, self._box_eval_functional, list(feats[:4]) + [rois5],
                                  self._box_slots_eval, box_kinds, False, alias_inputs=4)
        oh, ow = image_sizes[0]
        h, w = in_hw
        back_h = float(torch.tensor(h, dtype=torch.float32) / torch.tensor(oh, dtype=torch.float32))
        back_w = float(torch.tensor(w, dtype=torch.float32) / torch.tensor(ow, dtype=torch.float32))
        nc = rh.box_predictor.cls_score.weight.shape[0]
        det = K.det_top1(head, padded.view(B * R, 4), B, R, nc, rh.box_coder.weights, rh.box_coder.bbox_xform_clip,
                         rh.score_thresh, 1e-2, float(ow), float(oh), back_w, back_h)
        mask_logits = self._graphed_call(("mask_eval", B) + fkey, self._mask_eval_functional, list(feats[:4]) + [det["roi"]],
                                         self._mask_slots_eval, mask_kinds, False, alias_inputs=4)
        return det, mask_logits

    def _frame_functional(self, img, stats, fallback, rnd, *theta):
        """A whole inference frame as a pure function of tensors with static shapes -- what the frame graph captures:
        transform, trunk, RPN head, proposal kernels, EXTEND / REPLACE boxes, box branch, arg-max detection, mask
        branch on the one RoI, paste / threshold / next-target tail.  Configuration (mode, thresholds, sizes) is read
        from self._frame_cfg and is part of the graph key."""
        slots = self._frame_slots
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        try:
            cfg = self._frame_cfg
            rpn, rh, tr = self.rpn, self.roi_heads, self.transform
            B, _, h, w = img.shape
            oh, ow, Hp, Wp = cfg["oh"], cfg["ow"], cfg["Hp"], cfg["Wp"]
            x8 = K.transform(img, oh, ow, Hp, Wp, tr.image_mean, tr.image_std, Cs=8)
            feats = self._backbone_eager(x8)
            head_outs = self._rpn_head(feats)
            image_sizes, image_shape = [(oh, ow)] * B, (B, 3, Hp, Wp)
            post, mode, Kc = cfg["post"], cfg["mode"], self.num_classes - 1
            if cfg["has_target"] and mode is not None:
                n_aug = post // 2 if mode == 'EXTEND' else post
                n_front = post // 2 if mode == 'EXTEND' else 0
                padded = torch.empty((B, n_front + n_aug * Kc, 4), device=img.device, dtype=torch.float32)
                count = torch.zeros((B,), device=img.device, dtype=torch.int32)
                if n_front:
                    _, count = self._rpn_fast(feats, image_shape, image_sizes, head_outs, n_front, padded, 0)
                rw = float(torch.tensor(ow, dtype=torch.float32) / torch.tensor(w, dtype=torch.float32))
                rh_ = float(torch.tensor(oh, dtype=torch.float32) / torch.tensor(h, dtype=torch.float32))
                K.extend_boxes(stats, fallback, rnd, n_aug, rw, rh_, float(Wp), float(Hp), 0.1, padded, n_front)
            else:
                padded, count = self._rpn_fast(feats, image_shape, image_sizes, head_outs, post)
            R = padded.shape[1]
            rois5 = torch.cat([self._image_index(B, R, img.device), padded], dim=2).view(B * R, 5)
            head = self._box_branch(feats[:4], rois5)
            back_h = float(torch.tensor(h, dtype=torch.float32) / torch.tensor(oh, dtype=torch.float32))
            back_w = float(torch.tensor(w, dtype=torch.float32) / torch.tensor(ow, dtype=torch.float32))
            nc = rh.box_predictor.cls_score.weight.shape[0]
            det = K.det_top1(head, padded.view(B * R, 4), B, R, nc, rh.box_coder.weights, rh.box_coder.bbox_xform_clip,
                             cfg["score_thresh"], 1e-2, float(ow), float(oh), back_w, back_h)
            mask_logits = self._mask_branch_rois(feats[:4], det["roi"])
            probs, tgt, stats_out = K.mask_paste_threshold(mask_logits, det["chan"], det["label"], det["box"], B, Kc, h, w,
                                                           0.5, want_target=True)
            return probs, det["box"], tgt, stats_out, padded, count, det["row"]
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    # ---- look-ahead over several frames: everything of a frame that does not depend on the previous frame's result
    #      (transform, trunk, RPN head, proposal selection + NMS) runs for a whole run of frames in ONE batched graph;
    #      the per-frame remainder (jittered boxes from the propagated target, box branch, detection, mask branch,
    #      paste / threshold) replays per frame on that graph's outputs (reference loop: helper_func.py:100-126).
    def _frame_geometry(self, h, w):
        oh, ow = self._resized_size(h, w)
        div = int(self.transform.size_divisible)
        return oh, ow, (oh + div - 1) // div * div, (ow + div - 1) // div * div

    def _frame_kinds(self):
        conv1, fc6, C = self.backbone.body.conv1, self.roi_heads.box_head.fc6, self.backbone.out_channels

        def kinds(m, n, t):
            if n != "weight":
                return ()
            if m is conv1:
                return ("stem",)
            if isinstance(m, nn.ConvTranspose2d):
                return ("dc",)
            if isinstance(m, nn.Conv2d) and t.shape[0] >= 64:
                return ("f",)
            if isinstance(m, nn.Linear) and t.shape[0] >= 64:
                return (("lf", C if m is fc6 else 0),)
            return ()
        return kinds

    def _with_theta(self, slots, theta):
        saved = [m._parameters[n] for m, n in slots]
        for (m, n), t in zip(slots, theta):
            m._parameters[n] = t
        if self._active_plan is not None:
            self._active_plan.launch()
            for k in [k for k in ops._scope if isinstance(k[1], tuple) and k[1][0] == "head"]:
                del ops._scope[k]
        return saved

    def _frames_pre_functional(self, imgs, *theta):
        """Target-independent part of F frames at once: features of the four RoI levels, and the padded proposal rows
        [F, R, 4] with the first n_front rows (post-NMS RPN proposals) filled in."""
        slots = self._pre_slots
        saved = self._with_theta(slots, theta)
        try:
            cfg = self._frame_cfg
            tr = self.transform
            F_ = imgs.shape[0]
            oh, ow, Hp, Wp = cfg["oh"], cfg["ow"], cfg["Hp"], cfg["Wp"]
            x8 = K.transform(imgs, oh, ow, Hp, Wp, tr.image_mean, tr.image_std, Cs=8)
            feats = self._backbone_eager(x8)
            head_outs = self._rpn_head(feats)
            padded = torch.zeros((F_, cfg["R"], 4), device=imgs.device, dtype=torch.float32)
            count = torch.zeros((F_,), device=imgs.device, dtype=torch.int32)
            if cfg["n_front"]:
                _, count = self._rpn_fast(feats, (F_, 3, Hp, Wp), [(oh, ow)] * F_, head_outs, cfg["n_front"], padded, 0)
            return feats[0], feats[1], feats[2], feats[3], padded, count
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    def _frame_tail_functional(self, f0, f1, f2, f3, padded, stats, fallback, rnd, *theta):
        """Per-frame remainder on one frame's slice of the batched features / proposal rows (all static shapes)."""
        slots = self._tail_slots
        saved = self._with_theta(slots, theta)
        try:
            cfg = self._frame_cfg
            rh = self.roi_heads
            h, w, oh, ow, Hp, Wp = cfg["h"], cfg["w"], cfg["oh"], cfg["ow"], cfg["Hp"], cfg["Wp"]
            Kc = self.num_classes - 1
            if cfg["has_target"]:
                rw = float(torch.tensor(ow, dtype=torch.float32) / torch.tensor(w, dtype=torch.float32))
                rh_ = float(torch.tensor(oh, dtype=torch.float32) / torch.tensor(h, dtype=torch.float32))
                K.extend_boxes(stats, fallback, rnd, cfg["n_aug"], rw, rh_, float(Wp), float(Hp), 0.1, padded,
                               cfg["n_front"])
            R = padded.shape[1]
            feats = [f0, f1, f2, f3]
            rois5 = torch.cat([self._image_index(1, R, f0.device), padded], dim=2).view(R, 5)
            head = self._box_branch(feats, rois5)
            back_h = float(torch.tensor(h, dtype=torch.float32) / torch.tensor(oh, dtype=torch.float32))
            back_w = float(torch.tensor(w, dtype=torch.float32) / torch.tensor(ow, dtype=torch.float32))
            nc = rh.box_predictor.cls_score.weight.shape[0]
            det = K.det_top1(head, padded.view(R, 4), 1, R, nc, rh.box_coder.weights, rh.box_coder.bbox_xform_clip,
                             cfg["score_thresh"], 1e-2, float(ow), float(oh), back_w, back_h)
            mask_logits = self._mask_branch_rois(feats, det["roi"])
            probs, tgt, stats_out = K.mask_paste_threshold(mask_logits, det["chan"], det["label"], det["box"], 1, Kc, h, w,
                                                           0.5, want_target=True)
            return probs, det["box"], tgt, stats_out, det["row"]
        finally:
            for (m, n), t in zip(slots, saved):
                m._parameters[n] = t

    def _lookahead_cfg(self, h, w, has_target):
        rpn, rh = self.rpn, self.roi_heads
        oh, ow, Hp, Wp = self._frame_geometry(h, w)
        mode = rpn._eval_augment_proposals_mode
        post = rpn.post_nms_top_n()
        Kc = self.num_classes - 1
        if has_target:
            n_aug = post // 2 if mode == 'EXTEND' else post
            n_front = post // 2 if mode == 'EXTEND' else 0
        else:
            n_aug, n_front = 0, post
        return dict(h=h, w=w, oh=oh, ow=ow, Hp=Hp, Wp=Wp, mode=mode, has_target=has_target, post=post, n_aug=n_aug,
                    n_front=n_front, R=n_front + n_aug * Kc, score_thresh=float(rh.score_thresh),
                    pre=rpn.pre_nms_top_n(), nms=float(rpn.nms_thresh))

    def lookahead_ok(self):
        return (not self.training and self.use_cuda_graphs and self.capture is None and self._fast_ok()
                and self.num_classes == 2 and self.roi_heads.detections_per_img == 1
                and self.fixed_proposals is None and os.environ.get("EOSVOS_FRAME_GRAPH", "1") != "0"
                and int(os.environ.get("EOSVOS_FRAME_BATCH", "8")) > 1)

    def prefetch_frames(self, frames, has_target):
        """frames: the next F >= 2 inputs ([1,3,h,w] device tensors) that `forward` will be called with, in order, while
        the parameters stay as they are; has_target: whether those calls carry a target with proposal augmentation.
        Runs the target-independent part of all F frames as one batched graph; the forward calls then only replay the
        per-frame remainder.  A forward call with any other input simply does not use the look-ahead."""
        self._lookahead = None
        if len(frames) < 2 or not self.lookahead_ok():
            return False
        mode = self.rpn._eval_augment_proposals_mode
        if has_target and mode not in ('EXTEND', 'REPLACE'):
            return False
        device = frames[0].device
        shape = tuple(frames[0].shape)
        if any(tuple(f.shape) != shape or f.shape[0] != 1 for f in frames):
            return False
        _, _, h, w = shape
        cfg = self._lookahead_cfg(h, w, bool(has_target))
        if getattr(self, "_pre_slots", None) is None:
            rpn, rh = self.rpn, self.roi_heads
            self._pre_slots = [(m, n) for mod in (self.backbone, rpn.head) for _, m in mod.named_modules()
                               for n, p in m._parameters.items() if p is not None]
            self._tail_slots = [(m, n) for mod in (rh.box_head, rh.box_predictor, rh.mask_head, rh.mask_predictor)
                                for _, m in mod.named_modules() for n, p in m._parameters.items() if p is not None]
        # a shorter run (the tail of a sequence) is padded with copies of its last frame up to the nominal run length:
        # one batched graph and one set of per-frame graphs per model instead of one set per run length (a capture
        # costs far more than the few wasted trunk passes)
        nominal = max(int(os.environ.get("EOSVOS_FRAME_BATCH", "5")), len(frames))
        F_ = nominal
        imgs = torch.cat([f.to(torch.float32) for f in frames] + [frames[-1].to(torch.float32)] * (nominal - len(frames)))
        key = ("frames_pre", F_, h, w, cfg["mode"], cfg["has_target"], cfg["post"], cfg["pre"], cfg["nms"], device.index)
        self._frame_cfg = cfg
        K.zero_pool.reset()
        outs = self._graphed_call(key, self._frames_pre_functional, [imgs], self._pre_slots, self._frame_kinds(), False)
        self._lookahead = dict(cfg=cfg, outs=outs, frames=list(frames), index={id(f): i for i, f in enumerate(frames)},
                               versions=[f._version for f in frames], ptrs=[f.data_ptr() for f in frames])
        return True

    def _lookahead_slot(self, inputs, has_target):
        la = getattr(self, "_lookahead", None)
        if la is None:
            return None
        i = la["index"].get(id(inputs))
        if (i is None or la["frames"][i] is not inputs or la["versions"][i] != inputs._version
                or la["ptrs"][i] != inputs.data_ptr()):
            return None
        cfg = la["cfg"]
        now = self._lookahead_cfg(cfg["h"], cfg["w"], has_target)
        if now != cfg:          # mode / thresholds / proposal counts changed since the look-ahead ran
            return None
        return i

    def _forward_eval_tail(self, i, inputs, dev_stats):
        la = self._lookahead
        cfg = la["cfg"]
        device = inputs.device
        f0, f1, f2, f3, padded, count = la["outs"]
        Kc = self.num_classes - 1
        if cfg["has_target"]:
            stats, fallback = dev_stats
            rnd = self._extend_rands(1, Kc, cfg["n_aug"], device)
            fallback = stats if fallback is None else fallback
        else:
            dummy = getattr(self, "_frame_dummy", None)
            if dummy is None or dummy[0].device != device:
                dummy = self._frame_dummy = (torch.zeros((1, Kc, 5), dtype=torch.int32, device=device),
                                             torch.zeros((1,), device=device))
            stats = fallback = dummy[0]
            rnd = dummy[1]
        sl = [t[i:i + 1] for t in (f0, f1, f2, f3, padded)]
        key = ("frame_tail", i, sl[0].data_ptr(), sl[4].data_ptr(), cfg["h"], cfg["w"], cfg["mode"], cfg["has_target"],
               cfg["post"], cfg["score_thresh"], device.index)
        self._frame_cfg = cfg
        K.zero_pool.reset()
        probs, box, tgt, stats_out, row = self._graphed_call(key, self._frame_tail_functional, sl + [stats, fallback, rnd],
                                                             self._tail_slots, self._frame_kinds(), False, alias_inputs=5)
        self._last_padded = (sl[4], count[i:i + 1] if cfg["n_front"] else None, cfg["n_front"])
        self._last_det = {"row": row}
        self.last_propagated_target = tgt.clone()
        self.last_target_stats = stats_out.clone()
        return probs.clone(), box.clone().view(1, Kc, 4)

    def _image_index(self, B, R, device):
        cache = self.__dict__.setdefault("_img_idx_cache", {})     # per signature, never freed (graphs bake addresses)
        img_idx = cache.get((B, R, str(device)))
        if img_idx is None:
            img_idx = torch.arange(B, device=device, dtype=torch.float32).view(B, 1, 1).expand(B, R, 1).contiguous()
            cache[(B, R, str(device))] = img_idx
        return img_idx

    def _forward_eval_frame_graph(self, inputs, targets, dev_stats):
        """Inference frame as ONE CUDA-graph replay (everything is shape-static at inference: see _frame_functional).
        Host work per frame: the reference's CPU random numbers for the jittered boxes (uploaded), three small input
        copies, one replay."""
        device = inputs.device
        B, _, h, w = inputs.shape
        rpn, rh = self.rpn, self.roi_heads
        oh, ow = self._resized_size(h, w)
        div = int(self.transform.size_divisible)
        Hp, Wp = (oh + div - 1) // div * div, (ow + div - 1) // div * div
        mode = rpn._eval_augment_proposals_mode
        has_target = targets is not None and mode is not None
        post = rpn.post_nms_top_n()
        Kc = self.num_classes - 1
        cfg = dict(oh=oh, ow=ow, Hp=Hp, Wp=Wp, mode=mode, has_target=has_target, post=post,
                   score_thresh=float(rh.score_thresh))
        if getattr(self, "_frame_slots", None) is None:
            mods = (self.backbone, rpn.head, rh.box_head, rh.box_predictor, rh.mask_head, rh.mask_predictor)
            self._frame_slots = [(m, n) for mod in mods for _, m in mod.named_modules()
                                 for n, p in m._parameters.items() if p is not None]
        conv1, fc6, C = self.backbone.body.conv1, rh.box_head.fc6, self.backbone.out_channels

        def kinds(m, n, t):
            if n != "weight":
                return ()
            if m is conv1:
                return ("stem",)
            if isinstance(m, nn.ConvTranspose2d):
                return ("dc",)
            if isinstance(m, nn.Conv2d) and t.shape[0] >= 64:
                return ("f",)
            if isinstance(m, nn.Linear) and t.shape[0] >= 64:
                return (("lf", C if m is fc6 else 0),)
            return ()
        if has_target:
            stats, fallback = dev_stats
            n_aug = post // 2 if mode == 'EXTEND' else post
            rnd = self._extend_rands(B, Kc, n_aug, device)
            fallback = stats if fallback is None else fallback
        else:
            dummy = getattr(self, "_frame_dummy", None)
            if dummy is None or dummy[0].device != device:
                dummy = self._frame_dummy = (torch.zeros((B, Kc, 5), dtype=torch.int32, device=device),
                                             torch.zeros((1,), device=device))
            stats = fallback = dummy[0]
            rnd = dummy[1]
        key = ("frame", (B, h, w), mode, has_target, post, rpn.pre_nms_top_n(), float(rh.score_thresh), float(rpn.nms_thresh),
               device.index)
        self._frame_cfg = cfg
        K.zero_pool.reset()
        probs, box, tgt, stats_out, padded, count, row = self._graphed_call(
            key, self._frame_functional, [inputs.to(torch.float32).contiguous(), stats, fallback, rnd], self._frame_slots,
            kinds, False)
        n_front = (post // 2 if mode == 'EXTEND' else 0) if has_target else post
        self._last_padded = (padded, count if n_front else None, n_front)
        self._last_det = {"row": row}
        # the graph's output buffers are rewritten by the next frame: hand out copies of what the caller keeps
        self.last_propagated_target = tgt.clone()
        self.last_target_stats = stats_out.clone()
        return probs.clone(), box.clone().view(B, Kc, 4)

    def _forward_eval_fast(self, inputs, targets, dev_stats):
        """Inference frame without a host synchronisation (helper_func.py:100-126 body): transform, trunk graph,
        proposal kernels, EXTEND / REPLACE boxes from the device-resident target box, box graph, arg-max detection,
        mask graph, fused paste / threshold / next-target tail."""
        device = inputs.device
        B, _, h, w = inputs.shape
        K.zero_pool.reset()
        pre = None
        if self._prefetched:
            pf = self._prefetched.pop(id(inputs), None)
            if pf is not None and pf[0] is inputs and pf[5] == inputs._version and pf[6] == inputs.data_ptr():
                pre = (pf[1], pf[2], pf[3])
                pre_feats = pf[4]
        x8, _, (oh, ow), (Hp, Wp) = self._transform(inputs, None, pre)
        image_sizes = [(oh, ow)] * B
        image_shape = (B, 3, Hp, Wp)
        with torch.no_grad():
            feats = pre_feats if pre is not None else self._backbone(x8)
            feats, head_outs = feats[:5], feats[-5:]
            rpn = self.rpn
            post = rpn.post_nms_top_n()
            mode = rpn._eval_augment_proposals_mode
            Kc = self.num_classes - 1
            if targets is not None and mode is not None:
                stats, fallback = dev_stats
                if mode not in ('EXTEND', 'REPLACE'):
                    raise NotImplementedError
                n_aug = post // 2 if mode == 'EXTEND' else post
                n_front = post // 2 if mode == 'EXTEND' else 0
                padded = torch.empty((B, n_front + n_aug * Kc, 4), device=device, dtype=torch.float32)
                count = None
                if n_front:
                    _, count = self._rpn_fast(feats, image_shape, image_sizes, head_outs, n_front, padded, 0)
                rnd = self._extend_rands(B, Kc, n_aug, device)
                rw = float(torch.tensor(ow, dtype=torch.float32) / torch.tensor(w, dtype=torch.float32))
                rh_ = float(torch.tensor(oh, dtype=torch.float32) / torch.tensor(h, dtype=torch.float32))
                K.extend_boxes(stats, fallback, rnd, n_aug, rw, rh_, float(Wp), float(Hp), 0.1, padded, n_front)
            else:
                n_front = post
                padded, count = self._rpn_fast(feats, image_shape, image_sizes, head_outs, post)
            self._last_padded = (padded, count, n_front)
            det, mask_logits = self._heads_eval_fast(feats, padded, image_sizes, (h, w))
        self._last_det = det
        probs, tgt, stats_out = K.mask_paste_threshold(mask_logits, det["chan"], det["label"], det["box"], B, Kc, h, w, 0.5,
                                                       want_target=True)
        self.last_propagated_target = tgt
        self.last_target_stats = stats_out
        return probs, det["box"].view(B, Kc, 4)

    @property
    def last_proposals(self):
        """Proposals that entered the RoI heads in the last forward, as the reference's per-image lists (observation
        point for checkers; converting the padded buffer synchronises)."""
        lp = getattr(self, "_last_padded", None)
        if lp is None:
            return getattr(self, "_last_proposal_list", None)
        padded, count, n_front = lp
        if count is None:
            return [p for p in padded]
        return self._proposal_list(padded, count, n_front)

    @last_proposals.setter
    def last_proposals(self, value):
        self._last_padded = None
        self._last_proposal_list = value

    @property
    def last_detection_rows(self):
        """Candidate row (proposal index * (num_classes - 1) + class - 1) of every kept detection, per image."""
        det = getattr(self, "_last_det", None)
        if det is None:
            return getattr(self, "_last_detection_rows", None)
        rows = det["row"].cpu()
        lp = self._last_padded
        out = []
        if lp is not None and lp[1] is not None:
            cnt = lp[1].tolist()
        else:
            cnt = None
        for b, r in enumerate(rows.tolist()):
            if r < 0:
                out.append(torch.zeros((0,), dtype=torch.int64))
                continue
            if cnt is not None and r >= lp[2]:       # rows of the EXTEND half: shift by the unused RPN slots
                r = r - (lp[2] - min(cnt[b], lp[2]))
            out.append(torch.tensor([r], dtype=torch.int64))
        return out

    @last_detection_rows.setter
    def last_detection_rows(self, value):
        self._last_det = None
        self._last_detection_rows = value

    def _roi_heads(self, feats, proposals, image_sizes, targets):
        rh = self.roi_heads
        if self.training:
            proposals, matched_idxs, labels, regression_targets, pos_in = self._select_training_samples(proposals,
                                                                                                        targets)
        # (only the regular case -- every image filled its 512-RoI sample -- is worth a graph: odd sizes, e.g. when
        # the RPN produced no proposal, would each cost a capture)
        graphed_box = (self.training and self.use_cuda_graphs and self.capture is None and torch.is_grad_enabled()
                       and os.environ.get("EOSVOS_GRAPH_BOX", "1") != "0"
                       and all(p.shape[0] == rh.fg_bg_sampler.batch_size_per_image for p in proposals))
        if graphed_box:
            # static shapes (512 sampled RoIs per image): the whole box branch, its loss and their backward replay as
            # a second pair of CUDA graphs; only the mask branch (n_pos RoIs) stays eager
            if self._box_slots is None:
                mods = (rh.box_head, rh.box_predictor)
                self._box_slots = [(m, n) for mod in mods for _, m in mod.named_modules()
                                   for n, p in m._parameters.items() if p is not None and p.requires_grad]
            rois5 = self._rois5(proposals)
            lab_c, reg_c = torch.cat(labels, dim=0), torch.cat(regression_targets, dim=0)
            C = feats[0].shape[-1]
            fc6 = rh.box_head.fc6

            def kinds(m, n, t):
                if n != "weight" or not isinstance(m, nn.Linear) or t.shape[0] < 64:
                    return ()
                inner = C if m is fc6 else 0
                return (("lf", inner), ("lt", inner))
            key = ("box", tuple(rois5.shape), tuple(tuple(f.shape) for f in feats[:4]), feats[0].device.index)
            loss_classifier, loss_box_reg = self._graphed_call(
                key, self._box_train_functional, list(feats[:4]) + [rois5, lab_c, reg_c], self._box_slots, kinds, True,
                alias_inputs=4)
            class_logits = box_regression = None
        else:
            o = self._box_branch(feats[:4], self._rois5(proposals))
            nc = rh.box_predictor.cls_score.weight.shape[0]
            class_logits, box_regression = o[:, :nc], o[:, nc:nc + 4 * nc]
        if self.capture is not None:
            self.capture.update(class_logits=class_logits.detach(), box_regression=box_regression.detach(),
                                sampled_proposals=[p.detach() for p in proposals])

        result, losses = [], {}
        if self.training:
            if not graphed_box:
                loss_classifier, loss_box_reg = fastrcnn_loss(class_logits, box_regression, labels, regression_targets)
            losses = dict(loss_classifier=loss_classifier, loss_box_reg=loss_box_reg)
            mask_proposals, pos_matched_idxs = [], []
            for img_id in range(len(proposals)):
                pos = pos_in[img_id]
                mask_proposals.append(proposals[img_id][pos])
                pos_matched_idxs.append(matched_idxs[img_id][pos])
        else:
            boxes, scores, labels = self._postprocess_detections(class_logits, box_regression, proposals, image_sizes)
            for i in range(len(boxes)):
                result.append(dict(boxes=boxes[i], labels=labels[i], scores=scores[i]))
            if self.capture is not None:
                self.capture.update(detections=[{k: v.detach().clone() for k, v in r.items()} for r in result])
            if self.fixed_detections is not None:      # test hook: mask branch on given detections
                result = [{k: v.to(feats[0].device).clone() for k, v in d.items()} for d in self.fixed_detections]
            mask_proposals = [p["boxes"] for p in result]

        n_mask = sum(p.shape[0] for p in mask_proposals)
        loss_mask_graphed = None
        if (graphed_box and os.environ.get("EOSVOS_GRAPH_MASK", "0") == "1"
                and all(t["masks"].shape[0] == targets[0]["masks"].shape[0] for t in targets)):
            # opt-in: the mask branch, its targets and the Lovasz loss on the positives padded to a bucket size as a
            # third graph pair.  Measured neutral on B200 (47 vs 47-49 iter/s): with the trunk and the box branch
            # graphed the iteration is GPU-bound, and the padded rows cost what the saved launches gain.
            loss_mask_graphed = self._mask_loss_graphed(feats, mask_proposals, pos_matched_idxs, targets)
        if loss_mask_graphed is not None:
            mask_logits = None
        elif n_mask > 0:
            mask_logits = self._mask_branch(feats, mask_proposals)
            if self.capture is not None:
                self.capture.update(mask_logits=mask_logits.detach())
        elif len(mask_proposals) > 1:
            raise NotImplementedError
        else:
            mask_logits = torch.zeros((0, self.num_classes, 2 * self._roi_sizes['mask'], 2 * self._roi_sizes['mask']),
                                      device=feats[0].device)

        if self.training:
            kind = rh.maskrcnn_loss
            if kind not in ('BCE', 'LOVASZ'):
                raise NotImplementedError
            if loss_mask_graphed is not None:
                loss_mask = loss_mask_graphed
            elif n_mask == 0:
                loss_mask = torch.zeros((), device=feats[0].device)
            else:
                gt_masks = [t["masks"] for t in targets]
                gt_labels = [t["labels"] for t in targets]
                lab = torch.cat([l[idxs] for l, idxs in zip(gt_labels, pos_matched_idxs)], dim=0)
                offs, rois = 0, []
                for m, p, i in zip(gt_masks, mask_proposals, pos_matched_idxs):
                    rois.append(torch.cat([(i + offs).to(p)[:, None], p], dim=1))
                    offs += m.shape[0]
                M = mask_logits.shape[-1]
                tg = K.mask_targets(torch.cat(gt_masks, 0).contiguous(),
                                    torch.cat(rois, 0).to(torch.float32).contiguous(), M)
                loss_mask = ops.mask_loss(mask_logits, lab.contiguous(), tg, kind)
            losses.update(dict(loss_mask=loss_mask))
        else:
            for r in result:
                r["mask_logits_all"] = mask_logits   # class select + sigmoid + paste are fused in the tail kernel
        return result, losses, mask_logits

    # ---- forward (reference mask_rcnn.py:572-775) -----------------------------------------------
    def forward(self, inputs, targets=None, box_coord_perm=None, flip_label=False):
        device = inputs.device
        if device.type != "cuda":
            raise RuntimeError("eosvos_b200.MaskRCNN runs on sm_100 CUDA devices only (no CPU fallback)")
        if self.training and targets is None:
            raise ValueError("targets should not be None in training mode")
        if targets is not None and targets.device != device:
            # the reference's callers hand over the first-frame label as a host tensor (evaluate.py:297-301): its
            # forward reads targets on the host anyway (mask_rcnn.py:593-632)
            targets = targets.to(device, non_blocking=True)
        fast = self._fast_ok()
        if not self.training and fast and self.use_cuda_graphs and self.num_classes == 2 \
                and self.roi_heads.detections_per_img == 1 and not flip_label:
            # sync-free inference frame: the target's box stays on the DEVICE (run_frames hands over the one the
            # previous frame's tail kernel produced)
            # (the target only contributes its bounding box at inference time: mask_rcnn.py:251-285)
            dev_stats = None
            if targets is not None and self.rpn._eval_augment_proposals_mode is not None:
                dev_stats = K.target_stats.get_device(targets)
                if dev_stats is None:
                    # box of the target on the device (mask_tail.cu::mask_to_bbox_kernel), never read back.  The
                    # reference asserts here that the target holds at least one object (mask_rcnn.py:623); without
                    # the read-back an empty target yields degenerate boxes and no detection instead of raising.
                    dev_stats = (K.mask_to_bbox(targets.to(torch.float32).contiguous(), self.num_classes - 1), None)
            if os.environ.get("EOSVOS_FRAME_GRAPH", "1") != "0":
                slot = self._lookahead_slot(inputs, dev_stats is not None)
                if slot is not None:
                    return self._forward_eval_tail(slot, inputs, dev_stats)
                return self._forward_eval_frame_graph(inputs, targets, dev_stats)
            return self._forward_eval_fast(inputs, targets, dev_stats)
        self._last_padded, self._last_det = None, None
        if targets is not None:
            targets = self._build_targets(targets, flip_label, device)
        B, _, h, w = inputs.shape
        K.zero_pool.reset()          # one zeroed block per forward(+backward) serves all accumulate-into outputs
        fast_train = (self.training and fast and self.use_cuda_graphs and os.environ.get("EOSVOS_GRAPH_HEAD", "2") == "2")
        self._prepare_operands(skip_box=fast_train and torch.is_grad_enabled()
                               and os.environ.get("EOSVOS_GRAPH_BOX", "1") != "0")
        pre = None
        if self._prefetched:
            pf = self._prefetched.pop(id(inputs), None)
            if (pf is not None and pf[0] is inputs and pf[5] == inputs._version and pf[6] == inputs.data_ptr()
                    and not self.training):
                pre = (pf[1], pf[2], pf[3])
                pre_feats = pf[4]
        x8, targets_t, (oh, ow), (Hp, Wp) = self._transform(inputs, targets, pre)

        image_sizes = [(oh, ow)] * B
        image_shape = (B, 3, Hp, Wp)

        grad_ctx = torch.enable_grad() if self.training else torch.no_grad()
        with grad_ctx:
            if fast_train:
                am = self._anchor_match_async(image_shape, image_sizes, targets_t, device)
            feats = pre_feats if pre is not None else self._backbone(x8)
            rpn_ts = feats[5:10] if len(feats) == 15 else None
            feats, head_outs = feats[:5], (feats[-5:] if len(feats) > 5 else None)
            if self.training and rpn_ts is not None and not fast_train:
                head_outs = None          # forward-only head outputs carry no gradient: recompute with autograd below
            if fast_train and head_outs is not None and head_outs[0].dtype == torch.float32:
                # statically shaped proposal / sampling pipeline: ONE host sync (RoI sampler counts) per iteration
                # the proposal kernels and the RoI matching only need the trunk: queue them first, label the anchors
                # on the host / side stream meanwhile (RNG order unchanged: RPN sampler, then RoI sampler)
                padded, count = self._rpn_fast(feats, image_shape, image_sizes, head_outs, self.rpn.post_nms_top_n())
                self._last_padded = (padded, count, padded.shape[1])
                mt = self._match_rois_fast(padded, count, targets_t)
                sampled = self._sample_anchors_fast(am)
                det_losses = self._roi_heads_train_fast(feats, mt, targets_t)
                raw = dict(det_losses)
                if rpn_ts is not None:
                    lo, lb = ops.rpn_loss_sparse(feats, rpn_ts, head_outs, self.rpn.head, sampled, am["labels"],
                                                 am["matched"], am["anchors"], am["gt_boxes"], am["gt_off"])
                else:
                    lo, lb = ops.rpn_loss(head_outs, [f.shape[1] * f.shape[2] for f in feats],
                                          self.rpn.head.cls_logits.weight.shape[0], sampled, am["labels"], am["matched"],
                                          am["anchors"], am["gt_boxes"], am["gt_off"])
                raw.update({"loss_objectness": lo, "loss_rpn_box_reg": lb})
                losses = {n: l for n, l in raw.items() if l.requires_grad}
                return sum([l for l in losses.values()]), losses
            proposals, rpn_losses = self._rpn(feats, image_shape, image_sizes, targets_t, head_outs)
            if self.fixed_proposals is not None:
                proposals = [p.to(device).clone() for p in self.fixed_proposals]
            self.last_proposals = proposals          # observation point for checkers (device tensors, no copy)
            self.last_num_positives = None
            if self.capture is not None:
                self.capture.update(feats=feats, proposals=[p.detach() for p in proposals], x8=x8)
            detections, det_losses, mask_logits = self._roi_heads(feats, proposals, image_sizes, targets_t)

        if self.training:
            raw = {}
            raw.update(det_losses)
            raw.update(rpn_losses)
            losses = {n: l for n, l in raw.items() if l.requires_grad}
            loss = sum([l for l in losses.values()])
            return loss, losses

        # eval: first detection of every class -> dense probability map + box (mask_rcnn.py:732-775),
        # boxes rescaled to the input frame (tv transform.py:257-278), paste/threshold fused on device
        Kc = self.num_classes - 1
        rh_ = float(torch.tensor(h, dtype=torch.float32) / torch.tensor(oh, dtype=torch.float32))
        rw_ = float(torch.tensor(w, dtype=torch.float32) / torch.tensor(ow, dtype=torch.float32))
        det_boxes, det_labels, chan, off = [], [], [], 0
        cls_ids = torch.arange(1, self.num_classes, device=device)
        for det in detections:
            b = det["boxes"]
            xmin, ymin, xmax, ymax = b.unbind(1)
            b = torch.stack((xmin * rw_, ymin * rh_, xmax * rw_, ymax * rh_), dim=1)
            lab = det["labels"]
            if lab.shape[0] > 0:
                eq = lab[None, :] == cls_ids[:, None]
                first = eq.to(torch.int32).argmax(dim=1)
                chan.append(torch.where(eq.any(dim=1), first + off, torch.full_like(first, -1)))
            else:
                chan.append(torch.full((Kc,), -1, dtype=torch.int64, device=device))
            det_boxes.append(b)
            det_labels.append(lab)
            off += lab.shape[0]
        det_boxes = torch.cat(det_boxes, 0).to(torch.float32).contiguous()
        det_labels = torch.cat(det_labels, 0).to(torch.int64).contiguous()
        chan = torch.cat(chan, 0).to(torch.int32).contiguous()
        probs, tgt, stats = K.mask_paste_threshold(mask_logits, chan, det_labels, det_boxes, B, Kc, h, w, 0.5,
                                                   want_target=True)
        safe = chan.clamp(min=0).to(torch.int64)
        if det_boxes.shape[0] > 0:
            out_boxes = torch.where((chan >= 0)[:, None], det_boxes[safe], torch.zeros_like(det_boxes[safe]))
        else:
            out_boxes = torch.zeros((B * Kc, 4), device=device)
        self.last_propagated_target = tgt        # threshold/argmax of helper_func.py:113-121, already on device
        self.last_target_stats = stats
        return probs, out_boxes.view(B, Kc, 4)
