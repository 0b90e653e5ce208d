"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the UNMODIFIED reference (dvl-tum/e-osvos, read-only at /root/reference) under the installed
torch 2.11 / torchvision 0.26 by monkey-patching only the API drift listed in SURVEY.md §8c.  Used in
the build container to pin oracle/ against the real reference and to generate tests/golden/*; the GPU
box has no /root/reference: there the unmodified copy under oracle/_ref/ (oracle/install_ref.py) is imported.
"""
import os
import sys
import types

import torch
import torchvision

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """$EOSVOS_REFERENCE, else the read-only checkout of the build container, else the unmodified copy that
    oracle/install_ref.py placed under oracle/_ref/ (what the GPU box has)."""
    cands = [os.environ.get("EOSVOS_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "src", "networks")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference()


class _CallableInt(int):
    """An int that can also be called: serves both `x` (torchvision 0.4 property style, what the
    reference uses) and `x()` (torchvision 0.26 method style)."""

    def __call__(self):
        return int(self)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "networks"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    # (1) torchvision.models.utils.load_state_dict_from_url was removed; COCO weights need network
    #     -> random init, which is what BASELINE.json's north_star specifies.
    utils_mod = types.ModuleType("torchvision.models.utils")
    utils_mod.load_state_dict_from_url = lambda *a, **k: {}
    sys.modules["torchvision.models.utils"] = utils_mod
    torchvision.models.utils = utils_mod

    # (4) properties that became methods
    from torchvision.models.detection.roi_heads import RoIHeads
    from torchvision.models.detection.rpn import RegionProposalNetwork

    def _has_mask(self):
        return _CallableInt(int(self.mask_roi_pool is not None and self.mask_head is not None
                                and self.mask_predictor is not None))

    def _has_keypoint(self):
        return _CallableInt(int(self.keypoint_roi_pool is not None and self.keypoint_head is not None
                                and self.keypoint_predictor is not None))

    RoIHeads.has_mask = property(_has_mask)
    RoIHeads.has_keypoint = property(_has_keypoint)
    RegionProposalNetwork.pre_nms_top_n = property(
        lambda self: _CallableInt(self._pre_nms_top_n["training" if self.training else "testing"]))
    RegionProposalNetwork.post_nms_top_n = property(
        lambda self: _CallableInt(self._post_nms_top_n["training" if self.training else "testing"]))

    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import networks.mask_rcnn as ref_mrcnn  # noqa: E402  (the reference's own module, unmodified)

    # (2) resnet_fpn_backbone(name, True) is keyword-only now and True would download
    from torchvision.models.detection.backbone_utils import resnet_fpn_backbone as tv_backbone
    ref_mrcnn.resnet_fpn_backbone = lambda name, pretrained: tv_backbone(
        backbone_name=name, weights=None, trainable_layers=5)

    # (3) featmap_names are strings now
    from torchvision.ops import MultiScaleRoIAlign as TvMSRA

    def _msra(featmap_names, output_size, sampling_ratio):
        return TvMSRA(featmap_names=[str(n) for n in featmap_names], output_size=output_size,
                      sampling_ratio=sampling_ratio)

    ref_mrcnn.MultiScaleRoIAlign = _msra
    _installed = True


def reference_modules():
    """-> (networks.mask_rcnn, networks.loss_lovasz, meta_optim.meta_optim, meta_optim.meta_model)."""
    install()
    import meta_optim.meta_model as mm
    import meta_optim.meta_optim as mo
    import networks.loss_lovasz as ll
    import networks.mask_rcnn as mr
    return mr, ll, mo, mm


def build_reference_model(seed=1, maskrcnn_loss="LOVASZ", min_size=None, max_size=None):
    """The reference's model exactly as cfgs/meta.yaml builds it (helper_func.py:349-355)."""
    mr, _, _, _ = reference_modules()
    torch.manual_seed(seed)
    model = mr.MaskRCNN("resnet50", num_classes=2,
                        batch_norm={"accum_stats": False, "learn_weight": False, "learn_bias": False},
                        train_encoder=True, roi_pool_output_sizes={"box": 7, "mask": 28},
                        eval_augment_rpn_proposals_mode="EXTEND", replace_batch_with_group_norms=True,
                        box_nms_thresh=0.5, maskrcnn_loss=maskrcnn_loss)
    if min_size is not None:
        model.transform.min_size = (min_size,)
        model.transform.max_size = max_size
    return model
