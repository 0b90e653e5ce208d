"""First-frame augmentation used while fine-tuning (host side, as in the reference):
RandomHorizontalFlip + RandomScaleNRotate(rots=(-30,30), scales=(.75,1.25)) -- reference
src/data/custom_transforms.py:9-89,188-211, composed in src/util/helper_func.py:254-263."""
import random

import cv2
import numpy as np


def random_flip(image, gt):
    if random.random() < 0.5:
        image, gt = cv2.flip(image, flipCode=1), cv2.flip(gt, flipCode=1)
    return image, gt


def _warp(a, rot, sc, label):
    h, w = a.shape[:2]
    M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
    return cv2.warpAffine(a, M, (w, h), flags=cv2.INTER_NEAREST if label else cv2.INTER_CUBIC)


def random_scale_rotate(image, gt, rots=(-30, 30), scales=(.75, 1.25)):
    num_labels = len(np.unique(gt))
    while True:
        rot = (rots[1] - rots[0]) * random.random() - (rots[1] - rots[0]) / 2
        sc = (scales[1] - scales[0]) * random.random() - (scales[1] - scales[0]) / 2 + 1
        aug_gt = _warp(gt, rot, sc, True)
        if not num_labels > 1 or len(np.unique(aug_gt)) == num_labels:
            break
    return _warp(image, rot, sc, False), aug_gt


def augment_first_frame(image_hwc, gt_hw):
    """image float32 [H,W,3] in [0,1], gt float32 [H,W] -> augmented copies (same dtypes)."""
    image, gt = random_flip(image_hwc, gt_hw)
    return random_scale_rotate(image, gt)


class DeviceAugmenter:
    """Same random draws, flip and label warp as augment_first_frame (host, cheap: nearest on one channel), but
    the expensive bicubic image warp runs on the GPU from the device-resident first frame
    (kernels.affine_warp_cubic).  Not bit-identical to cv2's fixed-point bicubic (1/32-px coordinate grid)."""

    def __init__(self, frame0_chw_device, gt_hw):
        self.src = frame0_chw_device.contiguous()
        self.gt = np.ascontiguousarray(gt_hw, dtype=np.float32)
        self.h, self.w = self.gt.shape

    def batch(self, batch_size, rots=(-30, 30), scales=(.75, 1.25)):
        import torch
        from .. import kernels as K
        h, w = self.h, self.w
        minv = np.empty((batch_size, 6), np.float32)
        flips = np.empty((batch_size,), np.int32)
        gts = np.empty((batch_size, 1, h, w), np.float32)
        for b in range(batch_size):
            do_flip = random.random() < 0.5
            gt = cv2.flip(self.gt, flipCode=1) if do_flip else self.gt

            def n_ids(a):     # == len(np.unique(a)) for integer-valued label maps, without the sort
                return int(np.count_nonzero(np.bincount(a.astype(np.uint8).ravel(), minlength=256)))

            num_labels = n_ids(gt)
            while True:
                rot = (rots[1] - rots[0]) * random.random() - (rots[1] - rots[0]) / 2
                sc = (scales[1] - scales[0]) * random.random() - (scales[1] - scales[0]) / 2 + 1
                M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
                aug_gt = cv2.warpAffine(gt, M, (w, h), flags=cv2.INTER_NEAREST)
                if not num_labels > 1 or n_ids(aug_gt) == num_labels:
                    break
            minv[b] = cv2.invertAffineTransform(M).reshape(6)
            flips[b] = int(do_flip)
            gts[b, 0] = aug_gt
        dev = self.src.device
        minv_d = torch.from_numpy(minv).to(dev, non_blocking=True)
        flip_d = torch.from_numpy(flips).to(dev, non_blocking=True)
        gts_d = torch.from_numpy(gts).to(dev, non_blocking=True)
        return K.affine_warp_cubic(self.src, minv_d, flip_d, batch_size), gts_d
