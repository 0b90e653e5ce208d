"""First-frame augmentation used while fine-tuning (host side, as in the reference):
RandomHorizontalFlip + RandomScaleNRotate(rots=(-30,30), scales=(.75,1.25)) -- reference
src/data/custom_transforms.py:9-89,188-211, composed in src/util/helper_func.py:254-263."""
import os
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

    def host_part(self, batch_size, rng=random, rots=(-30, 30), scales=(.75, 1.25), out=None):
        """Random draws (reference order: flip, then rot/scale with the rejection loop) + label warp on the host.
        Returns (minv [B,6] f32, flips [B] i32, gts [B,1,H,W] f32) as numpy views (of `out` when given)."""
        h, w = self.h, self.w
        if out is None:
            minv = np.empty((batch_size, 6), np.float32)
            flips = np.empty((batch_size,), np.int32)
            gts = np.empty((batch_size, 1, h, w), np.float32)
        else:
            minv, flips, gts = out

        def n_ids(a):     # == len(np.unique(a)) for integer-valued label maps, without the sort
            return int(np.count_nonzero(np.bincount(a.astype(np.uint8).ravel(), minlength=256)))

        for b in range(batch_size):
            do_flip = rng.random() < 0.5
            gt = cv2.flip(self.gt, flipCode=1) if do_flip else self.gt
            num_labels = n_ids(gt)
            while True:
                rot = (rots[1] - rots[0]) * rng.random() - (rots[1] - rots[0]) / 2
                sc = (scales[1] - scales[0]) * rng.random() - (scales[1] - scales[0]) / 2 + 1
                M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
                aug_gt = cv2.warpAffine(gt, M, (w, h), flags=cv2.INTER_NEAREST)
                if not num_labels > 1 or n_ids(aug_gt) == num_labels:
                    break
            minv[b] = cv2.invertAffineTransform(M).reshape(6)
            flips[b] = int(do_flip)
            gts[b, 0] = aug_gt
        return minv, flips, gts

    @staticmethod
    def cv_inverse(M):
        """The inverse cv2.warpAffine forms from a forward 2x3 matrix (imgwarp.cpp), float64, same operation order."""
        M = np.array(M, dtype=np.float64).reshape(6).copy()
        D = M[0] * M[4] - M[1] * M[3]
        D = 1.0 / D if D != 0 else 0.0
        A11, A22 = M[4] * D, M[0] * D
        M[0] = A11
        M[1] *= -D
        M[3] *= -D
        M[4] = A22
        b1 = -M[0] * M[2] - M[1] * M[5]
        b2 = -M[3] * M[2] - M[4] * M[5]
        M[2], M[5] = b1, b2
        return M

    def device_labels(self, batch_size, rng=random, rots=(-30, 30), scales=(.75, 1.25)):
        """host_part with the label work on the GPU (current stream): same random draws in the same order, the nearest
        label warp by kernels.label_warp_nearest (== cv2.warpAffine INTER_NEAREST bit for bit), the rejection test
        (`len(np.unique(aug_gt)) == num_labels`, custom_transforms.py:74-78) from the per-id pixel counts of the warped
        label, which also are the boxes / counts MaskRCNN needs for its targets.  Only the counts (20 bytes per sample)
        come back to the host.  Returns (minv [B,6] f32 numpy, flips [B] i32 numpy, gts [B,1,H,W] device, stats
        int32 [B,K,5] host)."""
        import torch
        from .. import kernels as K
        dev = self.src.device
        h, w = self.h, self.w
        if getattr(self, "gt_dev", None) is None:
            self.gt_dev = torch.from_numpy(self.gt).to(dev)
            self.num_ids = max(int(self.gt.max()), 1)
            self.num_labels = int(np.count_nonzero(np.bincount(self.gt.astype(np.uint8).ravel(), minlength=256)))
            self._flags = [torch.zeros(1, dtype=torch.int32, device=dev), torch.ones(1, dtype=torch.int32, device=dev)]
            self._stager = K.PinnedStager(slot_bytes=256, slots=32)       # private: this runs on a worker thread
        Kn = self.num_ids
        minv = np.empty((batch_size, 6), np.float32)
        flips = np.empty((batch_size,), np.int32)
        gts = torch.empty((batch_size, 1, h, w), device=dev, dtype=torch.float32)
        stats = torch.empty((batch_size, Kn, 5), dtype=torch.int32)
        for b in range(batch_size):
            do_flip = rng.random() < 0.5
            while True:
                rot = (rots[1] - rots[0]) * rng.random() - (rots[1] - rots[0]) / 2
                sc = (scales[1] - scales[0]) * rng.random() - (scales[1] - scales[0]) / 2 + 1
                M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
                m64 = self._stager.put(self.cv_inverse(M).reshape(1, 6), dev)
                K.label_warp_nearest(self.gt_dev, m64, self._flags[int(do_flip)], out=gts[b:b + 1])
                st = K.mask_to_bbox(gts[b:b + 1], Kn).cpu()                 # (waits for this sample's two kernels)
                counts = st[0, :, 4]
                present = int((counts > 0).sum()) + int(h * w - int(counts.sum()) > 0)
                if not self.num_labels > 1 or present == self.num_labels:
                    break
            minv[b] = cv2.invertAffineTransform(M).reshape(6)
            flips[b] = int(do_flip)
            stats[b] = st[0]
        return minv, flips, gts, stats

    @staticmethod
    def target_stats(gts, num_ids=1):
        """Per-sample, per-id (xmin, ymin, xmax, ymax, count) of label maps [B,1,H,W] -- what
        MaskRCNN._build_targets would otherwise derive on the device and read back (a stream sync)."""
        import torch
        B = gts.shape[0]
        st = np.empty((B, num_ids, 5), np.int32)
        for b in range(B):
            for k in range(num_ids):
                m = (gts[b, 0] == (k + 1)).astype(np.uint8)
                n = cv2.countNonZero(m)
                if n == 0:
                    st[b, k] = (2147483647, 2147483647, -1, -1, 0)
                else:
                    x, y, w, h = cv2.boundingRect(m)
                    st[b, k] = (x, y, x + w - 1, y + h - 1, n)
        return torch.from_numpy(st), torch.zeros(B, dtype=torch.int32)

    def device_part(self, minv, flips, gts, stats=None):
        import torch
        from .. import kernels as K
        dev = self.src.device
        as_t = (lambda a: a) if isinstance(minv, torch.Tensor) else torch.from_numpy
        minv_d = as_t(minv).to(dev, non_blocking=True)
        flip_d = as_t(flips).to(dev, non_blocking=True)
        gts_d = as_t(gts).to(dev, non_blocking=True)
        if stats is not None:
            K.target_stats.put(gts_d, *stats)        # host-side boxes/counts: lets forward() skip a stream sync
        return K.affine_warp_cubic(self.src, minv_d, flip_d, minv_d.shape[0]), gts_d

    def batch(self, batch_size, rots=(-30, 30), scales=(.75, 1.25)):
        minv, flips, gts = self.host_part(batch_size, random, rots, scales)
        return self.device_part(minv, flips, gts, self.target_stats(gts))


class PrefetchingAugmenter:
    """DeviceAugmenter whose host half (random draws + nearest label warp, cv2 releases the GIL) runs in a
    background thread a few iterations ahead, into a ring of pinned buffers, while the GPU works.  Every epoch uses
    its own random.Random(seed_for_epoch(epoch)) -- the same stream the reference gets from
    set_random_seeds(seed + epoch + round) (src/util/evaluate.py:221-222) followed by the dataset's transforms."""

    # pinned rings and the worker thread are shared by successive augmenters (one per object / adaptation block):
    # cudaHostAlloc of ~20 MB costs milliseconds, which would otherwise be paid at the start of every block
    _rings = {}
    _pool = None
    _side = None

    def __init__(self, frame0_chw_device, gt_hw, batch_size, seed_for_epoch, depth=3, first_epoch=None,
                 device_labels=None):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        self.aug = DeviceAugmenter(frame0_chw_device, gt_hw)
        self.batch_size, self.seed_for_epoch, self.depth = batch_size, seed_for_epoch, depth
        h, w = self.aug.h, self.aug.w
        # label warp + rejection test + boxes on the GPU (default): the worker thread then only draws the random
        # numbers and launches kernels on its own stream -- the host keeps no per-pixel work
        if device_labels is None:
            device_labels = os.environ.get("EOSVOS_DEVICE_LABELS", "1") != "0" and float(np.max(gt_hw)) <= 64
        self.device_labels = bool(device_labels) and frame0_chw_device.is_cuda
        if self.device_labels:
            if PrefetchingAugmenter._side is None:
                PrefetchingAugmenter._side = torch.cuda.Stream(device=frame0_chw_device.device)
            self.side = PrefetchingAugmenter._side
            self.ready = torch.cuda.Event()
            self.ready.record(torch.cuda.current_stream(frame0_chw_device.device))   # frame 0 is on the device
            frame0_chw_device.record_stream(self.side)     # read by the worker's kernels until the last one has run
        key = (batch_size, h, w, depth)
        rings = PrefetchingAugmenter._rings.setdefault(key, [])
        # two rings alternate, so an augmenter created while its predecessor still has copies in flight never
        # writes into that predecessor's buffers
        if len(rings) < 2:
            rings.append([(torch.empty((batch_size, 6), dtype=torch.float32).pin_memory(),
                           torch.empty((batch_size,), dtype=torch.int32).pin_memory(),
                           torch.empty((batch_size, 1, h, w), dtype=torch.float32).pin_memory())
                          for _ in range(depth + 1)])
            self.ring = rings[-1]
        else:
            rings.append(rings.pop(0))
            self.ring = rings[-1]
        if PrefetchingAugmenter._pool is None:
            PrefetchingAugmenter._pool = ThreadPoolExecutor(max_workers=1)
        self.pool = PrefetchingAugmenter._pool
        self.futures = {}
        self.next_slot = 0
        if first_epoch is not None:          # start the host half now (e.g. while the previous block still runs)
            for e in range(first_epoch, first_epoch + depth):
                self._submit(e)

    def _submit(self, epoch):
        slot = self.ring[self.next_slot % len(self.ring)]
        self.next_slot += 1

        def work():
            self.aug.host_part(self.batch_size, random.Random(self.seed_for_epoch(epoch)),
                               out=tuple(t.numpy() for t in slot))
            return slot + (self.aug.target_stats(slot[2].numpy()),)

        def work_device():
            import torch
            from .. import kernels as K
            with K.capture_lock, torch.cuda.stream(self.side):          # (never while the main thread captures a graph)
                self.side.wait_event(self.ready)
                minv, flips, gts, stats = self.aug.device_labels(self.batch_size, random.Random(self.seed_for_epoch(epoch)))
                slot[0].copy_(torch.from_numpy(minv))
                slot[1].copy_(torch.from_numpy(flips))
                dev = gts.device
                imgs = K.affine_warp_cubic(self.aug.src, slot[0].to(dev, non_blocking=True),
                                           slot[1].to(dev, non_blocking=True), self.batch_size)
                done = torch.cuda.Event()
                done.record(self.side)
            return imgs, gts, stats, done

        self.futures[epoch] = self.pool.submit(work_device if self.device_labels else work)

    def get(self, epoch):
        for e in range(epoch, epoch + self.depth):
            if e not in self.futures:
                self._submit(e)
        slot = self.futures.pop(epoch).result()
        if self.device_labels:
            import torch
            from .. import kernels as K
            imgs, gts, stats, done = slot
            cur = torch.cuda.current_stream(imgs.device)
            cur.wait_event(done)
            imgs.record_stream(cur)          # allocated on the worker's stream, consumed (and freed) on this one
            gts.record_stream(cur)
            K.target_stats.put(gts, stats, torch.zeros(stats.shape[0], dtype=torch.int32))
            return imgs, gts
        return self.aug.device_part(*slot)

    def close(self):
        for f in self.futures.values():
            f.cancel()
        self.futures.clear()
