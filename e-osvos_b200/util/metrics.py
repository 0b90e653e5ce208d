"""DAVIS region similarity J and contour accuracy F with their per-sequence statistics, and the PNG writer for the
predicted object-id masks -- the stage after the hot path.

The reference computes J / F through the external `davis` package (github.com/timmeinhardt/davis-2017@e-osvos,
README.md:22; call sites src/util/helper_func.py:444-458 `db_eval_sequence(segmentations, annotations, measure)`),
which is not vendored under /root/reference.  What follows restates the published DAVIS-2017 definitions; it is
therefore NOT pinned against reference outputs ("parity unpinned" for F; J is the plain Jaccard index):
  J = |A n S| / |A u S|  (1 when both are empty)
  F = 2PR / (P + R) on boundary maps (seg2bmap: a pixel is boundary when its east / south / south-east neighbour
      differs), matched after dilation with a disk of radius ceil(0.008 * ||(H, W)||_2)
  per object: mean over the frames excluding the first and the last, recall = fraction > 0.5,
  decay = mean(first quarter) - mean(last quarter).
Predictions are written as in src/util/evaluate.py:332-343: one single-channel uint8 PNG of object ids per frame at
{save_dir}/{seq_name}/{frame_name}.png.
"""
import math
import os

import numpy as np


def jaccard(pred, gt):
    pred, gt = np.asarray(pred, dtype=bool), np.asarray(gt, dtype=bool)
    union = np.logical_or(pred, gt).sum()
    return 1.0 if union == 0 else float(np.logical_and(pred, gt).sum()) / float(union)


def seg2bmap(seg):
    """Boundary map of a binary mask (same size): pixels whose east, south or south-east neighbour differs."""
    seg = np.asarray(seg, dtype=bool)
    e, s, se = np.zeros_like(seg), np.zeros_like(seg), np.zeros_like(seg)
    e[:, :-1] = seg[:, 1:]
    s[:-1, :] = seg[1:, :]
    se[:-1, :-1] = seg[1:, 1:]
    b = (seg ^ e) | (seg ^ s) | (seg ^ se)
    b[-1, :] = seg[-1, :] ^ e[-1, :]
    b[:, -1] = seg[:, -1] ^ s[:, -1]
    b[-1, -1] = False
    return b


def _disk(radius):
    yy, xx = np.ogrid[-radius:radius + 1, -radius:radius + 1]
    return (xx * xx + yy * yy <= radius * radius).astype(np.uint8)


def f_measure(pred, gt, bound_th=0.008):
    """Contour accuracy of one frame / object."""
    import cv2
    pred, gt = np.asarray(pred, dtype=bool), np.asarray(gt, dtype=bool)
    bound_pix = bound_th if bound_th >= 1 else math.ceil(bound_th * np.linalg.norm(pred.shape))
    fg_b, gt_b = seg2bmap(pred), seg2bmap(gt)
    k = _disk(int(bound_pix))
    fg_dil = cv2.dilate(fg_b.astype(np.uint8), k).astype(bool)
    gt_dil = cv2.dilate(gt_b.astype(np.uint8), k).astype(bool)
    n_fg, n_gt = int(fg_b.sum()), int(gt_b.sum())
    if n_fg == 0 and n_gt > 0:
        precision, recall = 1.0, 0.0
    elif n_fg > 0 and n_gt == 0:
        precision, recall = 0.0, 1.0
    elif n_fg == 0 and n_gt == 0:
        precision, recall = 1.0, 1.0
    else:
        precision = float((fg_b & gt_dil).sum()) / n_fg
        recall = float((gt_b & fg_dil).sum()) / n_gt
    return 0.0 if precision + recall == 0 else 2.0 * precision * recall / (precision + recall)


def sequence_statistics(per_frame):
    """(mean, recall, decay) of per-frame values (already without the first / last frame)."""
    v = np.asarray(per_frame, dtype=np.float64)
    if v.size == 0:
        return float("nan"), float("nan"), float("nan")
    mean = float(np.nanmean(v))
    recall = float(np.nanmean(v > 0.5))
    ids = (np.round(np.linspace(1, len(v), 5) + 1e-10) - 1).astype(np.int64)
    bins = [v[ids[i]:ids[i + 1] + 1] for i in range(4)]
    decay = float(np.nanmean(bins[0]) - np.nanmean(bins[3]))
    return mean, recall, decay


def evaluate_sequence_jf(pred, labels, num_objects, measures=("J", "F")):
    """pred, labels: [T,H,W] object ids -> {"J": [...], "F": [...]} with one (mean, recall, decay) per object,
    frames 1 .. T-2 (the DAVIS protocol drops the first and the last frame)."""
    pred, labels = np.asarray(pred), np.asarray(labels)
    T = pred.shape[0]
    frames = range(1, max(T - 1, 2))
    out = {"J": [], "F": []}
    for k in range(1, num_objects + 1):
        if "J" in measures:
            out["J"].append(sequence_statistics([jaccard(pred[f] == k, labels[f] == k) for f in frames]))
        if "F" in measures:
            out["F"].append(sequence_statistics([f_measure(pred[f] == k, labels[f] == k) for f in frames]))
    return out


def evaluate_sequence_jf_device(pred, labels, num_objects, measures=("J", "F"), bound_th=0.008):
    """`evaluate_sequence_jf` with the per-pixel work on the GPU: pred, labels are [T,H,W] uint8 CUDA tensors of object
    ids; two launches count, per (frame, object), the integers J and F are ratios of (csrc/jf_measure.cu), the ratios
    and the sequence statistics are formed here in float64 exactly as on the host path."""
    from .. import kernels as K
    T, H, W = pred.shape
    last = max(T - 1, 2)                       # frames 1 .. T-2, as the host path
    radius = bound_th if bound_th >= 1 else math.ceil(bound_th * math.hypot(H, W))
    counts = K.jf_counts(pred[1:last].contiguous(), labels[1:last].contiguous(), num_objects, int(radius))
    c = counts.cpu().numpy().astype(np.float64)
    out = {"J": [], "F": []}
    for k in range(num_objects):
        inter, union, n_fg, n_gt, m_fg, m_gt = (c[:, k, i] for i in range(6))
        if "J" in measures:
            j = np.where(union == 0, 1.0, inter / np.maximum(union, 1.0))
            out["J"].append(sequence_statistics(j))
        if "F" in measures:
            prec = np.where(n_fg == 0, np.where(n_gt > 0, 1.0, 1.0), m_fg / np.maximum(n_fg, 1.0))
            rec = np.where(n_gt == 0, 1.0, np.where(n_fg == 0, 0.0, m_gt / np.maximum(n_gt, 1.0)))
            prec = np.where((n_fg > 0) & (n_gt == 0), 0.0, prec)
            f = np.where(prec + rec == 0, 0.0, 2.0 * prec * rec / np.maximum(prec + rec, 1e-300))
            out["F"].append(sequence_statistics(f))
    return out


def save_predictions(pred, save_dir, seq_name, frame_names=None):
    """pred [T,H,W] uint8 object ids -> {save_dir}/{seq_name}/{frame}.png (single-channel ids, evaluate.py:338-342)."""
    import cv2
    pred = np.asarray(pred, dtype=np.uint8)
    out_dir = os.path.join(save_dir, seq_name)
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for f in range(pred.shape[0]):
        name = frame_names[f] if frame_names is not None else f"{f:05d}"
        path = os.path.join(out_dir, f"{name}.png")
        if not cv2.imwrite(path, pred[f]):
            raise IOError(f"could not write {path}")
        paths.append(path)
    return paths
