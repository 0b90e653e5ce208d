"""DAVIS-layout sequence I/O and a per-dataset evaluation driver around `evaluate_sequence` -- the data format on
the input side of the hot path and the host loop of reference src/util/evaluate.py:20-439.

Layout (reference src/data/davis.py:50-58, src/util/helper_func.py:265-273):
    {root}/JPEGImages/480p/<seq>/<frame>.jpg     RGB frames
    {root}/Annotations/480p/<seq>/<frame>.png    object-id label maps (palette PNG; ids 0..K)
    {root}/<split>_seqs.txt                      one sequence name per line
Pixels reach the model exactly as in src/data/vos_dataset.py:233,276-279 with `normalize: False`
(cfgs/meta.yaml:114): cv2 BGR -> RGB, float32 / 255; labels as float ids via PIL (`np.atleast_3d(label)[..., 0]`).
"""
import os

import numpy as np
import torch


def list_sequences(root, split):
    with open(os.path.join(root, f"{split}_seqs.txt")) as f:
        return [ln.strip() for ln in f if ln.strip()]


def load_sequence(root, seq, resolution="480p", pin=False):
    """-> (frames float32 [T,3,H,W] RGB in [0,1], labels uint8 [T,H,W] (zeros where a frame has no annotation),
    frame names, has_label [T] bool)."""
    import cv2
    from PIL import Image
    img_dir = os.path.join(root, "JPEGImages", *([resolution] if resolution else []), seq)
    lab_dir = os.path.join(root, "Annotations", *([resolution] if resolution else []), seq)
    names = sorted(os.path.splitext(f)[0] for f in os.listdir(img_dir) if f.lower().endswith((".jpg", ".jpeg", ".png")))
    if not names:
        raise FileNotFoundError(f"no frames under {img_dir}")
    frames, labels, has = [], [], []
    for n in names:
        path = next(os.path.join(img_dir, n + e) for e in (".jpg", ".jpeg", ".png") if os.path.exists(os.path.join(img_dir, n + e)))
        bgr = cv2.imread(path, cv2.IMREAD_COLOR)
        if bgr is None:
            raise IOError(f"could not read {path}")
        frames.append(np.ascontiguousarray(bgr[..., ::-1]))
        lp = os.path.join(lab_dir, n + ".png")
        if os.path.exists(lp):
            labels.append(np.atleast_3d(np.array(Image.open(lp)))[..., 0].astype(np.uint8))
            has.append(True)
        else:
            labels.append(None)
            has.append(False)
    H, W = frames[0].shape[:2]
    lab = np.zeros((len(names), H, W), np.uint8)
    for i, l in enumerate(labels):
        if l is not None:
            lab[i] = l
    fr = torch.from_numpy(np.stack(frames)).permute(0, 3, 1, 2).float().div_(255.0).contiguous()
    if pin and torch.cuda.is_available():
        fr = fr.pin_memory()
    return fr, torch.from_numpy(lab), names, np.array(has, dtype=bool)


class VideoSequence:
    """One sequence in memory, with the per-object label rules of the reference's datasets under
    multi_object='single_id' (src/data/vos_dataset.py:193-339, src/data/youtube.py:107-185).
    frames float32 [T,3,H,W] RGB in [0,1]; labels uint8 [T,H,W] ids (zeros where a frame has no annotation);
    annotated [T] bool; objects: None (DAVIS: ids 1..K of the first annotation) or [(label id, first annotated
    frame)] sorted by id (YouTube-VOS meta.json); test_mode: only first annotations exist (youtube.py:47-48)."""

    def __init__(self, frames, labels, names=None, annotated=None, objects=None, test_mode=None, multi_object=True):
        self.frames, self.labels = frames, labels
        T = frames.shape[0]
        self.names = names if names is not None else [f"{i:05d}" for i in range(T)]
        self.annotated = np.ones(T, bool) if annotated is None else np.asarray(annotated, bool)
        self.label_frames = [i for i in range(T) if self.annotated[i]]
        self.objects = objects
        self.test_mode = (not self.annotated.all()) if test_mode is None else test_mode
        self.multi_object = multi_object
        if not multi_object:
            self.num_objects = 1
        elif objects is None:
            self.num_objects = int((torch.unique(labels[self.label_frames[0]]) != 0).sum())
        else:
            self.num_objects = len(objects)

    def __len__(self):
        return self.frames.shape[0]

    def gt_frame_id(self, obj):
        """(frame id, label-file index) of the object's first annotation: frame 0 for DAVIS (vos_dataset.py:193-194),
        meta.json's first frame for YouTube-VOS (youtube.py:131-143)."""
        if self.objects is None:
            return 0, None
        f = self.objects[obj][1]
        return f, self.label_frames.index(f)

    def label(self, idx, obj, label_idx=None):
        """Binary float label [H,W] of object `obj` as the dataset would hand it out for frame `idx`
        (vos_dataset.py:236-245: which label file; :288-339: single-id selection)."""
        if label_idx is not None:
            raw = self.labels[self.label_frames[label_idx]]
        elif self.test_mode:
            raw = self.labels[self.label_frames[0]]
        else:
            raw = self.labels[idx]
        if self.multi_object and self.num_objects > 1:
            want = obj + 1 if self.objects is None else self.objects[obj][0]
            return (raw == want).float()
        return (raw != 0).float()


class Dataset:
    """The sequences of one split in the reference's on-disk layouts (src/data/davis.py:30-66,
    src/data/youtube.py:27-95), rooted at `data/<name>` under the working directory (helper_func.py:265-273)."""

    def __init__(self, name, split, multi_object='single_id', full_resolution=False, root=None):
        self.name, self.split, self.multi_object = name, split, multi_object
        self.root = root or os.path.join('data', name)
        self.youtube = name == 'YouTube-VOS'
        if not self.youtube and not name.startswith('DAVIS'):
            raise NotImplementedError(name)
        seqs_file = os.path.join(self.root, f"{split}.txt")
        if os.path.exists(seqs_file):
            with open(seqs_file) as f:
                self.seq_names = [ln.strip() for ln in f if ln.strip()]
        elif self.youtube:
            raise NotImplementedError
        else:
            self.seq_names = [split]
        self.all_frames = False
        if self.youtube:
            import json
            self.part = split.split('_')[0]
            self.test_mode = self.part in ('valid', 'test', 'valid-all-frames', 'test-all-frames')
            self.all_frames = 'all-frames' in self.part
            with open(os.path.join(self.root, self.part, 'meta.json')) as f:
                self.meta = json.load(f)
            self.resolution = None
        else:
            self.test_mode = 'test' in split
            year = int(''.join(c for c in name if c.isdigit()))
            self.resolution = '480p' if not full_resolution else ('1080p' if year == 2016 else 'Full-Resolution')
            if year == 2016:
                self.multi_object = False

    def load(self, seq_name, pin=False):
        if not self.youtube:
            fr, lab, names, has = load_sequence(self.root, seq_name, self.resolution, pin=pin)
            return VideoSequence(fr, lab, names, has, None, self.test_mode, multi_object=bool(self.multi_object))
        base = os.path.join(self.root, self.part)
        fr, lab, names, has = load_sequence(base, seq_name, None, pin=pin)
        info = self.meta['videos'][seq_name]['objects']
        objects = []
        for oid in sorted(info.keys()):
            first = info[oid][0] if 'test' in self.split else info[oid]["frames"][0]
            objects.append((int(oid), names.index(first)))
        if self.all_frames:           # labels of un-annotated frames repeat the first one (youtube.py:78-79)
            pass
        return VideoSequence(fr, lab, names, has, objects, self.test_mode, multi_object=bool(self.multi_object))


def open_dataset(name, split, multi_object='single_id', full_resolution=False, root=None):
    return Dataset(name, split, multi_object, full_resolution, root)


def sequence_meta(root, seq, resolution="480p"):
    """(number of frames, number of objects in the first annotation) without decoding any frame."""
    from PIL import Image
    img_dir = os.path.join(root, "JPEGImages", resolution, seq)
    lab_dir = os.path.join(root, "Annotations", resolution, seq)
    T = sum(1 for f in os.listdir(img_dir) if f.lower().endswith((".jpg", ".jpeg", ".png")))
    labs = sorted(f for f in os.listdir(lab_dir) if f.lower().endswith(".png"))
    K = int(np.atleast_3d(np.array(Image.open(os.path.join(lab_dir, labs[0]))))[..., 0].max()) if labs else 0
    return T, K


def write_sequence(root, seq, frames_u8, labels_u8, resolution="480p", names=None, jpeg_quality=95):
    """frames_u8 [T,H,W,3] RGB uint8, labels_u8 [T,H,W] ids -> DAVIS layout on disk (labels as palette PNGs)."""
    import cv2
    from PIL import Image
    img_dir = os.path.join(root, "JPEGImages", resolution, seq)
    lab_dir = os.path.join(root, "Annotations", resolution, seq)
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(lab_dir, exist_ok=True)
    palette = [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0, 0, 0, 128, 128, 0, 128, 0, 128, 128, 128, 128, 128]
    palette += [0] * (768 - len(palette))
    for i in range(frames_u8.shape[0]):
        n = names[i] if names is not None else f"{i:05d}"
        cv2.imwrite(os.path.join(img_dir, n + ".jpg"), np.ascontiguousarray(frames_u8[i][..., ::-1]),
                    [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
        im = Image.fromarray(np.asarray(labels_u8[i], dtype=np.uint8), mode="P")
        im.putpalette(palette)
        im.save(os.path.join(lab_dir, n + ".png"))


def write_split(root, split, seqs):
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, f"{split}_seqs.txt"), "w") as f:
        f.write("\n".join(seqs) + "\n")


def evaluate_dataset(model, meta_optim, meta_optim_state_dict, root, split, save_dir=None, rank=0, world_size=1,
                     evaluate_fn=None, **cfg):
    """Per-dataset loop of reference evaluate.py:95-439 on this rank's share of the sequences (whole videos are
    sharded longest-first over ranks, no exchange): load -> evaluate_sequence (fine-tune + propagate, every object)
    -> PNGs -> J / F.  Returns {seq: {"J": [...], "F": [...], "time_per_frame": s, "num_objects": K}}.
    `cfg` is passed to evaluate_sequence (num_epochs_eval, online_adapt_step, ...)."""
    from . import metrics
    from .shard import unit_cost
    if evaluate_fn is None:
        from .evaluate import evaluate_sequence as evaluate_fn
    seqs = list_sequences(root, split)
    sched_cfg = {k: cfg[k] for k in ("num_epochs_eval", "online_adapt_step", "online_adapt_epochs") if k in cfg}
    # the cost model needs only (frame count, object count): read them from the directory listing and the first
    # annotation instead of decoding every sequence on every rank
    costs = []
    for s in seqs:
        T, K = sequence_meta(root, s)
        costs.append((K * unit_cost(T, batch=cfg.get("batch_size", 3), **sched_cfg) if sched_cfg else K * T, s))
    loads = [0.0] * world_size
    mine = []
    for c, s in sorted(costs, reverse=True):
        r = min(range(world_size), key=lambda j: (loads[j], j))
        loads[r] += c
        if r == rank:
            mine.append(s)
    results = {}
    for s in mine:
        fr, lab, names, has = load_sequence(root, s, pin=True)     # only this rank's share, one sequence at a time
        pred, stats = evaluate_fn(model, meta_optim, meta_optim_state_dict, fr, lab[0], **cfg)
        pred_np = pred.cpu().numpy() if isinstance(pred, torch.Tensor) else np.asarray(pred)
        if save_dir is not None:
            metrics.save_predictions(pred_np, save_dir, s, names)
        K = int(lab[0].max())
        ann = np.where(has)[0]
        if ann.size >= 3:
            dev = next((p.device for p in getattr(model, "parameters", lambda: [])()), torch.device("cpu"))
            if dev.type == "cuda":
                jf = metrics.evaluate_sequence_jf_device(torch.from_numpy(pred_np[ann]).to(torch.uint8).to(dev),
                                                         lab[ann].to(torch.uint8).to(dev), K)
            else:                                   # a host-only caller of the dataset driver (no model on a GPU)
                jf = metrics.evaluate_sequence_jf(pred_np[ann], lab.numpy()[ann], K)
        else:
            jf = {"J": [], "F": []}
        results[s] = {"J": jf["J"], "F": jf["F"], "time_per_frame": stats.get("time_per_frame"), "num_objects": K}
    return results
