"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Runs the UNMODIFIED reference workers -- `util.evaluate.evaluate` (src/util/evaluate.py:20-439) and, for the
meta-training path, the per-task loop of `util.meta_run.meta_run` -- on a synthetic DAVIS / YouTube-VOS tree, on the
CPU or on a CUDA device, either with the reference's own model classes or with drop-in classes swapped in by import
(the two-line change INTEGRATION.md describes).  Nothing of the reference is edited: the API drift of the installed
torch/torchvision is bridged by oracle/ref_shims.py, and the third-party modules the reference imports but this image
lacks are provided as small stand-ins:

  davis       -> `cfg` attribute dict + `Segmentation` / `Annotation` / `db_eval_sequence` RESTATED from the published
                 DAVIS-2017 definitions (the package is github.com/timmeinhardt/davis-2017@e-osvos, not vendored;
                 call sites src/util/helper_func.py:444-458, src/data/davis.py:76-93) -- "parity unpinned" for F
  imageio     -> imsave through PIL (single-channel uint8 object-id PNGs, evaluate.py:338-342)
  prettytable, matplotlib(.pyplot), sacred, visdom -> inert stubs (never executed on the evaluated path)

Also here: composition of the reference's Sacred configuration from cfgs/*.yaml with PyYAML
(train_meta.py:21-31: meta.yaml <- torch.yaml <- named configs <- `key=value` overrides), and the synthetic dataset
trees.
"""
import contextlib
import copy
import json
import math
import os
import sys
import types

import numpy as np
import torch

from . import ref_shims


# ------------------------------------------------------------------------------------------------
# stand-ins for missing third-party modules
# ------------------------------------------------------------------------------------------------
class AttrDict(dict):
    """easydict-like: attribute access, nested dicts wrapped on the fly."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _jaccard(pred, gt):
    union = np.logical_or(pred, gt).sum()
    return 1.0 if union == 0 else float(np.logical_and(pred, gt).sum()) / float(union)


def _seg2bmap(seg):
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


def _f_measure(pred, gt, bound_th=0.008):
    import cv2
    bound_pix = bound_th if bound_th >= 1 else math.ceil(bound_th * np.linalg.norm(pred.shape))
    fg_b, gt_b = _seg2bmap(pred), _seg2bmap(gt)
    r = int(bound_pix)
    yy, xx = np.ogrid[-r:r + 1, -r:r + 1]
    k = (xx * xx + yy * yy <= r * r).astype(np.uint8)
    fg_dil = cv2.dilate(fg_b.astype(np.uint8), k).astype(bool)
    gt_dil = cv2.dilate(gt_b.astype(np.uint8), k).astype(bool)
    n_fg, n_gt = int(fg_b.sum()), int(gt_b.sum())
    if n_fg == 0 and n_gt > 0:
        p, rcl = 1.0, 0.0
    elif n_fg > 0 and n_gt == 0:
        p, rcl = 0.0, 1.0
    elif n_fg == 0 and n_gt == 0:
        p, rcl = 1.0, 1.0
    else:
        p, rcl = float((fg_b & gt_dil).sum()) / n_fg, float((gt_b & fg_dil).sum()) / n_gt
    return 0.0 if p + rcl == 0 else 2.0 * p * rcl / (p + rcl)


def _stats(v):
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0:
        return float("nan"), float("nan"), float("nan")
    ids = (np.round(np.linspace(1, len(v), 5) + 1e-10) - 1).astype(np.int64)
    bins = [v[ids[i]:ids[i + 1] + 1] for i in range(4)]
    with np.errstate(all="ignore"):
        return float(np.nanmean(v)), float(np.nanmean(v > 0.5)), float(np.nanmean(bins[0]) - np.nanmean(bins[3]))


def _make_davis_module():
    m = types.ModuleType("davis")
    cfg = AttrDict(MULTIOBJECT=True, YEAR=2017, PHASE="val", RESOLUTION="480p", SEQUENCES={},
                   PATH=AttrDict(ROOT="", DATA="", SEQUENCES="", ANNOTATIONS="", PALETTE=""))
    m.cfg = cfg
    m.phase = {"TRAIN": "train", "VAL": "val"}

    class _MaskSet:
        def __init__(self, directory, single_object):
            from PIL import Image
            self.directory = directory
            names = sorted(f for f in os.listdir(directory) if f.endswith(".png"))
            self.names = [os.path.splitext(f)[0] for f in names]
            self.masks = [np.atleast_3d(np.array(Image.open(os.path.join(directory, f))))[..., 0] for f in names]
            if single_object:
                self.masks = [(a != 0).astype(np.uint8) for a in self.masks]
            self.n_objects = int(max([a.max() for a in self.masks] + [0]))

    class Segmentation(_MaskSet):
        pass

    class Annotation(_MaskSet):
        def __init__(self, seq_name, single_object):
            # the reference anchors the annotation root at its own checkout (src/data/davis.py:93-97); the data
            # tree of a run lives under the working directory (helper_func.py:265-273), so resolve it there
            rel = os.path.relpath(cfg.PATH.ANNOTATIONS, cfg.PATH.ROOT)
            super().__init__(os.path.join(os.getcwd(), rel, seq_name), single_object)

    def db_eval_sequence(segmentations, annotations, measure="J"):
        """Per-object statistics over the frames excluding the first and the last (DAVIS-2017 protocol)."""
        fn = _jaccard if measure == "J" else _f_measure
        ann = dict(zip(annotations.names, annotations.masks))
        seg = dict(zip(segmentations.names, segmentations.masks))
        names = [n for n in annotations.names if n in seg][1:-1]
        out = {"mean": [], "recall": [], "decay": [], "raw": []}
        for k in range(1, int(annotations.n_objects) + 1):
            raw = [fn(seg[n] == k, ann[n] == k) for n in names]
            mean, recall, decay = _stats(raw)
            out["mean"].append(mean)
            out["recall"].append(recall)
            out["decay"].append(decay)
            out["raw"].append(raw)
        return out

    def _unavailable(*a, **k):
        raise NotImplementedError("davis stand-in: only db_eval_sequence is restated")

    m.Segmentation, m.Annotation, m.db_eval_sequence = Segmentation, Annotation, db_eval_sequence
    m.DAVISLoader, m.db_eval = _unavailable, _unavailable
    return m


def _make_imageio_module():
    m = types.ModuleType("imageio")

    def imsave(path, arr):
        from PIL import Image
        a = np.asarray(arr)
        if a.ndim == 3 and a.shape[2] == 1:
            a = a[..., 0]
        Image.fromarray(a).save(path)

    m.imsave = m.imwrite = imsave
    return m


def _inert(name, attrs=()):
    m = types.ModuleType(name)

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, k):
            return _Any()

    for a in attrs:
        setattr(m, a, _Any)
    m.__getattr__ = lambda k: _Any()
    return m


_full_installed = False


def install_full():
    """ref_shims.install() + the stand-ins above; afterwards `import util.evaluate`, `util.meta_run`, `util.radam`,
    `data` resolve to the reference's own (unmodified) modules."""
    global _full_installed
    if _full_installed:
        return
    for name, mod in (("davis", _make_davis_module()), ("imageio", _make_imageio_module()),
                      ("prettytable", _inert("prettytable", ["PrettyTable"])),
                      ("matplotlib", _inert("matplotlib")), ("matplotlib.pyplot", _inert("matplotlib.pyplot")),
                      ("sacred", _inert("sacred", ["Experiment"])), ("visdom", _inert("visdom", ["Visdom"]))):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mod
    if "matplotlib" in sys.modules and isinstance(sys.modules["matplotlib"], types.ModuleType):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # removed aliases the reference still uses at call time
    import collections
    import collections.abc
    if not hasattr(collections, "Mapping"):
        collections.Mapping = collections.abc.Mapping
    ref_shims.install()
    _full_installed = True


class _RandOnDevice:
    """`torch` as seen by the reference's networks.mask_rcnn when it runs on a CUDA device.  The reference mixes
    `torch.rand(n)` (CPU) with 0-dim CUDA tensors (mask_rcnn.py:270-273), which torch 1.2 accepted and torch 2 rejects:
    draw from the same CPU generator (same random stream), then move the numbers to the model's device."""

    def __init__(self, device):
        self._device = torch.device(device)

    def rand(self, *a, **k):
        return torch.rand(*a, **k).to(self._device)

    def arange(self, *a, **k):
        # mask_rcnn.py:381 builds a CPU index vector and indexes it with a CUDA tensor (accepted by torch 1.2)
        k.setdefault("device", self._device)
        return torch.arange(*a, **k)

    def __getattr__(self, k):
        return getattr(torch, k)


@contextlib.contextmanager
def reference_on_device(device):
    """Context in which the reference's model code can run on `device` (no-op for the CPU)."""
    install_full()
    import networks.mask_rcnn as ref_mr
    saved = ref_mr.torch
    if torch.device(device).type == "cuda":
        ref_mr.torch = _RandOnDevice(device)
    try:
        yield
    finally:
        ref_mr.torch = saved


def reference_workers():
    """-> (util.evaluate, util.helper_func, util.meta_run, util.radam) of the unmodified reference."""
    install_full()
    import util.evaluate as ev
    import util.helper_func as hf
    import util.meta_run as mr
    import util.radam as radam
    return ev, hf, mr, radam


# ------------------------------------------------------------------------------------------------
# Sacred-free composition of the reference configuration
# ------------------------------------------------------------------------------------------------
NAMED_CONFIGS = {"DAVIS-2017": "meta_davis-2017.yaml", "YouTube-VOS": "meta_youtube-vos.yaml",
                 "e-OSVOS": "eval_e-osvos.yaml", "e-OSVOS-OnA": "eval_e-osvos-OnA.yaml"}


def _deep_update(d, u):
    for k, v in u.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            _deep_update(d[k], v)
        else:
            d[k] = copy.deepcopy(v)
    return d


def compose_config(named=(), overrides=None, cfg_dir=None):
    """train_meta.py:21-31 without Sacred: cfgs/meta.yaml <- cfgs/torch.yaml <- named configs (in order) <-
    overrides given as {"a.b.c": value}."""
    import yaml
    cfg_dir = cfg_dir or os.path.join(ref_shims.REFERENCE_ROOT, "cfgs")
    cfg = {}
    for f in ("meta.yaml", "torch.yaml") + tuple(NAMED_CONFIGS[n] for n in named):
        with open(os.path.join(cfg_dir, f)) as fh:
            _deep_update(cfg, yaml.safe_load(fh) or {})
    for key, val in (overrides or {}).items():
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return cfg


# ------------------------------------------------------------------------------------------------
# synthetic dataset trees
# ------------------------------------------------------------------------------------------------
def _smooth_noise(rs, h, w, c, cell):
    gh, gw = h // cell + 2, w // cell + 2
    g = rs.rand(gh, gw, c).astype(np.float32)
    ys = np.linspace(0, gh - 1.001, h, dtype=np.float32)
    xs = np.linspace(0, gw - 1.001, w, dtype=np.float32)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
    b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def synthetic_video(seed, num_frames, height, width, num_objects, appear=None):
    """Smooth random background + K moving textured ellipses (>= 1 % area each).  appear[k] = first frame in which
    object k exists (YouTube-VOS: objects may appear late).  -> frames uint8 [T,H,W,3] RGB, labels uint8 [T,H,W]."""
    rs = np.random.RandomState(seed)
    bg = _smooth_noise(rs, height, width, 3, max(height // 10, 4)) * 0.6 + _smooth_noise(rs, height, width, 3, 8) * 0.2 + 0.1
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    objs = []
    for k in range(num_objects):
        objs.append(dict(cx=rs.uniform(0.25, 0.75) * width, cy=rs.uniform(0.3, 0.7) * height,
                         rx=rs.uniform(0.08, 0.16) * width, ry=rs.uniform(0.12, 0.22) * height,
                         vx=rs.uniform(-6, 6) * width / 854.0, vy=rs.uniform(-3, 3) * height / 480.0,
                         color=rs.uniform(0.2, 1.0, 3).astype(np.float32), tex=_smooth_noise(rs, height, width, 3, 6)))
    frames = np.zeros((num_frames, height, width, 3), np.uint8)
    labels = np.zeros((num_frames, height, width), np.uint8)
    for t in range(num_frames):
        img = bg + 0.03 * np.sin(0.05 * (xx + 3 * t))[..., None]
        lab = np.zeros((height, width), np.uint8)
        for k, o in enumerate(objs):
            if appear is not None and t < appear[k]:
                continue
            cx = np.clip(o["cx"] + o["vx"] * t, 0.15 * width, 0.85 * width)
            cy = np.clip(o["cy"] + o["vy"] * t, 0.2 * height, 0.8 * height)
            inside = ((xx - cx) / o["rx"]) ** 2 + ((yy - cy) / o["ry"]) ** 2 <= 1.0
            img = np.where(inside[..., None], 0.55 * o["color"] + 0.45 * o["tex"], img)
            lab[inside] = k + 1
        frames[t] = np.clip(img * 255.0, 0, 255).astype(np.uint8)
        labels[t] = lab
    return frames, labels


_PALETTE = [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0, 0, 0, 128, 128, 0, 128, 0, 128, 128, 128, 128, 128]


def _write_frames(img_dir, lab_dir, frames, labels, names, label_mask=None, lossless=False):
    import cv2
    from PIL import Image
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(lab_dir, exist_ok=True)
    pal = _PALETTE + [0] * (768 - len(_PALETTE))
    for i, n in enumerate(names):
        bgr = np.ascontiguousarray(frames[i][..., ::-1])
        # ".jpg" paths as in DAVIS; `lossless` stores PNG-coded bytes under that name (cv2.imread sniffs the content)
        # so that a test can rebuild the exact pixels from the generator without shipping images
        if lossless:
            ok, buf = cv2.imencode(".png", bgr)
            with open(os.path.join(img_dir, n + ".jpg"), "wb") as fh:
                fh.write(buf.tobytes())
        else:
            cv2.imwrite(os.path.join(img_dir, n + ".jpg"), bgr, [cv2.IMWRITE_JPEG_QUALITY, 95])
        if label_mask is None or label_mask[i]:
            im = Image.fromarray(np.asarray(labels[i], dtype=np.uint8), mode="P")
            im.putpalette(pal)
            im.save(os.path.join(lab_dir, n + ".png"))


def make_davis_tree(root, videos, split="val_seqs", year=2017, height=480, width=854, lossless=True):
    """videos: list of (name, seed, num_frames, num_objects).  Writes {root}/data/DAVIS-<year>/... (the CWD-relative
    layout of helper_func.py:265-273, davis.py:30-66) and returns {name: (frames, labels)}."""
    base = os.path.join(root, "data", f"DAVIS-{year}")
    out = {}
    for name, seed, T, K in videos:
        frames, labels = synthetic_video(seed, T, height, width, K)
        names = [f"{i:05d}" for i in range(T)]
        _write_frames(os.path.join(base, "JPEGImages", "480p", name), os.path.join(base, "Annotations", "480p", name),
                      frames, labels, names, lossless=lossless)
        out[name] = (frames, labels)
    with open(os.path.join(base, f"{split}.txt"), "w") as fh:
        fh.write("\n".join(v[0] for v in videos) + "\n")
    return out


def make_youtube_tree(root, videos, split="valid_seqs", height=720, width=1280, lossless=True):
    """videos: list of (name, seed, num_frames, num_objects, appear list).  Writes {root}/data/YouTube-VOS/<split
    prefix>/{JPEGImages,Annotations}/<seq>/ + meta.json (youtube.py:41-95,131-143): frames are named at the
    annotated rate (every 5th), every object's first annotation is the frame in which it appears."""
    part = split.split("_")[0]
    base = os.path.join(root, "data", "YouTube-VOS")
    meta = {"videos": {}}
    out = {}
    for name, seed, T, K, appear in videos:
        frames, labels = synthetic_video(seed, T, height, width, K, appear=appear)
        names = [f"{5 * i:05d}" for i in range(T)]
        # valid split: only the first annotation of every object is given (test_mode, youtube.py:47-48)
        ann = [i in set(appear) for i in range(T)]
        # (an annotated frame shows every object present in it; the reference selects one id, vos_dataset.py:323-339)
        _write_frames(os.path.join(base, part, "JPEGImages", name), os.path.join(base, part, "Annotations", name),
                      frames, labels, names, label_mask=ann, lossless=lossless)
        meta["videos"][name] = {"objects": {str(k + 1): {"category": "synthetic",
                                                          "frames": [names[i] for i in range(appear[k], T)]}
                                            for k in range(K)}}
        out[name] = (frames, labels, appear)
    os.makedirs(os.path.join(base, part), exist_ok=True)
    with open(os.path.join(base, part, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    with open(os.path.join(base, f"{split}.txt"), "w") as fh:
        fh.write("\n".join(v[0] for v in videos) + "\n")
    return out


# ------------------------------------------------------------------------------------------------
# running the reference's evaluate() once
# ------------------------------------------------------------------------------------------------
class _StopEvaluation(Exception):
    pass


class OneShotSharedDict(dict):
    """The worker loops forever, waiting for `meta_iter` to be cleared by the parent (evaluate.py:34-36, 439):
    let exactly one evaluation through, then stop it when it comes back to wait."""

    def __getitem__(self, k):
        if k == "meta_iter" and dict.get(self, "meta_iter") is not None:
            raise _StopEvaluation()
        return dict.__getitem__(self, k)


class _TorchProxy:
    """`torch` as seen by util.evaluate, with `device('cuda:<rank>')` redirected (evaluate.py:44 hard-codes CUDA)."""

    def __init__(self, device):
        self._device = torch.device(device)

    def device(self, *a, **k):
        return self._device

    def __getattr__(self, k):
        return getattr(torch, k)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class _Log:
    def __init__(self):
        self.lines = []

    def info(self, msg):
        self.lines.append(str(msg))

    warning = info


def run_reference_evaluate(config, dataset_key, workdir, device="cpu", meta_optim_state_dict=None, model_cls=None,
                           optim_cls=None, save_dir="out", spy=None):
    """Runs the reference's `evaluate` worker once with CWD = workdir (which must hold data/<DATASET>/...).
    model_cls / optim_cls: drop-in classes installed by rebinding the names the worker modules imported
    (`util.evaluate.MaskRCNN`, `util.evaluate.MetaOptimizer`, `util.helper_func.MaskRCNN`) -- no reference file is
    touched.  meta_optim_state_dict None => the state dict of a MetaOptimizer built exactly as the worker builds it
    (same seed => same random theta_0 / lambda).  spy(name, obj): observation hook, called with ("model", model) and
    ("meta_optim", meta_optim) right after the worker constructed them.
    -> (shared_dict with the worker's results, {seq: [T,H,W] uint8 predicted ids}, log lines)."""
    ev, hf, _, _ = reference_workers()
    import meta_optim.meta_optim as ref_mo
    import networks.mask_rcnn as ref_mr
    from PIL import Image
    saved = (ev.torch, ev.MaskRCNN, ev.MetaOptimizer, hf.MaskRCNN, ev.init_parent_model)
    model_cls = model_cls or ref_mr.MaskRCNN
    optim_cls = optim_cls or ref_mo.MetaOptimizer
    captured = {}

    def optim_spy(model, *a, **k):
        opt = optim_cls(model, *a, **k)
        captured["model"], captured["meta_optim"] = model, opt
        if spy is not None:
            spy("model", model)
            spy("meta_optim", opt)
        return opt

    try:
        ev.torch = _TorchProxy(device)
        ev.MaskRCNN = hf.MaskRCNN = model_cls
        ev.MetaOptimizer = optim_spy
        with _cwd(workdir), reference_on_device(device):
            if meta_optim_state_dict is None:
                hf.set_random_seeds(config["seed"])
                model, _ = hf.init_parent_model(**config["parent_model"])
                meta_optim_state_dict = copy.deepcopy(optim_cls(model, **config["meta_optim_cfg"]).state_dict())
                del model
            shared = OneShotSharedDict(meta_iter=None, best_mean_J=0.0)
            log = _Log()
            try:
                ev.evaluate(0, dataset_key, meta_optim_state_dict, {"meta_iter": 0, "meta_epoch": 0}, config, shared,
                            save_dir, {}, True, log)
            except _StopEvaluation:
                pass
            ds = config["datasets"][dataset_key]
            pred_dir = os.path.join(save_dir, "best_eval_preds", f"{ds['name']}", f"{ds['split']}")
            preds = {}
            for seq in sorted(os.listdir(pred_dir)):
                files = sorted(f for f in os.listdir(os.path.join(pred_dir, seq)) if f.endswith(".png"))
                if files:
                    preds[seq] = np.stack([np.atleast_3d(np.array(Image.open(os.path.join(pred_dir, seq, f))))[..., 0]
                                           for f in files]).astype(np.uint8)
    finally:
        ev.torch, ev.MaskRCNN, ev.MetaOptimizer, hf.MaskRCNN, ev.init_parent_model = saved
    return dict(shared), preds, log.lines, captured
