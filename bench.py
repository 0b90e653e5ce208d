#!/usr/bin/env python
"""Benchmark of the e-OSVOS hot path (BASELINE.json: "fine-tune iters/s + inference frames/s, 480p
e-OSVOS-100-OnA").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (unmodified, oracle/_ref)

One STEP = one online-adaptation block of the e-OSVOS-100-OnA schedule on one 854x480 object (reference
src/util/evaluate.py:140-317, reset_model_mode FIRST_STEP): restore the model state saved after the first fine-tuning
round (`model.load_state_dict`, evaluate.py:200-205), ITERS_PER_STEP fine-tune iterations at batch 3 (forward +
backward + fused MetaOptimizer update), FRAMES_PER_STEP inference frames with target propagation -- 10 iterations
and the 5 frames up to the next adaptation (cfgs/eval.yaml e-OSVOS-OnA: online_adapt step 5, num_epochs 10).  The first round (FIRST_ROUND_ITERS iterations
from the random initialisation) runs once, untimed, so that the timed frames carry a real detection through the mask
branch and the paste kernel (`n_det` in the output line).
`value` = fine-tune iterations/s (device-resident inputs); `frames_per_s` = inference object-frames/s; `e2e` = the same
block driven from HOST buffers: first frame + every inference frame H2D from pinned memory, first-frame augmentation
per iteration (random draws on the host in the reference's order; label warp, rejection test and bicubic image warp on
the GPU), every loss and probability map read back (D2H) inside the timed region.  With N > 1 every rank runs its own objects (weak scaling, no data-path
collective); time = max over ranks.  Extra blocks in the same JSON line: `roofline_layers` (every distinct contraction
of an iteration, timed alone with an L2 flush between launches), `roofline_update` / `roofline_loss` / `roofline_tail`
(HBM-bound kernels), `gpu_eager_baseline` (the unmodified reference on the same GPU through cuDNN / ATen),
`sharded_set` (a DAVIS-2017-val-shaped set, LPT-sharded over the ranks: makespan) and `meta_iteration` (BASELINE
config 5: per-rank tasks -> one NCCL all-reduce of the flat meta-gradient -> fused RAdam).
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ITERS_PER_STEP = 10
FRAMES_PER_STEP = 5
FIRST_ROUND_ITERS = 40
BATCH = 3
H, W = 480, 854
# box_score_thresh of the benchmark model (the reference passes `box_nms_thresh` as box_score_thresh,
# mask_rcnn.py:452).  A random-initialised network fine-tuned for FIRST_ROUND_ITERS iterations does not reach the
# reference default 0.5, and a frame without a detection would skip the mask branch + paste: state the value used.
SCORE_THRESH = float(os.environ.get("EOSVOS_BENCH_SCORE_THRESH", "0.05"))
METRIC = "finetune_iters_per_s"
UNIT = "iter/s (batch 3, 854x480)"
WORKLOAD = ("e-OSVOS-100-OnA block on synthetic DAVIS-2017-val-shaped 854x480 video: state restore + 10 fine-tune iters "
            "(batch 3, LOVASZ) + 5 inference frames per step, Mask R-CNN R50-GN-FPN random init")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md recipe), read through NVML inside this
    process every 250 ms (a light query; an `nvidia-smi -lms` child was measured to perturb the timed region by up to
    12 %).  Falls back to one nvidia-smi query per second when NVML is unavailable."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (sm_mhz, reason bitmask or list)
        self.first = 0
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.t = None
        self.nvml = None

    def _gpu_handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.gpu
        if vis:
            ent = vis.split(",")[self.gpu].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent)
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            self.nvml, h = self._gpu_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(h, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll_nvml, args=(h,), daemon=True)
        except Exception:
            self.nvml = None
            self.t = threading.Thread(target=self._poll_smi, daemon=True)
        self.t.start()

    def _poll_nvml(self, h):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            self.stop_flag.wait(float(os.environ.get("EOSVOS_CLOCK_POLL_S", "0.25")))

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.max_mhz = float(f[1])
                self.samples.append((float(f[0]), [nm for (nm, _), v in zip(self.REASONS, f[2:6])
                                                   if v.lower().startswith("active")]))
            except Exception:
                pass
            self.stop_flag.wait(1.0)

    def mark(self):
        """Samples taken from here on belong to the timed region (the sampler itself starts before the warm-up)."""
        self.first = len(self.samples)

    def stop(self):
        self.stop_flag.set()
        if self.t is not None:
            self.t.join(timeout=3)
        sel = self.samples[self.first:] or self.samples[-1:]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        reasons = set()
        for _, r in sel:
            if isinstance(r, int):
                for nm, attr in self.REASONS:
                    if r & int(getattr(self.nvml, attr)):
                        reasons.add(nm)
            else:
                reasons.update(r)
        return {"sm_mhz": float(np.median([c for c, _ in sel])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sel)}


# ------------------------------------------------------------------------------------------------
def build_workload(seed):
    """One synthetic object: first frame + label, pre-augmented fine-tune batches, inference frames (host)."""
    import random
    from eosvos_b200.util import augment, synthetic
    frames, labels = synthetic.make_video(seed, num_frames=1 + FRAMES_PER_STEP, height=H, width=W, num_objects=1)
    fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
    gt0 = torch.from_numpy((labels[0] == 1).astype(np.float32))
    f0 = fr[0].permute(1, 2, 0).contiguous().numpy()
    random.seed(seed)
    batches = []
    for _ in range(4):      # a small pool of augmented first-frame batches (evaluate.py:224-225)
        imgs, gts = [], []
        for _ in range(BATCH):
            im, g = augment.augment_first_frame(f0, gt0.numpy())
            imgs.append(torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))))
            gts.append(torch.from_numpy(np.ascontiguousarray(g))[None])
        batches.append((torch.stack(imgs).contiguous(), torch.stack(gts).contiguous()))
    return fr, gt0, batches


def build_model(device, score_thresh=SCORE_THRESH):
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    from eosvos_b200.networks.mask_rcnn import MaskRCNN
    torch.manual_seed(1)
    model = MaskRCNN('resnet50', num_classes=2,
                     batch_norm={'accum_stats': False, 'learn_weight': False, 'learn_bias': False},
                     train_encoder=True, roi_pool_output_sizes={'box': 7, 'mask': 28},
                     eval_augment_rpn_proposals_mode='EXTEND', replace_batch_with_group_norms=True,
                     box_nms_thresh=score_thresh, maskrcnn_loss='LOVASZ')
    meta_optim = MetaOptimizer(model, init_lr=1e-3, learn_model_init=True, second_order_gradients=False,
                               lr_hierarchy_level='NEURON', use_log_init_lr=False, max_lr=None)
    model.to(device)
    meta_optim.to(device)
    meta_optim.reset()
    meta_optim.eval()
    model.roi_heads.detections_per_img = 1
    return model, meta_optim


_LOSS_PIN = {}


def _pinned_losses():
    if "t" not in _LOSS_PIN:
        _LOSS_PIN["t"] = torch.zeros(ITERS_PER_STEP, dtype=torch.float32).pin_memory()
    return _LOSS_PIN["t"]


def run_block(model, meta_optim, state_first, get_batch, get_frame, start_target, step_idx, ev=None, read_back=False,
              counts=None):
    """One step: FIRST_STEP restore + ITERS_PER_STEP fine-tune iterations + FRAMES_PER_STEP propagated frames."""
    from eosvos_b200.util import evaluate as E
    if ev is not None:
        ev[0].record()
    sink = 0.0
    loss_host = _pinned_losses() if read_back else None
    model.load_state_dict(state_first)          # evaluate.py:200-205 (the optimizer's theta_0 / lambda are untouched)
    meta_optim.eval()

    def on_iter(epoch, loss):
        nonlocal sink
        if counts is not None and model.last_num_positives is not None:
            counts["n_pos"].append(model.last_num_positives)
        if read_back:
            # evaluate.py:263 appends train_loss.item() per iteration; with early stopping off (the eval configs'
            # default, cfgs/meta.yaml:97-99) nothing consumes the value before the round ends, so the D2H copy is
            # issued asynchronously into pinned memory and read after the block (still inside the timed region)
            loss_host[epoch - 1].copy_(loss.detach(), non_blocking=True)

    E.finetune(model, meta_optim, lambda epoch: get_batch(step_idx * ITERS_PER_STEP + epoch), ITERS_PER_STEP,
               seed=1, round_idx=1 + step_idx, on_iter=on_iter)
    if ev is not None:
        ev[1].record()
    rows = []
    probs, boxes = E.run_frames(model, (get_frame(i) for i in range(FRAMES_PER_STEP)), start_target,
                                on_frame=(lambda i, t, p, b: rows.append(b)) if counts is not None else None)
    if counts is not None:
        counts["n_det"] += [int(b.abs().sum().item() > 0) for b in rows]
    if read_back:
        sink += float(E.to_host(probs).sum())     # evaluate.py:302 probs_frame_range.cpu()  (synchronises)
        sink += float(loss_host.sum())
    if ev is not None:
        ev[2].record()
    return sink


# ------------------------------------------------------------------------------------------------ rooflines
def _time_launch(fn, flush, reps):
    """Median CUDA-event time of fn() over `reps` launches, L2 flushed (a write larger than L2) before each."""
    ts = []
    for _ in range(2):
        fn()
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def record_contractions(model, meta_optim, batch):
    """Every contraction launched by one eager fine-tune iteration: [(kind, shape tuple, count)], by wrapping the
    C-ABI wrappers of kernels.py (the same entry points the graphed path replays)."""
    from eosvos_b200 import kernels as K
    calls = {}
    real = {n: getattr(K, n) for n in ("conv2d_fprop", "conv2d_dgrad", "conv2d_wgrad", "gemm_wgrad", "deconv2x2_fprop",
                                       "deconv2x2_dgrad", "deconv2x2_wgrad")}

    def note(key):
        calls[key] = calls.get(key, 0) + 1

    def fprop(x, w, bias=None, res=None, **kw):
        note(("fprop", tuple(x.shape), tuple(w.shape), kw.get("stride", 1), kw.get("pad", 0)))
        return real["conv2d_fprop"](x, w, bias, res, **kw)

    def dgrad(dy, wt, in_hw, **kw):
        note(("dgrad", tuple(dy.shape), tuple(wt.shape), tuple(in_hw), kw.get("stride", 1), kw.get("pad", 0)))
        return real["conv2d_dgrad"](dy, wt, in_hw, **kw)

    def wgrad(x, dy, ksize, **kw):
        note(("wgrad", tuple(x.shape), tuple(dy.shape), tuple(ksize), kw.get("stride", 1), kw.get("pad", 0)))
        return real["conv2d_wgrad"](x, dy, ksize, **kw)

    def gwgrad(x, dy, out, **kw):
        note(("gemm_wgrad", tuple(x.shape), tuple(dy.shape)))
        return real["gemm_wgrad"](x, dy, out, **kw)

    def dfprop(x, wd, bias4=None, **kw):
        note(("deconv_fprop", tuple(x.shape), tuple(wd.shape)))
        return real["deconv2x2_fprop"](x, wd, bias4, **kw)

    def ddgrad(dy, wdt, **kw):
        note(("deconv_dgrad", tuple(dy.shape), tuple(wdt.shape)))
        return real["deconv2x2_dgrad"](dy, wdt, **kw)

    def dwgrad(x, dy, **kw):
        note(("deconv_wgrad", tuple(x.shape), tuple(dy.shape)))
        return real["deconv2x2_wgrad"](x, dy, **kw)

    patched = dict(conv2d_fprop=fprop, conv2d_dgrad=dgrad, conv2d_wgrad=wgrad, gemm_wgrad=gwgrad, deconv2x2_fprop=dfprop,
                   deconv2x2_dgrad=ddgrad, deconv2x2_wgrad=dwgrad)
    graphs = model.use_cuda_graphs
    try:
        for n, f in patched.items():
            setattr(K, n, f)
        model.use_cuda_graphs = False
        model.train_without_dropout()
        loss, _ = model(*batch)
        meta_optim.set_train_loss(loss)
        meta_optim.step(loss)
        meta_optim.meta_model.detach_param_groups()
    finally:
        for n, f in real.items():
            setattr(K, n, f)
        model.use_cuda_graphs = graphs
    return calls


def layer_rooflines(device, calls, peaks, reps=5):
    """Every distinct contraction of the iteration timed ALONE (L2 flushed before each launch): algorithmic FLOPs and
    bytes (16-bit operands + result), achieved TFLOP/s, the per-layer bound min(tensor peak, AI x HBM bandwidth) and
    the fraction of that bound; FLOP-weighted aggregates per kernel family."""
    from eosvos_b200 import kernels as K
    act = K.ACT_DTYPE
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    bw = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def rnd(*shape):
        return (torch.randn(*shape, device=device) * 0.1).to(act)

    rows = []
    for key, count in calls.items():
        kind = key[0]
        if kind == "fprop":
            _, xs, ws, stride, pad = key
            x, w = rnd(*xs), rnd(*ws)
            N, Hh, Ww, Cin = xs
            Cout, KH, KW, _ = ws
            Ho, Wo = (Hh + 2 * pad - KH) // stride + 1, (Ww + 2 * pad - KW) // stride + 1
            flops = 2.0 * N * Ho * Wo * Cout * Cin * KH * KW
            nbytes = 2.0 * (x.numel() + w.numel() + N * Ho * Wo * Cout)
            fn = lambda: K.conv2d_fprop(x, w, stride=stride, pad=pad)
            fam = "conv_fprop (fprop + dgrad launches)"
        elif kind == "dgrad":
            _, ds, ws, in_hw, stride, pad = key
            dy, wt = rnd(*ds), rnd(*ws)
            N, Ho, Wo, Cout = ds
            Cin, KH, KW, _ = ws
            flops = 2.0 * N * Ho * Wo * Cout * Cin * KH * KW
            nbytes = 2.0 * (dy.numel() + wt.numel() + N * in_hw[0] * in_hw[1] * Cin)
            fn = lambda: K.conv2d_dgrad(dy, wt, in_hw, stride=stride, pad=pad)
            fam = "conv_fprop (fprop + dgrad launches)"
        elif kind == "wgrad":
            _, xs, ds, ksize, stride, pad = key
            x, dy = rnd(*xs), rnd(*ds)
            N, Ho, Wo, Cout = ds
            Cin = xs[-1]
            flops = 2.0 * N * Ho * Wo * Cout * Cin * ksize[0] * ksize[1]
            nbytes = 2.0 * (x.numel() + dy.numel()) + 4.0 * Cout * Cin * ksize[0] * ksize[1]
            # the destination is accumulated into (split over pixels): a preallocated buffer in the layout the
            # backward uses (channels_last strides for KxK filters)
            if ksize[0] * ksize[1] > 1:
                dw = torch.zeros((Cout, ksize[0], ksize[1], Cin), device=device).permute(0, 3, 1, 2)
            else:
                dw = torch.zeros((Cout, Cin, 1, 1), device=device)
            fn = lambda: K.conv2d_wgrad(x, dy, ksize, stride=stride, pad=pad, out=dw)
            fam = "conv_wgrad"
        elif kind == "gemm_wgrad":
            _, xs, ds = key
            x, dy = rnd(*xs), rnd(*ds)
            out = torch.zeros((ds[1], xs[1]), device=device)
            flops = 2.0 * xs[0] * xs[1] * ds[1]
            nbytes = 2.0 * (x.numel() + dy.numel()) + 4.0 * out.numel()
            fn = lambda: K.gemm_wgrad(x, dy, out, s_m=xs[1])
            fam = "conv_wgrad"
        elif kind == "deconv_fprop":
            _, xs, ws = key
            x, w = rnd(*xs), rnd(*ws)
            flops = 2.0 * xs[0] * xs[1] * xs[2] * xs[3] * ws[0]
            nbytes = 2.0 * (x.numel() + w.numel() + xs[0] * xs[1] * xs[2] * ws[0])
            fn = lambda: K.deconv2x2_fprop(x, w)
            fam = "conv_fprop (fprop + dgrad launches)"
        elif kind == "deconv_dgrad":
            _, ds, ws = key
            dy, w = rnd(*ds), rnd(*ws)
            flops = 2.0 * ds[0] * ds[1] * ds[2] * ds[3] * ws[0]
            nbytes = 2.0 * (dy.numel() + w.numel() + ds[0] * ds[1] * ds[2] // 4 * ws[0])
            fn = lambda: K.deconv2x2_dgrad(dy, w)
            fam = "conv_fprop (fprop + dgrad launches)"
        else:
            _, xs, ds = key
            x, dy = rnd(*xs), rnd(*ds)
            flops = 2.0 * xs[0] * xs[1] * xs[2] * xs[3] * ds[3] * 4
            nbytes = 2.0 * (x.numel() + dy.numel()) + 16.0 * xs[3] * ds[3]
            K.zero_pool.reset()
            K.zero_pool.hint = 0
            fn = lambda: (K.zero_pool.reset(), K.deconv2x2_wgrad(x, dy))       # (+ a 1 MB memset of its pool block)
            fam = "conv_wgrad"
        dur = _time_launch(fn, flush, reps)
        ai = flops / nbytes
        bound = min(peak_tf, ai * bw * 1e-3)
        rows.append(dict(kind=kind, shape=[list(k) if isinstance(k, tuple) else k for k in key[1:]], launches=count,
                         gflop=round(flops * 1e-9, 3), us=round(dur * 1e6, 1), tflops=round(flops / dur * 1e-12, 1),
                         bound_tflops=round(bound, 1), frac_of_bound=round(flops / dur * 1e-12 / bound, 3), family=fam))
    fams = {}
    for r in rows:
        f = fams.setdefault(r["family"], dict(gflop=0.0, us=0.0, bound_us=0.0, launches=0))
        f["gflop"] += r["gflop"] * r["launches"]
        f["us"] += r["us"] * r["launches"]
        f["bound_us"] += r["gflop"] / r["bound_tflops"] * 1e3 * r["launches"]
        f["launches"] += r["launches"]
    for kind in ("fprop", "dgrad", "wgrad"):
        sel = [r for r in rows if r["kind"].endswith(kind)]
        g = sum(r["gflop"] * r["launches"] for r in sel)
        t = sum(r["us"] * r["launches"] for r in sel)
        fams[f"all {kind}"] = dict(gflop=g, us=t, launches=sum(r["launches"] for r in sel),
                                   bound_us=sum(r["gflop"] / r["bound_tflops"] * 1e3 * r["launches"] for r in sel))
    total_g = sum(r["gflop"] * r["launches"] for r in rows)
    total_us = sum(r["us"] * r["launches"] for r in rows)
    agg = {k: dict(gflop=round(v["gflop"], 1), us=round(v["us"], 1), launches=v["launches"],
                   tflops=round(v["gflop"] / max(v["us"], 1e-9) * 1e3, 1),
                   frac_of_tensor_peak=round(v["gflop"] / max(v["us"], 1e-9) * 1e3 / peak_tf, 3),
                   frac_of_layer_bounds=round(v["bound_us"] / max(v["us"], 1e-9), 3)) for k, v in fams.items()}
    agg["all contractions"] = dict(gflop=round(total_g, 1), us=round(total_us, 1),
                                   tflops=round(total_g / total_us * 1e3, 1),
                                   frac_of_tensor_peak=round(total_g / total_us * 1e3 / peak_tf, 3),
                                   frac_of_sustained_peak=round(total_g / total_us * 1e3
                                                                / float(peaks.get("bf16_tflops_sustained", peak_tf)), 3))
    rows.sort(key=lambda r: -r["us"] * r["launches"])
    return rows, agg


def ncu_traffic(kernel_substr, path=os.path.join(ROOT, "profiles", "r02_top_kernels_raw.csv")):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the first kernel whose name contains
    `kernel_substr` in the committed `ncu --set full` capture (tools/prof_final.py); None when the file or row is absent.
    The capture is of the roofline layers (3x3 256->256 at 3x192x336; the 201-tensor update), not of this run."""
    import csv
    try:
        with open(path, newline="") as f:
            rows = list(csv.reader(f))
        H, units = rows[0], rows[1]
        ki, ri, wi = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_substr in r[ki]:
                return (float(r[ri].replace(",", "")) * scale.get(units[ri], 1.0)
                        + float(r[wi].replace(",", "")) * scale.get(units[wi], 1.0))
    except Exception:
        pass
    return None


def update_roofline(device, meta_optim, model, peaks, peak_kind, reps=20):
    """HBM roofline of the fused MetaOptimizer update on the real 201-tensor parameter set (528.1 MB / step)."""
    from eosvos_b200 import kernels as K
    params = [p.detach() for *_, p in meta_optim.meta_model.param_groups()]
    # gradients in the layouts the backward produces: KxK filter gradients arrive channels_last (wgrad epilogue)
    grads = [torch.randn_like(p).contiguous(memory_format=torch.channels_last) if (p.dim() == 4 and p.shape[-1] > 1)
             else torch.randn_like(p) for p in params]
    lrs = [l.detach() for l in meta_optim.state["log_lr"]]
    outs = [torch.empty_like(p) for p in params]
    plan = K.MetaUpdatePlan(params, grads, lrs, outs)
    nbytes = 12.0 * sum(p.numel() for p in params) + 4.0 * sum(l.numel() for l in lrs)
    for _ in range(3):
        K.meta_update(plan)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        K.meta_update(plan)
    e1.record()
    torch.cuda.synchronize(device)
    dur = e0.elapsed_time(e1) / reps * 1e-3
    peak = float(peaks["hbm_gbs"])
    return {"bound": "hbm", "kernel": "meta_update_kernel (201 tensors, 43,975,515 params)",
            "achieved": round(nbytes / dur / 1e9, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": round(nbytes / dur / 1e9 / peak, 4), "bytes_per_launch": nbytes, "us_per_launch": round(dur * 1e6, 1),
            "traffic": ncu_traffic("meta_update_kernel"),
            "note": "528 MB working set > 126 MB L2; back-to-back launches; traffic = DRAM bytes of one launch in the "
                    "committed ncu capture (profiles/r02_top_kernels_raw.csv)"}


def small_kernel_rooflines(device, peaks, n_pos=32):
    """K7 (fused mask loss fwd+grad) and K8 (inference tail) as GB/s of their algorithmic bytes: both
    are HBM-nominal but latency-bound at these sizes (KBs .. 1.6 MB per launch) -- the figure is reported, not a target."""
    from eosvos_b200 import kernels as K
    peak = float(peaks["hbm_gbs"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    g = torch.Generator(device="cpu").manual_seed(0)
    logits = torch.randn(n_pos, 2, 56, 56, generator=g).to(device)
    labels = torch.ones(n_pos, dtype=torch.int64, device=device)
    tg = (torch.rand(n_pos, 56, 56, generator=g) > 0.5).float().to(device)
    dur = _time_launch(lambda: K.mask_loss(logits, labels, tg, "LOVASZ"), flush, 10)
    nb = n_pos * 3136 * 12.0
    loss = {"bound": "hbm (latency-bound)", "kernel": f"lovasz_hinge_kernel, {n_pos} RoIs x 56x56", "achieved": round(nb / dur / 1e9, 2),
            "peak": peak, "unit": "GB/s", "frac": round(nb / dur / 1e9 / peak, 5), "bytes_per_launch": nb,
            "us_per_launch": round(dur * 1e6, 1)}
    ml = torch.randn(1, 2, 56, 56, generator=g).to(device)
    chan = torch.zeros(1, dtype=torch.int32, device=device)
    lab = torch.ones(1, dtype=torch.int64, device=device)
    box = torch.tensor([[200.0, 100.0, 600.0, 400.0]], device=device)
    dur = _time_launch(lambda: K.mask_paste_threshold(ml, chan, lab, box, 1, 1, H, W, 0.5), flush, 10)
    nb = 2 * 56 * 56 * 4.0 + H * W * 4.0
    tail = {"bound": "hbm (latency-bound)", "kernel": "paste_threshold_kernel, 1 detection -> 480x854", "achieved": round(nb / dur / 1e9, 2),
            "peak": peak, "unit": "GB/s", "frac": round(nb / dur / 1e9 / peak, 5), "bytes_per_launch": nb,
            "us_per_launch": round(dur * 1e6, 1), "note": "also writes the propagated target (1.64 MB) and box statistics"}
    return loss, tail


# ------------------------------------------------------------------------------------------------ reference arms
def _reference_model(device, tf32=False):
    """The UNMODIFIED reference (oracle/_ref or /root/reference under the import shims) exactly as cfgs/meta.yaml
    builds it, with its own MetaOptimizer."""
    from oracle import ref_harness as RH, ref_shims
    RH.install_full()
    import meta_optim.meta_optim as mo
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    model = ref_shims.build_reference_model(seed=1)
    model.roi_heads.score_thresh = SCORE_THRESH
    opt = mo.MetaOptimizer(model, init_lr=1e-3, learn_model_init=True, second_order_gradients=False,
                           lr_hierarchy_level='NEURON', use_log_init_lr=False, max_lr=None)
    model.to(device)
    opt.to(device)
    opt.reset()
    opt.eval()
    model.roi_heads.detections_per_img = 1
    return model, opt


def _reference_block(model, opt, batches, frames, tgt, n_iters, n_frames, device, sync):
    """n_iters fine-tune iterations (evaluate.py:255-274 body) + n_frames propagated frames (helper_func.py:100-126)
    on the reference's own classes -> (seconds per iteration, seconds per frame)."""
    model.train_without_dropout()
    sync()
    t0 = time.perf_counter()
    for i in range(n_iters):
        inputs, gts = batches[i % len(batches)]
        loss, _ = model(inputs.to(device), gts.to(device))
        loss.item()
        model.zero_grad()
        opt.set_train_loss(loss)
        opt.step(loss)
        opt.meta_model.detach_param_groups()
    sync()
    t_it = (time.perf_counter() - t0) / max(n_iters, 1)
    model.eval()
    targets = tgt.to(device)
    t0 = time.perf_counter()
    with torch.no_grad():
        for f in range(n_frames):
            probs, _ = model(frames[f % len(frames)].to(device), targets)
            bg = probs.max(dim=1, keepdim=True)[0].lt(0.5)
            nxt = probs.argmax(dim=1, keepdim=True).float() + 1.0
            nxt[bg] = 0.0
            targets = tgt.to(device) if nxt.sum().item() == 0 else nxt
    sync()
    t_fr = (time.perf_counter() - t0) / max(n_frames, 1)
    return t_it, t_fr


def reference_available():
    try:
        from oracle import ref_shims
        return ref_shims.available()
    except Exception:
        return False


def cpu_baseline(sample_iters=2, sample_frames=2, warm_iters=1):
    """The reference's own CPU path on a bounded sample of the same workload: `sample_iters` fine-tune iterations at
    batch 3 and `sample_frames` inference frames after `warm_iters` untimed iteration(s) (oneDNN primitive creation,
    thread pool spin-up).  kind "reference" = the unmodified reference code (oracle/_ref); "port" = the oracle
    restatement when that copy is absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fr, gt0, batches = build_workload(seed=1)
    frames = [fr[1 + i:2 + i] for i in range(FRAMES_PER_STEP)]
    cpu = torch.device("cpu")
    if reference_available():
        model, opt = _reference_model(cpu)
        kind, what = "reference", "unmodified reference (oracle/_ref) on torch CPU fp32"
    else:
        from oracle import model_oracle as MO
        model = MO.build_oracle_model(seed=1, maskrcnn_loss="LOVASZ")
        model.roi_heads.detections_per_img = 1
        torch.manual_seed(3)
        opt = MO.OracleMetaOptimizer(model, 1e-3)
        opt.reset()
        kind, what = "port", "oracle/model_oracle.py on torch CPU fp32"
    if warm_iters:
        _reference_block(model, opt, batches, frames, gt0[None, None], warm_iters, 1, cpu, lambda: None)
    t_it, t_fr = _reference_block(model, opt, batches, frames, gt0[None, None], sample_iters, sample_frames, cpu,
                                  lambda: None)
    return {"value": 1.0 / t_it, "unit": UNIT, "cores": cores, "kind": kind, "frames_per_s": 1.0 / t_fr,
            "sample": f"{sample_iters} fine-tune iteration(s) at batch {BATCH} ({t_it:.1f} s each) + {sample_frames} "
                      f"inference frame(s) ({t_fr:.1f} s each) after {warm_iters} warm-up iteration(s), {what}"}


def gpu_eager_baseline(device):
    """BASELINE.md §2's bar on the GPU: the unmodified reference on the SAME device through cuDNN / ATen / torchvision
    CUDA ops, fp32 (the reference's setting) and with TF32 allowed, for the same block (10 iterations at batch 3 + 3
    frames after 3 warm-up iterations)."""
    if not reference_available():
        return {"unavailable": "no unmodified reference copy (oracle/_ref) on this machine"}
    fr, gt0, batches = build_workload(seed=1)
    frames = [fr[1 + i:2 + i] for i in range(FRAMES_PER_STEP)]
    out = {"what": "unmodified reference (oracle/_ref) on this GPU, eager PyTorch + cuDNN/ATen + torchvision ops",
           "unit": UNIT}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        from oracle import ref_harness as RH
        model, opt = _reference_model(device, tf32)
        sync = lambda: torch.cuda.synchronize(device)
        with RH.reference_on_device(device):
            _reference_block(model, opt, batches, frames, gt0[None, None], 3, 2, device, sync)
            t_it, t_fr = _reference_block(model, opt, batches, frames, gt0[None, None], ITERS_PER_STEP,
                                          2 * FRAMES_PER_STEP, device, sync)
        out[name] = {"value": round(1.0 / t_it, 2), "frames_per_s": round(1.0 / t_fr, 2)}
        del model, opt
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return out


# ------------------------------------------------------------------------------------------------ multi-GPU blocks
def sharded_set(device, rank, world, dist, scale_videos=8, seed=1):
    """BASELINE config 3 at reduced scale: a DAVIS-2017-val-shaped set (T ~ U[34,104] / 4, K in 1..5, 854x480),
    (video, object) units LPT-sharded over the ranks (util/shard.py), every unit evaluated with
    e-OSVOS-20-OnA (20 initial + 4 adaptation iterations every 5 frames, batch 3) through evaluate_sequence.
    Strong scaling: the SET is fixed, the makespan is the max over ranks."""
    from eosvos_b200.util import evaluate as E
    from eosvos_b200.util import shard, synthetic
    spec = synthetic.davis_val_shaped_set(num_videos=scale_videos, seed=seed)
    spec = [(max(T // 4, 8), K) for T, K in spec]
    cfg = dict(num_epochs_eval=20, online_adapt_step=5, online_adapt_epochs=4)
    # units = (video, object) pairs whose object is visible in the first frame (the reference asserts as much,
    # mask_rcnn.py:623); deterministic, so every rank derives the same list
    units = []
    for v, (T, K) in enumerate(spec):
        _, lab0 = synthetic.make_video(100 + v, 1, H, W, K)
        units += [(v, k, T) for k in range(K) if (lab0[0] == k + 1).any()]
    shards, loads = shard.shard_units(units, world, **cfg)
    mine = shards[rank]
    # this rank's videos are decoded / generated BEFORE the timed region (pinned host memory), like a prefetching
    # loader would; the reference times disk reads inside, which is not what this block measures
    videos = {v: _unit_video(v, T, spec[v][1]) for v, _, T in mine}
    model, meta_optim = build_model(device)
    state = copy.deepcopy(meta_optim.state_dict())
    # warm the graphs outside the timed region (they are shape-keyed; all units share 854x480)
    fr, labels = _unit_video(0, 8, 1)
    E.evaluate_sequence(model, meta_optim, state, fr, (labels[0] == 1).to(torch.uint8), num_epochs_eval=2, online_adapt_step=5,
                        online_adapt_epochs=1, batch_size=BATCH, seed=1)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    timers = {}
    for v, k, T in mine:
        fr, labels = videos[v]
        E.evaluate_sequence(model, meta_optim, state, fr, (labels[0] == k + 1).to(torch.uint8), batch_size=BATCH, seed=1,
                            timers=timers, **cfg)
    torch.cuda.synchronize(device)
    mine_s = time.perf_counter() - t0
    t = torch.tensor([mine_s, float(timers.get("finetune_iters", 0)), float(timers.get("infer_frames", 0))],
                     device=device, dtype=torch.float64)
    if dist is not None:
        all_t = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
    else:
        all_t = [t]
    per_rank = [x.tolist() for x in all_t]
    makespan = max(p[0] for p in per_rank)
    return {"set": f"{scale_videos} synthetic DAVIS-2017-val-shaped videos ({len(units)} objects, 854x480), "
                   "e-OSVOS-20-OnA (20 + 4 iters every 5 frames, batch 3)",
            "scaling": "strong", "units": len(units), "makespan_s": round(makespan, 3),
            "per_rank_s": [round(p[0], 3) for p in per_rank], "per_rank_cost_model": [round(l, 1) for l in loads],
            "iterations": int(sum(p[1] for p in per_rank)), "object_frames": int(sum(p[2] for p in per_rank)),
            "sum_rank_s": round(sum(p[0] for p in per_rank), 3),
            "load_balance": round(sum(p[0] for p in per_rank) / (world * makespan), 3)}


def _unit_video(v, T, K):
    from eosvos_b200.util import synthetic
    frames, labels = synthetic.make_video(100 + v, T, H, W, K)
    fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous().pin_memory()
    return fr, torch.from_numpy(labels)


def meta_iteration_block(device, rank, world, dist, iters=2):
    """BASELINE config 5: one meta-iteration = meta_batch_size tasks (5 fine-tune steps at batch 1 with the update
    kept in the autograd graph + meta loss on another frame, first-order BPTT) spread over the ranks, ONE NCCL
    all-reduce of the flat 44,066,513-float meta-gradient over NVLink, fused RAdam replicated on every rank
    (reference meta_run.py:96-238, train_meta.py:361-373)."""
    from eosvos_b200.util import meta_train, synthetic
    mbs = 4 if world <= 4 else world
    per_rank = mbs // world
    model, meta_optim = build_model(device, score_thresh=0.5)
    meta_optim.train()
    radam = meta_train.FusedRAdam(meta_optim)
    tasks = []
    for t in range(per_rank):
        frames, labels = synthetic.make_video(500 + rank * 16 + t, 6, H, W, 1)
        fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).contiguous()
        gt = torch.from_numpy((labels == 1).astype(np.float32))
        tasks.append(((fr[0:1].to(device), gt[0][None, None].to(device)), (fr[4:5].to(device), gt[4][None, None].to(device))))
    import torch.distributed as tdist
    real_all_reduce = tdist.all_reduce
    ar = {"ms": [], "bytes": 0}

    def timed_all_reduce(tensor, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = real_all_reduce(tensor, *a, **k)
        e1.record()
        e1.synchronize()
        ar["ms"].append(e0.elapsed_time(e1))
        ar["bytes"] = tensor.numel() * tensor.element_size()
        return r

    times = []
    tdist.all_reduce = timed_all_reduce
    try:
        for it in range(iters + 1):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            meta_train.meta_iteration(model, meta_optim, radam, tasks, mbs, num_epochs=5, bptt_epochs=5, seed=1, meta_iter=it)
            torch.cuda.synchronize(device)
            times.append(time.perf_counter() - t0)
    finally:
        tdist.all_reduce = real_all_reduce
    t = torch.tensor([min(times[1:])], device=device, dtype=torch.float64)
    iso = None
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the exchange alone (ranks aligned by a barrier first; inside a meta-iteration the call also waits for the
        # slowest rank's tasks): 5 all-reduces of the flat fp32 meta-gradient
        flat, _ = meta_train.pack_meta_gradients(meta_optim)
        for _ in range(2):
            real_all_reduce(flat)
        dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            real_all_reduce(flat)
        e1.record()
        e1.synchronize()
        iso = torch.tensor([e0.elapsed_time(e1) / 5], device=device, dtype=torch.float64)
        dist.all_reduce(iso, op=dist.ReduceOp.MAX)
        iso = float(iso.item())
    out = {"config": f"meta_batch_size {mbs} ({per_rank} task(s) per rank), 5 fine-tune steps + meta loss per task, batch 1, 854x480",
           "s_per_meta_iteration": round(float(t.item()), 4), "tasks_per_s": round(mbs / float(t.item()), 2),
           "params": sum(p.numel() for p in meta_optim.parameters())}
    if ar["ms"]:
        ms = float(np.median(ar["ms"][1:] or ar["ms"]))
        # ring all-reduce moves 2 (N-1)/N x the buffer per rank: report the algorithmic (bus) bandwidth
        out["all_reduce"] = {"bytes": ar["bytes"], "ms_inside_iteration_incl_rank_skew": round(ms, 3)}
        if iso:
            out["all_reduce"].update(ms=round(iso, 3), algbw_GBps=round(ar["bytes"] / iso * 1e-6, 1),
                                     busbw_GBps=round(ar["bytes"] / iso * 1e-6 * 2 * (world - 1) / world, 1),
                                     note="NCCL over NVLink, ranks aligned; bus bandwidth = 2 (N-1)/N x bytes / time")
    return out


# ------------------------------------------------------------------------------------------------
def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, each step a bounded
    sample of the workload (2 iterations + 2 frames), W warm-up steps (at least 1 iteration) and K timed steps."""
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    vals, fvals = [], []
    t0 = time.perf_counter()
    base = None
    for s in range(steps):
        base = cpu_baseline(2, 2, warm_iters=warm if s == 0 else 0)
        vals.append(base["value"])
        fvals.append(base["frames_per_s"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    # ms_per_step of THIS arm's step definition (10 iterations + 5 frames), from the measured per-unit times
    ms_step = 1e3 * (ITERS_PER_STEP / v + FRAMES_PER_STEP / float(np.mean(fvals)))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "frames_per_s": float(np.mean(fvals)),
            "config": {"workload": WORKLOAD, "sampled": "each step = 2 iterations + 2 frames; ms_per_step extrapolates "
                                                        "the measured per-iteration / per-frame time to 10 + 5"},
            "gpu_launches": 0, "wall_s": wall,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": base["cores"], "kind": base["kind"],
                             "sample": base["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_layers / gpu_eager_baseline / sharded_set / "
                                                             "meta_iteration (the headline numbers only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if world > 1:
        # keep the ranks' host threads (augmentation worker, Python) off each other's cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(len(cores) // world, 1)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
            torch.set_num_threads(max(per // 2, 1))
        except Exception:
            pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    from eosvos_b200 import _lib, kernels
    from eosvos_b200.util import evaluate as E
    ACT_NAME = "f16" if kernels.ACT_DTYPE == torch.float16 else "bf16"
    peaks, peak_kind = load_peaks()
    model, meta_optim = build_model(device)
    fr, gt0, batches = build_workload(seed=1 + rank)      # every rank fine-tunes on its own object (weak scaling)

    # ---- device-resident inputs (kernel-level `value`)
    dev_batches = [(a.to(device), b.to(device)) for a, b in batches]
    dev_frames = [fr[1 + i:2 + i].to(device) for i in range(FRAMES_PER_STEP)]
    dev_target = gt0[None, None].to(device)
    # ---- pinned host inputs (`e2e`)
    pin_frames = [fr[1 + i:2 + i].pin_memory() for i in range(FRAMES_PER_STEP)]

    def dev_batch(i):
        return dev_batches[i % len(dev_batches)]

    def dev_frame(i):
        return dev_frames[i]

    from eosvos_b200.util import augment
    pin_frame0 = fr[0].contiguous().pin_memory()
    gt0_np = gt0.numpy()
    e2e_state = {}

    def host_batch(i):
        # end to end: frame 0 goes H2D once per step (block); every iteration draws fresh random flips/rotations/
        # scales (host, reference RNG order) and warps label (nearest, == cv2) and image (bicubic) on the GPU
        # (the host half of the NEXT block's first iterations is started while this block finishes, as a per-object
        # driver would do for the next object)
        if i % ITERS_PER_STEP == 1 or "aug" not in e2e_state:
            if "aug" in e2e_state:
                e2e_state["aug"].close()
            e2e_state["aug"] = e2e_state.pop("next", None) or augment.PrefetchingAugmenter(
                pin_frame0.to(device, non_blocking=True), gt0_np, BATCH, lambda e: 1 + e, first_epoch=i)
        out = e2e_state["aug"].get(i)
        if i % ITERS_PER_STEP == 0:
            e2e_state["next"] = augment.PrefetchingAugmenter(pin_frame0.to(device, non_blocking=True), gt0_np, BATCH,
                                                             lambda e: 1 + e, first_epoch=i + 1)
        return out

    def host_frame(i):
        return pin_frames[i].to(device, non_blocking=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- first round (untimed): FIRST_ROUND_ITERS iterations from the initialisation, state kept for the restores
    E.finetune(model, meta_optim, lambda epoch: dev_batch(epoch), FIRST_ROUND_ITERS, seed=1, round_idx=0)
    state_first = copy.deepcopy(model.state_dict())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        run_block(model, meta_optim, state_first, dev_batch, dev_frame, dev_target, s)
    # ---- timed region 1: device-resident inputs
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    if rank == 0:
        sampler.mark()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for s in range(args.steps):
        run_block(model, meta_optim, state_first, dev_batch, dev_frame, dev_target, args.warmup + s, evs[s])
    barrier()
    wall = time.perf_counter() - t0
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ft_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    inf_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
    total_ms = evs[0][0].elapsed_time(evs[-1][2])

    # ---- timed region 2: end to end from pinned host memory, results read back
    for s in range(1):
        run_block(model, meta_optim, state_first, host_batch, host_frame, dev_target, 1000 + s, read_back=True)
    evs2 = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        run_block(model, meta_optim, state_first, host_batch, host_frame, dev_target, 2000 + s, evs2[s], read_back=True)
    barrier()
    ft2_ms = sum(e[0].elapsed_time(e[1]) for e in evs2)
    inf2_ms = sum(e[1].elapsed_time(e[2]) for e in evs2)
    # ---- one more (untimed) block that counts the positives / detections the timed blocks carry
    counts = {"n_pos": [], "n_det": []}
    run_block(model, meta_optim, state_first, dev_batch, dev_frame, dev_target, args.warmup, counts=counts)

    times = torch.tensor([ft_ms, inf_ms, total_ms, ft2_ms, inf2_ms], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ft_ms, inf_ms, total_ms, ft2_ms, inf2_ms = times.tolist()

    line = None
    if rank == 0:
        n_it = args.steps * ITERS_PER_STEP * world
        n_fr = args.steps * FRAMES_PER_STEP * world
        # per iteration: the three samples' warp matrices (fp32 for the image, fp64 for the label) and flip flags go
        # up, the per-id pixel counts / boxes of the warped labels (rejection test, 20 B per sample) come back; with
        # EOSVOS_DEVICE_LABELS=0 the host-warped labels go up instead
        dev_labels = os.environ.get("EOSVOS_DEVICE_LABELS", "1") != "0"
        per_iter_up = BATCH * (24 + 4 + 48) if dev_labels else batches[0][1].numel() * 4 + BATCH * 28
        h2d = fr[0].numel() * 4 + ITERS_PER_STEP * per_iter_up + FRAMES_PER_STEP * fr[0:1].numel() * 4
        d2h = ITERS_PER_STEP * (4 + (BATCH * 20 if dev_labels else 0)) + FRAMES_PER_STEP * H * W * 4
        line = {
            "metric": METRIC, "value": n_it / (ft_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ACT_NAME, "data": "synthetic",
            "frames_per_s": n_fr / (inf_ms * 1e-3),
            "config": {"workload": WORKLOAD, "iters_per_step": ITERS_PER_STEP, "frames_per_step": FRAMES_PER_STEP,
                       "batch": BATCH, "first_round_iters": FIRST_ROUND_ITERS, "box_score_thresh": SCORE_THRESH,
                       "l2": "inputs+activations per iteration (>2 GB) exceed the 126 MB L2",
                       "parallelism": f"objects sharded over {world} GPU(s), no data-path collective"},
            "n_pos": {"mean": float(np.mean(counts["n_pos"])) if counts["n_pos"] else None, "per_iter": counts["n_pos"]},
            "n_det": {"frames_with_detection": int(sum(counts["n_det"])), "frames": len(counts["n_det"])},
            "e2e": {"value": n_it / (ft2_ms * 1e-3), "unit": UNIT, "frames_per_s": n_fr / (inf2_ms * 1e-3),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "wall_s": wall, "clocks": clocks,
            "gpu_busy_hint": {"device_ms_per_step": total_ms / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps},
        }
        line["roofline_update"] = update_roofline(device, meta_optim, model, peaks, peak_kind)
        line["roofline_loss"], line["roofline_tail"] = small_kernel_rooflines(device, peaks)
        if not args.no_extras:
            calls = record_contractions(model, meta_optim, dev_batch(1))
            rows, agg = layer_rooflines(device, calls, peaks)
            line["roofline_layers"] = {"peak_tflops": float(peaks.get("bf16_tflops", 1590.0)), "peak_kind": peak_kind + " burst",
                                       "hbm_gbs": float(peaks.get("hbm_gbs", 6650.0)), "timing": "each launch alone, L2 flushed, median of 5",
                                       "aggregate": agg, "layers": rows}
            fam = max((k for k in agg if k.startswith("conv_")), key=lambda k: agg[k]["us"])
            a = agg[fam]
            line["roofline"] = {"bound": "tensor", "kernel": f"{fam}: {a['launches']} launches per iteration (time-dominant family)",
                                "achieved": a["tflops"], "peak": float(peaks.get("bf16_tflops", 1590.0)),
                                "peak_kind": f"{peak_kind} burst (kernels timed alone)", "unit": "TFLOP/s",
                                "frac": a["frac_of_tensor_peak"], "frac_of_layer_bounds": a["frac_of_layer_bounds"],
                                "flops_per_iteration": a["gflop"] * 1e9, "us_per_iteration": a["us"],
                                # DRAM bytes of ONE launch of the family's largest layer (3x3 256->256 at 3x192x336:
                                # 198 MB algorithmic in + out) in the committed ncu capture
                                "traffic": ncu_traffic("conv_fprop_kernel<256"),
                                "traffic_of": "conv_fprop_kernel<256>, 3x3 256->256 @3x192x336, one launch (ncu --set full)"}
        else:
            line["roofline"] = line["roofline_update"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(2, 2, warm_iters=1)
        if world == 1 and not args.no_extras:
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(device)
            except Exception as e:          # a reported baseline must not take the bench line down
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    del model, meta_optim, state_first
    torch.cuda.empty_cache()
    if not args.no_extras:
        # collective blocks: every rank takes part; a failure is reported in the line instead of losing it
        for name, fn in (("sharded_set", sharded_set), ("meta_iteration", meta_iteration_block)):
            try:
                res = fn(device, rank, world, dist)
            except Exception as e:
                import traceback
                res = {"error": f"{type(e).__name__}: {e}"[:300], "where": traceback.format_exc()[-600:]}
                if world > 1:
                    raise
            torch.cuda.empty_cache()
            if rank == 0:
                line[name] = res
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
