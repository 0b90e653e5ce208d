"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on seeded synthetic inputs.  Run in the build
container:  python -m oracle.make_golden
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MIN_SIZE, MAX_SIZE = 160, 266      # transform.min/max_size override that keeps the CPU suite fast


def synthetic_frame(seed, h=96, w=170):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, h, w, generator=g)
    tgt = torch.zeros(1, 1, h, w)
    tgt[0, 0, 30:70, 50:120] = 1
    return img, tgt


EVAL_OVERRIDES = {"parent_model.train.val_split_files": [], "parent_model.val.val_split_files": [],
                  "parent_model.test.val_split_files": []}
# tiny end-to-end cases for the reference's evaluate() worker (CPU): (fixture, dataset, named configs, overrides, videos)
EVAL_CASES = [
    ("evaluate_davis_ona", "DAVIS-2017", ["DAVIS-2017", "e-OSVOS-OnA"],
     {"num_epochs.eval": 4, "eval_online_adapt.step": 2, "eval_online_adapt.num_epochs": 2,
      "parent_model.box_nms_thresh": 0.05},
     dict(videos=[("synth_a", 3, 5, 2)], height=96, width=170)),
    ("evaluate_youtube_late", "YouTube-VOS", ["YouTube-VOS", "e-OSVOS-OnA"],
     {"num_epochs.eval": 3, "eval_online_adapt.step": 2, "eval_online_adapt.num_epochs": 2,
      "parent_model.box_nms_thresh": 0.05, "datasets.val.split": "valid_seqs", "datasets.val.eval": True},
     dict(videos=[("synth_y", 5, 5, 2, [0, 2])], height=90, width=160)),
]


def small_transform_spy(name, obj):
    if name == "model":
        obj.transform.min_size, obj.transform.max_size = (MIN_SIZE,), MAX_SIZE


def evaluate_goldens(only=None):
    """Runs the UNMODIFIED reference worker util.evaluate.evaluate (evaluate.py:20-439) on tiny synthetic trees and
    stores what it returns through shared_dict plus the PNG predictions it wrote."""
    import tempfile
    from oracle import ref_harness as RH
    for name, dataset, named, over, tree in EVAL_CASES:
        if only and name not in only:
            continue
        over = dict(EVAL_OVERRIDES, **over)
        cfg = RH.compose_config(named, over)
        with tempfile.TemporaryDirectory() as wd:
            if dataset == "DAVIS-2017":
                RH.make_davis_tree(wd, tree["videos"], split=cfg["datasets"]["val"]["split"], height=tree["height"],
                                   width=tree["width"])
            else:
                RH.make_youtube_tree(wd, tree["videos"], split=cfg["datasets"]["val"]["split"], height=tree["height"],
                                     width=tree["width"])
            shared, preds, log, _ = RH.run_reference_evaluate(cfg, "val", wd, "cpu", spy=small_transform_spy)
        keep = {k: shared[k] for k in ("J_seq", "F_seq", "J_recall_seq", "J_decay_seq", "train_loss_seq",
                                       "train_losses_seq", "init_J_seq")}
        torch.save({"named": named, "overrides": over, "config": cfg, "tree": tree, "dataset": dataset, "shared": keep,
                    "preds": {k: torch.from_numpy(v) for k, v in preds.items()}, "min_size": MIN_SIZE,
                    "max_size": MAX_SIZE}, os.path.join(OUT, f"{name}.pt"))
        print(name, "J", keep["J_seq"], "train_loss_seq", keep["train_loss_seq"])


def radam_golden():
    """The reference's own optimizer (src/util/radam.py:28-94) with the per-group lr / weight decay of
    train_meta.py:110-127 over 8 steps (crosses the N_sma >= 5 switch at step 6)."""
    from oracle import ref_harness as RH
    _, _, _, radam = RH.reference_workers()
    g = torch.Generator().manual_seed(17)
    shapes = [(7, 5), (11,), (3, 4, 2, 2)]
    groups_cfg = [(1e-5, 1e-3), (1e-5, 0.0), (1e-3, 0.0)]         # model_init / log_init_lr / default
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    p0 = [p.detach().clone() for p in params]
    opt = radam.RAdam([{"params": [p], "lr": lr, "weight_decay": wd} for p, (lr, wd) in zip(params, groups_cfg)], lr=1e-3)
    grads, traj = [], []
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(8):
            gs = [torch.randn(s, generator=g) for s in shapes]
            for p, gg in zip(params, gs):
                p.grad = gg.clone()
            opt.step()
            grads.append(gs)
            traj.append([p.detach().clone() for p in params])
    torch.save({"p0": p0, "groups": groups_cfg, "grads": grads, "traj": traj}, os.path.join(OUT, "radam.pt"))
    print("radam golden: final norms", [t.norm().item() for t in traj[-1]])


def meta_run_golden():
    """ONE meta-iteration of the UNMODIFIED reference worker util.meta_run.meta_run (meta_run.py:14-243) on a tiny
    synthetic DAVIS-2017 train tree, CPU: meta_batch_size 1, num_epochs.train 2, bptt_epochs 2.  Stored: the batches
    the worker fed to the model (so the oracle can be driven with the very same tensors), its losses, and the
    meta-gradients it accumulated into `shared_meta_optim_grads` (norm + first 32 values per parameter)."""
    import tempfile
    from oracle import ref_harness as RH
    ev, hf, mrun, _ = RH.reference_workers()
    import meta_optim.meta_optim as ref_mo
    over = dict(EVAL_OVERRIDES, **{"meta_batch_size": 1, "num_epochs.train": 2, "bptt_epochs": 2,
                                   "datasets.train.split": "train_seqs", "eval_datasets": False,
                                   "num_meta_processes_per_gpu": 1,
                                   # per-task colour jitter calls a torchvision-0.4-only API (ColorJitter.get_params
                                   # returning a transform); it precedes the model and is not on the path
                                   "random_frame_transform_per_task": False})
    cfg = RH.compose_config(["DAVIS-2017"], over)
    seen = []

    def spy_init(**kw):
        model, states = hf.init_parent_model(**kw)
        model.transform.min_size, model.transform.max_size = (MIN_SIZE,), MAX_SIZE
        model.register_forward_pre_hook(
            lambda m, args: seen.append((args[0].clone(), args[1].clone(), torch.get_rng_state().clone())))
        return model, states

    class OneShot(dict):
        def __getitem__(self, k):
            if k == "sub_iter_done" and dict.get(self, k):
                raise RH._StopEvaluation()
            return dict.__getitem__(self, k)

    saved = (mrun.init_parent_model, mrun.device_for_process)
    try:
        mrun.init_parent_model = spy_init
        mrun.device_for_process = lambda *a, **k: (torch.device("cpu"), torch.device("cpu"))
        with tempfile.TemporaryDirectory() as wd, RH._cwd(wd):
            RH.make_davis_tree(wd, [("synth_t", 9, 4, 1)], split="train_seqs", height=96, width=170)
            hf.set_random_seeds(cfg["seed"])
            model, _ = hf.init_parent_model(**cfg["parent_model"])
            meta_optim = ref_mo.MetaOptimizer(model, **cfg["meta_optim_cfg"])
            sd = {k: v.detach().clone() for k, v in meta_optim.state_dict().items()}
            grads = {n: torch.zeros_like(p) for n, p in meta_optim.named_parameters()}
            shared = OneShot(sub_iter_done=False, meta_epoch_done=False)
            try:
                mrun.meta_run(0, model.state_dict(), sd, torch.get_rng_state(), cfg, cfg["datasets"]["train"], shared,
                              {"meta_iter": 0, "meta_epoch": 0}, grads, None, 1)
            except RH._StopEvaluation:
                pass
    finally:
        mrun.init_parent_model, mrun.device_for_process = saved
    metrics = dict.__getitem__(shared, "seqs_metrics")
    packed = []
    for a, b, r in seen:
        u8 = (a * 255.0).round().to(torch.uint8)
        assert torch.equal(u8.float() / 255.0, a), "frames are k/255 exactly (no colour jitter / warp in this case)"
        packed.append((u8, b.to(torch.uint8), r))
    torch.save({"config": cfg, "batches": packed,
                "train_loss": metrics["train_loss"], "meta_loss": metrics["meta_loss"],
                "grad_norms": {n: g_.norm().item() for n, g_ in grads.items()},
                "grad_samples": {n: g_.flatten()[:32].clone() for n, g_ in grads.items()},
                "min_size": MIN_SIZE, "max_size": MAX_SIZE}, os.path.join(OUT, "meta_run.pt"))
    print("meta_run golden:", len(seen), "forwards; train", metrics["train_loss"], "meta", metrics["meta_loss"])


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "evaluate":
        return evaluate_goldens(sys.argv[2:])
    if len(sys.argv) > 1 and sys.argv[1] == "meta":
        radam_golden()
        return meta_run_golden()
    mr, ll, mo, mm = ref_shims.reference_modules()

    # --- Lovasz hinge (loss_lovasz.py:78-126) ------------------------------------------------
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn(4, 56, 56, generator=g) * 2).requires_grad_(True)
    labels = (torch.rand(4, 56, 56, generator=g) > 0.5).float()
    labels[0, :8] = torch.rand(8, 56, generator=g)
    labels[1, 3:6] = 255.0
    loss = ll.lovasz_hinge(logits, labels, per_image=True, ignore=255.0)
    (grad,) = torch.autograd.grad(loss, logits)
    torch.save({"logits": logits.detach(), "labels": labels, "loss": loss.detach(), "grad": grad},
               os.path.join(OUT, "lovasz.pt"))

    # --- MetaOptimizer.step on a small module (meta_optim.py:177-214) --------------------------
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.GroupNorm(4, 8), torch.nn.Flatten(),
                              torch.nn.Linear(8 * 6 * 6, 5))
    for use_log in (False, True):
        torch.manual_seed(6)
        opt = mo.MetaOptimizer(net, init_lr=1e-2, learn_model_init=True, second_order_gradients=False,
                               lr_hierarchy_level='NEURON', use_log_init_lr=use_log, max_lr=None)
        opt.reset()
        opt.eval()
        x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(9))
        l = net(x).square().mean()
        params = [p.detach().clone() for p in net.parameters()]
        grads = torch.autograd.grad(l, list(net.parameters()), retain_graph=True)
        opt.step(l)
        new = [p.detach().clone() for _, _, _, p in opt.meta_model.param_groups()]
        torch.save({"params": params, "grads": [g_.detach() for g_ in grads],
                    "lrs": [l_.detach().clone() for l_ in opt.state["log_lr"]], "new": new, "use_log": use_log},
                   os.path.join(OUT, f"meta_update_{'log' if use_log else 'lin'}.pt"))
        opt.reset()

    # --- model: train forward + grads + one step + eval forward (mask_rcnn.py:572-775) ----------
    for kind in ("LOVASZ", "BCE"):
        model = ref_shims.build_reference_model(seed=1, maskrcnn_loss=kind, min_size=MIN_SIZE, max_size=MAX_SIZE)
        img, tgt = synthetic_frame(11)
        torch.manual_seed(3)
        opt = mo.MetaOptimizer(model, init_lr=1e-3, learn_model_init=True, second_order_gradients=False,
                               lr_hierarchy_level='NEURON', use_log_init_lr=False, max_lr=None)
        opt.reset()
        opt.eval()
        model.train_without_dropout()
        torch.manual_seed(21)
        loss, losses = model(img, tgt)
        names = [n for n, p in model.named_parameters() if p.requires_grad]
        grads = torch.autograd.grad(loss, [p for p in model.parameters() if p.requires_grad], retain_graph=True)
        gnorm = {n: g_.norm().item() for n, g_ in zip(names, grads)}
        keep = ["backbone.body.conv1.weight", "backbone.body.layer2.0.conv2.weight", "backbone.fpn.layer_blocks.0.0.weight",
                "rpn.head.conv.0.0.weight", "roi_heads.box_head.fc6.weight", "roi_heads.mask_predictor.conv5_mask.weight",
                "roi_heads.mask_predictor.mask_fcn_logits.bias"]
        gsample = {n: g_.flatten()[:64].clone() for n, g_ in zip(names, grads) if n in keep}
        opt.set_train_loss(loss)
        opt.step(loss)
        opt.meta_model.detach_param_groups()
        pnorm = {f"{n_m}.{n_p}": p.norm().item() for n_m, _, n_p, p in opt.meta_model.param_groups()}
        model.eval()
        torch.manual_seed(22)
        with torch.no_grad():
            probs, boxes = model(img, tgt)
        torch.save({"img": img, "tgt": tgt, "loss": loss.detach(), "losses": {k: v.detach() for k, v in losses.items()},
                    "grad_norms": gnorm, "grad_samples": gsample, "param_norms_after_step": pnorm,
                    "eval_probs": probs.half(), "eval_boxes": boxes, "min_size": MIN_SIZE, "max_size": MAX_SIZE},
                   os.path.join(OUT, f"model_small_{kind.lower()}.pt"))
        opt.reset()
    evaluate_goldens()
    radam_golden()
    meta_run_golden()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
