"""Pins oracle/ (the CPU restatement) to the REAL reference: every fixture under tests/golden/ was produced
by running the unmodified /root/reference code (oracle/make_golden.py).  CPU only, no GPU needed.
Same CPU arithmetic on both sides => near bit-exact tolerances (1e-6 relative)."""
import os

import pytest
import torch

from oracle import model_oracle as MO
from oracle import ops_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_lovasz_matches_reference():
    g = load("lovasz.pt")
    logits = g["logits"].clone().requires_grad_(True)
    loss = O.lovasz_hinge_per_image(logits, g["labels"], ignore=255.0)
    (grad,) = torch.autograd.grad(loss, logits)
    assert torch.allclose(loss, g["loss"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(grad, g["grad"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("mode", ["lin", "log"])
def test_meta_update_matches_reference(mode):
    g = load(f"meta_update_{mode}.pt")
    new = O.meta_update(g["params"], g["grads"], g["lrs"], g["use_log"])
    for a, b in zip(new, g["new"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["LOVASZ", "BCE"])
def test_model_matches_reference(kind):
    g = load(f"model_small_{kind.lower()}.pt")
    model = MO.build_oracle_model(seed=1, maskrcnn_loss=kind, min_size=g["min_size"], max_size=g["max_size"])
    torch.manual_seed(3)
    opt = MO.OracleMetaOptimizer(model, init_lr=1e-3)
    opt.reset()
    model.train_without_dropout()
    torch.manual_seed(21)
    loss, losses = model(g["img"], g["tgt"])
    assert torch.allclose(loss, g["loss"], rtol=1e-5), (loss.item(), g["loss"].item())
    for k, v in g["losses"].items():
        assert torch.allclose(losses[k], v, rtol=1e-5, atol=1e-7), k
    grads = opt.step(loss)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    for n, gr in zip(names, grads):
        ref = g["grad_norms"][n]
        assert abs(gr.norm().item() - ref) <= 1e-4 * max(ref, 1e-6), n
        if n in g["grad_samples"]:
            assert torch.allclose(gr.flatten()[:64], g["grad_samples"][n], rtol=1e-4, atol=1e-8), n
    for n_m, _, n_p, p in opt.groups():
        ref = g["param_norms_after_step"][f"{n_m}.{n_p}"]
        assert abs(p.norm().item() - ref) <= 1e-5 * max(ref, 1e-6)
    model.eval()
    torch.manual_seed(22)
    with torch.no_grad():
        probs, boxes = model(g["img"], g["tgt"])
    assert torch.allclose(boxes, g["eval_boxes"], rtol=1e-4, atol=1e-3)
    assert (probs - g["eval_probs"].float()).abs().max().item() < 2e-3   # fixture stored as fp16


def _oracle_evaluate_case(g):
    """Rebuilds the synthetic sequence of a tests/golden/evaluate_*.pt fixture and runs the oracle's
    evaluate_sequence with the configuration the reference worker was given."""
    from oracle import evaluate_oracle as EO
    from oracle import ref_harness as RH
    cfg = g["config"]            # the composed Sacred configuration the reference worker ran with
    tree = g["tree"]
    out = {}
    for video in tree["videos"]:
        name, seed, T, K = video[:4]
        appear = video[4] if len(video) > 4 else None
        frames, labels = RH.synthetic_video(seed, T, tree["height"], tree["width"], K, appear=appear)
        if appear is None:
            seq = EO.OracleSequence(frames, labels)
        else:
            seq = EO.OracleSequence(frames, labels, annotated=[i in set(appear) for i in range(T)],
                                    objects=[(k + 1, appear[k]) for k in range(K)], test_mode=True)
        EO.set_random_seeds(cfg["seed"])
        pm = cfg["parent_model"]
        model = MO.OracleMaskRCNN(pm["encoder"], 2, pm["roi_pool_output_sizes"], pm["eval_augment_rpn_proposals_mode"],
                                  pm["replace_batch_with_group_norms"], pm["box_nms_thresh"], pm["maskrcnn_loss"])
        model.transform.min_size, model.transform.max_size = (g["min_size"],), g["max_size"]
        opt = MO.OracleMetaOptimizer(model, cfg["meta_optim_cfg"]["init_lr"], cfg["meta_optim_cfg"]["use_log_init_lr"])
        import copy
        sd = copy.deepcopy(opt.state_dict())
        ona = cfg["eval_online_adapt"]
        pred, rec = EO.evaluate_sequence(
            model, opt, sd, seq, seed=cfg["seed"], num_epochs_eval=cfg["num_epochs"]["eval"], step=ona["step"],
            ona_epochs=ona["num_epochs"], min_prop=ona["min_prop"], batch_size=cfg["data_cfg"]["batch_sizes"]["train"],
            random_train_transform=cfg["data_cfg"]["random_train_transform"], reset_model_mode=ona["reset_model_mode"])
        out[name] = (pred, rec)
    return out


@pytest.mark.parametrize("case", ["evaluate_davis_ona", "evaluate_youtube_late"])
def test_evaluate_sequence_matches_reference(case):
    """oracle/evaluate_oracle.py (data path + fine-tune rounds + online adaptation + run_loader + merge) against the
    UNMODIFIED reference worker `util.evaluate.evaluate` (evaluate.py:20-439) run on the same synthetic tree: the
    final loss of every fine-tuning round and every predicted object-id mask.  Same CPU arithmetic and the same
    random streams on both sides => equal up to float summation order."""
    g = load(f"{case}.pt")
    res = _oracle_evaluate_case(g)
    losses = []
    for name, (pred, rec) in res.items():
        losses += rec["train_loss_seq"]
        ref = g["preds"][name]
        diff = (pred != ref).float().mean().item()
        assert pred.shape == ref.shape and diff < 1e-4, (name, diff)
    ref_losses = g["shared"]["train_loss_seq"]
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-4 * abs(b), (losses, ref_losses)


def test_radam_matches_reference():
    """oracle radam_reference against the reference's own optimizer (src/util/radam.py:28-94) over 8 steps with the
    per-group lr / weight decay of train_meta.py:110-127."""
    g = load("radam.pt")
    ps = [p.clone() for p in g["p0"]]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for step, (grads, want) in enumerate(zip(g["grads"], g["traj"]), start=1):
        for i, (lr, wd) in enumerate(g["groups"]):
            ps[i], ms[i], vs[i] = MO.radam_reference(ps[i], grads[i], ms[i], vs[i], step, lr=lr, wd=wd)
            assert torch.allclose(ps[i], want[i], rtol=1e-6, atol=1e-8), (step, i)


def test_meta_gradients_match_reference():
    """oracle_meta_gradients (first-order BPTT) against ONE meta-iteration of the unmodified reference worker
    util.meta_run.meta_run (meta_run.py:96-238): same theta_0 / lambda (same seeds), the very batches the worker fed
    to the model and the RNG state it had at every forward."""
    from oracle import evaluate_oracle as EO
    g = load("meta_run.pt")
    cfg = g["config"]
    EO.set_random_seeds(cfg["seed"])
    pm = cfg["parent_model"]
    model = MO.OracleMaskRCNN(pm["encoder"], 2, pm["roi_pool_output_sizes"], pm["eval_augment_rpn_proposals_mode"],
                              pm["replace_batch_with_group_norms"], pm["box_nms_thresh"], pm["maskrcnn_loss"])
    model.transform.min_size, model.transform.max_size = (g["min_size"],), g["max_size"]
    opt = MO.OracleMetaOptimizer(model, cfg["meta_optim_cfg"]["init_lr"], cfg["meta_optim_cfg"]["use_log_init_lr"])
    lrs = [l.clone().requires_grad_(True) for l in opt.lrs]
    batches = [(a.float() / 255.0, b.float()) for a, b, _ in g["batches"]]
    rng = [r for _, _, r in g["batches"]]
    n = cfg["num_epochs"]["train"]
    assert len(batches) == n + 1
    meta_loss, dtheta, dlr = MO.oracle_meta_gradients(model, lrs, batches[:n], batches[n], num_epochs=n, rng_states=rng)
    assert abs(meta_loss.item() - g["meta_loss"]["synth_t"][0]) <= 1e-4 * abs(meta_loss.item())
    names = [k for k, p in model.named_parameters() if p.requires_grad]
    assert len(names) == len(dtheta) == len(dlr)
    for nme, gt, gl in zip(names, dtheta, dlr):
        for prefix, got in (("model_init_", gt), ("log_init_lr_", gl)):
            key = prefix + nme.replace(".", "-")
            ref_n = g["grad_norms"][key]
            got = torch.zeros(1) if got is None else got
            assert abs(got.norm().item() - ref_n) <= 2e-3 * max(ref_n, 1e-6) + 1e-9, (key, got.norm().item(), ref_n)
            if ref_n > 0:
                assert torch.allclose(got.flatten()[:32], g["grad_samples"][key], rtol=5e-3, atol=1e-6 * max(ref_n, 1e-12) + 1e-10), key


def test_label_warp_restatement_matches_opencv():
    """The integer-arithmetic restatement of cv2.warpAffine(INTER_NEAREST) -- the specification of the label-warp
    kernel -- against OpenCV itself: every pixel of 150 random flip / rotate / scale draws on id maps and on a noise
    image at three sizes, plus the inverse the augmenter hands to the kernel."""
    import random
    import cv2
    import numpy as np
    from oracle import ops_oracle as O
    from eosvos_b200.util import augment
    rng = random.Random(3)
    for h, w in ((480, 854), (720, 1280), (97, 131)):
        yy, xx = np.mgrid[:h, :w]
        ids = np.zeros((h, w), np.float32)
        ids[(yy - h // 2) ** 2 + (xx - w // 3) ** 2 < (h // 4) ** 2] = 1
        ids[h // 10:h // 4, w // 2:w - 5] = 2
        noise = (np.random.RandomState(h).rand(h, w) * 6).astype(np.int32).astype(np.float32)
        for src in (ids, noise):
            for _ in range(25):
                rot, sc, fl = 60 * rng.random() - 30, 0.5 * rng.random() + 0.75, rng.random() < 0.5
                M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
                ref = cv2.warpAffine(cv2.flip(src, 1) if fl else src, M, (w, h), flags=cv2.INTER_NEAREST)
                assert np.array_equal(O.label_warp_nearest(src, M, fl), ref)
        # the augmenter's inverse == the one formed inside the restatement (same operations, same order)
        M = cv2.getRotationMatrix2D((w / 2, h / 2), 17.3, 1.11)
        inv = augment.DeviceAugmenter.cv_inverse(M)
        assert np.allclose(inv.reshape(2, 3) @ np.vstack([M, [0, 0, 1]])[:, :], np.hstack([np.eye(2), np.zeros((2, 1))]), atol=1e-9)
