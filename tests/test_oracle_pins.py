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
