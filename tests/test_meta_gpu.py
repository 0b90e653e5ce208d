"""Meta-training path (BASELINE config 5) on the GPU: fused-update autograd, fused RAdam, first-order BPTT
meta-gradients vs the CPU oracle (oracle/model_oracle.py::oracle_meta_gradients)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model_oracle as MO  # noqa: E402
from tests.conftest import statistical  # noqa: E402


def test_fused_update_autograd_matches_torch():
    """d/d theta and d/d lambda of the fused update == autograd of p - g * lr (exact formula, fp32)."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import _FusedUpdateFn
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(8, 3, 3, 3), (8,), (5, 16), (1, 1)]
    for use_log in (False, True):
        ps = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
        gs = [torch.randn(s, generator=g).to(dev) for s in shapes]
        raw = [torch.rand((s[0],) + (1,) * (len(s) - 1), generator=g).mul(1e-2).add(1e-3) for s in shapes]
        ls = [(r.log() if use_log else r).to(dev).requires_grad_(True) for r in raw]
        outs = _FusedUpdateFn.apply(use_log, len(ps), None, True, *ps, *gs, *ls)
        ref = [p - gg * (l.exp() if use_log else l) for p, gg, l in zip(ps, gs, ls)]
        ws = [torch.randn(s, generator=g).to(dev) for s in shapes]
        got = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs, ws)), ps + ls)
        exp = torch.autograd.grad(sum((o * w).sum() for o, w in zip(ref, ws)), ps + ls)
        for a, b in zip(got, exp):
            # d / d lambda is a row sum of up to 27 products of O(1) terms: the fused kernel sums in another order
            assert torch.allclose(a, b, rtol=1e-5, atol=5e-6)


def test_fused_radam_matches_reference():
    import eosvos_b200  # noqa: F401
    from eosvos_b200 import kernels as K
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(1000, generator=g)
    p, m, v = p0.clone().to(dev), torch.zeros(1000, device=dev), torch.zeros(1000, device=dev)
    rp, rm, rv = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    import math
    for step in range(1, 9):                       # crosses the N_sma >= 5 switch (step 6 for beta2 = 0.999)
        grad = torch.randn(1000, generator=g)
        rp, rm, rv = MO.radam_reference(rp, grad, rm, rv, step, lr=1e-3, wd=1e-3)
        beta2_t = 0.999 ** step
        n_max = 2 / (1 - 0.999) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        rect = n_sma >= 5
        ss = (math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2))
              if rect else 1.0) / (1 - 0.9 ** step)
        K.radam_step(p, grad.to(dev), m, v, gscale=1.0, clip=None, beta1=0.9, beta2=0.999, eps=1e-8, lr=1e-3, wd=1e-3,
                     step_size=ss * 1e-3, rectified=rect)
        assert torch.allclose(p.cpu(), rp, rtol=1e-5, atol=1e-7), step
        assert torch.allclose(m.cpu(), rm, rtol=1e-5, atol=1e-8) and torch.allclose(v.cpu(), rv, rtol=1e-5, atol=1e-10)


def test_bptt_chain_exact_on_small_model():
    """First-order BPTT through the fused update (theta_{k+1} = theta_k - lambda * g_k, g_k detached) on a small
    fp32 torch model: meta-gradients w.r.t. theta_0 and lambda equal the hand-unrolled torch computation."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    from eosvos_b200.util import meta_train
    dev = torch.device("cuda:0")

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(6, 8)
            self.b = torch.nn.Linear(8, 3)

        def forward(self, x, y):
            loss = (self.b(torch.tanh(self.a(x))) - y).square().mean()
            return loss, {"l": loss}

        def train_without_dropout(self):
            self.train()

    torch.manual_seed(0)
    net = Net().to(dev)
    opt = MetaOptimizer(net, 5e-2, True, False, 'NEURON', False, None)
    opt.to(dev)
    g = torch.Generator().manual_seed(3)
    tb = (torch.randn(16, 6, generator=g).to(dev), torch.randn(16, 3, generator=g).to(dev))
    mb = (torch.randn(16, 6, generator=g).to(dev), torch.randn(16, 3, generator=g).to(dev))
    theta0 = [p.detach().clone().requires_grad_(True) for p in opt._model_init.values()]
    lrs = [l.detach().clone().requires_grad_(True) for l in opt.log_init_lr]
    # hand-unrolled reference
    import torch.nn.functional as Fn
    cur = theta0
    for _ in range(3):
        loss = (Fn.linear(torch.tanh(Fn.linear(tb[0], cur[0], cur[1])), cur[2], cur[3]) - tb[1]).square().mean()
        gr = torch.autograd.grad(loss, cur)
        cur = [p - gg.detach() * lr for p, gg, lr in zip(cur, gr, lrs)]
    ml = (Fn.linear(torch.tanh(Fn.linear(mb[0], cur[0], cur[1])), cur[2], cur[3]) - mb[1]).square().mean()
    ref = torch.autograd.grad(ml, theta0 + lrs)
    opt.zero_grad()
    from unittest import mock
    with mock.patch.object(meta_train, "set_random_seeds", lambda s: None):
        _, got_ml = meta_train.task_meta_gradients(net, opt, tb, mb, num_epochs=3, bptt_epochs=3)
    assert torch.allclose(got_ml, ml.detach(), rtol=1e-5)
    got = [p.grad for p in opt._model_init.values()] + [l.grad for l in opt.log_init_lr]
    for a_, b_ in zip(got, ref):
        assert torch.allclose(a_, b_, rtol=1e-4, atol=1e-7)


@statistical()
def test_meta_gradients_vs_oracle():
    """First-order BPTT through 2 fine-tune steps + meta frame on the full model.  To keep the comparison about
    arithmetic (not about which RoIs a noisy NMS lets through), both sides get the same fixed proposal set in every
    forward and the same sampler permutations.  Gradients of a random-init net still only agree statistically under
    16-bit noise (ReLU gates flip): cosine similarity of the flat meta-gradients >= 0.5 (measured
    0.75 .. 0.9), meta loss within 15 % (measured 3 .. 9 %).  The BPTT chain itself is checked exactly in test_bptt_chain_exact_on_small_model."""
    from unittest import mock
    from tests.test_model_gpu import build_pair, det_randperm, frame
    from eosvos_b200.util import meta_train
    model, opt, oracle, oopt, dev, _ = build_pair("BCE")
    img, tgt = frame()
    img2, tgt2 = frame(seed=12)
    g = torch.Generator().manual_seed(77)
    ctr = torch.rand(600, 2, generator=g) * torch.tensor([266.0, 150.0])
    wh = torch.rand(600, 2, generator=g) * torch.tensor([120.0, 80.0]) + 4
    props = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(min=0)
    props[:, 2].clamp_(max=266.0)
    props[:, 3].clamp_(max=150.0)
    jit = torch.rand(200, 4, generator=g) * 16 - 8          # boxes around the object so that positives exist
    gtbox = torch.tensor([79.8, 45.0, 186.2, 109.4])
    props = torch.cat([props, (gtbox[None] + jit).clamp(min=0)], 0)
    oracle.fixed_proposals = [props]
    model.fixed_proposals = [props]
    oracle.train_without_dropout()
    olrs = [l.clone().requires_grad_(True) for l in oopt.lrs]
    with mock.patch("torch.randperm", det_randperm(5)), mock.patch.object(meta_train, "set_random_seeds", lambda s: None):
        torch.manual_seed(21)
        oml, odt, odl = MO.oracle_meta_gradients(oracle, olrs, (img, tgt), (img2, tgt2), num_epochs=2)
    opt.zero_grad()
    with mock.patch("torch.randperm", det_randperm(5)), mock.patch.object(meta_train, "set_random_seeds", lambda s: None):
        torch.manual_seed(21)
        _, ml = meta_train.task_meta_gradients(model, opt, (img.to(dev), tgt.to(dev)), (img2.to(dev), tgt2.to(dev)),
                                               num_epochs=2, bptt_epochs=2)
    model.fixed_proposals = None
    assert abs(ml.item() - oml.item()) <= 0.15 * abs(oml.item()), (ml.item(), oml.item())
    names = [n for n, _ in opt.named_parameters()]
    gl = torch.cat([p.grad.flatten().cpu() for n, p in opt.named_parameters() if n.startswith("log_init_lr")])
    gt = torch.cat([p.grad.flatten().cpu() for n, p in opt.named_parameters() if n.startswith("model_init")])
    ol = torch.cat([torch.zeros_like(l).flatten() if g_ is None else g_.flatten() for g_, l in zip(odl, olrs)])
    ot = torch.cat([g_.flatten() for g_ in odt])
    assert len(names) == 402 and gl.numel() == ol.numel() and gt.numel() == ot.numel()
    assert torch.isfinite(gl).all() and torch.isfinite(gt).all()
    cos_t = torch.nn.functional.cosine_similarity(gt, ot, dim=0).item()
    cos_l = torch.nn.functional.cosine_similarity(gl, ol, dim=0).item()
    print("meta loss", ml.item(), oml.item(), "meta-gradient cosine: theta0", cos_t, "lambda", cos_l)
    assert cos_t >= 0.5 and cos_l >= 0.5
    # outer step runs and changes the parameters
    radam = meta_train.FusedRAdam(opt)
    before = [p.detach().clone() for _, p in opt.named_parameters()]
    flat, offs = meta_train.pack_meta_gradients(opt)
    radam.step(flat, offs, gscale=1.0, grad_clip=None, lr_clamp=(0.0, float("inf")))
    assert any(not torch.equal(a, p.detach()) for a, (_, p) in zip(before, opt.named_parameters()))
    assert all(float(p.min()) >= 0.0 for n, p in opt.named_parameters() if n.startswith("log_init_lr"))


@statistical()
def test_meta_run_worker_follows_reference_worker():
    """`eosvos_b200.util.meta_run.meta_run` (reference signature and shared-memory protocol) on the synthetic DAVIS
    train tree of tests/golden/meta_run.pt, with the configuration the UNMODIFIED reference worker ran under: the
    worker must select the same task and feed the model the very same batches (same seeds => same frames: bit-equal
    tensors), produce losses close to the reference's (its samplers use the CUDA generator, the golden run the CPU
    one: 20 %), and hand back finite meta-gradients for all 402 tensors through `shared_meta_optim_grads`."""
    import os
    import tempfile
    import eosvos_b200  # noqa: F401
    from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
    from eosvos_b200.util import helper_func, meta_run as MR
    from eosvos_b200.util.evaluate import set_random_seeds
    from oracle import ref_harness as RH
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meta_run.pt"), weights_only=False)
    cfg = g["config"]
    seen = []
    real_init = helper_func.init_parent_model

    def small_init(**kw):
        model, states = real_init(**kw)
        model.transform.min_size, model.transform.max_size = (g["min_size"],), g["max_size"]
        model.register_forward_pre_hook(lambda m, a: seen.append((a[0].detach().cpu().clone(), a[1].detach().cpu().clone())))
        return model, states

    with tempfile.TemporaryDirectory() as wd, RH._cwd(wd):
        RH.make_davis_tree(wd, [("synth_t", 9, 4, 1)], split="train_seqs", height=96, width=170)
        set_random_seeds(cfg["seed"])
        model, _ = real_init(**cfg["parent_model"])
        opt = MetaOptimizer(model, **cfg["meta_optim_cfg"])
        sd = {k: v.detach().clone() for k, v in opt.state_dict().items()}
        grads = {n: torch.zeros_like(p) for n, p in opt.named_parameters()}
        shared = {"sub_iter_done": False, "meta_epoch_done": False}
        from unittest import mock
        with mock.patch.object(MR, "init_parent_model", small_init):
            MR.meta_run(0, model.state_dict(), sd, torch.get_rng_state(), cfg, cfg["datasets"]["train"], shared,
                        {"meta_iter": 0, "meta_epoch": 0}, grads, None, 1, once=True)
    assert shared["sub_iter_done"] is True
    assert len(seen) == len(g["batches"]) == cfg["num_epochs"]["train"] + 1
    for (x, y), (gx, gy, _) in zip(seen, g["batches"]):
        assert torch.equal(x, gx.float() / 255.0) and torch.equal(y, gy.float())
    tl, ml = shared["seqs_metrics"]["train_loss"]["synth_t"][0], shared["seqs_metrics"]["meta_loss"]["synth_t"][0]
    rtl, rml = g["train_loss"]["synth_t"][0], g["meta_loss"]["synth_t"][0]
    print("train loss", tl, "reference", rtl, "meta loss", ml, "reference", rml)
    assert abs(tl - rtl) <= 0.2 * rtl and abs(ml - rml) <= 0.2 * rml
    assert len(grads) == 402 and all(bool(torch.isfinite(v).all()) for v in grads.values())
    assert sum(float(v.abs().sum()) > 0 for v in grads.values()) > 350


def test_lr_grad_kernel_on_real_tensor_set():
    """lr_grad_kernel (d L / d lambda through the fused update) on the model's 201 tensors with NEURON-level rates,
    K x K filter gradients in channels-last memory order as the wgrad epilogue emits them, lin and log mode."""
    import eosvos_b200  # noqa: F401
    from eosvos_b200 import kernels as K
    from tests.test_model_gpu import build_pair
    model, opt, _, _, dev, _ = build_pair()
    params = [p.detach() for *_, p in opt.meta_model.param_groups()]
    g = torch.Generator(device="cpu").manual_seed(0)
    grads = [torch.randn(p.shape, generator=g).to(dev) for p in params]
    grads = [x.contiguous(memory_format=torch.channels_last) if (x.dim() == 4 and x.shape[-1] > 1) else x for x in grads]
    douts = [torch.randn(p.shape, generator=g).to(dev) for p in params]
    douts = [x.contiguous(memory_format=torch.channels_last) if (x.dim() == 4 and x.shape[-1] > 1 and i % 2) else x
             for i, x in enumerate(douts)]
    for use_log in (False, True):
        lrs = [(l.detach().log() if use_log else l.detach()) for l in opt.log_init_lr]
        got = K.lr_grad(douts, grads, lrs, use_log)
        for d, gg, l, r in zip(douts, grads, lrs, got):
            red = [k for k in range(d.dim()) if l.shape[k] == 1 and d.shape[k] != 1]
            want = -(d.double() * gg.double())
            want = want.sum(dim=red, keepdim=True) if red else want
            if use_log:
                want = want * l.double().exp()
            assert r.shape == l.shape
            assert torch.allclose(r.double(), want, rtol=1e-4, atol=1e-4 * float(want.abs().max() + 1e-12))
