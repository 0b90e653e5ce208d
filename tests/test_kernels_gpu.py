"""Parity of every CUDA kernel (called through the C ABI) against the CPU oracle (oracle/ops_oracle.py).

Tolerances (stated per test): operands are 16-bit (fp16 by default, bf16 optional), accumulation fp32.  The oracle is
evaluated in fp32 on the SAME bf16-rounded operands, so the remaining error is accumulation order plus
one bf16 rounding of the stored result: rel. Frobenius error <= 4e-3 for bf16 outputs, <= 2e-5 for
fp32 outputs.  Integer / index results (boxes, counts, targets) must match exactly.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ops_oracle as O  # noqa: E402


def K():
    from eosvos_b200 import kernels
    return kernels


def dev():
    return torch.device("cuda:0")


def bf(x):
    """round to the library's activation storage type (fp16 by default, bf16 with -DEOSVOS_ACT_BF16)"""
    return x.to(K().ACT_DTYPE)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def make_conv(N, H, W, Cin, Cout, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn(N, Cin, H, W, generator=g)).float()
    w = bf(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).float()
    return x, w


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, bn_hint
    (1, 16, 16, 64, 64, 1, 1, 0, 0),
    (2, 12, 21, 64, 128, 1, 1, 0, 0),
    (1, 24, 42, 256, 256, 3, 1, 1, 0),
    (2, 12, 21, 128, 64, 3, 1, 1, 0),
    (1, 48, 84, 64, 256, 3, 1, 1, 256),
    (1, 48, 84, 64, 256, 3, 1, 1, 128),
    (1, 48, 84, 64, 256, 3, 1, 1, 64),
    (1, 24, 42, 128, 128, 3, 2, 1, 0),
    (2, 24, 42, 256, 512, 1, 2, 0, 0),
    (1, 28, 28, 256, 256, 3, 1, 1, 0),
    (1, 7, 9, 64, 16, 1, 1, 0, 0),
    (1, 17, 23, 64, 64, 3, 2, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fprop(case):
    N, H, W, Cin, Cout, k, s, p, bn = case
    x, w = make_conv(N, H, W, Cin, Cout, k)
    ref = O.conv2d(x, w, None, s, p)
    y = K().conv2d_fprop(bf(nhwc(x)).to(dev()), bf(w.permute(0, 2, 3, 1).contiguous()).to(dev()), stride=s, pad=p,
                         bn_hint=bn)
    torch.cuda.synchronize()
    got = nchw(y.float().cpu())
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 4e-3, (case, rel_err(got, ref))


def test_conv2d_fprop_epilogue_variants():
    N, H, W, Cin, Cout = 2, 24, 42, 64, 128
    x, w = make_conv(N, H, W, Cin, Cout, 3, seed=1)
    g = torch.Generator().manual_seed(5)
    bias = torch.randn(Cout, generator=g)
    res = bf(torch.randn(N, Cout, H, W, generator=g)).float()
    res_half = bf(torch.randn(N, Cout, H // 2, W // 2, generator=g)).float()
    xd, wd = bf(nhwc(x)).to(dev()), bf(w.permute(0, 2, 3, 1).contiguous()).to(dev())
    base = O.conv2d(x, w, bias, 1, 1)
    # bias + relu, fp32 out
    y = K().conv2d_fprop(xd, wd, bias.to(dev()), stride=1, pad=1, relu=True, out_fp32=True)
    assert rel_err(nchw(y.cpu()), F.relu(base)) < 2e-5
    # bias + residual (same grid)
    y = K().conv2d_fprop(xd, wd, bias.to(dev()), bf(nhwc(res)).to(dev()), stride=1, pad=1, out_fp32=True)
    assert rel_err(nchw(y.cpu()), base + res) < 2e-5
    # bias + residual on the 2x coarser grid (FPN top-down: nearest upsample + add)
    y = K().conv2d_fprop(xd, wd, bias.to(dev()), bf(nhwc(res_half)).to(dev()), stride=1, pad=1, out_fp32=True,
                         res_half=True)
    up = F.interpolate(res_half, size=(H, W), mode="nearest")
    assert rel_err(nchw(y.cpu()), base + up) < 2e-5
    # GroupNorm statistics from the epilogue
    gn = torch.zeros(N, 32, 2, device=dev())
    y = K().conv2d_fprop(xd, wd, stride=1, pad=1, gn_sum=gn)
    ref = O.conv2d(x, w, None, 1, 1).reshape(N, 32, -1)
    torch.cuda.synchronize()
    assert rel_err(gn[..., 0].cpu(), ref.sum(-1)) < 1e-3
    assert rel_err(gn[..., 1].cpu(), (ref * ref).sum(-1)) < 1e-3
    # 1x1 flat path with statistics, two images
    x1, w1 = make_conv(2, 12, 21, 64, 256, 1, seed=2)
    gn = torch.zeros(2, 32, 2, device=dev())
    y = K().conv2d_fprop(bf(nhwc(x1)).to(dev()), bf(w1.permute(0, 2, 3, 1).contiguous()).to(dev()), gn_sum=gn)
    ref = O.conv2d(x1, w1).reshape(2, 32, -1)
    assert rel_err(nchw(y.float().cpu()), O.conv2d(x1, w1)) < 4e-3
    assert rel_err(gn[..., 0].cpu(), ref.sum(-1)) < 1e-3
    assert rel_err(gn[..., 1].cpu(), (ref * ref).sum(-1)) < 1e-3


@pytest.mark.parametrize("shape", [(3, 34, 64, 128, 256, 3), (2, 32, 64, 64, 512, 1), (1, 16, 128, 256, 256, 3)])
def test_conv2d_fprop_cta_pair(shape):
    """cta_group::2 kernel (bn_hint 512): 256 x 256 tiles on CTA pairs, incl. an odd trailing M tile (51 tiles),
    two column tiles (Cout 512), and every epilogue (bias/residual/ReLU, GroupNorm statistics)."""
    N, H, W, Cin, Cout, k = shape
    x, w = make_conv(N, H, W, Cin, Cout, k, seed=7)
    g = torch.Generator().manual_seed(6)
    bias = torch.randn(Cout, generator=g)
    res = bf(torch.randn(N, Cout, H, W, generator=g)).float()
    xd, wd = bf(nhwc(x)).to(dev()), bf(w.permute(0, 2, 3, 1).contiguous()).to(dev())
    base = O.conv2d(x, w, None, 1, k // 2)
    y = K().conv2d_fprop(xd, wd, stride=1, pad=k // 2, bn_hint=512)
    y1 = K().conv2d_fprop(xd, wd, stride=1, pad=k // 2, bn_hint=256)
    assert rel_err(nchw(y.float().cpu()), base) < 4e-3
    assert torch.equal(y, y1)                      # same products, same fp32 accumulation order per k block
    y = K().conv2d_fprop(xd, wd, bias.to(dev()), bf(nhwc(res)).to(dev()), stride=1, pad=k // 2, relu=True,
                         out_fp32=True, bn_hint=512)
    assert rel_err(nchw(y.cpu()), F.relu(base + bias[None, :, None, None] + res)) < 2e-5
    gn = torch.zeros(N, 32, 2, device=dev())
    K().conv2d_fprop(xd, wd, stride=1, pad=k // 2, gn_sum=gn, bn_hint=512)
    ref = base.reshape(N, 32, -1)
    assert rel_err(gn[..., 0].cpu(), ref.sum(-1)) < 1e-3
    assert rel_err(gn[..., 1].cpu(), (ref * ref).sum(-1)) < 1e-3


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_dgrad(case):
    N, H, W, Cin, Cout, k, s, p, bn = case
    if Cout % 64 != 0:
        pytest.skip("dgrad reduction (Cout) must be a multiple of 64")
    x, w = make_conv(N, H, W, Cin, Cout, k, seed=3)
    x.requires_grad_(True)
    y = O.conv2d(x, w, None, s, p)
    g = torch.Generator().manual_seed(9)
    dy = bf(torch.randn(y.shape, generator=g)).float()
    (ref,) = torch.autograd.grad(y, x, dy)
    wt = bf(w.permute(1, 2, 3, 0).contiguous()).to(dev())  # [Cin,KH,KW,Cout]
    dx = K().conv2d_dgrad(bf(nhwc(dy)).to(dev()), wt, (H, W), stride=s, pad=p, bn_hint=bn if Cin % max(bn, 1) == 0 else 0)
    got = nchw(dx.float().cpu())
    assert rel_err(got, ref) < 4e-3, (case, rel_err(got, ref))
    if s == 1:      # gradient of another branch accumulated in the epilogue (fp32 add, one rounding)
        acc = bf(torch.randn(ref.shape, generator=g)).float()
        dx2 = K().conv2d_dgrad(bf(nhwc(dy)).to(dev()), wt, (H, W), stride=s, pad=p, acc=bf(nhwc(acc)).to(dev()))
        assert rel_err(nchw(dx2.float().cpu()), ref + acc) < 4e-3


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_wgrad(case):
    N, H, W, Cin, Cout, k, s, p, bn = case
    x, w = make_conv(N, H, W, Cin, Cout, k, seed=4)
    w.requires_grad_(True)
    y = O.conv2d(x, w, None, s, p)
    g = torch.Generator().manual_seed(11)
    dy = bf(torch.randn(y.shape, generator=g)).float()
    (ref,) = torch.autograd.grad(y, w, dy)
    xd, dyd = bf(nhwc(x)).to(dev()), bf(nhwc(dy)).to(dev())
    dw = K().conv2d_wgrad(xd, dyd, (k, k), stride=s, pad=p)
    assert dw.shape == ref.shape
    if k > 1:   # KxK filters: channels_last memory order (vector-RED epilogue), same logical tensor
        assert dw.is_contiguous(memory_format=torch.channels_last) and not dw.is_contiguous()
    assert rel_err(dw.cpu(), ref) < 1e-4, (case, rel_err(dw.cpu(), ref))
    # torch-contiguous destination (scalar-RED epilogue) on request
    dw2 = K().conv2d_wgrad(xd, dyd, (k, k), stride=s, pad=p, channels_last=False)
    assert dw2.is_contiguous()
    assert rel_err(dw2.cpu(), ref) < 1e-4, (case, rel_err(dw2.cpu(), ref))


def test_gemm_wgrad_layouts():
    g = torch.Generator().manual_seed(2)
    R, Kin, Nout = 300, 49 * 64, 128  # fc6-like: x columns are (s, c) with c inner (64), torch wants (c, s)
    x = bf(torch.randn(R, Kin, generator=g)).float()
    dy = bf(torch.randn(R, Nout, generator=g)).float()
    ref = dy.t() @ x  # [Nout, (s,c)]
    ref_torch = ref.reshape(Nout, 49, 64).permute(0, 2, 1).reshape(Nout, Kin)  # [Nout, (c,s)]
    out = torch.zeros(Nout, Kin, device=dev())
    K().gemm_wgrad(bf(x).to(dev()), bf(dy).to(dev()), out, s_m=Kin, n_inner=64, s_n_inner=49, s_n_outer=1)
    assert rel_err(out.cpu(), ref_torch) < 1e-4
    out = torch.zeros(Nout, Kin, device=dev())
    K().gemm_wgrad(bf(x).to(dev()), bf(dy).to(dev()), out, s_m=Kin)
    assert rel_err(out.cpu(), ref) < 1e-4
    # stem layout: x = im2col with K = 49 taps x 3 channels zero-padded to 192 columns; the padding columns must not
    # be written at all (their destination offsets fall outside dw) -- checked with a guard band and non-finite dy,
    # which is what torch's CUDA-graph warm-up feeds through the backward
    R, T_, I_, Cout_ = 500, 49, 3, 64
    xs = torch.zeros(R, 192)
    xs[:, :T_ * I_] = bf(torch.randn(R, T_ * I_, generator=g)).float()
    dys = bf(torch.randn(R, Cout_, generator=g)).float()
    refs = (dys.t() @ xs[:, :T_ * I_]).reshape(Cout_, T_, I_).permute(0, 2, 1).reshape(Cout_, I_ * T_)
    buf = torch.zeros(Cout_ * I_ * T_ + 256, device=dev())
    outs = buf[:Cout_ * I_ * T_].view(Cout_, I_ * T_)
    K().gemm_wgrad(bf(xs).to(dev()), bf(dys).to(dev()), outs, s_m=I_ * T_, n_inner=I_, s_n_inner=T_, s_n_outer=1,
                   n_valid=T_ * I_)
    assert rel_err(outs.cpu(), refs) < 1e-4
    dys[0, :] = float("inf")
    buf.zero_()
    K().gemm_wgrad(bf(xs).to(dev()), bf(dys).to(dev()), outs, s_m=I_ * T_, n_inner=I_, s_n_inner=T_, s_n_outer=1,
                   n_valid=T_ * I_)
    assert float(buf[Cout_ * I_ * T_:].abs().max()) == 0.0          # nothing (not even NaN = 0 * inf) past the end


def test_deconv2x2():
    g = torch.Generator().manual_seed(6)
    N, h, w, Cin, Cout = 3, 28, 28, 256, 256
    x = bf(torch.randn(N, Cin, h, w, generator=g)).float().requires_grad_(True)
    wt = bf(torch.randn(Cin, Cout, 2, 2, generator=g) / 16).float().requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    y = F.relu(F.conv_transpose2d(x, wt, b, stride=2))
    dy = bf(torch.randn(y.shape, generator=g)).float()
    dy_eff = dy * (y > 0)
    rx, rw = torch.autograd.grad(y, (x, wt), dy)
    wd = bf(wt.detach().permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()).to(dev())      # [(dy,dx,co)][ci]
    wdt = bf(wt.detach().permute(0, 2, 3, 1).reshape(Cin, 4 * Cout).contiguous()).to(dev())     # [ci][(dy,dx,co)]
    xd = bf(nhwc(x.detach())).to(dev())
    yd = K().deconv2x2_fprop(xd, wd, b.repeat(4).to(dev()), relu=True)
    assert rel_err(nchw(yd.float().cpu()), y.detach()) < 4e-3
    dyd = bf(nhwc(dy_eff)).to(dev())
    dx = K().deconv2x2_dgrad(dyd, wdt)
    assert rel_err(nchw(dx.float().cpu()), rx) < 4e-3
    dw = K().deconv2x2_wgrad(xd, dyd)
    assert rel_err(dw.cpu(), rw) < 1e-4


@pytest.mark.parametrize("C,H,W,N", [(64, 48, 84, 2), (256, 24, 42, 1), (2048, 6, 11, 3), (128, 13, 7, 2)])
def test_groupnorm_fwd_bwd(C, H, W, N):
    g = torch.Generator().manual_seed(C)
    x = bf(torch.randn(N, C, H, W, generator=g) * 2 + 0.5).float().requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).requires_grad_(True)
    res = bf(torch.randn(N, C, H, W, generator=g)).float()
    k = K()
    xd = bf(nhwc(x.detach())).to(dev())
    sums = k.gn_stats(xd)
    for relu, use_res in [(True, False), (False, False), (True, True)]:
        y = O.group_norm(x, gamma, beta)
        if use_res:
            y = y + res
        if relu:
            y = F.relu(y)
        yd = k.gn_apply(xd, sums, gamma.detach().to(dev()), beta.detach().to(dev()),
                        bf(nhwc(res)).to(dev()) if use_res else None, relu=relu)
        assert rel_err(nchw(yd.float().cpu()), y.detach()) < 4e-3
        dy = bf(torch.randn(y.shape, generator=g)).float()
        rx, rg, rb = torch.autograd.grad(y, (x, gamma, beta), dy)
        mode = 0 if not relu else (2 if use_res else 1)
        # the kernel takes its ReLU mask from the bf16 output it produced itself (mode 2) or recomputes it
        # (mode 1); pixels where the oracle's pre-activation is within bf16 rounding of 0 may flip, hence 2e-2
        dx, dres, dg, db = k.gn_backward(xd, sums, gamma.detach().to(dev()), beta.detach().to(dev()),
                                         bf(nhwc(dy)).to(dev()), yout=yd if mode == 2 else None, mask_mode=mode,
                                         want_dres=use_res)
        assert rel_err(nchw(dx.float().cpu()), rx) < 2e-2, (relu, use_res, rel_err(nchw(dx.float().cpu()), rx))
        assert rel_err(dg.cpu(), rg) < 2e-2
        assert rel_err(db.cpu(), rb) < 2e-2
        if use_res:
            assert rel_err(nchw(dres.float().cpu()), dy * (y.detach() > 0)) < 2e-2


def test_roi_align_fwd_bwd():
    g = torch.Generator().manual_seed(3)
    N, C = 2, 256
    shapes = [(48, 84), (24, 42), (12, 21), (6, 11)]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    feats = [bf(torch.randn(N, C, h, w, generator=g)).float().requires_grad_(True) for h, w in shapes]
    R = 64
    ctr = torch.rand(R, 2, generator=g) * torch.tensor([336.0, 192.0])
    wh = torch.exp(torch.rand(R, 2, generator=g) * 5.0) + 2
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(min=0)
    boxes[:, 2].clamp_(max=336)
    boxes[:, 3].clamp_(max=192)
    rois = torch.cat([torch.randint(0, N, (R, 1), generator=g).float(), boxes], 1)
    k = K()
    for P in (7, 14):
        ref = O.roi_align_multiscale(feats, scales, rois, P)
        out = k.roi_align_fwd([bf(nhwc(f.detach())).to(dev()) for f in feats], scales, rois.to(dev()), P)
        got = out.float().cpu().permute(0, 3, 1, 2)
        assert rel_err(got, ref.detach()) < 4e-3
        dy = bf(torch.randn(ref.shape, generator=g)).float()
        refg = torch.autograd.grad(ref, feats, dy, allow_unused=True)
        outs = k.roi_align_bwd(bf(dy.permute(0, 2, 3, 1).contiguous()).to(dev()), [(N, h, w, C) for h, w in shapes],
                               scales, rois.to(dev()), P)
        for gg, rr in zip(outs, refg):
            rr = torch.zeros_like(feats[0]) if rr is None else rr
            if rr.abs().sum() == 0:
                assert gg.abs().sum().item() == 0
            else:
                assert rel_err(nchw(gg.cpu()), rr) < 1e-4


def test_mask_targets():
    g = torch.Generator().manual_seed(8)
    G, H, W, R, M = 3, 60, 90, 40, 28
    masks = (torch.rand(G, H, W, generator=g) > 0.6).to(torch.uint8)
    ctr = torch.rand(R, 2, generator=g) * torch.tensor([float(W), float(H)])
    wh = torch.rand(R, 2, generator=g) * 60 + 1
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    rois = torch.cat([torch.randint(0, G, (R, 1), generator=g).float(), boxes], 1)
    ref = O.mask_targets(masks, rois, M)
    got = K().mask_targets(masks.to(dev()), rois.to(dev()), M)
    assert (got.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("kind", ["LOVASZ", "BCE"])
def test_mask_loss(kind):
    g = torch.Generator().manual_seed(12)
    R, Cc, M = 5, 2, 56
    logits = (torch.randn(R, Cc, M, M, generator=g) * 2).requires_grad_(True)
    labels = torch.ones(R, dtype=torch.int64)
    targets = (torch.rand(R, M, M, generator=g) > 0.5).float()
    targets[0, :10] = torch.rand(10, M, generator=g)          # fractional RoIAlign targets
    if kind == "LOVASZ":
        targets[1, 5:9] = 37.0                                # > 1 -> ignored (255)
    ref = O.mask_loss(logits, labels, targets, kind)
    (rg,) = torch.autograd.grad(ref, logits)
    loss, dl = K().mask_loss(logits.detach().to(dev()), labels.to(dev()), targets.to(dev()), kind)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert rel_err(dl.cpu(), rg) < 1e-3


def test_mask_paste_threshold_and_bbox():
    g = torch.Generator().manual_seed(21)
    H, W, M = 480, 854, 56
    logits = torch.randn(2, 2, M, M, generator=g) * 3
    # smooth blob so that the thresholded mask is non-trivial
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, M), torch.linspace(-1, 1, M), indexing="ij")
    logits[:, 1] += 6 * (1 - 2 * (xx ** 2 + yy ** 2))
    labels = torch.ones(2, dtype=torch.int64)
    boxes = torch.tensor([[100.3, 50.7, 400.9, 300.2], [-20.5, 200.0, 120.25, 500.75]])
    k = K()
    for d in range(2):
        ref = O.paste_probs(logits[d:d + 1], labels[d:d + 1], boxes[d:d + 1], H, W)  # [1,1,H,W]
        ref_t = O.threshold_targets(ref)
        chan = torch.tensor([d], dtype=torch.int32)
        probs, target, stats = k.mask_paste_threshold(logits.to(dev()), chan.to(dev()), labels.to(dev()),
                                                      boxes.to(dev()), 1, 1, H, W)
        assert (probs.cpu() - ref).abs().max().item() < 1e-5
        flips = (target.cpu() != ref_t) & ((ref - 0.5).abs() > 1e-5)
        assert flips.sum().item() == 0
        rb = O.mask_to_boxes(target.cpu()[0, 0], [1.0])[0]
        st = stats.cpu()[0, 0]
        assert [st[0].item(), st[1].item(), st[2].item() + 1, st[3].item() + 1] == rb.tolist()
        assert st[4].item() == int((target.cpu() == 1).sum())
        st2 = k.mask_to_bbox(target, 1).cpu()[0, 0]
        assert st2.tolist() == st.tolist()
    # no detection: all background
    probs, target, stats = k.mask_paste_threshold(logits.to(dev()), torch.tensor([-1], dtype=torch.int32).to(dev()),
                                                  labels.to(dev()), boxes.to(dev()), 1, 1, H, W)
    assert probs.abs().sum().item() == 0 and target.abs().sum().item() == 0 and stats.cpu()[0, 0, 4].item() == 0


def test_meta_update_bit_exact():
    g = torch.Generator().manual_seed(1)
    shapes = [(64, 3, 7, 7), (64,), (256, 64, 1, 1), (1024, 12544), (1024,), (2, 1024), (256, 256, 3, 3), (5,), (3, 1)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [torch.randn(s, generator=g) for s in shapes]
    lrs = [torch.rand((s[0],) + (1,) * (len(s) - 1), generator=g) * 1e-3 for s in shapes]
    k = K()
    for use_log in (False, True):
        lr_in = [l.log() for l in lrs] if use_log else lrs
        ref = O.meta_update(params, grads, lr_in, use_log)
        dp = [p.to(dev()) for p in params]
        dg = [x.to(dev()) for x in grads]
        dl = [x.to(dev()) for x in lr_in]
        do = [torch.empty_like(p) for p in dp]
        k.meta_update(k.MetaUpdatePlan(dp, dg, dl, do), use_log)
        # same update with the KxK filter gradients in channels_last memory order (what conv2d_wgrad emits) and
        # written IN PLACE over the parameters (what MetaOptimizer.step does from the second step on)
        dg_cl = [x.contiguous(memory_format=torch.channels_last) if x.dim() == 4 else x for x in dg]
        assert not dg_cl[6].is_contiguous()
        dp2 = [p.clone() for p in dp]
        k.meta_update(k.MetaUpdatePlan(dp2, dg_cl, dl, dp2), use_log)
        for o, o2, r in zip(do, dp2, ref):
            assert torch.equal(o, o2)
            if use_log:  # expf on device vs CPU differs in the last ulp of lr
                assert torch.allclose(o.cpu(), r, rtol=1e-6, atol=1e-9)
            else:        # bit exact: same fp32 multiply and subtract
                assert torch.equal(o.cpu(), r)


def test_weight_prep_layouts():
    """Tiled multi-tensor operand preparation (one launch) == torch permute + cast, for every layout kind."""
    from eosvos_b200 import ops
    g = torch.Generator().manual_seed(8)
    act = K().ACT_DTYPE

    def rnd(*s):
        return torch.randn(*s, generator=g).to(dev())
    w3, w1, w7 = rnd(96, 40, 3, 3), rnd(130, 72, 1, 1), rnd(64, 3, 7, 7)
    fc6, fc7, dc = rnd(72, 49 * 24), rnd(40, 72), rnd(24, 16, 2, 2)
    reqs = [(w3, "f"), (w3, "t"), (w1, "f"), (w1, "t"), (w7, "stem"), (fc6, ("lf", 24)), (fc6, ("lt", 24)),
            (fc7, ("lf", 0)), (fc7, ("lt", 0)), (dc, "dc")]
    ops.clear_prep_cache()
    ops.prep_many(reqs)
    got = [ops._cache_get(w, k) for w, k in reqs]
    ref = [w3.permute(0, 2, 3, 1), w3.permute(1, 2, 3, 0), w1.permute(0, 2, 3, 1), w1.permute(1, 2, 3, 0)]
    for a, b in zip(got[:4], ref):
        assert torch.equal(a, b.contiguous().to(act))
    stem = torch.zeros(64, 192, device=dev())
    stem[:, :147] = w7.permute(0, 2, 3, 1).reshape(64, 147)
    assert torch.equal(got[4].view(64, 192), stem.to(act))
    f6 = fc6.view(72, 24, 49).permute(0, 2, 1).reshape(72, 49 * 24)           # (c, s) -> (s, c)
    assert torch.equal(got[5].view(72, -1), f6.to(act))
    assert torch.equal(got[6].view(49 * 24, 72), f6.t().contiguous().to(act))
    assert torch.equal(got[7].view(40, 72), fc7.to(act))
    assert torch.equal(got[8].view(72, 40), fc7.t().contiguous().to(act))
    wf, wt = got[9]
    assert torch.equal(wf.view(4, 16, 24), dc.permute(2, 3, 1, 0).reshape(4, 16, 24).to(act))
    assert torch.equal(wt.view(24, 4, 16), dc.permute(0, 2, 3, 1).reshape(24, 4, 16).to(act))
    # the same layouts through a persistent plan (what the CUDA-graphed trunk captures)
    plan, vals = ops.build_prep_plan(reqs[:5])
    plan.launch()
    for i, (w, k) in enumerate(reqs[:5]):
        assert torch.equal(vals[(i, k)], got[i])


def test_transform_and_stem():
    g = torch.Generator().manual_seed(4)
    img = torch.rand(2, 3, 120, 214, generator=g)
    ref, (oh, ow) = O.transform_image(img, min_size=200, max_size=333)
    Hp, Wp = ref.shape[-2:]
    k = K()
    out = k.transform(img.to(dev()), oh, ow, Hp, Wp, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    got = out.float().cpu()[..., :3].permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() < 0.02  # one bf16 rounding of values up to ~2.6
    assert out.float().cpu()[..., 3:].abs().sum().item() == 0
    # stem: im2col + flat GEMM == 7x7/2 conv
    w = bf(torch.randn(64, 3, 7, 7, generator=g) / 12).float()
    col, Ho, Wo = k.im2col_stem(out)
    w2 = torch.zeros(64, 192)
    w2[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)
    y = k.conv2d_fprop(col.view(1, 1, -1, 192), bf(w2).view(64, 1, 1, 192).to(dev()))
    refc = O.conv2d(got, w, None, 2, 3)
    assert rel_err(y.float().cpu().view(2, Ho, Wo, 64).permute(0, 3, 1, 2), refc) < 4e-3
    # pooling helpers
    x = bf(torch.randn(2, 64, 23, 31, generator=g)).float().requires_grad_(True)
    yp = F.max_pool2d(x, 3, 2, 1)
    xd = bf(nhwc(x.detach())).to(dev())
    yd, arg = k.maxpool_fwd(xd)
    assert torch.equal(nchw(yd.float().cpu()), yp.detach())
    dy = bf(torch.randn(yp.shape, generator=g)).float()
    (rx,) = torch.autograd.grad(yp, x, dy)
    dx = k.maxpool_bwd(arg, bf(nhwc(dy)).to(dev()), tuple(xd.shape))
    assert rel_err(nchw(dx.float().cpu()), rx) < 4e-3
    s = k.subsample2(xd)
    assert torch.equal(nchw(s.float().cpu()), F.max_pool2d(x.detach(), 1, 2, 0))
    u = k.sum2x2(bf(nhwc(x.detach()[:, :, :22, :30])).to(dev()))
    assert rel_err(nchw(u.float().cpu()), F.avg_pool2d(x.detach()[:, :, :22, :30], 2) * 4) < 4e-3
    c = k.colsum(xd.view(-1, 64))
    assert rel_err(c.cpu(), x.detach().sum((0, 2, 3))) < 1e-4


def test_nms_segments_matches_torchvision():
    """Segmented NMS == torchvision CPU nms run independently per segment (exact index sets)."""
    import torchvision
    g = torch.Generator().manual_seed(31)
    sizes = [2000, 2000, 1500, 700, 63, 1, 0, 130]
    boxes, offs = [], [0]
    for n in sizes:
        ctr = torch.rand(n, 2, generator=g) * torch.tensor([1333.0, 749.0])
        wh = torch.exp(torch.rand(n, 2, generator=g) * 4.5) + 1
        b = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        if n:
            b[::7] = b[:1].clone()                # clusters of identical boxes
        boxes.append(b)
        offs.append(offs[-1] + n)
    allb = torch.cat(boxes).contiguous()
    flags = K().nms_segments(allb.to(dev()), torch.tensor(offs, dtype=torch.int32, device=dev()), len(sizes), max(sizes),
                             0.7).cpu().bool()
    for i, n in enumerate(sizes):
        scores = torch.arange(n, 0, -1, dtype=torch.float32)     # already sorted: descending score = position
        ref = torchvision.ops.nms(boxes[i], scores, 0.7) if n else torch.zeros(0, dtype=torch.int64)
        got = torch.nonzero(flags[offs[i]:offs[i + 1]]).flatten()
        assert torch.equal(got, ref.sort()[0]), (i, n, got.numel(), ref.numel())


def test_device_augmentation_matches_cv2_family():
    """GPU bicubic flip+rotate/scale vs cv2.warpAffine(INTER_CUBIC): same transform; cv2 snaps source coordinates
    to a 1/32-px grid, so agreement is to ~1e-2 on a smooth image (mean abs < 2e-3), labels identical."""
    import random
    import numpy as np
    from eosvos_b200.util import augment, synthetic
    frames, labels = synthetic.make_video(3, 1, 240, 427, 1)
    img = frames[0].astype(np.float32) / 255.0
    gt = (labels[0] == 1).astype(np.float32)
    random.seed(7)
    ref = [augment.augment_first_frame(img, gt) for _ in range(3)]
    random.seed(7)
    aug = augment.DeviceAugmenter(torch.from_numpy(img.transpose(2, 0, 1).copy()).to(dev()), gt)
    x, g = aug.batch(3)
    for b in range(3):
        ri = torch.from_numpy(np.ascontiguousarray(ref[b][0].transpose(2, 0, 1)))
        assert torch.equal(g[b, 0].cpu(), torch.from_numpy(np.ascontiguousarray(ref[b][1])))
        d = (x[b].cpu() - ri).abs()
        assert d.mean().item() < 2e-3 and d.max().item() < 0.25, (d.mean().item(), d.max().item())


def _label_map(h, w, seed, ids=3):
    import numpy as np
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:h, :w]
    gt = np.zeros((h, w), np.float32)
    for k in range(ids):
        cy, cx, r = rng.randint(0, h), rng.randint(0, w), rng.randint(min(h, w) // 10, min(h, w) // 3)
        gt[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = k + 1
    return gt


@pytest.mark.parametrize("h,w", [(480, 854), (720, 1280), (97, 131)])
def test_label_warp_nearest_equals_cv2(h, w):
    """label_warp_nearest_kernel == cv2.warpAffine(flip(gt), M, flags=INTER_NEAREST) bit for bit (OpenCV's 10-bit
    fixed-point coordinates, int16 saturation, zero border) on id maps and on a noise image, 24 random transforms."""
    import random
    import cv2
    import numpy as np
    from eosvos_b200.util import augment
    rng = random.Random(h)
    for src in (_label_map(h, w, 1), (np.random.RandomState(2).rand(h, w) * 5).astype(np.int32).astype(np.float32)):
        src_d = torch.from_numpy(src).to(dev())
        Ms, flips, refs = [], [], []
        for _ in range(12):
            rot, sc, fl = 60 * rng.random() - 30, 0.5 * rng.random() + 0.75, rng.random() < 0.5
            M = cv2.getRotationMatrix2D((w / 2, h / 2), rot, sc)
            refs.append(cv2.warpAffine(cv2.flip(src, 1) if fl else src, M, (w, h), flags=cv2.INTER_NEAREST))
            Ms.append(augment.DeviceAugmenter.cv_inverse(M))
            flips.append(int(fl))
        out = K().label_warp_nearest(src_d, torch.from_numpy(np.stack(Ms)).to(dev()),
                                     torch.tensor(flips, dtype=torch.int32, device=dev()))
        for b, ref in enumerate(refs):
            assert torch.equal(out[b, 0].cpu(), torch.from_numpy(ref)), b


def test_device_label_augmentation_equals_host_path():
    """DeviceAugmenter.device_labels (draws on the host, label warp / rejection test / boxes on the GPU) against
    host_part + target_stats (cv2 + numpy): same flips, matrices, labels and boxes from the same random stream --
    including a label whose small corner object is pushed out of the frame by some draws (rejection loop,
    custom_transforms.py:74-78) -- and the prefetching augmenter's two modes hand out identical batches."""
    import random
    import numpy as np
    from eosvos_b200.util import augment
    h, w = 240, 427
    gt = np.zeros((h, w), np.float32)
    gt[4:22, 6:30] = 1                                   # leaves the frame under many rotations / scalings
    img = np.random.RandomState(0).rand(h, w, 3).astype(np.float32)
    src = torch.from_numpy(img.transpose(2, 0, 1).copy()).to(dev())
    aug = augment.DeviceAugmenter(src, gt)
    draws = []

    class Counting(random.Random):
        def random(self):
            draws.append(1)
            return super().random()
    minv_h, flips_h, gts_h = aug.host_part(6, Counting(11))
    n_host = len(draws)
    del draws[:]
    minv_d, flips_d, gts_d, stats_d = aug.device_labels(6, Counting(11))
    assert len(draws) == n_host and n_host > 6 * 3       # the rejection loop ran, and ran equally often
    assert np.array_equal(minv_h, minv_d) and np.array_equal(flips_h, flips_d)
    assert torch.equal(gts_d.cpu(), torch.from_numpy(gts_h))
    st_h, _ = aug.target_stats(gts_h)
    assert torch.equal(stats_d, st_h)
    # multi-id label map: every id must survive the warp
    gt3 = _label_map(h, w, 5)
    aug3 = augment.DeviceAugmenter(src, gt3)
    a = aug3.host_part(4, random.Random(3))
    b = aug3.device_labels(4, random.Random(3))
    assert np.array_equal(a[0], b[0]) and torch.equal(b[2].cpu(), torch.from_numpy(a[2]))
    assert torch.equal(b[3], aug3.target_stats(a[2], num_ids=3)[0])
    # the prefetching augmenter (worker thread + its own stream) in both modes
    pa = augment.PrefetchingAugmenter(src, gt, 3, lambda e: 100 + e, device_labels=True)
    pb = augment.PrefetchingAugmenter(src, gt, 3, lambda e: 100 + e, device_labels=False)
    for e in (1, 2, 3, 4, 5):
        xa, ga = pa.get(e)
        xb, gb = pb.get(e)
        torch.cuda.synchronize()
        assert torch.equal(ga, gb) and torch.equal(xa, xb)
        sa, sb = K().target_stats.get_host(ga), K().target_stats.get_host(gb)
        assert sa is not None and torch.equal(sa[0], sb[0])
    pa.close()
    pb.close()


# ------------------------------------------------------------------------------------------------ K5/K6 (rpn.cu)
def _rpn_case(N, feat_shapes, seed):
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    g = torch.Generator().manual_seed(seed)
    A = 3
    heads = [(torch.randn(N * h * w, 16, generator=g) * torch.tensor([2.0] * 3 + [0.3] * 12 + [0.0])).to(dev())
             for h, w in feat_shapes]
    ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)

    class IL:
        pass
    il = IL()
    stride = 4
    Hp, Wp = feat_shapes[0][0] * stride, feat_shapes[0][1] * stride
    il.tensors = torch.empty((N, 3, Hp, Wp), device="meta")
    il.image_sizes = [(Hp - 19, Wp - 11)] * N
    fms = [torch.empty((N, 1, h, w), device=dev()) for h, w in feat_shapes]
    anchors = ag(il, fms)[0].contiguous()
    return heads, anchors, il.image_sizes, A


@pytest.mark.parametrize("N,pre,post", [(3, 2000, 2000), (1, 1000, 500)])
def test_rpn_select_nms_postnms_match_torchvision(N, pre, post):
    """rpn_select + nms_segments + rpn_postnms (csrc/rpn.cu) against tv rpn.py filter_proposals restated with torch /
    torchvision ops (decode of ALL anchors, per-level top-k, clip, filters, batched_nms, top-n): same boxes in the same
    order.  Box arithmetic follows the ATen op sequence (no FMA contraction) => equal to float rounding of exp()."""
    import torchvision
    from torchvision.models.detection._utils import BoxCoder
    feat_shapes = [(192, 336), (96, 168), (48, 84), (24, 42), (12, 21)]
    heads, anchors, image_sizes, A = _rpn_case(N, feat_shapes, 5)
    hw = [h * w for h, w in feat_shapes]
    clip = math.log(1000.0 / 16)
    boxes_c, scores_c, valid_c, ks = K().rpn_select(heads, hw, A, N, anchors, image_sizes, pre, clip, 1e-3, 0.0)
    C = sum(ks)
    offs = [0]
    for _ in range(N):
        for k in ks:
            offs.append(offs[-1] + k)
    seg = torch.tensor(offs, dtype=torch.int32, device=dev())
    keep = K().nms_segments(boxes_c.view(-1, 4), seg, N * len(ks), max(ks), 0.7)
    out, _, count = K().rpn_postnms(hw, A, N, pre, boxes_c, scores_c, valid_c, keep, post)
    # reference
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0))
    obj = torch.cat([h[:, :A].reshape(N, -1) for h in heads], 1)
    dlt = torch.cat([h[:, A:5 * A].reshape(N, -1, 4) for h in heads], 1)
    for n in range(N):
        props = coder.decode_single(dlt[n], anchors)
        lvl_boxes, lvl_scores, lvl_ids = [], [], []
        o = 0
        for l, cnt in enumerate(hw):
            cnt *= A
            k = min(pre, cnt)
            sc, idx = obj[n, o:o + cnt].topk(k)
            lvl_boxes.append(props[o + idx])
            lvl_scores.append(torch.sigmoid(sc))
            lvl_ids.append(torch.full((k,), l, device=dev()))
            # the kernel's per-level selection: same set, same order (descending objectness)
            got = boxes_c[n, sum(ks[:l]):sum(ks[:l]) + k]
            o += cnt
        b, s, ids = torch.cat(lvl_boxes), torch.cat(lvl_scores), torch.cat(lvl_ids)
        b = torchvision.ops.clip_boxes_to_image(b, image_sizes[n])
        kp = torchvision.ops.remove_small_boxes(b, 1e-3)
        b, s, ids = b[kp], s[kp], ids[kp]
        kp = torchvision.ops.batched_nms(b, s, ids, 0.7)[:post]
        want = b[kp]
        c = int(count[n])
        assert c == want.shape[0], (c, want.shape[0])
        d = (out[n, :c] - want).abs().max().item()
        assert d <= 2e-3, d
        assert out[n, c:].abs().sum().item() == 0


def test_extend_boxes_bit_equal_to_reference_arithmetic():
    """extend_boxes_kernel against reference mask_rcnn.py:262-285 evaluated with torch CPU ops on the same uniforms."""
    g = torch.Generator().manual_seed(3)
    B, G, n_aug = 2, 1, 500
    stats = torch.tensor([[[100, 50, 420, 330, 9000]], [[0, 0, 853, 479, 1]]], dtype=torch.int32)
    rnd = torch.rand(B, G, 4, n_aug, generator=g)
    h, w, oh, ow, Hp, Wp = 480, 854, 749, 1333, 768, 1344
    rw = float(torch.tensor(ow, dtype=torch.float32) / torch.tensor(w, dtype=torch.float32))
    rh = float(torch.tensor(oh, dtype=torch.float32) / torch.tensor(h, dtype=torch.float32))
    out = torch.zeros(B, 1000, 4, device=dev())
    K().extend_boxes(stats.to(dev()), None, rnd.to(dev()), n_aug, rw, rh, Wp, Hp, 0.1, out, 500)
    for b in range(B):
        st = stats[b, 0]
        box = torch.tensor([st[0], st[1], st[2] + 1, st[3] + 1], dtype=torch.float32)
        box = torch.stack((box[0] * rw, box[1] * rh, box[2] * rw, box[3] * rh))
        bw, bh = box[2] - box[0], box[3] - box[1]
        x_mins = box[0] - rnd[b, 0, 0] * bw * 0.1
        y_mins = box[1] - rnd[b, 0, 1] * bh * 0.1
        x_maxs = box[2] + rnd[b, 0, 2] * bw * 0.1
        y_maxs = box[3] + rnd[b, 0, 3] * bh * 0.1
        want = torch.stack([x_mins.clamp(0, Wp), y_mins.clamp(0, Hp), x_maxs.clamp(0, Wp), y_maxs.clamp(0, Hp)], dim=1)
        assert torch.equal(out[b, 500:].cpu(), want)
        assert out[b, :500].abs().sum().item() == 0
    # empty target -> fallback box
    empty = torch.tensor([[[2147483647, 2147483647, -1, -1, 0]]], dtype=torch.int32).to(dev())
    out2 = torch.zeros(1, 500, 4, device=dev())
    K().extend_boxes(empty, stats[:1].to(dev()), rnd[:1].to(dev()), n_aug, rw, rh, Wp, Hp, 0.1, out2, 0)
    assert torch.equal(out2[0], out[0, 500:])


def test_det_top1_matches_postprocess_detections():
    """det_top1_kernel against tv-style postprocess_detections (softmax, decode, clip, score / size filter, NMS, first
    detection) evaluated with torch / torchvision ops."""
    import torchvision
    from torchvision.models.detection._utils import BoxCoder
    g = torch.Generator().manual_seed(9)
    B, R, nc = 2, 1000, 2
    head = torch.zeros(B * R, 16)
    head[:, :nc] = torch.randn(B * R, nc, generator=g) * 2
    head[:, nc:nc + 4 * nc] = torch.randn(B * R, 4 * nc, generator=g)
    ctr = torch.rand(B * R, 2, generator=g) * torch.tensor([1333.0, 749.0])
    wh = torch.rand(B * R, 2, generator=g) * 300 + 2
    props = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(min=0)
    props[7] = 0                       # a padding row
    coder = BoxCoder((10.0, 10.0, 5.0, 5.0))
    for thr in (0.05, 0.5, 0.9999):
        det = K().det_top1(head.to(dev()), props.to(dev()), B, R, nc, (10.0, 10.0, 5.0, 5.0), math.log(1000.0 / 16), thr,
                           1e-2, 1333.0, 749.0, 0.64, 0.64)
        for b in range(B):
            hb, pb = head[b * R:(b + 1) * R], props[b * R:(b + 1) * R]
            boxes = coder.decode(hb[:, nc:nc + 4 * nc], [pb]).reshape(R, nc, 4)
            scores = F.softmax(hb[:, :nc], -1)
            boxes = torchvision.ops.clip_boxes_to_image(boxes, (749, 1333))[:, 1:].reshape(-1, 4)
            scores = scores[:, 1:].flatten()
            inds = torch.nonzero(scores > thr).squeeze(1)
            bx, sc = boxes[inds], scores[inds]
            kp = torchvision.ops.remove_small_boxes(bx, 1e-2)
            bx, sc, inds = bx[kp], sc[kp], inds[kp]
            if sc.numel() == 0:
                assert int(det["row"][b]) == -1 and int(det["chan"][b]) == -1 and float(det["roi"][b, 0]) == -1.0
                assert det["box"][b].abs().sum().item() == 0
                continue
            top = sc.argsort(descending=True, stable=True)[0]
            assert int(det["row"][b]) == int(inds[top])
            assert abs(float(det["score"][b]) - float(sc[top])) <= 1e-6
            assert (det["roi"][b, 1:].cpu() - bx[top]).abs().max().item() <= 1e-3
            assert (det["box"][b].cpu() - bx[top] * 0.64).abs().max().item() <= 1e-3
            assert int(det["label"][b]) == 1 and int(det["chan"][b]) == b and float(det["roi"][b, 0]) == float(b)


def test_roi_match_and_encode_match_torchvision():
    """roi_match_kernel / roi_encode_kernel against tv roi_heads.py assign_targets_to_proposals (box_iou + Matcher 0.5 /
    0.5) and box_coder.encode on the same padded proposal list with the ground-truth boxes appended."""
    import torchvision
    from torchvision.models.detection._utils import BoxCoder, Matcher
    g = torch.Generator().manual_seed(4)
    B, P = 3, 2000
    gts = [torch.tensor([[100.0, 80.0, 400.0, 300.0]]), torch.tensor([[50.0, 60.0, 300.0, 500.0], [600.0, 100.0, 900.0, 400.0]]),
           torch.tensor([[700.0, 300.0, 1200.0, 700.0]])]
    gls = [torch.tensor([1]), torch.tensor([1, 1]), torch.tensor([1])]
    counts = torch.tensor([2000, 1500, 37], dtype=torch.int32)
    props = torch.zeros(B, P, 4)
    for b in range(B):
        n = int(counts[b])
        gt = gts[b][torch.randint(0, gts[b].shape[0], (n,), generator=g)]
        props[b, :n] = (gt + torch.randn(n, 4, generator=g) * 60).clamp(min=0)
        props[b, :n, 2:] = torch.maximum(props[b, :n, 2:], props[b, :n, :2] + 1)
    gt_cat, gl_cat = torch.cat(gts), torch.cat(gls)
    gt_off = torch.tensor([0, 1, 3, 4], dtype=torch.int32)
    all_boxes, labels, matched, cnt = K().roi_match(props.to(dev()), counts.to(dev()), gt_cat.to(dev()), gl_cat.to(dev()),
                                                    gt_off.to(dev()), 2, 0.5)
    matcher = Matcher(0.5, 0.5, allow_low_quality_matches=False)
    for b in range(B):
        n, G = int(counts[b]), gts[b].shape[0]
        pl = torch.cat([props[b, :n], gts[b]])
        m = matcher(torchvision.ops.box_iou(gts[b], pl))
        lab = gls[b][m.clamp(min=0)]
        lab[m == Matcher.BELOW_LOW_THRESHOLD] = 0
        got_l = torch.cat([labels[b, :n], labels[b, P:P + G]]).cpu()
        got_m = torch.cat([matched[b, :n], matched[b, P:P + G]]).cpu()
        assert torch.equal(got_l, lab)
        fg = lab > 0
        assert torch.equal(got_m[fg], m[fg])
        assert (labels[b, n:P] == -1).all() and (labels[b, P + G:] == -1).all()
        assert cnt[b].tolist() == [int(fg.sum()), int((lab == 0).sum())]
    inds = torch.stack([torch.randperm(int(counts[b]), generator=g)[:32] for b in range(B)]).to(dev())
    rois5, ol, om, reg = K().roi_encode(all_boxes, labels, matched, gt_cat.to(dev()), gt_off.to(dev()), inds,
                                       (10.0, 10.0, 5.0, 5.0))
    coder = BoxCoder((10.0, 10.0, 5.0, 5.0))
    for b in range(B):
        pb = props[b][inds[b].cpu()]
        mb = matched[b].cpu()[inds[b].cpu()]
        want = coder.encode_single(gts[b][mb], pb)
        got = reg[b * 32:(b + 1) * 32].cpu()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
        assert torch.equal(rois5[b * 32:(b + 1) * 32, 1:].cpu(), pb) and (rois5[b * 32:(b + 1) * 32, 0] == b).all()


def test_rpn_anchor_match_sampler_and_loss_match_torchvision():
    """anchor_match_kernel (tv rpn.py assign_targets_to_anchors: box_iou + Matcher(0.7, 0.3, low-quality matches)),
    roi_sample_kernel as the RPN sampler over 257,796 anchors per image, and rpn_loss_kernel (forward + gradient, with
    the regression targets encoded on the fly) against the torchvision / torch formulation on concatenated tensors."""
    import torchvision
    from torchvision.models.detection._utils import BoxCoder, Matcher
    from eosvos_b200 import ops
    N, A = 2, 3
    feat_shapes = [(192, 336), (96, 168), (48, 84), (24, 42), (12, 21)]
    heads, anchors, image_sizes, _ = _rpn_case(N, feat_shapes, 6)
    heads = [h.requires_grad_(True) for h in heads]
    hw = [h * w for h, w in feat_shapes]
    gts = [torch.tensor([[300.0, 200.0, 700.0, 520.0]]), torch.tensor([[50.0, 60.0, 180.0, 400.0], [900.0, 100.0, 1300.0, 700.0]])]
    gt_cat = torch.cat(gts).to(dev())
    gt_off = torch.tensor([0, 1, 3], dtype=torch.int32, device=dev())
    labels, matched, counts = K().rpn_anchor_match(anchors, gt_cat, gt_off, N, 0.7, 0.3)
    matcher = Matcher(0.7, 0.3, allow_low_quality_matches=True)
    ref_labels, ref_reg = [], []
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0))
    for b in range(N):
        g = gts[b].to(dev())
        m = matcher(torchvision.ops.box_iou(g, anchors))
        lab = (m >= 0).to(torch.int64)
        lab[m == Matcher.BELOW_LOW_THRESHOLD] = 0
        lab[m == Matcher.BETWEEN_THRESHOLDS] = -1
        assert torch.equal(labels[b], lab)
        fg = lab == 1
        assert torch.equal(matched[b][fg].to(torch.int64), m[fg])
        assert counts[b].tolist() == [int(fg.sum()), int((lab == 0).sum())]
        ref_labels.append(lab)
        ref_reg.append(coder.encode_single(g[m.clamp(min=0)], anchors))
    # sampler: the reference's where(mask) of randperm draws
    gen = torch.Generator().manual_seed(2)
    perms, npos_l, nneg_l, want = [], [], [], []
    A_total = anchors.shape[0]
    for b in range(N):
        npos, nneg = counts[b].tolist()
        num_pos = min(npos, 128)
        num_neg = min(nneg, 256 - num_pos)
        pp, pn = torch.randperm(npos, generator=gen).to(dev()), torch.randperm(nneg, generator=gen).to(dev())
        perms.append((pp, pn))
        npos_l.append(num_pos)
        nneg_l.append(num_neg)
        positive = torch.where(ref_labels[b] >= 1)[0]
        negative = torch.where(ref_labels[b] == 0)[0]
        mask = torch.zeros(A_total, dtype=torch.bool, device=dev())
        mask[positive[pp[:num_pos]]] = True
        mask[negative[pn[:num_neg]]] = True
        want.append(torch.where(mask)[0])
    inds, pos_in = K().roi_sample(labels, perms, npos_l, nneg_l, 256, 128)
    for b in range(N):
        sz = npos_l[b] + nneg_l[b]
        assert torch.equal(inds[b, :sz], want[b]) and (inds[b, sz:] == -1).all()
        assert torch.equal(pos_in[b, :npos_l[b]], torch.where(ref_labels[b][want[b]] >= 1)[0])
    sampled = torch.cat([want[b] + b * A_total for b in range(N)])
    lo, lb = ops.rpn_loss(heads, hw, A, sampled, labels, matched, anchors, gt_cat, gt_off)
    (lo * 1.3 + lb * 0.7).backward()
    got = [h.grad.clone() for h in heads]
    ref_heads = [h.detach().clone().requires_grad_(True) for h in heads]
    obj = torch.cat([h[:, :A].reshape(N, -1, 1) for h in ref_heads], 1).flatten(0, -2)
    dlt = torch.cat([h[:, A:5 * A].reshape(N, -1, 4) for h in ref_heads], 1).flatten(0, -2)
    lab_c = torch.cat(ref_labels).float()
    reg_c = torch.cat(ref_reg)
    pos = sampled[lab_c[sampled] >= 1]
    rb = F.smooth_l1_loss(dlt[pos], reg_c[pos], beta=1 / 9, reduction="sum") / sampled.numel()
    ro = F.binary_cross_entropy_with_logits(obj.flatten()[sampled], lab_c[sampled])
    (ro * 1.3 + rb * 0.7).backward()
    assert torch.allclose(lo, ro, rtol=1e-5) and torch.allclose(lb, rb, rtol=1e-4)
    for a_, b_ in zip(got, ref_heads):
        assert torch.allclose(a_, b_.grad, rtol=1e-3, atol=1e-8)


# ------------------------------------------------------------------------------------------- J / F on the device
def _blob_sequence(T, H, W, K, seed):
    """Object-id maps made of drifting discs (some touching the frame border, one object absent in some frames)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:H, :W]
    seq = np.zeros((T, H, W), np.uint8)
    cen = rng.rand(K, 2) * [H, W]
    vel = rng.randn(K, 2) * 3
    rad = rng.randint(6, max(7, min(H, W) // 4), K)
    for t in range(T):
        for k in range(K):
            if k == 1 and t % 3 == 0:
                continue
            c = cen[k] + vel[k] * t
            seq[t][(yy - c[0]) ** 2 + (xx - c[1]) ** 2 <= rad[k] ** 2] = k + 1
    return seq


@pytest.mark.parametrize("T,H,W,K", [(6, 120, 213, 3), (4, 97, 131, 10), (3, 480, 854, 2), (5, 64, 64, 1)])
def test_jf_device_matches_host(T, H, W, K):
    """DAVIS J / F with the pixel work on the device == util.metrics' host path (integer counts => exact)."""
    import numpy as np
    from eosvos_b200.util import metrics
    gt = _blob_sequence(T, H, W, K, 1)
    pred = _blob_sequence(T, H, W, K, 2)
    pred[1] = gt[1]                                   # a perfect frame
    pred[2][pred[2] == 1] = 0                         # object 1 missing from the prediction in one frame
    host = metrics.evaluate_sequence_jf(pred, gt, K)
    dev = metrics.evaluate_sequence_jf_device(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), K)
    for m in ("J", "F"):
        assert len(dev[m]) == K
        np.testing.assert_allclose(np.asarray(dev[m]), np.asarray(host[m]), rtol=0, atol=1e-12, equal_nan=True)


def test_jf_counts_rejects_bad_arguments():
    p = torch.zeros((2, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        K().jf_counts(p, p[:1], 1, 2)
    with pytest.raises(RuntimeError):
        K().jf_counts(p, p, 1, 100)
